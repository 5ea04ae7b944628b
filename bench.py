#!/usr/bin/env python
"""bench.py -- likelihood evaluations per second on the ST-U NSX workload.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step is one pass of the hot path (pulse integration of both hot regions,
energy integration, response fold, background-marginalised likelihood) over a
batch of parameter vectors.  ``value`` times the kernels with the batch already
resident in HBM; ``e2e`` times the public call with pinned HOST buffers (H2D of
the integrator inputs and D2H of lnL/status inside the timed region).

For N>1 the driver launches one rank per GPU (torchrun); the batch is sharded
by contiguous blocks (weak scaling: the per-GPU batch is fixed), there is no
data-path collective, and the per-rank lnL blocks are collected with one NCCL
all_gather after the timed region's last kernel.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "likelihood_evals_per_sec"
UNIT = "evals/s"

# algorithmic flop model, SURVEY.md s8d
C_ATM_NUM4D, C_GEOM, C_LEAF, C_INIT, C_EVAL, C_INTEG, C_LOG = 800, 150, 20, 25, 12, 30, 24


def load_workload():
    """M2 (ST-U NSX) constants and data.  Parameter vectors are drawn from the
    closed-form prior in xpsi_b200/synthetic.py (seed 20261017); the synthetic
    Poisson data set is the committed fixture.  Mesh and rays are built on the GPU.
    """
    from xpsi_b200 import synthetic as syn
    m2 = np.load(os.path.join(ROOT, "tests", "golden", "m2_stu_nsx.npz"))
    matrix, edges, channels, ch_edges = syn.nicer_like_response()
    return dict(m2=m2, matrix=matrix, edges=edges, table=syn.nsx_like_table(),
                exposure=syn.M2_EXPOSURE, n_theta=int(m2["n_theta"]))


def theta_block(first, count):
    """Rows [first, first+count) of the deterministic ST-U parameter-vector list; row 0 and 1 are
    the two golden parameter vectors so that parity against the reference can be asserted."""
    from xpsi_b200 import synthetic as syn
    m2 = np.load(os.path.join(ROOT, "tests", "golden", "m2_stu_nsx.npz"))
    head = np.array([m2["t0_theta"], m2["t1_theta"]])
    body = syn.m2_theta_batch(first + count)
    full = np.vstack([head, body])
    return np.ascontiguousarray(full[first:first + count])


def make_pipeline(w, max_batch):
    from xpsi_b200.pipeline import BatchedLikelihood
    m2 = w["m2"]
    pad = 64                     # max_sqrt_num_cells: ST-U spots allocate 36-42 rings, polar caps up to 64
    return BatchedLikelihood(member_component=[0, 1], max_rings=pad, max_azi=pad, n_rays=200,
                             energies=m2["t0_int0_energies"], leaves=m2["t0_int0_leaves"],
                             phases=m2["t0_int0_phases"], hot_atm_ext=2, hot_atmosphere=w["table"],
                             image_order_limit=3, response=w["matrix"], energy_edges=w["edges"],
                             counts=m2["counts"], data_phases=np.linspace(0.0, 1.0, 33),
                             exposure_time=w["exposure"], max_batch=max_batch)


def fill_batch(w, batch, first_index):
    m2 = w["m2"]
    for b in range(batch.B):
        t = (first_index + b) % w["n_theta"]
        batch.omega[b] = m2["t%d_int0_omega" % t]
        batch.inclination[b] = m2["t%d_int0_inclination" % t]
        batch.d_sq[b] = m2["t%d_d_sq" % t]
        batch.phase_shifts[b] = m2["t%d_marg_phase_shifts" % t]
        for m in range(2):
            g = lambda k: m2["t%d_int%d_%s" % (t, m, k)]
            batch.set_member(b, m, g("cellArea"), g("theta"), g("phi"), g("radialCoords_of_parallels"),
                             g("r_s_over_r"), g("srcCellParams"), g("deflection"), g("cos_alpha"),
                             g("lag"), g("maxDeflection"), g("cos_gammaArray"))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def flops_per_eval(work, n_evals, shape, n_regions, n_quad=65, n_newton=3):
    """F_total of SURVEY.md s8d from the integrator's measured work counters."""
    N_E, N_P, N_L = shape["n_energies"], shape["n_phases"], shape["n_phases"]
    N_in, N_chan, N_bins = shape["n_in"], shape["n_chan"], shape["n_bins"]
    H, V, RI, K = (work[k] / n_evals for k in ("H", "V", "RI", "K"))
    f_int = H * C_GEOM + V * (C_LEAF + N_E * (C_ATM_NUM4D + 4)) + RI * N_E * N_L * C_INIT + K * N_E * N_P * C_EVAL
    f_fold = n_regions * (N_P * (N_E * C_INIT + N_in * C_INTEG) + N_in * N_P + 2 * N_chan * N_in * N_P)
    f_like = n_regions * N_chan * (N_P * C_INIT + N_bins * C_INTEG) + \
        N_chan * (n_newton * N_bins * 6 + n_quad * N_bins * C_LOG + 60)
    return dict(integrate=f_int, fold=f_fold, likelihood=f_like, total=f_int + f_fold + f_like)


# --------------------------------------------------------------------------- reference arm
def _ref_worker(args):
    """Evaluate the reference's likelihood(theta, force=True) -- embed, integrate,
    register, likelihood: the same span our arm times -- n times in one process."""
    n, first = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        import ref_env
        ref_env.import_reference()
        import make_golden as mg
        rec = mg.Recorder()
        m2 = np.load(os.path.join(ROOT, "tests", "golden", "m2_stu_nsx.npz"))
        like, signal, instrument, hots = mg.build_m2(rec, m2["counts"])
        thetas = theta_block(first, n)
        like(list(thetas[0]), force=True)                 # warm-up
        t_total = 0.0
        lnLs = []
        for k in range(n):
            t0 = time.perf_counter()
            v = like(list(thetas[k]), force=True)
            t_total += time.perf_counter() - t0
            lnLs.append(float(v))
    return t_total, n, lnLs


def reference_available():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_env
        return ref_env.available()
    except Exception:
        return False


def run_reference_sample(per_proc, procs):
    """nproc independent processes x threads=1, the reference's recommended mode
    (xpsi/Likelihood.py:44-53); returns evals/s over the wall time of the slowest."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_worker, [(per_proc, i * per_proc) for i in range(procs)])
    wall_incl_setup = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    n_total = sum(r[1] for r in res)
    lnLs = [v for r in res for v in r[2]]                  # theta rows 0 .. n_total-1 in order
    return n_total / slowest, n_total, slowest, wall_incl_setup, lnLs


_JSON_FD = 1


def emit(line):
    """The one JSON line, written to the process's original stdout."""
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    procs = len(os.sched_getaffinity(0))
    if not reference_available():
        emit({"impl": "reference", "unavailable": "oracle/_ref not built on this box"})
        return 0
    per_proc = 4
    vals = []
    for step in range(args.warmup + args.steps):
        v, n_total, slowest, wall, lnLs = run_reference_sample(per_proc, procs)
        if step >= args.warmup:
            vals.append((v, slowest))
    value = float(np.mean([v for v, _ in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean([s for _, s in vals]) * 1e3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "M2 ST-U NSX-shaped Num4D (35,14,67,166), 2 hot regions, 128 energies, "
                               "100 leaves/phases, 200 rays, 270x1500 response, 32 phase bins; "
                               "likelihood(theta) including embed (mesh + rays)",
                   "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference",
                         "sample": "%d processes x %d evaluations of likelihood(theta, force=True), "
                                   "threads=1 each (X-PSI 3.3.0 sources on the GSL-subset shim)" % (procs, per_proc)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------- our arm
def main_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    from xpsi_b200 import _lib
    _lib.check(_lib.lib.xpsi_b200_set_device(local))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    B = args.batch
    w = load_workload()
    from xpsi_b200 import synthetic as syn
    pipe = make_pipeline(w, B)
    thetas = theta_block(rank * B, B)                    # this rank's contiguous block of the theta list
    spots = syn.m2_spot_batch(pipe, thetas)
    stream = torch.cuda.ExternalStream(_lib.lib.xpsi_b200_stream(), device=torch.device("cuda", local))

    peak = np.zeros(1)
    _lib.check(_lib.lib.xpsi_b200_fp64_peak_tflops(_lib.dptr(peak)))

    # ---- kernels only: parameter vectors resident on the device, embed + four stages per step ----
    pipe.embed_spots(spots)
    pipe.count_work(True)
    for _ in range(max(args.warmup, 3)):
        pipe.eval_spots_resident(B)
    torch.cuda.synchronize()
    work = pipe.count_work(False)
    lnL, status = pipe.download(B)
    k0 = _lib.counters()[0]
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    stage = dict(embed=0.0, integrate=0.0, energy=0.0, fold=0.0, marginal=0.0)
    barrier()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ev[0].record()
        for _ in range(args.steps):
            pipe.eval_spots_resident(B)
        ev[1].record()
    torch.cuda.synchronize()
    barrier()
    ms_total = ev[0].elapsed_time(ev[1])
    launches = _lib.counters()[0] - k0
    for _ in range(3):                       # per-stage split (separate runs, not part of `value`)
        pipe.eval_spots_resident(B)
        for k, v in pipe.stage_ms().items():
            stage[k] += v / 3.0
    clocks = sampler.stop()

    # ---- end to end: host parameter arrays in, lnL out (H2D + embed + stages + D2H per step) ----
    for _ in range(2):
        pipe.eval_spots(spots)
    c0 = _lib.counters()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev[0].record()
    for _ in range(args.steps):
        lnL_e2e, status_e2e = pipe.eval_spots(spots)
    with torch.cuda.stream(stream):
        ev[1].record()
    torch.cuda.synchronize()
    wall_e2e = time.perf_counter() - t0
    ms_e2e = max(ev[0].elapsed_time(ev[1]), wall_e2e * 1e3)
    c1 = _lib.counters()
    h2d = (c1[1] - c0[1]) // args.steps
    d2h = (c1[2] - c0[2]) // args.steps

    # ---- max over ranks, gather lnL ------------------------------------------------------
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        from xpsi_b200.sharding import gather_blocks
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the path's only collective: one all_gather of B/G lnL + status per rank (NCCL over NVLink)
        all_lnL, all_status = gather_blocks(lnL, status, B * world, device=torch.device("cuda", local))
    else:
        all_lnL, all_status = lnL, status
    ms_total, ms_e2e = float(t[0]), float(t[1])
    n_evals_step = B * world
    value = n_evals_step * args.steps / (ms_total * 1e-3)
    e2e_value = n_evals_step * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        return 0
    # 11 = the reference's "slim" early exit (model exceeds the data by > 20 sigma in a channel,
    # default_background_marginalisation.pyx:677-684): normal for prior draws far from the truth
    codes, cnts = np.unique(all_status, return_counts=True)
    status_counts = {str(int(c)): int(n) for c, n in zip(codes, cnts)}
    ok = bool(np.isin(all_status, (0, 11)).all())
    refs = [float(w["m2"]["t%d_lnL_total" % b]) for b in range(min(B, 2))]
    parity = float(max(abs(all_lnL[b] - refs[b]) for b in range(len(refs))))
    fl = flops_per_eval(work, B, pipe.shape, 2)
    int_ms = stage["integrate"]
    achieved = fl["integrate"] * B / (int_ms * 1e-3) / 1e12
    whole = fl["total"] * value / 1e12
    ws_bytes = 4.3e6 * B                       # leaf + slab workspaces and intermediates per step
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "M2 ST-U NSX-shaped Num4D (35,14,67,166), 2 hot regions, 128 energies, "
                               "100 leaves/phases, 200 rays, 270x1500 response, 32 phase bins",
                   "batch_per_gpu": B,
                   "theta": "distinct ST-U parameter vectors from the closed-form prior (seed 20261017, polar caps "
                            "included); mesh + rays embedded on the GPU inside the timed region",
                   "parallelism": "theta-sharded x%d, no data-path collective" % world,
                   "l2": "per-step working set larger than L2 (%.0f MB of workspaces/intermediates); only the "
                         "theta-independent atmosphere table and response stay L2-resident" % (ws_bytes / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "kernel": "k_azinv_flux<2,0,0> (+ k_azinv_geometry, k_azinv_slab, k_azinv_slab_member, k_azinv_moments: the integrate stage)",
                     "achieved": achieved, "peak": float(peak[0]),
                     "unit": "TFLOP/s", "frac": achieved / float(peak[0]),
                     "traffic": 7.86e6 * B,
                     "traffic_source": "ncu --set full at batch 32 (dram read+write 251.4 MB per launch, profiles/r01l_k_azinv_flux_ncu_full.txt), scaled to this batch",
                     "peak_source": "in-run DFMA microbenchmark (MEASURED_PEAKS.json has no fp64 entry)",
                     "algorithmic_gflop_per_eval": {k: v / 1e9 for k, v in fl.items()},
                     "whole_path_tflops": whole, "whole_path_frac": whole / float(peak[0]),
                     "stage_ms": stage,
                     "hbm_sanity_gbs": 4.64e6 * B / (int_ms * 1e-3) / 1e9},
        "parity": {"max_abs_lnL_diff_vs_reference_golden": parity, "no_unexpected_status": ok,
                   "status_counts": status_counts},
    }
    if world == 1 and not args.no_cpu_baseline:
        if reference_available():
            procs = len(os.sched_getaffinity(0))
            per = 2
            v, n_total, slowest, wall, ref_lnL = run_reference_sample(per, procs)
            n_cmp = min(n_total, B)
            diffs = [abs(ref_lnL[k] - lnL[k]) for k in range(n_cmp) if status[k] == 0 and ref_lnL[k] > -1e80]
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": procs, "kind": "reference",
                                    "sample": "%d processes x %d evaluations of likelihood(theta, force=True) on the "
                                              "first %d parameter vectors of this batch (threads=1 each), %.1f s of "
                                              "CPU work; X-PSI 3.3.0 sources on the GSL-subset shim"
                                              % (procs, per, n_total, slowest * procs),
                                    "max_abs_lnL_diff_vs_gpu": float(max(diffs)) if diffs else None,
                                    "n_compared": len(diffs)}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                    "sample": "oracle/_ref absent on this box"}
    emit(line)
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="parameter vectors per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    # stdout carries exactly one JSON line: whatever libraries print on fd 1 (NCCL's version banner under
    # torchrun, warnings of the reference package) is sent to stderr
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rc = main_reference(a) if a.impl == "reference" else main_ours(a)
    try:                                   # leave NCCL cleanly under torchrun
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()
    except Exception:
        pass
    sys.exit(rc)
