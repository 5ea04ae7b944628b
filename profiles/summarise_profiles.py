#!/usr/bin/env python
"""Turn gpurun_out/<tag>_full.ncu-rep + <tag>_launches.csv into profiles/<tag>_kernels.json (read by bench.py for
``roofline.traffic`` and the per-kernel ncu numbers), profiles/<tag>_launches.csv and one text summary per kernel
that takes >= 1 % of the step.   python profiles/summarise_profiles.py r02 [batch]"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rep = os.path.join(ROOT, "gpurun_out", tag + "_full.ncu-rep")
launch_csv = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")

KEYS = {
    "gpu__time_duration.sum": "duration",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__shared_mem_per_block_static": "smem_static",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions",
    "sm__inst_executed_pipe_fp64.sum": "fp64_pipe_instructions",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active": "fmaheavy_pipe_active_pct",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active": "shared_pipe_active_pct",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_pipe_active_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
}
UNIT_SCALE = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def clean(full):
    """'void xb::k_azinv_flux_mma<(int)2, ...>(xb::AzinvArgs, ...)' -> 'k_azinv_flux_mma'"""
    name = re.sub(r"<.*", "", full).split("(")[0].strip()
    name = name.split()[-1] if name else name
    return name.split("::")[-1]


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main():
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kernels = {}
    for r in rows[2:]:
        full = r[col["Kernel Name"]]
        name = clean(full)
        k = {"instantiation": full}
        for h, key in KEYS.items():
            if h not in col:
                continue
            v = num(r[col[h]])
            u = units[col[h]]
            if v is None:
                continue
            if key == "duration":
                k["duration_ms"] = v * UNIT_SCALE.get(u, 1.0)
            elif key in ("dram_read", "dram_write"):
                k[key] = v * UNIT_SCALE.get(u, 1.0)
            else:
                k[key] = v
        k["dram_bytes"] = k.get("dram_read", 0.0) + k.get("dram_write", 0.0)
        # several launches of one kernel in a step (slab / slab-member run per atmosphere): keep the longest
        if name not in kernels or k.get("duration_ms", 0) > kernels[name].get("duration_ms", 0):
            kernels[name] = k
    # the launch list: share of the step per kernel
    step = {}
    if os.path.isfile(launch_csv):
        lines = [l for l in open(launch_csv) if not l.startswith("==")]
        lr = list(csv.reader(lines))
        lh = {h: i for i, h in enumerate(lr[0])}
        for r in lr[1:]:
            if len(r) <= lh.get("Metric Value", 0) or r[lh["Metric Name"]] != "gpu__time_duration.sum":
                continue
            name = clean(r[lh["Kernel Name"]])
            v = num(r[lh["Metric Value"]]) * UNIT_SCALE.get(r[lh["Metric Unit"]], 1.0)
            step[name] = step.get(name, 0.0) + v
        tot = sum(step.values())
        for name, v in step.items():
            kernels.setdefault(name, {})["step_ms_cold"] = v
            kernels[name]["step_share"] = v / tot

    out = {"tag": tag, "batch": batch, "source": "ncu --set full --clock-control none, one launch per kernel of one "
           "bench step (profiles/make_profiles.sh); step_share from the gpu__time_duration launch list of the same command",
           "kernels": kernels}
    outdir = os.path.join(ROOT, sys.argv[3]) if len(sys.argv) > 3 else os.path.join(ROOT, "profiles")
    with open(os.path.join(outdir, tag + "_kernels.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for name, k in sorted(kernels.items(), key=lambda kv: -kv[1].get("step_share", 0)):
        print("%-26s %6.2f%% of step  %8.3f ms  fp64 %5.1f%%  dmma %5.1f%%  issue %5.1f%%  regs %3d  dram %8.1f MB"
              % (name, 100 * k.get("step_share", 0), k.get("duration_ms", 0), k.get("fp64_pipe_active_pct", 0),
                 k.get("dmma_pipe_active_pct", 0), k.get("issue_active_pct", 0), int(k.get("registers", 0)), k.get("dram_bytes", 0) / 1e6))


if __name__ == "__main__":
    main()
