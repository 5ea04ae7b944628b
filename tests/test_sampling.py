"""Host logic of the sampler-facing driver on CPU: vectorised prior, sharding,
world_size-2 gloo sweep + importance reweighting with a stand-in evaluator, vectorised loglike conventions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_box_prior_matches_scalar_rules():
    from xpsi_b200 import synthetic as syn
    prior = syn.m2_prior()
    rng = np.random.default_rng(3)
    P = prior.inverse_sample(rng.random((200, len(prior))))
    assert P.shape == (200, 11)
    lp = prior(P)
    assert set(np.unique(lp)) <= {0.0, -np.inf}
    # the scalar form of the same prior (xpsi/Prior.py:84-100 + the compactness rule of TestRun_Num.py)
    for row, v in zip(P, lp):
        inside = all(lo <= x <= hi for x, (lo, hi) in zip(row, syn.M2_BOUNDS)) and \
            row[1] * syn.KM >= 3.0 * row[0] * syn.GM_SUN
        assert (v == 0.0) == inside
    outside = P.copy()
    outside[::2, 0] = 0.5                                # mass below its bound
    assert np.all(np.isinf(prior(outside)[::2])) and np.all(prior(outside)[1::2] == lp[1::2])
    # a single hypercube point maps like the reference's list-returning inverse_sample
    u = rng.random(len(prior))
    assert np.allclose(prior.inverse_sample(u), syn.M2_BOUNDS[:, 0] + u * (syn.M2_BOUNDS[:, 1] - syn.M2_BOUNDS[:, 0]))
    draws = prior.draw(50, rng)
    assert draws.shape == (50, 11) and np.all(prior(draws) == 0.0)
    # the bench's parameter-vector list is a stream of draws from this prior
    assert np.all(prior(syn.m2_bench_thetas(2, 64)) == 0.0)


def test_shard_indices_partition_and_balance():
    from xpsi_b200.sampling import shard_indices
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 1000, 100003):
        for world in (1, 2, 3, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    n, world = 1000, 8
    cost = rng.uniform(36, 64, n)
    parts = [shard_indices(n, r, world, cost) for r in range(world)]
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(n))
    sums = np.array([cost[p].sum() for p in parts])
    assert sums.max() - sums.min() <= cost.max()          # snake dealing: spread bounded by one row's cost
    rr = np.array([cost[shard_indices(n, r, world)].sum() for r in range(world)])
    assert sums.max() - sums.min() <= rr.max() - rr.min()


def _fake_eval(P):
    """Stand-in for the GPU evaluation: a smooth function of the row and a status pattern."""
    lnL = -0.5 * np.sum(P ** 2, axis=1)
    status = (np.floor(np.abs(P[:, 0]) * 10).astype(np.int64) % 7 == 0).astype(np.int32) * 11
    return lnL, status


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from xpsi_b200 import sampling
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)                      # same rows on every rank
    P = rng.normal(size=(37, 4))                        # ragged: 19 + 18
    info = {}
    lnL, st = sampling.sweep(None, P, evaluate=_fake_eval, info=info)
    cost = np.abs(P[:, 1])
    lnL_c, st_c = sampling.sweep(None, P, cost=cost, evaluate=_fake_eval)
    samples = np.column_stack([rng.uniform(0.0, 1.0, 37), rng.uniform(10, 20, 37), P])
    rew, norm = sampling.importance(None, None, samples, weight_threshold=0.2, evaluate=_fake_eval)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, info["rows"], lnL, st, lnL_c, st_c, rew, norm))


def test_gloo_world_size_2_sweep_and_importance():
    import torch.multiprocessing as mp
    from xpsi_b200 import sampling
    ctx = mp.get_context("spawn")
    world = 2
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    P = rng.normal(size=(37, 4))
    e_lnL, e_st = _fake_eval(P)
    e_lnL = np.where(e_st != 0, np.nan, e_lnL)
    samples = np.column_stack([rng.uniform(0.0, 1.0, 37), rng.uniform(10, 20, 37), P])
    one_rank, one_norm = sampling.importance(None, None, samples, weight_threshold=0.2, evaluate=_fake_eval)
    assert sorted(r[1] for r in res) == [18, 19]
    for rank, rows, lnL, st, lnL_c, st_c, rew, norm in res:
        for got_l, got_s in ((lnL, st), (lnL_c, st_c)):          # every rank holds the full result, in row order
            assert np.array_equal(got_s, e_st)
            assert np.array_equal(np.isnan(got_l), np.isnan(e_lnL))
            assert np.array_equal(got_l[e_st == 0], e_lnL[e_st == 0])
        assert np.allclose(rew, one_rank, rtol=0, atol=0, equal_nan=True) and norm == one_norm
    # the reweighting itself (xpsi/Sample.py:306-336): w *= exp(lnL_target - lnL_importance), renormalised
    keep = samples[:, 0] / samples[:, 0].max() >= 0.2
    t_lnL, t_st = _fake_eval(samples[keep, 2:])
    wexp = np.where(t_st != 0, 0.0, samples[keep, 0] * np.exp(t_lnL + 0.5 * samples[keep, 1]))
    assert np.allclose(one_rank[:, 0], wexp / wexp.sum())
    assert np.allclose(one_rank[t_st == 0, 1], -2.0 * t_lnL[t_st == 0])


class _FakeLikelihood:
    llzero = -1.0e90

    def __init__(self):
        self.calls = []

    def batch(self, P):
        self.calls.append(P.shape[0])
        lnL, status = _fake_eval(P)
        lnL = np.where(status != 0, np.nan, lnL)
        return lnL, status


def test_vectorized_loglike_conventions():
    from xpsi_b200 import sampling
    prior = sampling.BoxPrior(("a", "b", "c", "d"), [(-2.0, 2.0)] * 4)
    like = _FakeLikelihood()
    loglike, transform = sampling.ultranest_callables(like, prior)
    rng = np.random.default_rng(11)
    U = rng.random((64, 4))
    P = transform(U)
    assert P.shape == (64, 4) and np.all(np.isfinite(prior(P)))
    P[5] = 3.0                                             # outside the prior: never evaluated
    out = loglike(P)
    assert like.calls == [63]
    e_lnL, e_st = _fake_eval(P)
    good = (e_st == 0) & (np.arange(64) != 5)
    assert np.array_equal(out[good], e_lnL[good])
    bad = ~good                                            # random values near llzero (xpsi/Likelihood.py:267-271)
    assert np.all(out[bad] <= 0.1 * like.llzero) and np.all(out[bad] >= like.llzero)
    assert len(np.unique(out[bad])) == bad.sum()
