"""Sampler-facing batched driver.

The reference feeds samplers one parameter vector at a time (``xpsi.Likelihood.__call__``,
xpsi/Likelihood.py:450-511) and spreads vectors over MPI ranks by scatter / gather
(``xpsi.Sample.importance``, xpsi/Sample.py:287-336).  Here the unit of work is a *block* of parameter
vectors, so the sampler-facing pieces are vectorised:

* :class:`BoxPrior` -- ``xpsi.Prior`` (xpsi/Prior.py:84-131) over an array ``P[B, d]``: hard bounds, optional
  joint rules, ``inverse_sample`` of a block of hypercube points;
* :func:`vectorized_loglike` / :func:`ultranest_callables` -- the ``loglike(P[B, d]) -> lnL[B]`` /
  ``transform(U[B, d])`` pair that ``ultranest.ReactiveNestedSampler(..., vectorized=True)`` expects
  (xpsi/UltranestSampler.py:56-57 passes the scalar versions), with the reference's return conventions
  (random value near ``llzero`` outside the prior, on numerical failure and on the ``slim`` early exit);
* :func:`shard_indices`, :func:`sweep` -- one rank per GPU: rows are dealt round-robin (or by predicted cost,
  heaviest first, in snake order) so that every rank gets the same mix, each rank evaluates its rows
  block by block with the parameter vectors resident on its device, and ONE all_gather of the per-rank
  ``[lnL | status]`` payload (NCCL over NVLink; gloo in the CPU tests) puts the full result on every rank;
* :func:`importance` -- the reweighting loop of ``xpsi.Sample.importance`` on top of ``sweep``.

torch is used for the process group only.
"""
import numpy as np


class BoxPrior:
    """Vectorised flat prior with hard bounds and optional joint rules.

    ``rules`` are callables ``rule(P) -> bool[B]`` (True = inside the support), the vectorised form of the
    ``if ...: return -np.inf`` lines of a ``CustomPrior.__call__`` (e.g. TestRun_Num.py's compactness rule).
    ``density`` (optional) is ``density(P) -> float[B]`` for ``importance(prior_change=True)``.
    """

    def __init__(self, names, bounds, rules=(), density=None):
        self.names = tuple(names)
        self.bounds = np.asarray(bounds, dtype=np.float64).reshape(len(self.names), 2)
        if not np.all(np.isfinite(self.bounds)) or np.any(self.bounds[:, 1] < self.bounds[:, 0]):
            raise ValueError('Compact support required.')          # xpsi/Prior.py:129
        self.rules = tuple(rules)
        self._density = density

    def __len__(self):
        return len(self.names)

    def __call__(self, P):
        """log-prior of every row: 0 inside the support, -inf outside (xpsi/Prior.py:84-100)."""
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        if P.shape[1] != len(self):
            raise ValueError("expected %d parameters per row, got %d" % (len(self), P.shape[1]))
        inside = np.all((P >= self.bounds[:, 0]) & (P <= self.bounds[:, 1]), axis=1)
        for rule in self.rules:
            inside &= np.asarray(rule(P), dtype=bool)
        return np.where(inside, 0.0, -np.inf)

    def inverse_sample(self, hypercube=None, n=1, rng=None):
        """Rows of the unit hypercube -> rows of parameter space (xpsi/Prior.py:102-131, flat between bounds)."""
        if hypercube is None:
            hypercube = (rng or np.random.default_rng()).random((n, len(self)))
        U = np.asarray(hypercube, dtype=np.float64)
        scalar = U.ndim == 1
        U = np.atleast_2d(U)
        P = self.bounds[:, 0] + (self.bounds[:, 1] - self.bounds[:, 0]) * U
        return P[0] if scalar else P

    def draw(self, n, rng=None):
        """``n`` rows from the prior (rejection against the rules)."""
        rng = rng or np.random.default_rng()
        out = np.empty((0, len(self)))
        while out.shape[0] < n:
            P = self.inverse_sample(rng.random((max(n, 64), len(self))))
            out = np.vstack([out, P[np.isfinite(self(P))]])
        return np.ascontiguousarray(out[:n])

    def density(self, P):
        if self._density is None:
            raise AttributeError("this prior has no density")
        return np.asarray(self._density(np.atleast_2d(P)), dtype=np.float64)


def random_near_llzero(llzero, n, rng=None):
    """xpsi/Likelihood.py:267-271, one draw per row."""
    r = rng.random(n) if rng is not None else np.random.rand(n)
    return llzero * (0.1 + 0.9 * r)


def vectorized_loglike(likelihood, prior=None, add_logprior=False):
    """``loglike(P[B, d]) -> lnL[B]`` over ``likelihood.batch`` with the scalar call's conventions
    (xpsi/Likelihood.py:476-511): rows outside the prior are not evaluated and get a random value near
    ``llzero``, so do rows that end numerically (status 1 / 11 / 12) or at or below ``llzero``."""
    llzero = float(likelihood.llzero)

    def loglike(P):
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        out = random_near_llzero(llzero, P.shape[0])
        logp = np.zeros(P.shape[0]) if prior is None else np.asarray(prior(P), dtype=np.float64)
        live = np.isfinite(logp)
        if live.any():
            lnL, status = likelihood.batch(P[live])
            good = (status == 0) & (lnL > llzero)
            val = lnL + (logp[live] if add_logprior else 0.0)
            rows = np.flatnonzero(live)[good]
            out[rows] = val[good]
        return out
    return loglike


def ultranest_callables(likelihood, prior):
    """``(loglike, transform)`` for ``ultranest.ReactiveNestedSampler(names, loglike, transform,
    vectorized=True)``: the batched form of xpsi/UltranestSampler.py:56-57."""
    return vectorized_loglike(likelihood, prior), prior.inverse_sample


# --------------------------------------------------------------------------- sharding
def shard_indices(n_total, rank, world, cost=None):
    """Row indices of ``rank`` (ascending).  Without ``cost`` rows are dealt round-robin; with a predicted
    cost per row they are dealt heaviest first in snake order (0..G-1, G-1..0, ...), which bounds the spread
    of the per-rank sums by one row's cost."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    n_total = int(n_total)
    if cost is None:
        return np.arange(rank, n_total, world, dtype=np.int64)
    cost = np.asarray(cost, dtype=np.float64)
    if cost.shape != (n_total,):
        raise ValueError("one cost per row is required")
    order = np.argsort(-cost, kind="stable")
    pos = np.arange(n_total)
    lap, k = divmod(pos, world)
    owner = np.where(lap % 2 == 0, k, world - 1 - k)
    return np.sort(order[owner == rank]).astype(np.int64)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def gather_rows(local_lnL, local_status, idx, n_total, world, device=None, index_of_rank=None):
    """all_gather of ragged per-rank results back into row order: every rank gets ``lnL[n_total]`` and
    ``status[n_total]``.  ``local_*`` are host arrays or torch tensors already on ``device``; the payload is one
    fp64 vector ``[lnL | status]`` per rank, padded to the widest share."""
    import torch
    dist = _dist()
    if dist is None or world == 1:
        lnL = np.empty(n_total)
        status = np.empty(n_total, dtype=np.int32)
        lnL[idx] = _to_host(local_lnL)
        status[idx] = _to_host(local_status)
        return lnL, status
    if index_of_rank is None:
        raise ValueError("index_of_rank(r) is required to place the other ranks' rows")
    shares = [index_of_rank(r) for r in range(world)]
    width = max(len(s) for s in shares)
    n_loc = len(idx)
    payload = torch.zeros(2 * width, dtype=torch.float64, device=device)
    payload[:n_loc] = torch.as_tensor(local_lnL, device=device).to(torch.float64)
    payload[width:width + n_loc] = torch.as_tensor(local_status, device=device).to(torch.float64)
    out = torch.empty(world * 2 * width, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, payload)
    host = out.cpu().numpy().reshape(world, 2, width)
    lnL = np.empty(n_total)
    status = np.empty(n_total, dtype=np.int32)
    for r, s in enumerate(shares):
        lnL[s] = host[r, 0, :len(s)]
        status[s] = host[r, 1, :len(s)].astype(np.int32)
    return lnL, status


def _to_host(a):
    try:
        import torch
        if isinstance(a, torch.Tensor):
            return a.cpu().numpy()
    except ImportError:
        pass
    return np.asarray(a)


def sweep(likelihood, P, cost=None, device=None, evaluate=None, info=None):
    """Evaluate every row of ``P[N, d]`` once across the ranks of the process group (one rank per GPU) and
    return ``(lnL[N], status[N])`` on every rank -- the batched form of the scatter / evaluate / gather loop
    of xpsi/Sample.py:287-336.

    ``cost``: optional predicted cost per row (see :func:`shard_indices`).  ``evaluate(P_local) -> (lnL,
    status)`` replaces the GPU evaluation (used by the CPU tests of the host logic).  ``info`` (a dict) receives
    ``rows`` (this rank's share), ``device_ms`` (this rank's evaluation, CUDA events on the library's stream)
    and ``gather_ms``.
    """
    P = np.atleast_2d(np.asarray(P, dtype=np.float64))
    N = P.shape[0]
    dist = _dist()
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    shares = [shard_indices(N, r, world, cost) for r in range(world)]
    idx_of = lambda r: shares[r]
    idx = idx_of(rank)
    stats = info if info is not None else {}
    stats["rows"] = int(len(idx))
    if evaluate is not None:
        lnL_loc, st_loc = evaluate(P[idx]) if len(idx) else (np.empty(0), np.empty(0, np.int32))
        lnL, status = gather_rows(lnL_loc, st_loc, idx, N, world, device=device, index_of_rank=idx_of)
        lnL[status != 0] = np.nan
        return lnL, status
    import torch
    from . import _lib
    stream = torch.cuda.ExternalStream(_lib.lib.xpsi_b200_stream(), device=device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    with torch.cuda.stream(stream):
        ev[0].record()
        if len(idx):
            likelihood.sweep_local(P[idx], download=False)
        ev[1].record()
        if world > 1:
            if len(idx):
                d_lnL, d_st = likelihood.pipeline.sweep_device_results()
                loc = (torch.as_tensor(d_lnL, device=device), torch.as_tensor(d_st, device=device))
            else:
                loc = (torch.empty(0, dtype=torch.float64, device=device), torch.empty(0, dtype=torch.int32, device=device))
            lnL, status = gather_rows(loc[0], loc[1], idx, N, world, device=device, index_of_rank=idx_of)
        else:
            lnL_loc, st_loc = likelihood.pipeline.sweep_download()
            lnL, status = gather_rows(lnL_loc, st_loc, idx, N, 1)
        ev[2].record()
    ev[2].synchronize()
    stats["device_ms"] = ev[0].elapsed_time(ev[1])
    stats["gather_ms"] = ev[1].elapsed_time(ev[2])
    lnL[status != 0] = np.nan                 # as Likelihood.batch: a row that ended numerically carries no value
    return lnL, status


def importance(target, importance_lnL, samples, weight_threshold=1.0e-3, prior_ratio=None, cost=None, device=None,
               evaluate=None):
    """Importance-reweight posterior samples under a changed likelihood (xpsi/Sample.py:182-336 with
    ``likelihood_change=True``).

    ``samples[n, 2 + d]``: weight, -2 lnL, parameters (the sample-file layout the reference loads).  Rows with
    ``weight / max(weight) >= weight_threshold`` are kept, the target log-likelihood is evaluated for all of
    them in one sharded sweep, and their weights are multiplied by ``exp(lnL_target - lnL_importance)`` (times
    ``prior_ratio(P)`` when given, the ``prior_change`` factor) and renormalised; column 1 becomes
    ``-2 lnL_target``.  ``importance_lnL``: the importance log-likelihoods of the kept rows, or ``None`` to take
    them from column 1.  Returns ``(reweighted_samples, normalisation)``.
    """
    samples = np.asarray(samples, dtype=np.float64)
    keep = samples[:, 0] / np.max(samples[:, 0]) >= weight_threshold
    ref = samples[keep].copy()
    P = ref[:, 2:]
    lnL_imp = -0.5 * ref[:, 1] if importance_lnL is None else np.asarray(importance_lnL, dtype=np.float64)
    lnL_t, status = sweep(target, P, cost=cost, device=device, evaluate=evaluate)
    bad = status != 0
    w = np.where(bad, 0.0, np.exp(np.where(bad, 0.0, lnL_t) - lnL_imp))      # a failed target evaluation carries no weight
    if prior_ratio is not None:
        w = w * np.asarray(prior_ratio(P), dtype=np.float64)
    ref[:, 0] *= w
    ref[:, 1] = np.where(bad, np.inf, -2.0 * np.where(bad, 0.0, lnL_t))
    norm = float(np.sum(ref[:, 0]))
    if norm > 0.0:
        ref[:, 0] /= norm
    return ref, norm
