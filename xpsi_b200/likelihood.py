"""Scalar / batched likelihood callable over the GPU pipeline, mirroring ``xpsi.Likelihood.__call__``
(xpsi/Likelihood.py:450-511) for models whose hot regions the parameter-level embed covers.

The reference's ``Likelihood`` walks a tree of parameter objects (``Star`` / ``Photosphere`` / ``HotRegion`` /
``Signal``); here the model is the ``BatchedLikelihood`` pipeline plus one user function that maps an array of
parameter vectors onto a ``SpotBatch`` (and, optionally, the per-batch extras) -- the vectorised equivalent of
``super(Likelihood, self).__call__(p)`` followed by ``Star.update``.  The return conventions are the reference's:
a float log-likelihood (plus the log-prior when a prior is given), or a random value near ``llzero`` when the prior
is not finite, when a stage reports a numerical failure (the reference's ``PulseError`` / ``RayError`` handling,
xpsi/Likelihood.py:346-358) or when the result is below ``llzero``.
"""
import numpy as np


class Likelihood:
    def __init__(self, pipeline, fill, prior=None, llzero=-1.0e90, max_batch=None):
        """
        :param pipeline: a :class:`xpsi_b200.pipeline.BatchedLikelihood`.
        :param fill: ``fill(pipeline, P) -> SpotBatch`` for an array ``P[B, n_params]``; it may call
                     ``pipeline.upload_extras`` for Elsewhere / interstellar inputs.
        :param prior: optional callable ``prior(p) -> log-prior`` (``xpsi.Prior.__call__``).
        """
        self._pipe, self._fill, self._prior = pipeline, fill, prior
        self.llzero = float(llzero)
        self._max_batch = int(max_batch or pipeline.max_batch)
        self._cached_p, self._cached_value = None, None
        self.externally_updated = False

    @property
    def pipeline(self):
        return self._pipe

    def _filled(self, P):
        """``fill`` may return the spot batch alone (it then uploads per-block extras itself) or
        ``(spots, dict(att_power=..., else_temperature=..., signal_shifts=...))``."""
        r = self._fill(self._pipe, P)
        return r if isinstance(r, tuple) else (r, None)

    def sweep_local(self, P, download=True):
        """All rows of ``P`` on this GPU with the parameter vectors resident on the device: one upload, blocks of
        ``max_batch``, one download (``download=False`` leaves the results on the device for a collective,
        ``pipeline.sweep_device_results()``).  No prior, statuses as in :meth:`batch`."""
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        spots, extras = self._filled(P)
        if extras is None and (self._pipe.shape.get("has_elsewhere") or self._pipe.shape.get("has_attenuation")):
            raise ValueError("sweeps of a pipeline with Elsewhere / interstellar components need a fill function "
                             "that returns (spots, dict(att_power=..., else_temperature=...))")
        self._pipe.sweep_upload(spots, **(extras or {}))
        self._pipe.sweep_run()
        if not download:
            return None
        lnL, status = self._pipe.sweep_download()
        lnL[status != 0] = np.nan
        return lnL, status

    @property
    def random_near_llzero(self):
        """xpsi/Likelihood.py:267-271"""
        return float(self.llzero * (0.1 + 0.9 * np.random.rand(1))[0])

    def clear_cache(self):
        self._cached_p, self._cached_value = None, None

    # statuses that the reference turns into a random value near llzero: numerical error in a compiled stage
    # (PulseError / RayError handling, xpsi/Likelihood.py:346-358), the slim early exit and a non-positive
    # marginal integral (default_background_marginalisation.pyx:677-684,702-706)
    NUMERICAL_STATUSES = (1, 11, 12)

    def batch(self, P, strict=True):
        """log-likelihoods of the rows of ``P`` (no prior added); rows whose evaluation ended numerically
        (statuses 1, 11, 12) carry ``nan`` and the status.  A row with status 3 is a configuration outside the
        kernels' coverage -- the reference would have evaluated it -- and raises ``NotImplementedError``
        (``strict=False`` returns it as ``nan`` / 3 instead, for callers that inspect ``status`` themselves)."""
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        lnL = np.empty(P.shape[0])
        status = np.empty(P.shape[0], dtype=np.int32)
        starts = list(range(0, P.shape[0], self._max_batch))
        # the host-side fill of block k + 1 (parameter walk, array packing) runs in a worker thread while block k is on
        # the GPU (the ctypes call releases the GIL) -- only for fill functions that hand their per-batch extras back
        # by value; one that uploads extras itself must not run ahead of the evaluation
        pool = None
        nxt = self._filled(P[0:self._max_batch]) if starts else None
        overlap = len(starts) > 1 and (nxt[1] is not None or not (self._pipe.shape.get("has_elsewhere") or
                                                                   self._pipe.shape.get("has_attenuation")))
        if overlap:
            from concurrent.futures import ThreadPoolExecutor
            pool = ThreadPoolExecutor(max_workers=1)
        try:
            for k, i in enumerate(starts):
                blk = P[i:i + self._max_batch]
                spots, extras = nxt.result() if hasattr(nxt, "result") else nxt
                if k + 1 < len(starts):
                    nb = P[starts[k + 1]:starts[k + 1] + self._max_batch]
                    nxt = pool.submit(self._filled, nb) if overlap else None
                if extras is not None:
                    extras = dict(extras)
                    sig = extras.pop("signal_shifts", None)
                    if extras:
                        self._pipe.upload_extras(blk.shape[0], **extras)
                    if sig is not None or len(getattr(self._pipe, "signals", ())) > 1:
                        self._pipe.upload_signal_shifts(blk.shape[0], sig)
                lnL[i:i + blk.shape[0]], status[i:i + blk.shape[0]] = self._pipe.eval_spots(spots)
                if not overlap and k + 1 < len(starts):
                    nxt = self._filled(P[starts[k + 1]:starts[k + 1] + self._max_batch])
        finally:
            if pool is not None:
                pool.shutdown(wait=True)
        bad = ~np.isin(status, (0,) + self.NUMERICAL_STATUSES)
        if strict and bad.any():
            k = int(np.flatnonzero(bad)[0])
            raise NotImplementedError("xpsi_b200: parameter vector %d of the batch needs a configuration the "
                                      "kernels do not cover (status %d; e.g. more mesh rings than max_rings, a "
                                      "zero-radius member, the Num4D slab budget): the result would not be the "
                                      "reference's" % (k, int(status[k])))
        lnL[status != 0] = np.nan
        return lnL, status

    def __call__(self, p=None, reinitialise=False, force=False):
        """Same signature and return convention as xpsi/Likelihood.py:450-511."""
        if reinitialise or force:
            self.clear_cache()
        if p is None:
            if self._cached_p is None:
                raise TypeError('Parameter values have not been updated.')
            p = self._cached_p
        p = np.asarray(p, dtype=np.float64)
        logprior = None
        if self._prior is not None:
            logprior = self._prior(p)
            if not np.isfinite(logprior):
                return self.random_near_llzero
        if self._cached_p is not None and np.array_equal(p, self._cached_p):
            loglikelihood = self._cached_value                      # memoised, Likelihood.py:489-490
        else:
            lnL, status = self.batch(p[None, :])
            if status[0] != 0:                  # numerical failure or slim early exit (status 3 raised in batch)
                return self.random_near_llzero
            loglikelihood = float(lnL[0])
            self._cached_p, self._cached_value = p.copy(), loglikelihood
        if loglikelihood <= self.llzero:
            return self.random_near_llzero
        return loglikelihood + logprior if logprior is not None else loglikelihood
