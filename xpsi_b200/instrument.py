"""GPU drop-in for the response contraction of ``xpsi.Instrument.__call__``."""
import numpy as np

from . import _lib


def fold(matrix, signal, irange, orange):
    """``numpy.dot(matrix[orange[0]:orange[1], irange[0]:irange[1]], signal)``
    (xpsi/Instrument.py:192-197) on the GPU."""
    matrix = _lib.as_f8(matrix, 2)
    signal = _lib.as_f8(signal, 2)
    i0, i1 = int(irange[0]), int(irange[1])
    o0, o1 = int(orange[0]), int(orange[1])
    if signal.shape[0] != i1 - i0:
        raise ValueError("shapes (%d,%d) and %r not aligned" % (o1 - o0, i1 - i0, signal.shape))
    out = np.empty((o1 - o0, signal.shape[1]), dtype=np.float64)
    _lib.check(_lib.lib.xpsi_b200_instrument_fold(
        _lib.dptr(matrix), matrix.shape[0], matrix.shape[1], i0, i1, o0, o1,
        _lib.dptr(signal), signal.shape[1], _lib.dptr(out)))
    return out


class Instrument:
    """Minimal mirror of ``xpsi.Instrument`` for the folding call: holds the
    response ``matrix`` and is callable as ``instrument(signal, irange, orange)``
    with the reference's caching of the last product (xpsi/Instrument.py:163-197).
    """

    def __init__(self, matrix, energy_edges=None, channels=None, channel_edges=None):
        self.matrix = _lib.as_f8(matrix, 2)
        if (self.matrix < 0.0).any():
            raise ValueError('Matrix elements must be positive.')    # Instrument.py:133-137
        self.energy_edges = energy_edges
        self.channels = channels
        self.channel_edges = channel_edges
        self._cached_signal = None

    def construct_matrix(self):
        return self.matrix

    def __call__(self, signal, irange, orange):
        self._cached_signal = fold(self.construct_matrix(), signal, irange, orange)
        return self._cached_signal

    @property
    def cached_signal(self):
        return self._cached_signal
