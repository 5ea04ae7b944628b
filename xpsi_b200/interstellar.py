"""GPU mirror of ``xpsi.Interstellar.__call__`` (xpsi/Interstellar.py:27-58): in-place row scaling of a
signal ``[n_energies(, n_phases)]`` by an attenuation factor per energy."""
import numpy as np

from . import _lib


def attenuate(attenuation, signal):
    """``signal[i, :] *= attenuation[i]`` on the GPU, in place (``signal`` must be C-contiguous float64)."""
    if not isinstance(signal, np.ndarray):
        raise TypeError('Signal must be a numpy.ndarray.')
    if signal.ndim not in (1, 2):
        raise ValueError('Invalid number of signal array dimensions.')
    if signal.dtype != np.float64 or not signal.flags.c_contiguous:
        raise TypeError('Signal must be a C-contiguous float64 array (it is modified in place).')
    att = _lib.as_f8(attenuation, 1)
    if att.shape[0] != signal.shape[0]:
        raise ValueError('One attenuation factor per signal row is required.')
    ncol = 1 if signal.ndim == 1 else signal.shape[1]
    _lib.check(_lib.lib.xpsi_b200_interstellar_attenuate(_lib.dptr(att), signal.shape[0], ncol, _lib.dptr(signal)))
    return None


class Interstellar:
    """Minimal mirror of ``xpsi.Interstellar``: subclass and implement ``attenuation(energies)``; calling
    the object attenuates a signal in place on the GPU."""

    def attenuation(self, energies):
        raise NotImplementedError('Implement the attenuation method.')

    def __call__(self, energies, signal):
        return attenuate(self.attenuation(np.asarray(energies)), signal)
