"""Averaging-window sample times and weights (host scalars).

Mirror of /root/reference/averaged_sw_explicit.py:77-95,233-235: ensemble member k uses
expt = rho*dt*s_k and weight w_k; here they become the per-column scalars of one batch.
"""
import math

import numpy as np

EIGS = (0.003465, 0.007274, 0.014955)          # spectral bounds, ref levels 3,4,5 (averaged_sw_explicit.py:77)


def natural_Mbar(eig, dt, rho, ppp=4.0):
    """Mbar = ceil(ppp*rho*dt*eig/(2 pi))  (averaged_sw_explicit.py:86)."""
    return int(math.ceil(ppp * rho * dt * eig / 2.0 / math.pi))


def sample_times_weights(Mbar, dt, rho):
    """(t_k, w_k), k < Mbar: midpoints of the window [-rho dt/2, rho dt/2] with the
    normalised bump weights exp(-1/(s(1-s)))  (averaged_sw_explicit.py:91-95,234-235)."""
    s = (np.arange(Mbar) + 0.5) / Mbar
    w = np.exp(-1.0 / s / (1.0 - s))
    return rho * dt * (s - 0.5), w / w.sum()
