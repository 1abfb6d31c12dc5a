"""TEST INFRASTRUCTURE (oracle).  Averaging-window sample times and weights.

/root/reference/averaged_sw_explicit.py:77-95,233-235:
  L = eig*dt*rho; Mbar = ceil(ppp*rho*dt*eig/2/pi); s_k=(k+1/2)/Mbar;
  w_k = exp(-1/(s(1-s))) normalised; expt_k = rho*dt*(s_k-1/2).
/root/reference/timestepping_methods.py:10-11: V = sum_k w_k DU_k.
"""
from math import ceil, pi
import numpy as np

EIGS = [0.003465, 0.007274, 0.014955]                  # averaged_sw_explicit.py:77 (ref 3,4,5)


def natural_Mbar(eig, dt, rho, ppp=4):
    return ceil(ppp * rho * dt * eig / 2 / pi)         # :86


def sample_times_weights(Mbar, dt, rho):
    svals = np.arange(0.5, Mbar) / Mbar                # :91
    weights = np.exp(-1.0 / svals / (1.0 - svals))     # :92
    weights = weights / np.sum(weights)                # :93
    svals = svals - 0.5                                # :95
    return rho * dt * svals, weights                   # :234-235


def weighted_average(DU, weights):
    """DU: N x B columns; ensemble.allreduce of DU_k*wt_k (timestepping_methods.py:10-11)."""
    return DU @ weights
