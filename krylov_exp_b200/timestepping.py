"""Batched averaged time-steppers: the callers of the hot path.

Same schemes as /root/reference/timestepping_methods.py (rk2 :4-29, rk4 :32-77, heuns :80-103,
ssprk3 :107-140, leapfrog :144-161), but the Mbar ensemble members are the columns of one device
batch instead of Mbar MPI ranks, so "apply on every rank; SlowSolver.solve(); apply back;
DU *= wt; ensemble.allreduce" becomes

    averaged(U) = cheby.apply_average( slow( cheby.apply_batch(U, expt, shared=True) ), -expt, wt )

(`slow` is the nonlinear right-hand side N(.) on a batch, i.e. SlowSolver.solve() per member: host
Firedrake in the reference, out of scope here; any callable on (B, n) arrays).  cheby2 applies
(identical on every rank in the reference) are computed once.  Vectors are NumPy arrays [u ; eta].
"""
import numpy as np


class AveragedStepper:
    def __init__(self, cheby, cheby2, slow, expt, wt, dt, ensemble=None, fused=True):
        """expt, wt: the LOCAL sample times / weights (all of them on one GPU).  fused: the backward half of an
        averaged evaluation only needs the weighted sum, so it runs as one Clenshaw recurrence
        (cheby_exp.apply_average(fused=True)); False forms the Mbar backward results first, as the reference does."""
        self.cheby, self.cheby2, self.slow = cheby, cheby2, slow
        self.fused = bool(fused)
        self.expt = np.ascontiguousarray(expt, dtype=np.float64)
        self.wt = np.ascontiguousarray(wt, dtype=np.float64)
        self.dt = float(dt)
        self.ensemble = ensemble

    # -- building blocks
    def averaged(self, U):
        """sum_k w_k exp(-t_k L) N(exp(t_k L) U)   (timestepping_methods.py:7-11)."""
        F = self.cheby.apply_batch(U, self.expt, shared=True)
        N = np.ascontiguousarray(self.slow(F))
        if self.ensemble is not None and self.ensemble.world > 1:
            return self.ensemble.averaged_apply(self.cheby, N, -self.expt, self.wt, fused=self.fused)
        return self.cheby.apply_average(N, -self.expt, self.wt, fused=self.fused)

    def exp2(self, X, t):
        """cheby2.apply(X, Y, t): one un-batched exponential."""
        Y = np.empty_like(X)
        self.cheby2.apply(X, Y, t)
        return Y

    # -- schemes: U is updated in place and returned
    def rk2(self, U):
        dt = self.dt
        V = U + 0.5 * self.averaged(U)                      # :7-13
        DU = self.exp2(V, dt / 2)                           # :16
        V = self.averaged(DU)                               # :19-23
        DU = self.exp2(U, dt / 2)                           # :25
        V = DU + V                                          # :26
        U[:] = self.exp2(V, dt / 2)                         # :29
        return U

    def rk4(self, U):
        dt = self.dt
        V1 = self.averaged(U)                               # :35-39
        U1 = U + 0.5 * V1                                   # :41
        DU = self.exp2(U1, dt / 2)                          # :44
        V2 = self.averaged(DU)                              # :45-49
        V = self.exp2(U, dt / 2)                            # :51
        U2 = V + 0.5 * V2                                   # :52
        V3 = self.averaged(U2)                              # :55-59
        U3 = V + V3                                         # :61
        U1 = V2 + V3                                        # :64
        U2 = self.exp2(U1, dt / 2)                          # :65
        DU = self.exp2(U3, dt / 2)                          # :67
        U3 = self.averaged(DU)                              # :68-72
        U1 = self.exp2(V1, dt)                              # :74
        V = self.exp2(U, dt)                                # :75
        U[:] = V + 1 / 6 * U1 + 1 / 3 * U2 + 1 / 6 * U3     # :77
        return U

    def heuns(self, U):
        dt = self.dt
        U1 = U + self.averaged(U)                           # :83-89
        DU = self.exp2(U1, dt)                              # :92
        U2 = self.averaged(DU)                              # :93-97
        U2 = self.exp2(U, dt) + U2                          # :99-100
        Un = self.exp2(U1, dt)                              # :103
        U[:] = 0.5 * Un + 0.5 * U2                          # :104
        return U

    def ssprk3(self, U):
        dt = self.dt
        U1 = self.exp2(U + self.averaged(U), dt)            # :110-117
        U2 = self.averaged(U1)                              # :120-124
        U2 = self.exp2(U1 + U2, -dt / 2)                    # :126-127
        U1 = self.exp2(U, dt / 2)                           # :128
        U2 = 0.75 * U1 + 0.25 * U2                          # :129
        U1 = self.averaged(U2)                              # :132-136
        U2 = self.exp2(U2 + U1, dt / 2)                     # :138-139
        U1 = self.exp2(U, dt)                               # :140
        U[:] = 1 / 3 * U1 + 2 / 3 * U2                      # :141
        return U

    def leapfrog(self, U, U_old, asselin):
        """U_old is updated in place as in the reference (:159); returns U."""
        dt = self.dt
        V = self.averaged(U)                                # :147-151
        V = self.exp2(U_old, dt) + 2 * V                    # :153-154
        U_new = self.exp2(V, dt)                            # :155
        U1 = self.exp2(U_old, dt)                           # :157
        U2 = self.exp2(U_new, -dt)                          # :158
        V = (U1 + U2) * 0.5 - U                             # :159
        U_old[:] = U + asselin * V                          # :160
        U[:] = U_new                                        # :162
        return U
