"""Batched averaged time-steppers: the callers of the hot path.

Same schemes as /root/reference/timestepping_methods.py (rk2 :4-29, rk4 :32-77, heuns :80-103,
ssprk3 :107-140, leapfrog :144-161), but the Mbar ensemble members are the columns of one device
batch instead of Mbar MPI ranks, so "apply on every rank; SlowSolver.solve(); apply back;
DU *= wt; ensemble.allreduce" becomes

    averaged(U) = cheby.apply_average( slow( cheby.apply_batch(U, expt, shared=True) ), -expt, wt )

`slow` is the nonlinear right-hand side N(.) on a batch, i.e. SlowSolver.solve() per member
(averaged_sw_explicit.py:177-213).  cheby2 applies (identical on every rank in the reference) are
computed once, and the INDEPENDENT ones of a scheme (rk4 :65/:67 and :74/:75, heuns :99/:103, ...)
go through `exp2_many` so that a device implementation runs them as one batch.

  AveragedStepper        host arrays (NumPy [u ; eta]); `slow` any callable on (B, n) arrays
  DeviceAveragedStepper  state, stage vectors, ensemble batches and cheby2 results stay in HBM (torch CUDA
                         tensors on the library's stream); `slow` is a DEVICE callback on a (B, n) tensor --
                         torch ops or a CUDA kernel such as krylov_exp_b200.slow_rhs.SlowRHS -- so a whole
                         time step makes no host<->device copy of a state vector and no host synchronisation.
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

    def exp2_many(self, pairs):
        """Independent cheby2 applies [(X, t), ...] of one scheme stage; here one after the other."""
        return [self.exp2(X, t) for X, t in pairs]

    # -- schemes: U is updated in place and returned
    def rk2(self, U):
        dt = self.dt
        V = U + 0.5 * self.averaged(U)                      # :7-13
        DU, DU0 = self.exp2_many([(V, dt / 2), (U, dt / 2)])  # :16, :25 (independent)
        V = self.averaged(DU)                               # :19-23
        V = DU0 + V                                         # :26
        U[:] = self.exp2(V, dt / 2)                         # :29
        return U

    def rk4(self, U):
        dt = self.dt
        V1 = self.averaged(U)                               # :35-39
        U1 = U + 0.5 * V1                                   # :41
        DU, V = self.exp2_many([(U1, dt / 2), (U, dt / 2)])  # :44, :51 (independent)
        V2 = self.averaged(DU)                              # :45-49
        U2 = V + 0.5 * V2                                   # :52
        V3 = self.averaged(U2)                              # :55-59
        U3 = V + V3                                         # :61
        U1 = V2 + V3                                        # :64
        U2, DU = self.exp2_many([(U1, dt / 2), (U3, dt / 2)])  # :65, :67
        U3 = self.averaged(DU)                              # :68-72
        U1, V = self.exp2_many([(V1, dt), (U, dt)])         # :74, :75
        U[:] = V + 1 / 6 * U1 + 1 / 3 * U2 + 1 / 6 * U3     # :77
        return U

    def heuns(self, U):
        dt = self.dt
        U1 = U + self.averaged(U)                           # :83-89
        DU, E0 = self.exp2_many([(U1, dt), (U, dt)])        # :92 (= :103, the same apply twice in the reference), :99
        U2 = self.averaged(DU)                              # :93-97
        U2 = E0 + U2                                        # :99-100
        U[:] = 0.5 * DU + 0.5 * U2                          # :103-104
        return U

    def ssprk3(self, U):
        dt = self.dt
        U1 = self.exp2(U + self.averaged(U), dt)            # :110-117
        U2 = self.averaged(U1)                              # :120-124
        U2, U1 = self.exp2_many([(U1 + U2, -dt / 2), (U, dt / 2)])  # :126-127, :128
        U2 = 0.75 * U1 + 0.25 * U2                          # :129
        U1 = self.averaged(U2)                              # :132-136
        U2, U1 = self.exp2_many([(U2 + U1, dt / 2), (U, dt)])  # :138-139, :140
        U[:] = 1 / 3 * U1 + 2 / 3 * U2                      # :141
        return U

    def leapfrog(self, U, U_old, asselin):
        """U_old is updated in place as in the reference (:159); returns U."""
        dt = self.dt
        V = self.averaged(U)                                # :147-151
        U1 = self.exp2(U_old, dt)                           # :153 (and :157: the reference applies it twice)
        V = U1 + 2 * V                                      # :154
        U_new = self.exp2(V, dt)                            # :155
        U2 = self.exp2(U_new, -dt)                          # :158
        V = (U1 + U2) * 0.5 - U                             # :159
        U_old[:] = U + asselin * V                          # :160
        U[:] = U_new                                        # :162
        return U


class DeviceAveragedStepper(AveragedStepper):
    """The same schemes with everything resident in HBM.

    U (and U_old) are torch CUDA float64 tensors of n values in the user dof order [u ; eta]; `slow` maps a (B, n)
    device tensor to a (B, n) device tensor.  All library calls and all torch operations of a step are enqueued on the
    library's stream (torch.cuda.ExternalStream), so nothing synchronises with the host and no state vector crosses
    PCIe: Operator.stats()['h2d_vector_bytes'] / ['d2h_vector_bytes'] stay 0 over a step.  With `ensemble` (one process
    per GPU, SampleEnsemble.init_library_comm) the weighted sum is all-reduced by NCCL on the same stream.
    """

    def __init__(self, cheby, cheby2, slow, expt, wt, dt, ensemble=None):
        import torch
        super().__init__(cheby, cheby2, slow, expt, wt, dt, ensemble=ensemble, fused=True)
        if cheby2.op is not cheby.op:
            raise ValueError("cheby and cheby2 must share one Operator")
        self.torch = torch
        self.op = op = cheby.op
        self.n = op.n
        B = self.expt.shape[0]
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(op.stream_ptr())
        self.f_in, self.f_out, self.b_in = op.batch(B), op.batch(B), op.batch(B)
        self._pairs = {}                                   # batches of the cheby2 applies, by size
        op.set_async(True)
        with torch.cuda.stream(self.stream):
            self.F = torch.empty((B, self.n), dtype=torch.float64, device=self.dev)

    def close(self):
        for b in [self.f_in, self.f_out, self.b_in] + [x for pr in self._pairs.values() for x in pr]:
            b.close()
        self._pairs = {}

    def averaged(self, U):
        torch = self.torch
        self.f_in.upload_dev(U, shared=True)
        self.cheby.apply_dev(self.f_in, self.f_out, self.expt)
        self.f_out.export_dev(self.F)
        N = self.slow(self.F)                               # device callback: SlowSolver.solve() for every member
        if N.dtype != torch.float64 or not N.is_contiguous():
            N = N.to(torch.float64).contiguous()
        self.b_in.upload_dev(N)
        v = torch.empty(self.n, dtype=torch.float64, device=self.dev)
        self.cheby.average_dev(self.b_in, -self.expt, self.wt, v.data_ptr())
        if self.ensemble is not None and self.ensemble.world > 1:
            if self.ensemble._lib_comm is not self.op:
                raise RuntimeError("device-resident stepping over several GPUs needs SampleEnsemble.init_library_comm(op)")
            self.op._check(self.op.lib.kexp_allreduce_sum_dev(self.op.ctx, v.data_ptr(), self.n), "kexp_allreduce_sum_dev")
        return v

    def exp2_many(self, pairs):
        """The independent cheby2 applies of a stage as ONE batch (columns = the inputs, each with its own t)."""
        torch = self.torch
        m = len(pairs)
        if m not in self._pairs:
            self._pairs[m] = (self.op.batch(m), self.op.batch(m))
        bi, bo = self._pairs[m]
        X = torch.stack([x for x, _ in pairs])
        bi.upload_dev(X)
        self.cheby2.apply_dev(bi, bo, np.array([t for _, t in pairs], dtype=np.float64))
        Y = torch.empty_like(X)
        bo.export_dev(Y)
        return [Y[k] for k in range(m)]

    def exp2(self, X, t):
        return self.exp2_many([(X, t)])[0]

    def _on_stream(self, fn, *args):
        with self.torch.cuda.stream(self.stream):
            return fn(*args)

    def rk2(self, U):
        return self._on_stream(super().rk2, U)

    def rk4(self, U):
        return self._on_stream(super().rk4, U)

    def heuns(self, U):
        return self._on_stream(super().heuns, U)

    def ssprk3(self, U):
        return self._on_stream(super().ssprk3, U)

    def leapfrog(self, U, U_old, asselin):
        return self._on_stream(super().leapfrog, U, U_old, asselin)

    def synchronize(self):
        """Joins the stream and raises if a mass solve of the queued applies missed its tolerance."""
        self.op.synchronize()
        if hasattr(self.slow, "synchronize"):
            self.slow.synchronize()
