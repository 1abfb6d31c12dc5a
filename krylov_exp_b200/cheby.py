"""Drop-in mirror of the reference's `cheby_exp` class on top of libkexp.so.

Same constructor and `apply(x, y, t)` as /root/reference/cheby_exp.py:9-10,110, so the
averaged / standard model scripts keep working unchanged (timestepping_methods.py calls
only `.apply`).  What changes is what sits behind it:

  * the coefficient table (cheby_exp.py:31-96) is computed here on the host with NumPy
    and collapsed to the real table a_n = Re(c_n (-i)^n): T_n(tA/(iL))x is purely real
    for even n and purely imaginary for odd n, so the reference's 2*ncheb-1 operator
    solves (half of them on zero vectors, cheby_exp.py:145-152) reduce to ncheb;
  * `operator_solver` may be a krylov_exp_b200.Operator (matrices already on the GPU) or
    a Firedrake LinearVariationalSolver, from which K and M are exported once as CSR
    (firedrake_adapter.py);
  * apply_batch / apply_average run all averaging-window sample times t_k
    (averaged_sw_explicit.py:233-235) as columns of one device batch.
"""
import ctypes as C

import numpy as np

from . import _lib
from .operator import Operator, DeviceBatch
from .function import as_array, write_back


def chebyshev_coefficients(ncheb, tol, L, filter=False, filter_val=0.75, filter_freq=False):
    """Complex Chebyshev coefficients of exp(ix) on [-L, L] and the truncated length.

    Follows cheby_exp.py:31-96: ncheb+2 Chebyshev nodes, optional (1+(x/fL)^2)^-4 filter
    (:41-46), FFT of the even extension scaled by 1/ncheb (:48-49), halve the end
    coefficients (:52-53), drop trailing terms while their summed magnitude stays below
    tol (:57-63, at least one term must go), rescale so the series is 1 at the origin
    (:69-83; T_n(0) = 0, -1, 0, 1, ... so only even terms contribute).
    """
    n0 = int(ncheb)
    theta = np.arange(np.pi, -np.pi / (n0 + 1) / 2, -np.pi / (n0 + 1))
    nodes = L * np.cos(theta)
    f = np.exp(1j * nodes)
    if filter_freq:
        f = f / (1 + (nodes / (0.003465 * 60 * 60)) ** 2) ** 4          # cheby_exp.py:37-38,43
    elif filter:
        f = f / (1 + (nodes / (filter_val * L)) ** 2) ** 4               # cheby_exp.py:46
    coeffs = np.fft.fft(np.concatenate((f[::-1], f[1:-1])))[:n0 + 2] / n0
    coeffs[0] *= 0.5
    coeffs[-1] *= 0.5
    mag = np.abs(coeffs)
    removed, n = 0.0, n0
    while removed + mag[n + 1] < tol:
        removed += mag[n + 1]
        n -= 1
    if n == n0:
        raise AssertionError("Chebyshev series was not compressed (cheby_exp.py:63)")
    at_zero = coeffs[0]
    sign = -1.0
    for k in range(2, n + 1, 2):           # T_2(0) = -1, T_4(0) = +1, ...
        at_zero = at_zero + sign * coeffs[k]
        sign = -sign
    return coeffs / at_zero, n


def real_table(coeffs, n):
    """a_k = Re(c_k (-i)^k): y = sum_k a_k S_k with S_0 = x, S_1 = Bx, S_k = 2B S_{k-1} + S_{k-2}."""
    k = np.arange(n + 1)
    return np.ascontiguousarray(np.real(coeffs[:n + 1] * (-1j) ** k))


def reexpand_series(a, s, eps=1e-13):
    """b with sum_n a_n S_n(s z) = sum_m b_m S_m(z) (|s| <= 1) and the truncated length for a dropped tail
    <= eps*sum|b|: the library's own host routine (kexp_series_reexpand), no GPU involved."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.empty_like(a)
    keep = C.c_int(0)
    rc = _lib.load().kexp_series_reexpand(_lib.ptr(a), a.shape[0] - 1, float(s), float(eps), _lib.ptr(b), C.byref(keep))
    if rc != 0:
        raise ValueError("kexp_series_reexpand: need |s| <= 1")
    return b, int(keep.value)


class cheby_exp(object):
    def __init__(self, operator_solver, operator_in, operator_out,
                 ncheb, tol, L, filter=False, filter_val=0.75, filter_freq=False, *, spectral_radius=None):
        """Same positional signature as the reference (cheby_exp.py:5-6).  `spectral_radius` (keyword only, an
        extension) is the bound L was built from; see set_spectral_radius."""
        if isinstance(operator_solver, Operator):
            self.op = operator_solver
        else:
            from .firedrake_adapter import operator_from_firedrake
            self.op = operator_from_firedrake(operator_solver, operator_in)
        self.operator_solver = operator_solver
        self.operator_in = operator_in
        self.operator_out = operator_out
        self.ChebCoeffs, self.ncheb = chebyshev_coefficients(ncheb, tol, L, filter, filter_val, filter_freq)
        self.L = L
        self.a = real_table(self.ChebCoeffs, self.ncheb)
        self.id = self.op.new_cheby_id()
        lib = self.op.lib
        self.op._check(lib.kexp_set_cheby(self.op.ctx, self.id, _lib.ptr(self.a), self.ncheb, float(L)), "kexp_set_cheby")
        self.t_valid = 0.0
        if spectral_radius is not None:
            self.set_spectral_radius(spectral_radius)

    def set_spectral_radius(self, eig):
        """State the bound `eig` >= rho(M^-1 K) that L was built from (the drivers' `eigs[ref_level-3]`,
        averaged_sw_explicit.py:77-82: L = eig*dt*rho).  The series is then valid for |t| <= L/eig and sample
        times below that evaluate the same polynomial through a shorter, exactly re-expanded series
        (include/kexp.h, kexp_set_cheby_time_range).  None/0 switches it off."""
        self.t_valid = float(self.L) / float(eig) if eig else 0.0
        self.op._check(self.op.lib.kexp_set_cheby_time_range(self.op.ctx, self.id, self.t_valid), "kexp_set_cheby_time_range")

    def series_terms(self, t):
        """Operator applications (Chebyshev terms) the apply spends on each sample time: the length of the economised series
        when a spectral radius is set, ncheb otherwise.  Host only (kexp_series_reexpand); used to balance samples over GPUs."""
        t = np.atleast_1d(np.asarray(t, dtype=np.float64))
        if not self.t_valid:
            return np.full(t.shape, int(self.ncheb), dtype=np.int64)
        out = np.empty(t.shape, dtype=np.int64)
        for k, tk in enumerate(t):
            s = abs(float(tk)) * 1.01 / self.t_valid
            out[k] = int(self.ncheb) if s > 1.0 else max(reexpand_series(self.a, s)[1] - 1, 1)
        return out

    # ---- the reference interface -------------------------------------------------
    def apply(self, x, y, t):
        """y <- exp(t M^-1 K) x   (cheby_exp.py:110-165).  x is left unchanged, x is not y."""
        if x is y:
            raise ValueError("x and y must be different (cheby_exp.py:119 zeroes y before reading x)")
        xin = np.ascontiguousarray(as_array(x), dtype=np.float64).reshape(-1)
        if xin.shape[0] != self.op.n:
            raise ValueError(f"x must hold n = {self.op.n} values, got {xin.shape[0]}")
        yout = np.empty_like(xin)
        lib = self.op.lib
        self.op._check(lib.kexp_cheby_apply(self.op.ctx, self.id, _lib.ptr(xin), _lib.ptr(yout), float(t)), "kexp_cheby_apply")
        write_back(y, yout)

    # ---- batched over the averaging samples --------------------------------------
    def apply_batch(self, X, t, shared=False, out=None, wait=True):
        """Rows of X (B x n) advanced by their own t[k]; shared=True: one x for every t[k].
        wait=False leaves the device-to-host copy of `out` in flight (it overlaps the calls that follow); `out` is
        complete after Operator.wait_transfers() and should be pinned memory."""
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(-1)
        B = t.shape[0]
        n = self.op.n
        X = np.asarray(X, dtype=np.float64)
        if shared:
            X = np.ascontiguousarray(X.reshape(-1))
            if X.shape[0] != n:
                raise ValueError(f"shared input must hold n = {n} values, got {X.shape[0]}")
            ldx = n
        else:
            X = np.ascontiguousarray(np.atleast_2d(X))
            if X.shape != (B, n):
                raise ValueError(f"X must have shape (B, n) = ({B}, {n}), got {X.shape}")
            ldx = n
        if out is None:
            out = np.empty((B, n))
        elif (not isinstance(out, np.ndarray) or out.dtype != np.float64 or out.shape != (B, n)
              or out.strides[1] != 8 or out.strides[0] % 8 or (B > 1 and out.strides[0] < 8 * n) or not out.flags.writeable):
            raise ValueError(f"out must be a writeable float64 array of shape ({B}, {n}) with unit inner stride")
        lib = self.op.lib
        fn = lib.kexp_cheby_apply_batch if wait else lib.kexp_cheby_apply_batch_async
        self.op._check(fn(self.op.ctx, self.id, _lib.ptr(X), ldx, int(shared),
                          _lib.ptr(out), out.strides[0] // 8 if B > 1 else n, _lib.ptr(t), B), "kexp_cheby_apply_batch")
        return out

    def apply_average(self, X, t, w, shared=False, allreduce=False, fused=False):
        """sum_k w_k exp(t_k L) X_k: `DU *= wt; ensemble.allreduce(DU, V)` (timestepping_methods.py:9-11).
        fused=True never forms the B individual results: the inputs are combined into the vector coefficients of
        one polynomial in M^-1 K, evaluated by a single Clenshaw recurrence (same sum to round-off)."""
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(-1)
        w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
        B = t.shape[0]
        if w.shape[0] != B:
            raise ValueError("t and w must have the same length")
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(-1) if shared else np.atleast_2d(X), dtype=np.float64)
        if X.shape != ((self.op.n,) if shared else (B, self.op.n)):
            raise ValueError(f"X must have shape {'(n,)' if shared else '(B, n)'} with n = {self.op.n}, B = {B}; got {X.shape}")
        v = np.empty(self.op.n)
        lib = self.op.lib
        fn = lib.kexp_cheby_apply_average_fused if fused else lib.kexp_cheby_apply_average
        self.op._check(fn(self.op.ctx, self.id, _lib.ptr(X), self.op.n, int(shared), _lib.ptr(t),
                          _lib.ptr(w), B, _lib.ptr(v), int(allreduce)), "kexp_cheby_apply_average")
        return v

    def average_dev(self, bin_: DeviceBatch, t, w, v_dev_ptr):
        """Fused weighted average of a device-resident batch into the device vector at `v_dev_ptr` (n doubles)."""
        t = np.ascontiguousarray(t, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert t.shape == (bin_.B,) and w.shape == (bin_.B,)
        self.op._check(self.op.lib.kexp_cheby_average_fused_dev(self.op.ctx, self.id, bin_.handle, _lib.ptr(t), _lib.ptr(w),
                                                                C.c_void_p(int(v_dev_ptr))), "kexp_cheby_average_fused_dev")

    def apply_dev(self, bin_: DeviceBatch, bout: DeviceBatch, t):
        """Device-resident batch in, device-resident batch out (no host copies)."""
        t = np.ascontiguousarray(t, dtype=np.float64)
        assert t.shape == (bin_.B,)
        self.op._check(self.op.lib.kexp_cheby_apply_dev(self.op.ctx, self.id, bin_.handle, bout.handle, _lib.ptr(t)), "kexp_cheby_apply_dev")
