"""Drop-in mirror of the reference's `krylov_exp` class (shift-and-invert Arnoldi).

Same constructor and `apply(x, y, t, its=-999)` as /root/reference/krylov_exp.py:6-7,50.
`solver` implements w = (M + gamma K)^-1 M v (test_krylov.py:34-62); on the GPU the depth
unknowns are eliminated on the host once (M_eta is block diagonal):

    (M_u + gamma K_uu - gamma^2 K_ue M_eta^-1 K_eu) w_u = M_u v_u - gamma K_ue v_eta
    w_eta = v_eta - gamma M_eta^-1 K_eu w_u

and the sparse, mildly non-symmetric system is solved by block-Jacobi BiCGStab on the
device.  The modified Gram-Schmidt sweep in the M inner product (krylov_exp.py:73-80),
the small Hessenberg inverse / exponential (:87-99) and the final combination (:102-109)
all run in libkexp.so; see csrc/kexp_krylov.cu.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from .operator import Operator, _csr, _csr_args, _blockdiag
from .function import as_array, write_back


class ShiftedOperator:
    """Host-side description of `solver` for the GPU: built from K, M_u, M_eta and gamma."""

    def __init__(self, op: Operator, K, Mu, Meta_blocks, gamma, rtol=1e-13, max_it=2000):
        K, Mu = _csr(K), _csr(Mu)
        n_u = Mu.shape[0]
        n = K.shape[0]
        Meta_blocks = np.ascontiguousarray(Meta_blocks, dtype=np.float64)
        Meta = _blockdiag(Meta_blocks)
        Meta_inv = _blockdiag(np.linalg.inv(Meta_blocks))
        Kuu, Kue, Keu = K[:n_u][:, :n_u], K[:n_u][:, n_u:], K[n_u:][:, :n_u]
        g = float(gamma)
        self.gamma = g
        self.Ash = _csr(Mu + g * Kuu - (g * g) * (Kue @ Meta_inv @ Keu))
        self.Brhs = _csr(sp.hstack([Mu, -g * Kue]))
        self.Eback = _csr(-g * (Meta_inv @ Keu))
        self.M = _csr(sp.block_diag([Mu, Meta]))
        self.op = op
        a, b, e, m = _csr_args(self.Ash), _csr_args(self.Brhs), _csr_args(self.Eback), _csr_args(self.M)
        op._check(op.lib.kexp_set_krylov_operator(
            op.ctx, g, _lib.ptr(a[0]), _lib.ptr(a[1]), _lib.ptr(a[2]), _lib.ptr(b[0]), _lib.ptr(b[1]), _lib.ptr(b[2]),
            _lib.ptr(e[0]), _lib.ptr(e[1]), _lib.ptr(e[2]), _lib.ptr(m[0]), _lib.ptr(m[1]), _lib.ptr(m[2]),
            float(rtol), int(max_it)), "kexp_set_krylov_operator")


def _shifted_from(solver, solver_in, gamma, operator_solver):
    """The reference's call form (test_krylov.py:44-62,89-90): `solver` is whatever solves (M + gamma K') w = M v.
    Accepted besides a ShiftedOperator:
      * an object with `kexp_matrices()` -> (K, Mu, Meta_blocks) -- the matrices of M^-1 K in this package's layout
        (any host assembler; the tests use it as the stand-in for Firedrake);
      * a Firedrake LinearVariationalSolver: the bilinear form of its problem IS M + gamma K and its right-hand side
        is M solver_in, so both are exported once as AIJ/CSR (firedrake_adapter.shifted_matrices_from_firedrake,
        following L_eigs.py:44-45) and K = (A_shifted - M) / gamma."""
    g = gamma.evaluate(0, 0, 0, 0) if hasattr(gamma, "evaluate") else float(gamma)   # krylov_exp.py:59
    if hasattr(solver, "kexp_matrices"):
        K, Mu, Meta_blocks = solver.kexp_matrices()
    else:
        from .firedrake_adapter import shifted_matrices_from_firedrake
        K, Mu, Meta_blocks = shifted_matrices_from_firedrake(solver, solver_in, g)
    op = operator_solver if isinstance(operator_solver, Operator) else Operator(K, Mu, np.linalg.inv(Meta_blocks))
    return ShiftedOperator(op, K, Mu, Meta_blocks, g)


class krylov_exp(object):
    def __init__(self, solver, solver_in, solver_out, gamma,
                 kdim, operator_solver, operator_in, operator_out):
        if not isinstance(solver, ShiftedOperator):
            solver = _shifted_from(solver, solver_in, gamma, operator_solver)
        self.solver = solver
        self.solver_in, self.solver_out = solver_in, solver_out
        g = gamma.evaluate(0, 0, 0, 0) if hasattr(gamma, "evaluate") else float(gamma)   # krylov_exp.py:59
        if abs(g - solver.gamma) > 1e-14 * abs(g):
            raise ValueError("gamma differs from the one the shifted operator was built with")
        self.gamma = gamma
        self.kdim = int(kdim)
        self.operator_solver, self.operator_in, self.operator_out = operator_solver, operator_in, operator_out
        self.op = solver.op
        self.H = np.zeros((self.kdim + 1, self.kdim))             # krylov_exp.py:45
        self.residuals = None
        self.u = None

    def apply(self, x, y, t, its=-999, residuals=True):
        """y <- exp(-t M^-1 K) x by SAI-Arnoldi (krylov_exp.py:50-109; note the sign, :89,96).

        `self.residuals` holds the reference's residual estimates (krylov_exp.py:95-101).  Deviation: the constant C of
        the estimate is ||M^-1 K w|| from THIS context's operator; the reference applies whatever `operator_solver` it was
        constructed with (its docstring and test_krylov.py pass the operator of the shifted problem there), so the printed
        estimates agree with the reference only when that argument is an M^-1 K solver, as in the golden vectors.  The
        result y does not depend on it."""
        if its < 0:
            its = self.kdim                                          # krylov_exp.py:66-67
        assert its <= self.kdim                                      # krylov_exp.py:68
        xin = np.ascontiguousarray(as_array(x), dtype=np.float64)
        yout = np.empty_like(xin)
        u = np.zeros(its)
        res = np.zeros(its) if residuals else None
        self.H[:] = 0.0
        lib = self.op.lib
        self.op._check(lib.kexp_krylov_apply(self.op.ctx, _lib.ptr(xin), _lib.ptr(yout), float(t), self.kdim, int(its),
                                             _lib.ptr(self.H), _lib.ptr(u), _lib.ptr(res)), "kexp_krylov_apply")
        self.u, self.residuals = u, res
        write_back(y, yout)

    def apply_batch(self, X, t, shared=False, its=-999):
        """exp(-t_k M^-1 K) for B sample times at once.  shared=True: one input x for every t_k -- the shift-and-invert
        Krylov space does not depend on t, so ONE Arnoldi process serves all sample times; otherwise row k of X (B x n)
        is advanced by t[k] (Arnoldi processes back to back on the device queue)."""
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(-1)
        B = t.shape[0]
        n = self.op.n
        if its < 0:
            its = self.kdim
        assert its <= self.kdim                                      # krylov_exp.py:68
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(-1) if shared else np.atleast_2d(X), dtype=np.float64)
        if X.shape != ((n,) if shared else (B, n)):
            raise ValueError(f"X must have shape {'(n,)' if shared else '(B, n)'} with n = {n}, B = {B}; got {X.shape}")
        Y = np.empty((B, n))
        self.op._check(self.op.lib.kexp_krylov_apply_batch(self.op.ctx, _lib.ptr(X), n, int(shared), _lib.ptr(Y), n, _lib.ptr(t), B,
                                                           self.kdim, int(its)), "kexp_krylov_apply_batch")
        return Y
