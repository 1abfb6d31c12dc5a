"""Golden vectors for the batched averaged time-steppers, produced by running the reference's
/root/reference/timestepping_methods.py UNMODIFIED: every ensemble member is a Python thread with
its own NumPy-backed Function objects and its own instance of the reference's cheby_exp class;
`ensemble.allreduce` is a barrier + sum.  The operator is the fixture's ref-level-1 assembly, the
nonlinear `SlowSolver` a fixed synthetic quadratic map (documented below; the real one is
host-side Firedrake and out of scope).

    python tests/golden/make_golden_timestepping.py
"""
import os
import sys
import threading

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as mg                      # the firedrake stand-in

REF = "/root/reference"


def slow_map(v, scale):
    """Synthetic N(.): quadratic, couples neighbouring dofs, O(1e-2) relative size."""
    return 1e-2 * v * np.roll(v, 1) / scale


class _Expr(mg._Expr):
    pass


def _patch_function_algebra():
    F = mg.Function
    val = lambda o: o.data if isinstance(o, F) else (o.arr if isinstance(o, mg._Expr) else o)
    F.__add__ = lambda a, b: mg._Expr(a.data + val(b))
    F.__radd__ = F.__add__
    F.__sub__ = lambda a, b: mg._Expr(a.data - val(b))
    F.__mul__ = lambda a, b: mg._Expr(a.data * val(b))
    F.__rmul__ = F.__mul__
    E = mg._Expr
    E.__add__ = lambda a, b: mg._Expr(a.arr + val(b))
    E.__radd__ = E.__add__
    E.__sub__ = lambda a, b: mg._Expr(a.arr - val(b))
    E.__mul__ = lambda a, b: mg._Expr(a.arr * val(b))
    E.__rmul__ = E.__mul__


class ThreadEnsemble:
    def __init__(self, size):
        self.size = size
        self.barrier = threading.Barrier(size)
        self.slots = [None] * size
        self.lock = threading.Lock()

    def member(self, rank):
        ens = self

        class _M:
            def allreduce(self, f, f_reduced):
                ens.slots[rank] = f.data.copy()
                ens.barrier.wait()
                total = np.zeros_like(f.data)
                for r in range(ens.size):            # fixed order: identical on every member
                    total += ens.slots[r]
                ens.barrier.wait()
                f_reduced.data[:] = total
        return _M()


def main():
    mg.install_shim()
    _patch_function_algebra()
    sys.path.insert(0, REF)
    import cheby_exp as ref_cheby
    import timestepping_methods as ref_ts
    g = np.load(os.path.join(HERE, "reference_golden.npz"))
    nu, ne, eig = g["op_meta"]
    n = int(nu + ne)
    K = sp.csr_matrix((g["op_K_data"], g["op_K_indices"], g["op_K_indptr"]), shape=(n, n)).toarray()
    M = sp.csr_matrix((g["op_M_data"], g["op_M_indices"], g["op_M_indptr"]), shape=(n, n)).toarray()
    A = np.linalg.solve(M, K)
    Mbar, dt, rho = 5, 3.0 / eig, 1.0
    Lc = 1.0005 * eig * dt * rho
    svals = np.arange(0.5, Mbar) / Mbar
    weights = np.exp(-1.0 / svals / (1.0 - svals))
    weights = weights / np.sum(weights)
    svals -= 0.5
    x1 = g["x1"]
    scale = np.abs(x1).max()
    fs = mg._FS(n)
    out = {"meta": np.array([Mbar, dt, rho, Lc, 0.3]), "U0": x1}

    def run(scheme):
        ens = ThreadEnsemble(Mbar)
        results = [None] * Mbar
        old_results = [None] * Mbar

        def member(rank):
            mk = lambda: mg.Function(fs)
            fin, fout = mk(), mk()
            solver = mg._Solver(A, fin, fout)
            cheby = ref_cheby.cheby_exp(solver, fin, fout, 10000, 1e-8, Lc)
            cheby2 = ref_cheby.cheby_exp(solver, fin, fout, 10000, 1e-8, Lc)
            U, USlow_in, USlow_out, DU, V = mk(), mk(), mk(), mk(), mk()
            U1, U2, U3, X1, X2, X3 = mk(), mk(), mk(), mk(), mk(), mk()
            U.data[:] = x1

            class Slow:
                def solve(self_inner):
                    USlow_out.data[:] = slow_map(USlow_in.data, scale)
            expt = rho * dt * svals[rank]
            wt = weights[rank]
            e = ens.member(rank)
            W = None
            if scheme == "rk2":
                ref_ts.rk2(U, USlow_in, USlow_out, DU, V, W, expt, e, cheby, cheby2, Slow(), wt, dt)
            elif scheme == "rk4":
                ref_ts.rk4(U, USlow_in, USlow_out, DU, U1, U2, U3, X1, X2, X3, V, W, expt, e, cheby, cheby2, Slow(), wt, dt)
            elif scheme == "heuns":
                ref_ts.heuns(U, USlow_in, USlow_out, DU, U1, U2, W, expt, e, cheby, cheby2, Slow(), wt, dt)
            elif scheme == "ssprk3":
                ref_ts.ssprk3(U, USlow_in, USlow_out, DU, U1, U2, W, expt, e, cheby, cheby2, Slow(), wt, dt)
            elif scheme == "leapfrog":
                U_old, U_new = mk(), mk()
                U_old.data[:] = 0.97 * x1
                ref_ts.leapfrog(U, USlow_in, USlow_out, U_old, U_new, DU, U1, U2, V, W, expt, e, cheby, cheby2, Slow(), wt, dt, 0.3)
                old_results[rank] = U_old.data.copy()
            results[rank] = U.data.copy()

        th = [threading.Thread(target=member, args=(r,)) for r in range(Mbar)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for r in range(1, Mbar):                     # every member ends with the same state
            assert np.array_equal(results[r], results[0])
        return results[0], old_results[0]

    for scheme in ("rk2", "rk4", "heuns", "ssprk3", "leapfrog"):
        U, Uold = run(scheme)
        out[f"{scheme}_U"] = U
        if Uold is not None:
            out[f"{scheme}_Uold"] = Uold
        print(scheme, np.linalg.norm(U - x1) / np.linalg.norm(x1))
    np.savez_compressed(os.path.join(HERE, "timestepping_golden.npz"), **out)


if __name__ == "__main__":
    main()
