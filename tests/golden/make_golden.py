"""Generate golden vectors by running the reference's OWN classes.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

/root/reference/cheby_exp.py and krylov_exp.py are imported UNMODIFIED.  Their only
dependency that is missing here is the `firedrake` namespace (Function, Constant,
assemble, inner, dx, real, imag, PETSc.Sys.Print).  This script installs a NumPy
stand-in for exactly those names, so the reference's coefficient code
(cheby_exp.py:31-96), its recurrence (cheby_exp.py:110-165) and its Arnoldi loop
(krylov_exp.py:50-109) execute line by line on NumPy vectors.  The operator K, M is the
oracle's ref-level-1 assembly (stored in the fixture, so the fixture is self-contained);
`operator_solver.solve()` and `solver.solve()` are dense LU solves.

What this pins: the algorithm text of the reference (everything in its repository).
What it cannot pin: Firedrake/PETSc assembly and KSP arithmetic (not installable here).
"""
import io
import os
import sys
import types
import contextlib
import numpy as np
import scipy.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


# ------------------------------------------------------------------ firedrake stand-in
class _Expr:
    def __init__(self, arr):
        self.arr = arr


class Constant:
    def __init__(self, v=0.0):
        self.v = float(v)

    def assign(self, v):
        self.v = float(np.real(v))

    def evaluate(self, *a):
        return self.v

    def __neg__(self):
        return Constant(-self.v)

    def __mul__(self, o):
        if isinstance(o, Function):
            return _Expr(self.v * o.data)
        return Constant(self.v * float(o))

    __rmul__ = __mul__

    def __float__(self):
        return self.v


class _FS:
    def __init__(self, n):
        self.n = n


class Function:
    def __init__(self, fs):
        self.fs = fs
        self.data = np.zeros(fs.n)

    def function_space(self):
        return self.fs

    def assign(self, o):
        if isinstance(o, Function):
            self.data[:] = o.data
        elif isinstance(o, _Expr):
            self.data[:] = o.arr
        else:
            self.data[:] = float(o)
        return self

    def _val(self, o):
        if isinstance(o, Function):
            return o.data
        if isinstance(o, Constant):
            return o.v
        if isinstance(o, _Expr):
            return o.arr
        return float(o)

    def __imul__(self, o):
        self.data *= self._val(o)
        return self

    def __itruediv__(self, o):
        self.data /= self._val(o)
        return self

    def __iadd__(self, o):
        self.data += self._val(o)
        return self

    def __isub__(self, o):
        self.data -= self._val(o)
        return self


class _Inner:
    def __init__(self, a, b):
        self.a, self.b = a, b

    def __mul__(self, measure):
        return self


_MASS = {"M": None}


def inner(a, b):
    return _Inner(a, b)


def assemble(form):
    return float(form.a.data @ (_MASS["M"] @ form.b.data))


class _Solver:
    def __init__(self, mat, fin, fout):
        self.mat, self.fin, self.fout = mat, fin, fout
        self.calls = 0

    def solve(self):
        self.calls += 1
        self.fout.data[:] = self.mat @ self.fin.data


def install_shim():
    fd = types.ModuleType("firedrake")
    fd.Function, fd.Constant = Function, Constant
    fd.inner, fd.assemble, fd.dx = inner, assemble, object()
    fd.real, fd.imag = np.real, np.imag
    fd.__all__ = ["Function", "Constant", "inner", "assemble", "dx", "real", "imag"]
    petsc = types.ModuleType("firedrake.petsc")

    class _Sys:
        Print = staticmethod(lambda *a, **k: None)

    class _PETSc:
        Sys = _Sys

    petsc.PETSc = _PETSc
    fd.petsc = petsc
    sys.modules["firedrake"] = fd
    sys.modules["firedrake.petsc"] = petsc


def main():
    install_shim()
    sys.path.insert(0, REF)
    import cheby_exp as ref_cheby          # the reference, unmodified
    import krylov_exp as ref_krylov
    from oracle.assemble_sw import assemble_sw, test_pulse, largest_imag_eig

    out = {}
    # ---------------- coefficient tables (cheby_exp.py:31-96) at the configs of SURVEY 8
    fs1 = _FS(1)
    cases = [("qjrms", 10000, 1e-8, 0.014955 * 900.0, {}),
             ("rho745", 10000, 1e-8, 0.014955 * 900.0 * 7.45, {}),
             ("test_cheby", 10000, 1e-6, 0.007274 * 21600.0, {}),
             ("filter", 10000, 1e-8, 0.014955 * 900.0, {"filter": True, "filter_val": 0.75}),
             ("filter_freq", 10000, 1e-8, 0.014955 * 900.0 * 7.45, {"filter_freq": True}),
             ("small", 200, 1e-10, 3.0, {})]
    for name, n0, tol, L, kw in cases:
        ce = ref_cheby.cheby_exp(None, Function(fs1), Function(fs1), n0, tol, L, **kw)
        keep = ce.ncheb + 40
        out[f"coef_{name}_c"] = np.asarray(ce.ChebCoeffs[:keep])
        out[f"coef_{name}_meta"] = np.array([n0, tol, L, ce.ncheb, float(kw.get("filter", False)),
                                              kw.get("filter_val", 0.75), float(kw.get("filter_freq", False))])
        out[f"coef_{name}_tailabs"] = np.array([np.abs(ce.ChebCoeffs[keep:]).sum()])

    # ---------------- operator: oracle assembly at ref level 1 (stored in the fixture)
    op = assemble_sw(1)
    K = op.K.toarray()
    M = op.M().toarray()
    n = op.n
    A = np.linalg.solve(M, K)
    eig = float(np.max(np.abs(np.linalg.eigvals(A).imag)))
    out["op_K_data"], out["op_K_indices"], out["op_K_indptr"] = op.K.data, op.K.indices, op.K.indptr
    Msp = op.M()
    out["op_M_data"], out["op_M_indices"], out["op_M_indptr"] = Msp.data, Msp.indices, Msp.indptr
    out["op_meta"] = np.array([op.n_u, op.n_eta, eig])
    _MASS["M"] = M
    fs = _FS(n)
    x0 = test_pulse(op)
    rng = np.random.default_rng(7)
    amp = 1e-1 * np.abs(x0).max()
    x1 = x0.copy()
    x1[op.n_u:] += amp * rng.standard_normal(op.n_eta)
    # BDM dofs are normal fluxes (velocity x length): scale to comparable energy
    x1[:op.n_u] += amp * np.sqrt(op.g / op.H) * op.radius * rng.standard_normal(op.n_u)
    out["x0"], out["x1"] = x0, x1

    # ---------------- cheby_exp.apply (cheby_exp.py:110-165)
    T = 12.0 / eig                               # L = 12.x on this coarse mesh
    L = 1.0005 * eig * T
    fin, fout = Function(fs), Function(fs)
    solver = _Solver(A, fin, fout)
    ce = ref_cheby.cheby_exp(solver, fin, fout, 10000, 1e-8, L)
    out["apply_meta"] = np.array([T, L, ce.ncheb])
    out["apply_c"] = np.asarray(ce.ChebCoeffs[:ce.ncheb + 2])
    ts = np.array([T, -T, 0.37 * T, -0.5 * T / 9, 0.0])
    ys = []
    for xin in (x0, x1):
        for t in ts:
            xf, yf = Function(fs), Function(fs)
            xf.data[:] = xin
            solver.calls = 0
            ce.apply(xf, yf, t)
            assert solver.calls == 2 * ce.ncheb - 1
            ys.append(yf.data.copy())
    out["apply_t"] = ts
    out["apply_y"] = np.array(ys)
    # filtered variant (averaged_sw_williamson6.py style)
    cef = ref_cheby.cheby_exp(solver, fin, fout, 10000, 1e-8, L, filter=True, filter_val=0.75)
    xf, yf = Function(fs), Function(fs)
    xf.data[:] = x1
    cef.apply(xf, yf, 0.8 * T)
    out["applyf_c"] = np.asarray(cef.ChebCoeffs[:cef.ncheb + 2])
    out["applyf_meta"] = np.array([0.8 * T, L, cef.ncheb])
    out["applyf_y"] = yf.data.copy()

    # ---------------- krylov_exp.apply (krylov_exp.py:50-109), test_krylov.py:27-30 scaling
    tk = 2.0 * T                                 # t*eig ~ 24
    gamma = tk / 10.0
    Sinv = np.linalg.solve(M + gamma * K, M)     # (M + gamma K)^-1 M, test_krylov.py:34-62
    sin, sout = Function(fs), Function(fs)
    ssolver = _Solver(Sinv, sin, sout)
    kdim = 30
    ke = ref_krylov.krylov_exp(ssolver, sin, sout, Constant(gamma), kdim, solver, fin, fout)
    for tag, its in (("full", -999), ("its12", 12)):
        xf, yf = Function(fs), Function(fs)
        xf.data[:] = x1
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ke.apply(xf, yf, tk, its=its) if its > 0 else ke.apply(xf, yf, tk)
        res = [float(l.split()[1]) for l in buf.getvalue().splitlines() if l.startswith("Residual")]
        out[f"krylov_{tag}_y"] = yf.data.copy()
        out[f"krylov_{tag}_H"] = ke.H.copy()
        out[f"krylov_{tag}_res"] = np.array(res)
    out["krylov_meta"] = np.array([tk, gamma, kdim])
    # independent truth for context: dense expm
    out["expm_x1_mtk"] = sla.expm(-tk * A) @ x1
    out["expm_x1_T"] = sla.expm(T * A) @ x1

    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"),
          {k: v.shape for k, v in out.items() if k.startswith(("apply_y", "krylov_full"))})
    print("ncheb", ce.ncheb, "eig0", eig, "krylov residuals", out["krylov_full_res"][[0, 10, 20, -1]])
    print("krylov vs expm", np.linalg.norm(out["krylov_full_y"] - out["expm_x1_mtk"]) / np.linalg.norm(out["expm_x1_mtk"]))
    print("cheby vs expm", np.linalg.norm(out["apply_y"][5] - out["expm_x1_T"]) / np.linalg.norm(out["expm_x1_T"]))


if __name__ == "__main__":
    main()
