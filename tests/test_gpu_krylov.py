"""GPU parity tests of the shift-and-invert Arnoldi path (krylov_exp.py:50-109)."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp
from scipy.sparse.linalg import splu

pytestmark = pytest.mark.gpu


def _relM(M, a, b):
    d = a - b
    return np.sqrt((d @ (M @ d)) / (b @ (M @ b)))


def test_dense_expm_matches_scipy():
    import krylov_exp_b200 as kx
    from krylov_exp_b200 import _lib
    lib = kx.load_library()
    import ctypes as C
    ctx = C.c_void_p()
    assert lib.kexp_create(C.byref(ctx), 0) == 0
    rng = np.random.default_rng(0)
    for m, scale in ((1, 1.0), (3, 0.1), (7, 3.0), (20, 30.0), (41, 80.0), (51, 200.0), (64, 10.0)):
        A = rng.standard_normal((m, m))
        A = (A - A.T) * scale / np.sqrt(m) + 0.05 * rng.standard_normal((m, m))
        E = np.empty((m, m))
        rc = lib.kexp_dense_expm(ctx, _lib.ptr(np.ascontiguousarray(A)), m, _lib.ptr(E))
        assert rc == 0, lib.kexp_last_error(ctx)
        ref = sla.expm(A)
        assert np.abs(E - ref).max() <= 2e-12 * max(1.0, np.abs(ref).max()), (m, np.abs(E - ref).max())
    assert lib.kexp_dense_expm(ctx, _lib.ptr(np.zeros((65, 65))), 65, _lib.ptr(np.zeros((65, 65)))) != 0
    lib.kexp_destroy(ctx)


def test_krylov_golden_from_reference(golden):
    """Fixture produced by the reference's own krylov_exp class (tests/golden/make_golden.py)."""
    import krylov_exp_b200 as kx
    g = golden
    nu, ne, _ = g["op_meta"]
    nu, ne = int(nu), int(ne)
    n = nu + ne
    K = sp.csr_matrix((g["op_K_data"], g["op_K_indices"], g["op_K_indptr"]), shape=(n, n))
    M = sp.csr_matrix((g["op_M_data"], g["op_M_indices"], g["op_M_indptr"]), shape=(n, n))
    Mu = M[:nu][:, :nu].tocsr()
    Meta = M[nu:][:, nu:].tobsr(blocksize=(3, 3)).data
    tk, gamma, kdim = g["krylov_meta"]
    kdim = int(kdim)
    op = kx.Operator(K, Mu, np.linalg.inv(Meta), rtol=1e-14, max_it=150)
    sh = kx.ShiftedOperator(op, K, Mu, Meta, gamma, rtol=1e-14)
    ke = kx.krylov_exp(sh, None, None, gamma, kdim, op, None, None)
    y = np.empty(n)
    ke.apply(g["x1"], y, float(tk))
    assert np.abs(ke.H - g["krylov_full_H"]).max() <= 1e-9 * np.abs(g["krylov_full_H"]).max()
    assert _relM(M, y, g["krylov_full_y"]) < 1e-9
    assert np.allclose(ke.residuals, g["krylov_full_res"], rtol=1e-5, atol=1e-16)
    y2 = np.empty(n)
    ke.apply(g["x1"], y2, float(tk), its=12, residuals=False)
    assert _relM(M, y2, g["krylov_its12_y"]) < 1e-9
    with pytest.raises(AssertionError):
        ke.apply(g["x1"], y2, float(tk), its=kdim + 1)
    op.close()


def test_krylov_matches_oracle_and_cheby(sw3):
    """M4-style: SAI-Arnoldi vs the oracle restatement, and vs Chebyshev at -t."""
    import krylov_exp_b200 as kx
    from oracle import krylov_ref
    from oracle.assemble_sw import test_pulse
    sw = sw3
    t = 7200.0
    gamma = t / 10.0                                    # test_krylov.py:27-30
    kdim = 25
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80)
    sh = kx.ShiftedOperator(op, sw.K, sw.Mu, sw.Meta_blocks, gamma)
    ke = kx.krylov_exp(sh, None, None, gamma, kdim, op, None, None)
    x = test_pulse(sw)
    y = np.empty(sw.n)
    ke.apply(x, y, t)
    M = sw.M()
    lu_s = splu((M + gamma * sw.K).tocsc())
    lu_m = splu(M.tocsc())
    yo, Ho, uo, reso = krylov_ref.krylov_apply(lambda v: lu_s.solve(M @ v), lambda v: lu_m.solve(sw.K @ v),
                                               lambda a, b: a @ (M @ b), x, t, gamma, kdim)
    assert _relM(M, y, yo) < 1e-8
    assert np.abs(ke.H - Ho).max() <= 1e-8 * np.abs(Ho).max()
    assert np.allclose(ke.residuals, reso, rtol=1e-4, atol=1e-18)
    # Krylov exp(-tA)x against Chebyshev apply(x, y, -t): agreement at the approximation level
    L = kx.EIGS[0] * t
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    yc = np.empty(sw.n)
    ce.apply(x, yc, -t)
    # kdim = 25 is not converged for t*eig = 25 (the reference only prints its residual estimate): the
    # GPU pair must show the same Krylov-vs-Chebyshev gap as the oracle pair does
    assert abs(_relM(M, y, yc) - _relM(M, yo, yc)) < 1e-7
    assert _relM(M, y, yc) < 5e-3
    op.close()


def test_krylov_ref4_kdim40_matches_oracle():
    """M4-sized Krylov space (kdim 40, test_krylov.py:89-90) at ref 4: the GPU Arnoldi follows the oracle.  The
    SAI-Krylov result itself agrees with Chebyshev only to ~1e-3 here -- with exact solves in the oracle too --
    so that gap is a property of the reference algorithm (gamma = t/10, MGS, inv(H)), not of this port."""
    import krylov_exp_b200 as kx
    from oracle import krylov_ref
    from oracle.assemble_sw import test_pulse
    from conftest import get_sw
    sw = get_sw(4)
    t, kdim = 7200.0, 40
    gamma = t / 10.0
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80)
    sh = kx.ShiftedOperator(op, sw.K, sw.Mu, sw.Meta_blocks, gamma, rtol=1e-13)
    ke = kx.krylov_exp(sh, None, None, gamma, kdim, op, None, None)
    x = test_pulse(sw)
    y = np.empty(sw.n)
    ke.apply(x, y, t, residuals=False)
    M = sw.M()
    lu_s = splu((M + gamma * sw.K).tocsc())
    lu_m = splu(M.tocsc())
    yo, Ho, uo, _ = krylov_ref.krylov_apply(lambda v: lu_s.solve(M @ v), lambda v: lu_m.solve(sw.K @ v),
                                            lambda a, b: a @ (M @ b), x, t, gamma, kdim, residuals=False)
    assert np.abs(ke.H - Ho).max() <= 1e-6 * np.abs(Ho).max()
    assert _relM(M, y, yo) < 1e-5
    op.close()


def test_one_kernel_bicgstab_matches_host_driven_loop(sw3, monkeypatch):
    """The cooperative one-kernel BiCGStab -- matrix resident in shared memory (default) or read from L2
    (KEXP_BICG_SMEM=0) -- and the host-driven loop (KEXP_BICG_HOST=1) solve the same shifted systems to the
    same tolerance: H and y agree far below the Arnoldi sensitivity."""
    import krylov_exp_b200 as kx
    from oracle.assemble_sw import test_pulse
    sw = sw3
    t, kdim = 3600.0, 12
    gamma = t / 10.0
    x = test_pulse(sw)
    out = {}
    for mode, env in (("smem", {}), ("l2", {"KEXP_BICG_SMEM": "0"}), ("host", {"KEXP_BICG_HOST": "1"})):
        for key in ("KEXP_BICG_SMEM", "KEXP_BICG_HOST"):
            monkeypatch.delenv(key, raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80)
        sh = kx.ShiftedOperator(op, sw.K, sw.Mu, sw.Meta_blocks, gamma, rtol=1e-13)
        ke = kx.krylov_exp(sh, None, None, gamma, kdim, op, None, None)
        y = np.empty(sw.n)
        op.stats_reset()
        ke.apply(x, y, t, residuals=False)
        out[mode] = (y, ke.H.copy(), op.stats()["bicg_iterations"])
        op.close()
    M = sw.M()
    for mode in ("smem", "l2"):
        assert _relM(M, out[mode][0], out["host"][0]) < 1e-9, mode
        assert np.abs(out[mode][1] - out["host"][1]).max() <= 1e-9 * np.abs(out["host"][1]).max(), mode
        assert out[mode][2] > 0 and abs(out[mode][2] - out["host"][2]) <= 0.2 * out["host"][2], mode


def test_krylov_batch_over_sample_times(sw3):
    """kexp_krylov_apply_batch: one Arnoldi process serves every sample time of a shared input (the shift-and-invert
    Krylov space does not depend on t); distinct inputs run back to back on the device queue.  Both must reproduce the
    un-batched apply, and the un-batched apply the oracle."""
    import krylov_exp_b200 as kx
    from oracle import krylov_ref
    from oracle.assemble_sw import test_pulse
    sw = sw3
    T = 3600.0
    gamma = T / 10.0
    kdim = 18
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80)
    sh = kx.ShiftedOperator(op, sw.K, sw.Mu, sw.Meta_blocks, gamma, rtol=1e-13)
    ke = kx.krylov_exp(sh, None, None, gamma, kdim, op, None, None)
    x = test_pulse(sw)
    t, w = kx.sample_times_weights(7, T, 1.0)
    M = sw.M()
    Ys = ke.apply_batch(x, t, shared=True)
    singles = []
    for k in range(len(t)):
        y = np.empty(sw.n)
        ke.apply(x, y, float(t[k]), residuals=False)
        singles.append(y)
        assert _relM(M, Ys[k], y) < 1e-12, k
    X = np.stack([x * (1.0 + 0.3 * k) + 0.05 * np.roll(x, 7 * k) for k in range(len(t))])
    Yd = ke.apply_batch(X, t)
    for k in (0, 3, 6):
        y = np.empty(sw.n)
        ke.apply(X[k], y, float(t[k]), residuals=False)
        assert _relM(M, Yd[k], y) < 1e-12, k
    X[2] = 0.0                                           # a zero column: exp(.) 0 = 0, no breakdown error
    Yz = ke.apply_batch(X, t)
    assert np.all(Yz[2] == 0.0) and _relM(M, Yz[5], Yd[5]) < 1e-12
    lu_s = splu((M + gamma * sw.K).tocsc())
    lu_m = splu(M.tocsc())
    yo, Ho, uo, _ = krylov_ref.krylov_apply(lambda v: lu_s.solve(M @ v), lambda v: lu_m.solve(sw.K @ v),
                                            lambda a, b: a @ (M @ b), x, float(t[5]), gamma, kdim, residuals=False)
    assert _relM(M, singles[5], yo) < 1e-8
    op.close()


def test_krylov_reference_call_form(sw3):
    """krylov_exp(solver, solver_in, solver_out, gamma, kdim, operator_solver, ...) with a `solver` that is NOT a
    ShiftedOperator (test_krylov.py:44-62,89-90): anything that can hand over the matrices (a Firedrake
    LinearVariationalSolver through firedrake_adapter, or an object with kexp_matrices()) is accepted."""
    import krylov_exp_b200 as kx
    from oracle.assemble_sw import test_pulse
    sw = sw3

    class HostSolver:                                    # stand-in for the Firedrake solver of the shifted problem
        def kexp_matrices(self):
            return sw.K, sw.Mu, sw.Meta_blocks

    T = 3600.0
    gamma = T / 10.0
    ke = kx.krylov_exp(HostSolver(), None, None, gamma, 12, None, None, None)
    x = test_pulse(sw)
    y = np.empty(sw.n)
    ke.apply(x, y, T, residuals=False)
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80)
    sh = kx.ShiftedOperator(op, sw.K, sw.Mu, sw.Meta_blocks, gamma)
    ke2 = kx.krylov_exp(sh, None, None, gamma, 12, op, None, None)
    y2 = np.empty(sw.n)
    ke2.apply(x, y2, T, residuals=False)
    assert _relM(sw.M(), y, y2) < 1e-10
    op.close()
    ke.op.close()
