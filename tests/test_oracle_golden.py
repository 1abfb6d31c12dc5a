"""CPU: the oracle restatements against the golden vectors produced by running the
reference's own cheby_exp.py / krylov_exp.py classes (tests/golden/make_golden.py)."""
import numpy as np
import scipy.sparse as sp

from oracle import coeffs, cheby_ref, krylov_ref

CASES = ("qjrms", "rho745", "test_cheby", "filter", "filter_freq", "small")


def _ops(g):
    nu, ne, eig = g["op_meta"]
    n = int(nu + ne)
    K = sp.csr_matrix((g["op_K_data"], g["op_K_indices"], g["op_K_indptr"]), shape=(n, n))
    M = sp.csr_matrix((g["op_M_data"], g["op_M_indices"], g["op_M_indptr"]), shape=(n, n))
    return K, M, n


def test_coefficients_match_reference(golden):
    g = golden
    for name in CASES:
        n0, tol, L, nc, f, fv, ff = g[f"coef_{name}_meta"]
        c, ncheb = coeffs.cheby_coeffs(int(n0), tol, L, filter=bool(f), filter_val=fv, filter_freq=bool(ff))
        gc = g[f"coef_{name}_c"]
        assert ncheb == int(nc)
        assert np.abs(c[:len(gc)] - gc).max() == 0.0
        assert abs(np.abs(c[len(gc):]).sum() - g[f"coef_{name}_tailabs"][0]) <= 1e-18
    # term counts quoted in SURVEY.md section 8
    assert int(g["coef_qjrms_meta"][3]) == 29 and int(g["coef_rho745_meta"][3]) == 130
    assert int(g["coef_test_cheby_meta"][3]) == 185


def test_real_table_is_exactly_the_reference_series(golden):
    g = golden
    T, L, nc = g["apply_meta"]
    nc = int(nc)
    c = g["apply_c"]
    assert np.abs(np.imag(c[:nc + 1] * (-1j) ** np.arange(nc + 1))).max() < 1e-11


def test_cheby_apply_literal_and_real(golden):
    g = golden
    K, M, n = _ops(g)
    A = np.linalg.solve(M.toarray(), K.toarray())
    op = lambda v: A @ v
    T, L, nc = g["apply_meta"]
    nc = int(nc)
    c = g["apply_c"]
    a = coeffs.real_recurrence_coeffs(c, nc)
    i = 0
    for xin in (g["x0"], g["x1"]):
        for t in g["apply_t"]:
            ref = g["apply_y"][i]
            yl = cheby_ref.apply_literal(c, nc, L, op, xin, t)
            yr = cheby_ref.apply_real(a, L, op, xin, t)
            assert np.array_equal(yl, ref)                       # same arithmetic, statement by statement
            assert np.linalg.norm(yr - ref) <= 1e-13 * max(np.linalg.norm(ref), 1e-300)
            i += 1
    # batch form == column-wise form
    X = np.stack([g["x0"], g["x1"], g["x1"] * 0.5], axis=1)
    t3 = g["apply_t"][:3]
    Y = cheby_ref.apply_batch(a, L, lambda V: A @ V, X, t3)
    for k in range(3):
        yk = cheby_ref.apply_real(a, L, op, X[:, k], t3[k])
        assert np.linalg.norm(Y[:, k] - yk) <= 1e-13 * np.linalg.norm(yk)


def test_cheby_vs_dense_expm(golden):
    g = golden
    K, M, n = _ops(g)
    T, L, nc = g["apply_meta"]
    y = g["apply_y"][5]                                         # x1, t = +T
    ref = g["expm_x1_T"]
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-7  # truncation tol 1e-8 (cheby_exp.py:57-63)


def test_krylov_restatement(golden):
    g = golden
    K, M, n = _ops(g)
    Md, Kd = M.toarray(), K.toarray()
    A = np.linalg.solve(Md, Kd)
    tk, gamma, kdim = g["krylov_meta"]
    kdim = int(kdim)
    Sinv = np.linalg.solve(Md + gamma * Kd, Md)
    y, H, u, res = krylov_ref.krylov_apply(lambda v: Sinv @ v, lambda v: A @ v, lambda a, b: a @ (Md @ b),
                                           g["x1"], tk, gamma, kdim)
    assert np.array_equal(H, g["krylov_full_H"])
    assert np.linalg.norm(y - g["krylov_full_y"]) <= 1e-13 * np.linalg.norm(y)
    assert np.allclose(res, g["krylov_full_res"], rtol=1e-9)
    y12, _, _, _ = krylov_ref.krylov_apply(lambda v: Sinv @ v, lambda v: A @ v, lambda a, b: a @ (Md @ b),
                                           g["x1"], tk, gamma, kdim, its=12)
    assert np.linalg.norm(y12 - g["krylov_its12_y"]) <= 1e-13 * np.linalg.norm(y12)
