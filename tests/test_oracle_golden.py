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


def test_exact_rewritings_on_the_cpu(sw2):
    """The three rewritings the GPU path relies on, restated in NumPy and checked against the reference recurrence
    (apply_batch, itself pinned by the golden vectors above): re-expansion onto a common operator, per-column
    economised series, and the weighted average as one Clenshaw recurrence."""
    from oracle import cheby_ref, coeffs, samples
    from oracle.assemble_sw import test_pulse
    from scipy.sparse.linalg import splu
    sw = sw2
    lu = splu(sw.M().tocsc())
    A = lambda V: lu.solve(sw.K @ V)
    eig, dt, rho, B = 0.0018, 3600.0, 6.0, 12                   # ref 2: largest |eig| ~ 1.7e-3
    L = eig * dt * rho
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    a = coeffs.real_recurrence_coeffs(c, n)
    t, w = samples.sample_times_weights(B, dt, rho)
    x = test_pulse(sw)
    rng = np.random.default_rng(3)
    X = x[:, None] * (1 + 0.1 * np.arange(B))[None, :] + 1e-3 * np.abs(x).max() * rng.standard_normal((sw.n, B))
    ref = cheby_ref.apply_batch(a, L, A, X, -t)
    M = sw.M()
    rel = lambda p, q: np.sqrt(((p - q) @ (M @ (p - q))) / (q @ (M @ q)))
    Ye, Me = cheby_ref.apply_batch_economised(a, L, A, X, -t, t_valid=dt * rho)
    assert Me < n and max(rel(Ye[:, k], ref[:, k]) for k in range(B)) < 1e-12
    v, Mc = cheby_ref.average_clenshaw(a, L, A, X, -t, w)
    assert rel(v, ref @ w) < 1e-12
    # shared input: the forward half, all columns from one basis
    tstar = np.abs(t).max()
    S = [x, (tstar / L) * A(x)]
    for m in range(2, n + 1):
        S.append(2.0 * (tstar / L) * A(S[-1]) + S[-2])
    S = np.array(S).T                                            # N x (n+1): U_m = S_m((t*/L) A) x
    F = cheby_ref.apply_batch(a, L, A, np.repeat(x[:, None], B, axis=1), t)
    for k in range(B):
        assert rel(S @ cheby_ref.reexpand(a, t[k] / tstar), F[:, k]) < 1e-12
