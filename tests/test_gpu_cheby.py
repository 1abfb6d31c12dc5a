"""GPU parity tests (C ABI -> CUDA kernels) against the CPU oracle.  Run with -m gpu."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import splu

pytestmark = pytest.mark.gpu

TOL = 1e-10          # north_star: relative L2 (M-weighted) agreement with the oracle


def _oracle_op(sw):
    lu = splu(sw.M().tocsc())
    return lambda V: lu.solve(sw.K @ V)


def _relM(sw, a, b):
    M = sw.M()
    d = a - b
    return np.sqrt(np.sum(d * (M @ d)) / np.sum(b * (M @ b)))


def _inputs(sw, B, seed=1234):
    """M1-style inputs: pulse + noise per column."""
    from oracle.assemble_sw import test_pulse
    x = test_pulse(sw)
    X = np.empty((B, sw.n))
    for k in range(B):
        rng = np.random.default_rng(seed + k)
        X[k] = x
        X[k, sw.n_u:] += 1e-3 * np.abs(x).max() * rng.standard_normal(sw.n_eta)
        X[k, :sw.n_u] += 1e-3 * np.abs(x).max() * np.sqrt(sw.g / sw.H) * sw.radius / 2 ** sw.ref_level * rng.standard_normal(sw.n_u)
    return x, X


@pytest.fixture(scope="module")
def ops(sw3):
    import krylov_exp_b200 as kx
    plain = kx.Operator(sw3.K, sw3.Mu, sw3.Meta_inv_blocks, condense=False, rtol=1e-13, max_it=120, solver="cg")
    cond = kx.Operator(sw3.K, sw3.Mu, sw3.Meta_inv_blocks, condense=True, rtol=1e-13, max_it=80, solver="cg")
    cheb = kx.Operator(sw3.K, sw3.Mu, sw3.Meta_inv_blocks, condense=True, rtol=1e-13, max_it=80, solver="chebyshev")
    assert cond.condensed and not plain.condensed and cheb.spectrum[1] / cheb.spectrum[0] < 8
    yield plain, cond, cheb
    plain.close(); cond.close(); cheb.close()


@pytest.mark.parametrize("B", [1, 2, 3, 8, 17, 64, 70])
def test_spmm_K_matches_scipy(sw3, ops, B):
    _, X = _inputs(sw3, B)
    Y = ops[0].spmm_K(X)
    ref = (sw3.K @ X.T).T
    assert np.abs(Y - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("which", [0, 1, 2])
@pytest.mark.parametrize("B", [1, 5, 64])
def test_operator_apply_matches_oracle(sw3, ops, which, B):
    _, X = _inputs(sw3, B)
    W = ops[which].apply(X)
    ref = _oracle_op(sw3)(X.T).T
    for k in range(B):
        assert _relM(sw3, W[k], ref[k]) < 1e-11
    st = ops[which].stats()
    assert st["last_max_rel_residual"] <= 1e-13


@pytest.mark.parametrize("which", [0, 1, 2])
def test_cheby_apply_matches_oracle(sw3, ops, which):
    import krylov_exp_b200 as kx
    from oracle import coeffs, cheby_ref
    dt, rho = 3600.0, 1.0
    L = kx.EIGS[0] * dt * rho
    ce = kx.cheby_exp(ops[which], None, None, 10000, 1e-8, L)
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    assert n == ce.ncheb
    a = coeffs.real_recurrence_coeffs(c, n)
    x, X = _inputs(sw3, 9)
    t, w = kx.sample_times_weights(9, dt, rho)
    op = _oracle_op(sw3)
    # single vector, reference signature
    y = kx.MixedVector(sw3.n_u, sw3.n_eta)
    ce.apply(kx.MixedVector(sw3.n_u, sw3.n_eta, x), y, t[0])
    ref = cheby_ref.apply_real(a, L, op, x, t[0])
    assert _relM(sw3, y.data, ref) < TOL
    # batch, distinct inputs
    Y = ce.apply_batch(X, -t)
    ref = cheby_ref.apply_batch(a, L, op, X.T, -t).T
    for k in range(9):
        assert _relM(sw3, Y[k], ref[k]) < TOL
    # shared input, weighted average
    v = ce.apply_average(x, t, w, shared=True)
    refv = cheby_ref.apply_batch(a, L, op, np.repeat(x[:, None], 9, axis=1), t) @ w
    assert _relM(sw3, v, refv) < TOL


def test_cheby_golden_from_reference(golden):
    """Fixture produced by the reference's own cheby_exp class (tests/golden/make_golden.py)."""
    import krylov_exp_b200 as kx
    g = golden
    nu, ne, _ = g["op_meta"]
    nu, ne = int(nu), int(ne)
    n = nu + ne
    K = sp.csr_matrix((g["op_K_data"], g["op_K_indices"], g["op_K_indptr"]), shape=(n, n))
    M = sp.csr_matrix((g["op_M_data"], g["op_M_indices"], g["op_M_indptr"]), shape=(n, n))
    Mu = M[:nu][:, :nu].tocsr()
    Meta = M[nu:][:, nu:].tobsr(blocksize=(3, 3)).data
    for condense, solver in ((False, "cg"), (True, "cg"), (True, "chebyshev"), (False, "chebyshev")):
        op = kx.Operator(K, Mu, np.linalg.inv(Meta), condense=condense, rtol=1e-14, max_it=150, solver=solver)
        T, L, nc = g["apply_meta"]
        ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
        assert ce.ncheb == int(nc)
        assert np.abs(ce.ChebCoeffs[:ce.ncheb + 2] - g["apply_c"]).max() < 1e-14
        i = 0
        for xin in (g["x0"], g["x1"]):
            for t in g["apply_t"]:
                y = np.empty(n)
                ce.apply(xin, y, float(t))
                ref = g["apply_y"][i]
                d = y - ref
                rel = np.sqrt((d @ (M @ d)) / (ref @ (M @ ref)))
                assert rel < TOL, (condense, i, rel)
                i += 1
        Tf, Lf, ncf = g["applyf_meta"]
        cef = kx.cheby_exp(op, None, None, 10000, 1e-8, Lf, filter=True, filter_val=0.75)
        assert cef.ncheb == int(ncf)
        y = np.empty(n)
        cef.apply(g["x1"], y, float(Tf))
        d = y - g["applyf_y"]
        assert np.sqrt((d @ (M @ d)) / (g["applyf_y"] @ (M @ g["applyf_y"]))) < TOL
        op.close()


def test_alias_and_errors(sw3, ops):
    import krylov_exp_b200 as kx
    ce = kx.cheby_exp(ops[1], None, None, 10000, 1e-8, 5.0)
    x = np.zeros(sw3.n)
    with pytest.raises(ValueError):
        ce.apply(x, x, 1.0)
    # zero input -> zero output, no NaN from 0/0 in CG
    y = np.ones(sw3.n)
    ce.apply(x, y, 100.0)
    assert np.all(y == 0.0)
    # exp(tL) exp(-tL) = I to the truncation tolerance
    _, X = _inputs(sw3, 4)
    t = np.array([900.0, -400.0, 1200.0, 10.0])
    Y = ce.apply_batch(X, t)
    Z = ce.apply_batch(Y, -t)
    for k in range(4):
        assert _relM(sw3, Z[k], X[k]) < 1e-7
