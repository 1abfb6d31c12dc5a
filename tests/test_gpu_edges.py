"""GPU: edge cases and error behaviour of the C ABI / Python mirror."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _relM(M, a, b):
    d = a - b
    return np.sqrt((d @ (M @ d)) / (b @ (M @ b)))


@pytest.fixture(scope="module")
def setup(sw2):
    import krylov_exp_b200 as kx
    op = kx.Operator(sw2.K, sw2.Mu, sw2.Meta_inv_blocks, rtol=1e-13, max_it=80)
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, 0.0017 * 3600.0)
    yield sw2, op, ce
    op.close()


def test_t_zero_is_identity(setup):
    sw, op, ce = setup
    from oracle.assemble_sw import test_pulse
    x = test_pulse(sw)
    y = np.empty(sw.n)
    ce.apply(x, y, 0.0)                       # series renormalised to 1 at the origin (cheby_exp.py:68-83)
    assert np.abs(y - x).max() <= 1e-14 * np.abs(x).max()
    Y = ce.apply_batch(x, np.zeros(3), shared=True)
    assert np.abs(Y - x[None, :]).max() <= 1e-14 * np.abs(x).max()


@pytest.mark.parametrize("B", [1, 2, 31, 33, 65, 97, 130])
def test_ragged_batches_match_single_applies(setup, B):
    """Multi-panel batches with a ragged last panel: every column equals its single-vector apply."""
    sw, op, ce = setup
    from oracle.assemble_sw import test_pulse
    x = test_pulse(sw)
    rng = np.random.default_rng(B)
    X = x[None, :] * (1.0 + 0.2 * rng.standard_normal((B, 1)))
    t = 1800.0 * (rng.random(B) - 0.5)
    Y = ce.apply_batch(X, t)
    M = sw.M()
    for k in {0, B // 2, B - 1}:
        y = np.empty(sw.n)
        ce.apply(X[k], y, float(t[k]))
        assert _relM(M, Y[k], y) < 1e-11
    w = rng.random(B)
    v = ce.apply_average(X, t, w)
    assert _relM(M, v, w @ Y) < 1e-12
    Ys = ce.apply_batch(x, t, shared=True)    # shared-input path, same ragged layout
    y = np.empty(sw.n)
    ce.apply(x, y, float(t[B - 1]))
    assert _relM(M, Ys[B - 1], y) < 1e-11


def test_strided_host_batches(setup):
    """ldx, ldy > n at the C ABI (vectors embedded in a wider host array)."""
    sw, op, ce = setup
    from krylov_exp_b200 import _lib
    from oracle.assemble_sw import test_pulse
    x = test_pulse(sw)
    B, pad = 5, 37
    Xw = np.zeros((B, sw.n + pad)); Yw = np.full((B, sw.n + pad), 7.0)
    Xw[:, :sw.n] = x[None, :] * np.arange(1, B + 1)[:, None]
    t = np.linspace(-900, 900, B)
    rc = op.lib.kexp_cheby_apply_batch(op.ctx, ce.id, _lib.ptr(Xw), sw.n + pad, 0, _lib.ptr(Yw), sw.n + pad, _lib.ptr(t), B)
    assert rc == 0
    assert np.all(Yw[:, sw.n:] == 7.0)                       # padding untouched
    ref = ce.apply_batch(np.ascontiguousarray(Xw[:, :sw.n]), t)
    assert np.abs(Yw[:, :sw.n] - ref).max() == 0.0            # bit-reproducible


def test_zero_and_mixed_zero_columns(setup):
    sw, op, ce = setup
    from oracle.assemble_sw import test_pulse
    x = test_pulse(sw)
    X = np.stack([x, np.zeros(sw.n), 2 * x])
    Y = ce.apply_batch(X, np.array([500.0, 500.0, 500.0]))
    assert np.all(Y[1] == 0.0) and np.all(np.isfinite(Y))
    assert _relM(sw.M(), Y[2], 2 * Y[0]) < 1e-12


def test_error_paths(setup, sw2):
    import krylov_exp_b200 as kx
    from krylov_exp_b200 import _lib
    sw, op, ce = setup
    lib = op.lib
    x = np.zeros(sw.n)
    t = np.zeros(1)
    # B < 1, unknown table id, aliasing, table not set
    assert lib.kexp_cheby_apply_batch(op.ctx, ce.id, _lib.ptr(x), sw.n, 0, _lib.ptr(np.empty(sw.n)), sw.n, _lib.ptr(t), 0) != 0
    assert lib.kexp_cheby_apply(op.ctx, 7, _lib.ptr(x), _lib.ptr(np.empty(sw.n)), 0.0) != 0
    assert b"not set" in lib.kexp_last_error(op.ctx)
    assert lib.kexp_cheby_apply(op.ctx, ce.id, _lib.ptr(x), _lib.ptr(x), 0.0) != 0
    assert lib.kexp_cheby_apply_batch(op.ctx, ce.id, _lib.ptr(x), sw.n - 1, 0, _lib.ptr(np.empty(sw.n)), sw.n, _lib.ptr(t), 1) != 0
    with pytest.raises(kx.KexpError):
        op.batch(0)
    b3, b4 = op.batch(3), op.batch(4)
    with pytest.raises(kx.KexpError):
        ce.apply_dev(b3, b4, np.zeros(3))
    # bad operators: column index out of range, n_u not a multiple of 3, bad permutation
    K = sp.csr_matrix(sw.K); Mu = sp.csr_matrix(sw.Mu)
    ctx = C.c_void_p()
    assert lib.kexp_create(C.byref(ctx), 0) == 0
    bad_col = K.indices.astype(np.int32).copy(); bad_col[0] = K.shape[0] + 5
    args = lambda col: (ctx, sw.n_u, sw.n_eta, _lib.ptr(K.indptr.astype(np.int32)), _lib.ptr(col), _lib.ptr(K.data),
                        _lib.ptr(Mu.indptr.astype(np.int32)), _lib.ptr(Mu.indices.astype(np.int32)), _lib.ptr(Mu.data),
                        _lib.ptr(np.ascontiguousarray(sw.Meta_inv_blocks)), sw.n_eta // 3)
    assert lib.kexp_set_operator(*args(bad_col)) != 0
    assert b"out of range" in lib.kexp_last_error(ctx)
    assert lib.kexp_set_operator(*args(K.indices.astype(np.int32))) == 0
    assert lib.kexp_set_mass_solver(ctx, -1.0, 10, 3) != 0
    assert lib.kexp_set_mass_solver_chebyshev(ctx, 1e-10, 2.0, 1.0, 30) != 0
    assert lib.kexp_set_cheby(ctx, 0, _lib.ptr(np.ones(3)), 2, -1.0) != 0
    assert lib.kexp_destroy(ctx) == 0
    # krylov: its > kdim asserts as the reference does (krylov_exp.py:68); apply before the operator is set errors
    ctx = C.c_void_p()
    assert lib.kexp_create(C.byref(ctx), 0) == 0
    assert lib.kexp_krylov_apply(ctx, _lib.ptr(x), _lib.ptr(np.empty(sw.n)), 1.0, 5, 5, None, None, None) != 0
    assert lib.kexp_destroy(ctx) == 0
