"""GPU: per-column economised series (kexp_set_cheby_time_range).  With the spectral bound the drivers build L
from (L = eigs*dt*rho, averaged_sw_explicit.py:77-82,168) every column evaluates the reference polynomial
sum_n a_n S_n((t_k/L) A) through its exact re-expansion on the sub-interval its spectrum fills -- fewer terms,
same result.  Checked against the oracle's full series and against the full-series GPU path."""
import numpy as np
import pytest
from scipy.sparse.linalg import splu

pytestmark = pytest.mark.gpu

DT, RHO = 900.0, 4.0          # L = eig * 3600: 45 terms at tol 1e-8 on the ref-3 operator


def _relM(M, a, b):
    d = a - b
    return np.sqrt((d @ (M @ d)) / (b @ (M @ b)))


@pytest.fixture(scope="module")
def setup(sw3):
    import krylov_exp_b200 as kx
    from oracle import coeffs
    from oracle.assemble_sw import test_pulse
    sw = sw3
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80)
    L = kx.EIGS[0] * DT * RHO
    full = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    econ = kx.cheby_exp(op, None, None, 10000, 1e-8, L, spectral_radius=kx.EIGS[0])
    assert econ.t_valid == pytest.approx(DT * RHO)
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    a = coeffs.real_recurrence_coeffs(c, n)
    lu = splu(sw.M().tocsc())
    A = lambda V: lu.solve(sw.K @ V)
    x = test_pulse(sw)
    yield sw, op, full, econ, a, L, A, x
    op.close()


def _inputs(sw, x, B):
    X = np.empty((B, sw.n))
    for k in range(B):
        rng = np.random.default_rng(77 + k)
        X[k] = x * (1 + 0.05 * k) + 1e-3 * np.abs(x).max() * rng.standard_normal(sw.n)
    return X


@pytest.mark.parametrize("B", [1, 9, 64, 70])
def test_economised_matches_oracle_and_full_series(setup, B):
    """B = 70 spans two panels, so the columns are stored sorted by |t| and come back in the caller's order."""
    import krylov_exp_b200 as kx
    from oracle import cheby_ref
    sw, op, full, econ, a, L, A, x = setup
    t, w = kx.sample_times_weights(B, DT, RHO)
    X = _inputs(sw, x, B)
    full.apply_batch(X, -t); econ.apply_batch(X, -t)              # settle the adaptive iteration count first
    op.stats_reset()
    Yf = full.apply_batch(X, -t)
    apps_full = op.stats()["op_applications"]
    op.stats_reset()
    Ye = econ.apply_batch(X, -t)
    apps_econ = op.stats()["op_applications"]
    assert apps_econ < apps_full                                  # fewer terms ...
    M = sw.M()
    ref = cheby_ref.apply_batch(a, L, A, X.T, -t).T
    for k in range(B):
        assert _relM(M, Ye[k], ref[k]) < 1e-10, k                 # ... the reference polynomial all the same
        assert _relM(M, Ye[k], Yf[k]) < 1e-11, k
    v = econ.apply_average(X, -t, w)
    assert _relM(M, v, w @ ref) < 1e-10


def test_economised_shared_input(setup):
    import krylov_exp_b200 as kx
    from oracle import cheby_ref
    sw, op, full, econ, a, L, A, x = setup
    t, w = kx.sample_times_weights(64, DT, RHO)
    F = econ.apply_batch(x, t, shared=True)
    ref = cheby_ref.apply_batch(a, L, A, np.tile(x, (64, 1)).T, t).T
    M = sw.M()
    assert max(_relM(M, F[k], ref[k]) for k in range(64)) < 1e-10


def test_times_outside_the_stated_range_run_the_full_series(setup):
    """|t| up to L/eig is legal for the reference series; beyond t_valid/1.01 the library must not economise."""
    from oracle import cheby_ref
    sw, op, full, econ, a, L, A, x = setup
    t = np.array([0.3, -0.999, 1.0]) * DT * RHO
    X = _inputs(sw, x, 3)
    op.stats_reset()
    Y = econ.apply_batch(X, t)
    apps = op.stats()["op_applications"]
    assert apps >= econ.ncheb and apps % econ.ncheb == 0          # one panel, all terms
    ref = cheby_ref.apply_batch(a, L, A, X.T, t).T
    assert max(_relM(sw.M(), Y[k], ref[k]) for k in range(3)) < 1e-10


def test_device_batch_order(setup):
    """kexp_batch_set_order: stored order differs, every interface keeps the caller's order."""
    import krylov_exp_b200 as kx
    sw, op, full, econ, a, L, A, x = setup
    B = 70
    t, w = kx.sample_times_weights(B, DT, RHO)
    X = _inputs(sw, x, B)
    ref = full.apply_batch(X, t)
    bi, bo = op.batch(B), op.batch(B)
    assert bi.num_panels == 2
    bi.set_order(np.argsort(np.abs(t), kind="stable"))
    bi.upload(X)
    assert np.array_equal(bi.download(), X)
    econ.apply_dev(bi, bo, t)
    Y = bo.download()
    M = sw.M()
    assert max(_relM(M, Y[k], ref[k]) for k in range(B)) < 1e-11
    with pytest.raises(kx.KexpError):
        bi.set_order(np.zeros(B, dtype=np.int32))
    bi.close(); bo.close()


@pytest.mark.parametrize("B", [1, 9, 64, 70])
@pytest.mark.parametrize("which", ["full", "econ"])
def test_fused_average_matches_oracle(setup, B, which):
    """sum_k w_k exp(t_k L) X_k by one Clenshaw recurrence (kexp_cheby_apply_average_fused) = the weighted sum of
    the oracle's B separate applications (timestepping_methods.py:9-11)."""
    import krylov_exp_b200 as kx
    from oracle import cheby_ref
    sw, op, full, econ, a, L, A, x = setup
    ce = full if which == "full" else econ
    t, w = kx.sample_times_weights(B, DT, RHO)
    X = _inputs(sw, x, B)
    ref = w @ cheby_ref.apply_batch(a, L, A, X.T, -t).T
    M = sw.M()
    v = ce.apply_average(X, -t, w, fused=True)
    assert _relM(M, v, ref) < 1e-10
    assert _relM(M, v, ce.apply_average(X, -t, w)) < 1e-11
    # shared input: (sum_k w_k p_k)(A) x
    vs = ce.apply_average(x, t, w, shared=True, fused=True)
    refs = w @ cheby_ref.apply_batch(a, L, A, np.tile(x, (B, 1)).T, t).T
    assert _relM(M, vs, refs) < 1e-10


def test_fused_average_device_batch_with_order(setup):
    import torch
    import krylov_exp_b200 as kx
    sw, op, full, econ, a, L, A, x = setup
    B = 70
    t, w = kx.sample_times_weights(B, DT, RHO)
    X = _inputs(sw, x, B)
    ref = econ.apply_average(X, -t, w)
    bi = op.batch(B)
    bi.set_order(np.argsort(np.abs(t), kind="stable"))
    bi.upload(X)
    v = torch.zeros(sw.n, dtype=torch.float64, device="cuda:0")
    econ.average_dev(bi, -t, w, v.data_ptr())
    op.synchronize()
    assert _relM(sw.M(), v.cpu().numpy(), ref) < 1e-11
    # zero weights / zero times
    econ.average_dev(bi, np.zeros(B), w, v.data_ptr())
    op.synchronize()
    assert _relM(sw.M(), v.cpu().numpy(), w @ X) < 1e-13
    bi.close()


def test_economised_edge_times(setup):
    """t = 0 (series of length 1), a mix of zero / tiny / maximal sample times in one panel, and B = 1."""
    from oracle import cheby_ref
    sw, op, full, econ, a, L, A, x = setup
    M = sw.M()
    y = np.empty(sw.n)
    econ.apply(x, y, 0.0)
    assert _relM(M, y, x) < 1e-14
    tmax = DT * RHO / 1.01 * 0.999                                  # just inside the stated range
    t = np.array([0.0, 1e-9, -1e-3, 17.0, -tmax, tmax, 0.5 * tmax])
    X = _inputs(sw, x, len(t))
    Y = econ.apply_batch(X, t)
    ref = cheby_ref.apply_batch(a, L, A, X.T, t).T
    assert max(_relM(M, Y[k], ref[k]) for k in range(len(t))) < 1e-10
    econ.apply(X[4], y, float(t[4]))
    assert _relM(M, y, ref[4]) < 1e-10
    # switching the range off again restores the full series
    econ.set_spectral_radius(None)
    op.stats_reset()
    econ.apply(X[3], y, float(t[3]))
    assert op.stats()["op_applications"] % econ.ncheb == 0 and _relM(M, y, ref[3]) < 1e-10
    import krylov_exp_b200 as kx
    econ.set_spectral_radius(kx.EIGS[0])
