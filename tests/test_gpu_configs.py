"""GPU: the other BASELINE.json configs as parity cases."""
import numpy as np
import pytest

from conftest import get_sw

pytestmark = pytest.mark.gpu


def _relM(M, a, b):
    d = a - b
    return np.sqrt((d @ (M @ d)) / (b @ (M @ b)))


def test_M3_test_cheby_config():
    """test_cheby.py as written: ref 4, c-scaled operator (:38-42), t = 6 h, L = 0.007274*21600 (:63-66),
    tol 1e-6 (:70-71) -> 185 terms; input u = 0, h = exp((x+y+z)/R) xyz/R^3 (:73-75)."""
    import krylov_exp_b200 as kx
    from oracle import coeffs
    from oracle.assemble_sw import test_pulse
    from oracle.cpu_port import CpuCheby
    sw = get_sw(4, c_scaled=True)
    t = 6 * 3600.0
    L = 0.007274 * t
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=60)
    ce = kx.cheby_exp(op, None, None, 10000, 1.0e-6, L)
    assert ce.ncheb == 185
    x = test_pulse(sw)
    y = kx.MixedVector(sw.n_u, sw.n_eta)
    ce.apply(kx.MixedVector(sw.n_u, sw.n_eta, x), y, t)
    c, n = coeffs.cheby_coeffs(10000, 1.0e-6, L)
    cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13)
    ref, _ = cpu.apply_batch(coeffs.real_recurrence_coeffs(c, n), n, L, x[None, :], np.array([t]))
    assert _relM(sw.M(), y.data, ref[0]) < 1e-10
    lit, _ = cpu.apply_batch(c, n, L, x[None, :], np.array([t]), literal=True)     # the reference's 2n-1-solve form
    assert _relM(sw.M(), y.data, lit[0]) < 1e-10
    # energy (c-scaled: <u,u> + <h,h>) is conserved to the truncation tolerance
    nu = sw.n_u
    E = lambda z: z[:nu] @ (sw.Mu @ z[:nu]) + z[nu:] @ (sw.Meta() @ z[nu:])
    assert abs(E(y.data) / E(x) - 1) < 1e-4
    op.close()


def test_M2_long_window_roundtrip():
    """rho = 7.45 (natural Mbar = 64): L = 100.27, 130 terms, 64 sample times; exp(-tL) exp(tL) = I."""
    import krylov_exp_b200 as kx
    from oracle.assemble_sw import test_pulse
    sw = get_sw(4)
    dt, rho = 900.0 * 2.056, 7.45            # same L as ref 5 / dt 900 (eig ratio 0.014955/0.007274)
    L = kx.EIGS[1] * dt * rho
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=60)
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    assert 125 <= ce.ncheb <= 135
    Mbar = kx.natural_Mbar(kx.EIGS[1], dt, rho)
    assert Mbar == 64
    t, w = kx.sample_times_weights(Mbar, dt, rho)
    x = test_pulse(sw)
    F = ce.apply_batch(x, t, shared=True)
    Bk = ce.apply_batch(F, -t)
    M = sw.M()
    assert max(_relM(M, Bk[k], x) for k in (0, 13, 31, 63)) < 5e-7
    v = ce.apply_average(F, -t, w)
    assert _relM(M, v, x) < 5e-7              # sum_k w_k exp(-t_k L) exp(t_k L) x = x
    op.close()


def test_M5_degree1_mesh_H8000():
    """averaged_sw_williamson6.py:42-47 flavour: degree-1 (flat) mesh, H = 8000, filtered coefficients."""
    import krylov_exp_b200 as kx
    from oracle import coeffs, cheby_ref
    from oracle.assemble_sw import test_pulse, largest_imag_eig
    from scipy.sparse.linalg import splu
    sw = get_sw(3, degree=1, H=8000.0)
    eig = 1.02 * largest_imag_eig(sw)         # the reference's eig table stops at ref 5: own eigen-solve
    dt, rho = 3600.0, 1.0
    L = eig * dt * rho
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80)
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L, filter=True, filter_val=0.75)
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L, filter=True, filter_val=0.75)
    assert n == ce.ncheb
    t, w = kx.sample_times_weights(7, dt, rho)
    x = test_pulse(sw)
    X = np.stack([x * (1 + 0.1 * k) for k in range(7)])
    Y = ce.apply_batch(X, t)
    lu = splu(sw.M().tocsc())
    ref = cheby_ref.apply_batch(coeffs.real_recurrence_coeffs(c, n), L, lambda V: lu.solve(sw.K @ V), X.T, t).T
    assert max(_relM(sw.M(), Y[k], ref[k]) for k in range(7)) < 1e-10
    op.close()


def test_spectral_bound_from_the_device_operator():
    """N4: the `eigs` table (averaged_sw_explicit.py:77-82) recomputed with ARPACK driving the device operator:
    equals the host eigen-solve of the same pencil and sits within the assembler's 1e-3 of eigs.dat:2."""
    import krylov_exp_b200 as kx
    from oracle.assemble_sw import largest_imag_eig
    sw = get_sw(3)
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80)
    lam = op.largest_imag_eig()
    op.close()
    assert abs(lam / largest_imag_eig(sw) - 1) < 1e-8
    assert abs(lam / 0.0034641819861161867 - 1) < 1e-3
    assert lam <= kx.EIGS[0] * 1.0005
