"""GPU parity on every kernel path a BASELINE.json config executes (round-2 additions).

  * the per-iteration Chebyshev mass solve (cheb_solve_init_kernel + bsr3_cheb_iter_kernel) that every ref >= 6
    run takes, forced at ref 3/4 with KEXP_COOP=0 and exercised for real at ref 6;
  * M1 at the EXACT bench settings (rtol 1e-11, per-term relaxation, economised series, panels sorted by |t_k|,
    shared-input forward batch) for all 64 + 64 results against the C port at CG 1e-13;
  * BASELINE configs[2]: krylov_exp at ref 5, kdim 40 and 50 (test_krylov.py:8,27-30,89-90) against the oracle
    restatement with sparse-LU solves, and the Krylov-vs-Chebyshev gap of both;
  * long series (ncheb > 320) with a shared input; mass solves that miss rtol are errors; asynchronous mode.
"""
import numpy as np
import pytest
from scipy.sparse.linalg import splu

from conftest import get_sw

pytestmark = pytest.mark.gpu

TOL = 1e-10          # north_star: relative (M-weighted) L2 / H(div) agreement with the oracle


def _relM(M, a, b):
    d = a - b
    return np.sqrt((d @ (M @ d)) / (b @ (M @ b)))


def _m1_inputs(sw, B, seed=1234):
    """M1 inputs: pulse + per-column noise, the same generator as bench.py."""
    from oracle.assemble_sw import test_pulse
    x = test_pulse(sw)
    X = np.empty((B, sw.n))
    amp = 1e-3 * np.abs(x).max()
    for k in range(B):
        rng = np.random.default_rng(seed + k)
        X[k] = x
        X[k, sw.n_u:] += amp * rng.standard_normal(sw.n_eta)
        X[k, :sw.n_u] += amp * np.sqrt(sw.g / sw.H) * sw.radius / 2 ** sw.ref_level * rng.standard_normal(sw.n_u)
    return x, X


# ------------------------------------------------------------------------------------------------
# (a) per-iteration mass-solve path and the other environment-switched variants
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env,path", [({"KEXP_COOP": "0", "KEXP_NARROW": "0"}, 1), ({"KEXP_COOP": "0", "KEXP_NO_GRAPH": "1", "KEXP_NARROW": "0"}, 1),
                                      ({"KEXP_NARROW": "0"}, 2), ({"KEXP_NARROW_MAXW": "8"}, 3), ({"KEXP_NARROW_MREG": "0"}, 3), ({"KEXP_NARROW_MAXW": "2"}, 3),
                                      ({"KEXP_FUSED_UPDATE": "1", "KEXP_NARROW": "0"}, 2), ({"KEXP_FUSED_UPDATE": "1", "KEXP_COOP": "0", "KEXP_NARROW": "0"}, 1),
                                      ({"KEXP_LANES": "2"}, None), ({"KEXP_PANEL_CAP": "16"}, None),
                                      ({"KEXP_RELAX": "0", "KEXP_PDL": "0"}, None)])
def test_switched_variants_ref3_against_oracle(sw3, monkeypatch, env, path):
    import krylov_exp_b200 as kx
    from oracle import coeffs, cheby_ref
    for k in ("KEXP_COOP", "KEXP_NO_GRAPH", "KEXP_LANES", "KEXP_PANEL_CAP", "KEXP_RELAX", "KEXP_PDL", "KEXP_NARROW", "KEXP_NARROW_MAXW", "KEXP_NARROW_MREG",
              "KEXP_FUSED_UPDATE"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sw = sw3
    dt, rho = 3600.0, 1.0
    L = kx.EIGS[0] * dt * rho
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=60, solver="chebyshev")
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L, spectral_radius=kx.EIGS[0])
    B = 40                                                # two panels: 32 + 8 columns
    t, w = kx.sample_times_weights(B, dt, rho)
    x, X = _m1_inputs(sw, B)
    Y = ce.apply_batch(X, -t)
    F = ce.apply_batch(x, t, shared=True)
    if path is not None:
        assert op.stats()["solve_path"] == path
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    a = coeffs.real_recurrence_coeffs(c, n)
    lu = splu(sw.M().tocsc())
    opm = lambda V: lu.solve(sw.K @ V)
    ref = cheby_ref.apply_batch(a, L, opm, X.T, -t).T
    reff = cheby_ref.apply_batch(a, L, opm, np.repeat(x[:, None], B, axis=1), t).T
    M = sw.M()
    assert max(_relM(M, Y[k], ref[k]) for k in range(B)) < TOL
    assert max(_relM(M, F[k], reff[k]) for k in range(B)) < TOL
    v = ce.apply_average(X, -t, w, fused=True)
    assert _relM(M, v, w @ ref) < TOL
    op.close()


def test_per_iteration_path_ref4_against_cpu_port(monkeypatch):
    import krylov_exp_b200 as kx
    from oracle import coeffs, norms
    from oracle.cpu_port import CpuCheby
    monkeypatch.setenv("KEXP_COOP", "0")
    sw = get_sw(4)
    dt, rho = 1800.0, 1.0
    L = kx.EIGS[1] * dt * rho
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-11, max_it=40, solver="chebyshev")
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L, spectral_radius=kx.EIGS[1])
    B = 64
    t, w = kx.sample_times_weights(B, dt, rho)
    x, X = _m1_inputs(sw, B)
    Y = ce.apply_batch(X, -t)
    assert op.stats()["solve_path"] == 1
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13)
    cols = list(range(0, B, 5))
    ref, _ = cpu.apply_batch(coeffs.real_recurrence_coeffs(c, n), n, L, X[cols], -t[cols])
    for i, k in enumerate(cols):
        assert max(norms.calculate_norm(sw, Y[k], ref[i])) < TOL, k
    op.close()


def test_ref6_against_cpu_port():
    """A true ref-6 case (BASELINE configs[3] size, N = 860 160): the cooperative solve does not fit there, so this is
    cheb_solve_init_kernel + bsr3_cheb_iter_kernel<.,0/1>.  3 columns of an 8-column batch (the per-GPU share of 64
    samples over 8 GPUs) against oracle/cpu_cheby.c at CG 1e-13, in the three calculate_norm.py:93-96 norms."""
    import krylov_exp_b200 as kx
    from oracle import coeffs, norms
    from oracle.cpu_port import CpuCheby
    sw = get_sw(6)
    dt, rho = 900.0, 1.0
    eig = kx.EIGS[2] * 2.0 * 1.03                          # bench.py's bound for levels the reference table lacks
    L = eig * dt * rho
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-11, max_it=40, solver="chebyshev")
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L, spectral_radius=eig)
    B = 8
    t_all, _ = kx.sample_times_weights(64, dt, rho)
    t = np.ascontiguousarray(t_all[3::8])                  # rank 3 of 8, cyclic
    x, X = _m1_inputs(sw, B)
    Y = ce.apply_batch(X, -t)
    assert op.stats()["solve_path"] == 1                   # one kernel per Chebyshev iteration
    F = ce.apply_batch(x, t, shared=True)
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    a = coeffs.real_recurrence_coeffs(c, n)
    cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13)
    cols = [0, 4, 7]
    ref, _ = cpu.apply_batch(a, n, L, X[cols], -t[cols])
    reff, _ = cpu.apply_batch(a, n, L, x[None, :], t[[7]])
    for i, k in enumerate(cols):
        assert max(norms.calculate_norm(sw, Y[k], ref[i])) < TOL, k
    assert max(norms.calculate_norm(sw, F[7], reff[0])) < TOL
    op.close()


# ------------------------------------------------------------------------------------------------
# (b) M1 at the exact bench settings, all 64 + 64 results
# ------------------------------------------------------------------------------------------------
def test_M1_exact_bench_settings_all_columns():
    """bench.py's step: Operator(rtol=1e-11, max_it=40, chebyshev), cheby_exp(spectral_radius=eig), device batches,
    backward panels sorted by |t_k|, shared-input forward batch; every one of the 128 delivered results against the
    C port of the reference recurrence with CG at 1e-13."""
    import krylov_exp_b200 as kx
    from oracle import coeffs, norms
    from oracle.cpu_port import CpuCheby
    sw = get_sw(5)
    DT, RHO, B = 900.0, 1.0, 64
    eig = kx.EIGS[2]
    L = eig * DT * RHO
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-11, max_it=40, solver="chebyshev")
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L, spectral_radius=eig)
    assert ce.ncheb == 29
    t, w = kx.sample_times_weights(B, DT, RHO)
    x, X = _m1_inputs(sw, B)
    b_fi, b_fo, b_bi, b_bo = op.batch(B), op.batch(B), op.batch(B), op.batch(B)
    b_fi.upload(x, shared=True)
    b_bi.set_order(np.argsort(np.abs(t), kind="stable"))
    b_bi.upload(X)
    op.set_async(True)                                     # as bench.py runs it
    for _ in range(2):
        ce.apply_dev(b_fi, b_fo, t)
        ce.apply_dev(b_bi, b_bo, -t)
    op.synchronize()
    F, Y = b_fo.download(), b_bo.download()
    st = op.stats()
    assert st["unconverged_solves"] == 0
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    a = coeffs.real_recurrence_coeffs(c, n)
    cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13)
    ref, _ = cpu.apply_batch(a, n, L, X, -t)
    reff, _ = cpu.apply_batch(a, n, L, x[None, :], t, shared=True)
    worst_b = max(max(norms.calculate_norm(sw, Y[k], ref[k])) for k in range(B))
    worst_f = max(max(norms.calculate_norm(sw, F[k], reff[k])) for k in range(B))
    print(f"M1 bench settings: worst backward {worst_b:.2e}, worst forward {worst_f:.2e} (eta L2 / u H(div) / u L2)")
    assert worst_b < TOL and worst_f < TOL
    # the weighted average of the bench step and its fused form
    import torch
    v = torch.zeros(sw.n, dtype=torch.float64, device="cuda:0")
    b_bo.average_into(w, v.data_ptr())
    op.synchronize()
    M = sw.M()
    assert _relM(M, v.cpu().numpy(), w @ ref) < TOL
    ce.average_dev(b_bi, -t, w, v.data_ptr())
    op.synchronize()
    assert _relM(M, v.cpu().numpy(), w @ ref) < TOL
    for b in (b_fi, b_fo, b_bi, b_bo):
        b.close()
    op.close()


# ------------------------------------------------------------------------------------------------
# (c) BASELINE configs[2]: krylov_exp at ref 5, kdim 40 and 50
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def krylov_ref5():
    """Oracle side, once: sparse LU of M + gamma K and of M at ref 5, Chebyshev apply(x, y, -t) on the C port."""
    from oracle import coeffs
    from oracle.assemble_sw import test_pulse
    from oracle.cpu_port import CpuCheby
    import krylov_exp_b200 as kx
    sw = get_sw(5)
    t = 7200.0
    gamma = t / 10.0                                       # test_krylov.py:27-30
    M = sw.M()
    lu_s = splu((M + gamma * sw.K).tocsc())
    lu_m = splu(M.tocsc())
    x = test_pulse(sw)
    L = kx.EIGS[2] * t
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13)
    yc, _ = cpu.apply_batch(coeffs.real_recurrence_coeffs(c, n), n, L, x[None, :], np.array([-t]))
    return sw, M, lu_s, lu_m, x, t, gamma, L, yc[0]


@pytest.mark.parametrize("kdim", [40, 50])
def test_config3_krylov_ref5(krylov_ref5, kdim):
    import krylov_exp_b200 as kx
    from oracle import krylov_ref
    sw, M, lu_s, lu_m, x, t, gamma, L, yc = krylov_ref5
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=60)
    sh = kx.ShiftedOperator(op, sw.K, sw.Mu, sw.Meta_blocks, gamma, rtol=1e-13)
    ke = kx.krylov_exp(sh, None, None, gamma, kdim, op, None, None)
    y = np.empty(sw.n)
    ke.apply(x, y, t, residuals=False)
    yo, Ho, uo, _ = krylov_ref.krylov_apply(lambda v: lu_s.solve(M @ v), lambda v: lu_m.solve(sw.K @ v),
                                            lambda a, b: a @ (M @ b), x, t, gamma, kdim, residuals=False)
    eH = np.abs(ke.H - Ho).max() / np.abs(Ho).max()
    ey = _relM(M, y, yo)
    # Chebyshev on the GPU for the same -t (test_krylov.py's comparison), and the gap both Krylov results show to it
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    ycg = np.empty(sw.n)
    ce.apply(x, ycg, -t)
    assert _relM(M, ycg, yc) < TOL
    gap_gpu, gap_oracle = _relM(M, y, yc), _relM(M, yo, yc)
    print(f"config 3 ref 5 kdim {kdim}: H {eH:.2e}, y {ey:.2e}; Krylov-vs-Chebyshev gap GPU {gap_gpu:.3e} oracle {gap_oracle:.3e}")
    # MGS in the M inner product amplifies the 1e-13 of the shifted solves; the Krylov space is far from
    # converged at this size (gap ~1e-1), which both implementations must show alike
    # measured on B200: H 4.0e-12 / 1.6e-11, y 2.5e-12 / 2.5e-11 for kdim 40 / 50
    assert eH < 1e-9
    assert ey < 1e-9
    assert abs(gap_gpu - gap_oracle) < 1e-6 * max(gap_oracle, 1e-3)
    op.close()


# ------------------------------------------------------------------------------------------------
# ADVICE items
# ------------------------------------------------------------------------------------------------
def test_long_series_with_shared_input(sw2):
    """ncheb > 320 with one shared input (the combination kernels tile over the terms): every column and the fused
    weighted average against the oracle."""
    import krylov_exp_b200 as kx
    from oracle import coeffs, cheby_ref
    sw = sw2
    L = 420.0
    eig = 0.0017
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80)
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    assert ce.ncheb > 330
    B = 9
    t, w = kx.sample_times_weights(B, L / eig, 1.0)
    from oracle.assemble_sw import test_pulse
    x = test_pulse(sw)
    F = ce.apply_batch(x, t, shared=True)
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    a = coeffs.real_recurrence_coeffs(c, n)
    lu = splu(sw.M().tocsc())
    ref = cheby_ref.apply_batch(a, L, lambda V: lu.solve(sw.K @ V), np.repeat(x[:, None], B, axis=1), t).T
    M = sw.M()
    assert max(_relM(M, F[k], ref[k]) for k in range(B)) < TOL
    for fused in (False, True):
        v = ce.apply_average(x, t, w, shared=True, fused=fused)
        assert _relM(M, v, w @ ref) < TOL, fused
    Xd = np.stack([x * (1 + 0.1 * k) for k in range(B)])
    refd = cheby_ref.apply_batch(a, L, lambda V: lu.solve(sw.K @ V), Xd.T, -t).T
    v = ce.apply_average(Xd, -t, w, fused=True)
    assert _relM(M, v, w @ refd) < TOL
    # the other entry points refuse a batch that only holds the shared vector
    bs, bo = op.batch(B), op.batch(B)
    bs.upload(x, shared=True)
    with pytest.raises(kx.KexpError):
        op._check(op.lib.kexp_operator_apply_dev(op.ctx, bs.handle, bo.handle), "kexp_operator_apply_dev")
    with pytest.raises(kx.KexpError):
        op._check(op.lib.kexp_spmm_K_dev(op.ctx, bs.handle, bo.handle), "kexp_spmm_K_dev")
    import torch
    v = torch.zeros(sw.n, dtype=torch.float64, device="cuda:0")
    with pytest.raises(kx.KexpError):
        bs.average_into(w, v.data_ptr())
    bs.close(); bo.close()
    op.close()


def test_unconverged_mass_solve_is_an_error(sw2):
    import krylov_exp_b200 as kx
    from oracle.assemble_sw import test_pulse
    sw = sw2
    x = test_pulse(sw)
    L = 0.0017 * 3600.0
    # PCG capped far too early
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=3, solver="cg")
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    with pytest.raises(kx.KexpError, match="did not reach rtol"):
        ce.apply_batch(np.stack([x, 2 * x]), np.array([1800.0, -900.0]))
    assert op.stats()["unconverged_solves"] > 0 and op.stats()["worst_rel_residual"] > 1e-12
    with pytest.raises(kx.KexpError, match="did not reach rtol"):
        op.apply(x[None, :])
    op.close()
    # Chebyshev iteration with a wrong spectral enclosure (lmin far too large): the measured residual catches it
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=12, solver="chebyshev", spectrum=(1.2, 2.05))
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    with pytest.raises(kx.KexpError, match="did not reach rtol"):
        ce.apply_batch(np.stack([x, 2 * x]), np.array([1800.0, -900.0]))
    op.close()
    # ... and with enough iterations allowed the same wrong enclosure is repaired by the retry (slower, not wrong)
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=400, solver="chebyshev", spectrum=(0.6, 2.05))
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    Y = ce.apply_batch(np.stack([x, 2 * x]), np.array([1800.0, -900.0]))
    good = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=60, solver="chebyshev")
    ce2 = kx.cheby_exp(good, None, None, 10000, 1e-8, L)
    Y2 = ce2.apply_batch(np.stack([x, 2 * x]), np.array([1800.0, -900.0]))
    assert _relM(sw.M(), Y[0], Y2[0]) < 1e-10
    assert op.stats()["unconverged_solves"] > 0
    op.close(); good.close()


def test_async_mode_defers_the_check_not_the_error(sw2):
    import krylov_exp_b200 as kx
    from oracle.assemble_sw import test_pulse
    sw = sw2
    x = test_pulse(sw)
    L = 0.0017 * 3600.0
    B = 6
    t, w = kx.sample_times_weights(B, 3600.0, 1.0)
    X = np.stack([x * (1 + 0.2 * k) for k in range(B)])
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=60)
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    bi, bo = op.batch(B), op.batch(B)
    bi.upload(X)
    ce.apply_dev(bi, bo, t)
    Y_sync = bo.download()
    op.set_async(True)
    for _ in range(5):
        ce.apply_dev(bi, bo, t)                           # queued back to back, no host synchronisation
    op.synchronize()
    assert np.array_equal(bo.download(), Y_sync)
    op.set_async(False)
    bi.close(); bo.close()
    op.close()


# ------------------------------------------------------------------------------------------------
# persistent narrow-panel apply kernel (kexp_narrow.cu) against the multi-launch path and the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [1, 3, 4, 7, 8])
def test_narrow_kernel_equals_multi_launch_path(sw3, monkeypatch, B):
    monkeypatch.setenv("KEXP_NARROW_MAXW", "8")           # default: up to 4 columns (8 is slower than the multi-launch path at ref 5)
    """Same rows, same order of operations: the persistent kernel (halo exchange through flagged inboxes instead of grid
    barriers) must reproduce the multi-launch path to round-off (forward recurrence, shared-input basis run, fused Clenshaw
    average), with the same iteration counts."""
    import krylov_exp_b200 as kx
    sw = sw3
    dt, rho = 3600.0, 1.0
    L = kx.EIGS[0] * dt * rho
    t, w = kx.sample_times_weights(B, dt, rho)
    x, X = _m1_inputs(sw, B)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("KEXP_NARROW", mode)
        op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=60, solver="chebyshev")
        ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L, spectral_radius=kx.EIGS[0])
        Y = ce.apply_batch(X, -t)
        path = op.stats()["solve_path"] if mode == "1" else None
        F = ce.apply_batch(x, t, shared=True)
        v = ce.apply_average(X, -t, w, fused=True)
        st = op.stats()
        out[mode] = (Y, F, v, path, st["cg_iterations"], st["cg_solves"], st["unconverged_solves"])
        op.close()
    assert out["1"][3] == 3
    assert out["1"][4:] == out["0"][4:]
    M = sw.M()
    for i in range(3):
        a, b = np.atleast_2d(out["1"][i]), np.atleast_2d(out["0"][i])
        assert max(_relM(M, a[k], b[k]) for k in range(a.shape[0])) < 1e-14, i


@pytest.mark.parametrize("condense", [True, False])
def test_narrow_kernel_against_oracle(sw3, condense):
    """Both operator plans (statically condensed; plain K, M_u, M_eta^-1 with the 3x3 depth blocks in the epilogue)."""
    import krylov_exp_b200 as kx
    from oracle import coeffs, cheby_ref
    sw = sw3
    dt, rho, B = 5400.0, 1.0, 6
    L = kx.EIGS[0] * dt * rho
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13, max_it=80, solver="chebyshev", condense=condense)
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    B = 4                                                  # default: panels of up to 4 columns take the persistent kernel
    t, w = kx.sample_times_weights(B, dt, rho)
    x, X = _m1_inputs(sw, B)
    Y = ce.apply_batch(X, -t)
    assert op.stats()["solve_path"] == 3
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    a = coeffs.real_recurrence_coeffs(c, n)
    lu = splu(sw.M().tocsc())
    ref = cheby_ref.apply_batch(a, L, lambda V: lu.solve(sw.K @ V), X.T, -t).T
    M = sw.M()
    assert max(_relM(M, Y[k], ref[k]) for k in range(B)) < TOL
    y1 = np.empty(sw.n)
    ce.apply(X[2], y1, -t[2])                              # the reference's own call form: B = 1
    assert _relM(M, y1, ref[2]) < TOL
    op.close()


@pytest.mark.parametrize("ncols", [4, 8])
def test_narrow_kernel_ref5_columns_against_cpu_port(monkeypatch, ncols):
    monkeypatch.setenv("KEXP_NARROW_MAXW", "8")
    """A strong-scaling share of M1 (4 or 8 of the 64 samples, and the shared-input forward pass) through the persistent
    kernel at full size."""
    import krylov_exp_b200 as kx
    from oracle import coeffs, norms
    from oracle.cpu_port import CpuCheby
    sw = get_sw(5)
    DT, RHO = 900.0, 1.0
    eig = kx.EIGS[2]
    L = eig * DT * RHO
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-11, max_it=40, solver="chebyshev")
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L, spectral_radius=eig)
    t_all, w_all = kx.sample_times_weights(64, DT, RHO)
    cols = np.arange(5, 64, 64 // ncols)
    t = np.ascontiguousarray(t_all[cols])
    x, Xall = _m1_inputs(sw, 64)
    X = Xall[cols]
    Y = ce.apply_batch(X, -t)
    assert op.stats()["solve_path"] == 3
    F = ce.apply_batch(x, t, shared=True)
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    a = coeffs.real_recurrence_coeffs(c, n)
    cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13)
    ref, _ = cpu.apply_batch(a, n, L, X, -t)
    reff, _ = cpu.apply_batch(a, n, L, x[None, :], t, shared=True)
    for k in range(len(cols)):
        assert max(norms.calculate_norm(sw, Y[k], ref[k])) < TOL, k
        assert max(norms.calculate_norm(sw, F[k], reff[k])) < TOL, k
    assert op.stats()["unconverged_solves"] == 0
    op.close()


@pytest.mark.parametrize("B", [1, 4])
def test_narrow_kernel_repeats_bit_identically(B):
    """The halo exchange of the persistent kernel has no barrier: a value read before its flag, a slot overwritten too early or a
    stale halo would show up as run-to-run differences.  60 applies of the same inputs at ref 5 (every CTA has ~7 neighbours and
    ~500 exchanges per apply) must agree bit for bit, forward recurrence and fused Clenshaw average."""
    import krylov_exp_b200 as kx
    sw = get_sw(5)
    DT, RHO = 900.0, 1.0
    eig = kx.EIGS[2]
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-11, max_it=40, solver="chebyshev")
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, eig * DT * RHO, spectral_radius=eig)
    t_all, w_all = kx.sample_times_weights(64, DT, RHO)
    cols = np.arange(3, 64, 64 // B)[:B]
    t = np.ascontiguousarray(t_all[cols]); w = np.ascontiguousarray(w_all[cols])
    x, Xall = _m1_inputs(sw, 64)
    X = np.ascontiguousarray(Xall[cols])
    Y0 = ce.apply_batch(X, -t).copy()
    assert op.stats()["solve_path"] == 3
    v0 = ce.apply_average(X, -t, w, fused=True).copy()
    for _ in range(60):
        assert np.array_equal(ce.apply_batch(X, -t), Y0)
    for _ in range(20):
        assert np.array_equal(ce.apply_average(X, -t, w, fused=True), v0)
    assert op.stats()["unconverged_solves"] == 0
    op.close()
