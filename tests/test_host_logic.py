"""CPU: host-side logic of the product package (no GPU compute) -- coefficient tables,
sample times, static condensation algebra, the C-ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import splu

import krylov_exp_b200 as kx
from conftest import ROOT


def test_product_coefficients_match_reference(golden):
    g = golden
    for name in ("qjrms", "rho745", "test_cheby", "filter", "filter_freq", "small"):
        n0, tol, L, nc, f, fv, ff = g[f"coef_{name}_meta"]
        c, n = kx.chebyshev_coefficients(int(n0), tol, L, filter=bool(f), filter_val=fv, filter_freq=bool(ff))
        gc = g[f"coef_{name}_c"]
        assert n == int(nc)
        assert np.abs(c[:len(gc)] - gc).max() <= 5e-15 * np.abs(gc).max()
    with pytest.raises(AssertionError):
        kx.chebyshev_coefficients(4, 1e-30, 50.0)        # cannot compress: reference asserts (cheby_exp.py:63)


def test_real_table_matches_oracle(golden):
    from oracle import coeffs
    c, n = kx.chebyshev_coefficients(10000, 1e-8, 13.4595)
    a = kx.real_table(c, n)
    co, no = coeffs.cheby_coeffs(10000, 1e-8, 13.4595)
    assert n == no == 29
    assert np.abs(a - coeffs.real_recurrence_coeffs(co, no)).max() < 1e-15


def test_sample_times_and_weights():
    from oracle import samples as osamples
    for Mbar, dt, rho in ((9, 900.0, 1.0), (64, 900.0, 7.45), (1, 450.0, 1.0)):
        t, w = kx.sample_times_weights(Mbar, dt, rho)
        to, wo = osamples.sample_times_weights(Mbar, dt, rho)
        assert np.array_equal(t, to) and np.array_equal(w, wo)
        assert abs(w.sum() - 1) < 1e-15 and np.all(np.abs(t) <= rho * dt / 2)
    assert kx.natural_Mbar(0.014955, 900.0, 1.0) == 9                 # qjrms setup
    assert kx.natural_Mbar(0.014955, 900.0, 7.45) == 64
    assert osamples.natural_Mbar(0.014955, 900.0, 7.45) == 64


def test_block_partition():
    assert list(kx.block_partition(64, 8)) == list(range(0, 65, 8))
    off = kx.block_partition(9, 4)
    assert off[0] == 0 and off[-1] == 9 and np.all(np.diff(off) >= 2) and np.all(np.diff(off) <= 3)
    e = kx.SampleEnsemble(9, rank=3, world=4)
    t, w = e.local_times_weights(900.0, 1.0)
    tt, ww = kx.sample_times_weights(9, 900.0, 1.0)
    assert np.array_equal(t, tt[e.local_slice]) and np.array_equal(w, ww[e.local_slice])
    # cyclic partition: every sample exactly once, |t| mix balanced across ranks
    seen = np.zeros(9, dtype=int)
    for r in range(4):
        c = kx.SampleEnsemble(9, rank=r, world=4, partition="cyclic")
        seen[c.local_slice] += 1
        assert c.n_local == len(tt[c.local_slice])
        assert np.array_equal(c.local_times_weights(900.0, 1.0)[0], tt[r::4])
    assert np.all(seen == 1)


def test_static_condensation_is_exact(sw2):
    """The condensed plan (Kt, As, Rt, perm) reproduces M^-1 K to round-off (pure SciPy)."""
    sw = sw2
    interior = kx.detect_interior_dofs(sw.Mu)
    assert interior is not None and interior.shape == (sw.n_cells, 3)
    assert np.array_equal(np.sort(interior.ravel()), np.sort(sw.cell_u_dofs[:, 9:12].ravel()))
    plan = kx.CondensedPlan(sw.K, sw.Mu, sw.Meta_inv_blocks, interior)
    assert plan.n_s == 90 * 4 ** 2 and plan.n_b == 60 * 4 ** 2 and plan.n_s % 3 == 0
    rng = np.random.default_rng(3)
    x = rng.standard_normal(sw.n)
    ref = splu(sw.M().tocsc()).solve(sw.K @ x)
    xi = np.empty(sw.n)
    xi[plan.perm] = x                                           # user -> internal
    T = plan.Kt @ xi
    ws = splu(plan.As.tocsc()).solve(T[:plan.n_s])
    wi = np.concatenate([ws, T[plan.n_s:plan.n_s + plan.n_b] + plan.Rt @ ws, T[plan.n_s + plan.n_b:]])
    w = wi[plan.perm]                                           # internal -> user
    assert np.linalg.norm(w - ref) <= 1e-12 * np.linalg.norm(ref)
    # the Schur complement is SPD and the 3x3 entity structure survives
    assert abs(plan.As - plan.As.T).max() < 1e-12
    assert plan.As.tobsr((3, 3)).indices.size * 9 == plan.As.nnz


def test_shifted_operator_algebra(sw2):
    """Eliminating eta from (M + gamma K) w = M v (test_krylov.py:34-62) is exact."""
    sw = sw2
    gamma = 720.0
    nu = sw.n_u
    K, Mu = sw.K.tocsr(), sw.Mu
    Meta = sw.Meta()
    Meta_inv = sp.bsr_matrix((sw.Meta_inv_blocks, np.arange(sw.n_cells), np.arange(sw.n_cells + 1))).tocsr()
    Kuu, Kue, Keu = K[:nu][:, :nu], K[:nu][:, nu:], K[nu:][:, :nu]
    Ash = Mu + gamma * Kuu - gamma ** 2 * (Kue @ Meta_inv @ Keu)
    rng = np.random.default_rng(5)
    v = rng.standard_normal(sw.n)
    M = sw.M()
    ref = splu((M + gamma * K).tocsc()).solve(M @ v)
    wu = splu(Ash.tocsc()).solve(Mu @ v[:nu] - gamma * (Kue @ v[nu:]))
    weta = v[nu:] - gamma * (Meta_inv @ (Keu @ wu))
    assert np.linalg.norm(np.concatenate([wu, weta]) - ref) <= 1e-11 * np.linalg.norm(ref)


def test_cpu_port_matches_scipy_oracle(sw2):
    from oracle import coeffs, cheby_ref
    from oracle.cpu_port import CpuCheby
    from oracle.assemble_sw import test_pulse
    sw = sw2
    cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13)
    L = 0.0017 * 3600
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    a = coeffs.real_recurrence_coeffs(c, n)
    x = test_pulse(sw)
    X = np.stack([x, 2 * x, x[::-1].copy()])
    t = np.array([1800.0, -900.0, 77.0])
    lu = splu(sw.M().tocsc())
    ref = cheby_ref.apply_batch(a, L, lambda V: lu.solve(sw.K @ V), X.T, t).T
    Y, its = cpu.apply_batch(a, n, L, X, t)
    Yl, _ = cpu.apply_batch(c, n, L, X, t, literal=True)
    assert np.abs(Y - ref).max() <= 1e-11 * np.abs(ref).max() and its > 0
    assert np.abs(Yl - ref).max() <= 1e-11 * np.abs(ref).max()


def test_norms_restatement(sw2):
    from oracle import norms
    sw = sw2
    rng = np.random.default_rng(0)
    b = rng.standard_normal(sw.n)
    a = b * (1 + 1e-6)
    e, h, l2 = norms.calculate_norm(sw, a, b)
    assert abs(e - 1e-6) < 1e-12 and abs(h - 1e-6) < 1e-12 and abs(l2 - 1e-6) < 1e-12
    assert abs(norms.rel_mixed(sw, a, b) - 1e-6) < 1e-12


# --------------------------------------------------------------------------- C ABI surface
def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "kexp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kexp_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = kx.load_library()
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"libkexp.so does not export {s}"
    from krylov_exp_b200._lib import PROTOTYPES
    assert set(PROTOTYPES) == set(syms), set(PROTOTYPES) ^ set(syms)
    assert b"sm_100a" in lib.kexp_version()


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (no CPU path, no oracle import)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = kx.load_library()
    ctx = C.c_void_p()
    assert lib.kexp_create(C.byref(ctx), 0) != 0
    with pytest.raises(kx.KexpError):
        kx.Operator(sp.identity(6, format="csr"), sp.identity(3, format="csr"), np.eye(3)[None])
    # the product package never imports the oracle
    import subprocess, sys
    code = "import sys; import krylov_exp_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for fn in os.listdir(os.path.join(ROOT, "krylov_exp_b200")):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "krylov_exp_b200", fn)).read().replace("the oracle re-exports", "").replace("CPU oracle", "") or fn == "assembly.py"


def test_series_reexpansion_against_numpy_chebyshev():
    """The coefficient algebra of the economised series / shared-input basis / fused average (host routine of
    libkexp.so, no GPU): on the spectrum z = i*mu the real recurrence is S_n(i mu) = i^n T_n(mu), so
    sum_n a_n S_n(s z) = sum_m b_m S_m(z) means f(mu) = sum_n a_n i^n T_n(s mu) has Chebyshev coefficients b_m i^m."""
    from numpy.polynomial import chebyshev as Ch
    from krylov_exp_b200.cheby import reexpand_series
    from oracle import coeffs
    L = kx.EIGS[2] * 900.0
    c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
    a = coeffs.real_recurrence_coeffs(c, n)
    fac = 1j ** np.arange(n + 1)
    mu = np.cos(np.pi * (np.arange(4 * n) + 0.5) / (4 * n))
    for s in (1.0, 0.505, -0.505, 0.25, -0.01, 0.0):
        b, keep = reexpand_series(a, s, 1e-13)
        f = Ch.chebval(s * mu, a * fac)                       # the reference polynomial on the scaled spectrum
        g = Ch.chebval(mu, b * fac)                           # its re-expansion
        assert np.abs(f - g).max() < 5e-15 * max(1.0, np.abs(f).max()), s
        g_cut = Ch.chebval(mu, (b * fac)[:keep])
        tail = np.abs(b[keep:]).sum()
        assert tail <= 1e-13 * np.abs(b).sum() and np.abs(f - g_cut).max() <= tail + 5e-15, s
        if abs(s) < 1:
            assert keep < n + 1                               # shorter than the reference series
    assert reexpand_series(a, 1.0, 0.0)[1] == n + 1 and np.allclose(reexpand_series(a, 1.0)[0], a, rtol=0, atol=1e-16)
    assert reexpand_series(a, 0.0)[1] == 1 and abs(reexpand_series(a, 0.0)[0][0] - 1.0) < 1e-12   # series is 1 at A = 0
    with pytest.raises(ValueError):
        reexpand_series(a, 1.5)


def test_checkpoint_roundtrip_and_norm_table(tmp_path, sw2):
    """averaged_sw_explicit.py:319-326 / calculate_norm.py:66-96: store two runs, read them back, norm table = oracle norms."""
    from krylov_exp_b200 import checkpoint
    from oracle import norms
    sw = sw2
    rng = np.random.default_rng(5)
    runs = {}
    for name in ("avg", "std"):
        for h in (1, 2):
            U = rng.standard_normal(sw.n)
            runs[(name, h)] = U
            checkpoint.store(str(tmp_path / name), h, sw.n_u, U, time=3600.0 * h, tdump=0.5, tcheck=0.25)
    U, attrs = checkpoint.load(str(tmp_path / "avg"), 2)
    assert np.array_equal(U, runs[("avg", 2)]) and attrs == {"time": 7200.0, "tdump": 0.5, "tcheck": 0.25}
    table = checkpoint.calculate_norm(sw, str(tmp_path / "avg"), str(tmp_path / "std"), (1, 2))
    for i, h in enumerate((1, 2)):
        assert np.allclose(table[i], norms.calculate_norm(sw, runs[("avg", h)], runs[("std", h)]), rtol=1e-13)


def test_balanced_partition_covers_every_sample_and_beats_cyclic():
    """Unequal |t|-sorted shares for the strong-scaling split of an averaging window (ensemble.balanced_partition): every sample
    exactly once, no empty rank, and under its own cost model never worse than the cyclic split; with the M1 series lengths the
    rank that holds the longest series gets the fewest samples."""
    from krylov_exp_b200.ensemble import balanced_partition, partition_cost
    # series lengths of the 64 M1 sample times (economised, ref 5, dt = 900 s): cheby_exp.series_terms
    half = [7, 9, 10, 11, 12, 13, 13, 14, 15, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 21, 22, 22, 23, 23, 23, 24, 24, 25, 25]
    terms = np.array(half[::-1] + half)
    for world in (1, 2, 3, 4, 8, 16):
        groups = balanced_partition(terms, world)
        assert len(groups) == world and all(len(g) > 0 for g in groups)
        assert sorted(np.concatenate(groups).tolist()) == list(range(64))
        cyc = [np.arange(r, 64, world) for r in range(world)]
        assert partition_cost(terms, groups) <= partition_cost(terms, cyc) * (1 + 1e-12)
    g8 = balanced_partition(terms, 8)
    longest = [g for g in g8 if terms[g].max() == 25][0]
    assert len(longest) == min(len(g) for g in g8)
    assert partition_cost(terms, g8) < 0.95 * partition_cost(terms, [np.arange(r, 64, 8) for r in range(8)])
    with pytest.raises(ValueError):
        balanced_partition(terms[:3], 4)
