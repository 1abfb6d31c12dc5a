"""Parity error of the DEFAULT product path (ref 5, 64 samples, economised per-column series, relaxed term
tolerances, rtol 1e-11) against the CPU port running the reference's full 29-term series with CG at 1e-14,
in the calculate_norm.py norms; and the same with the economisation switched off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import krylov_exp_b200 as kx
from krylov_exp_b200.assembly import assemble_sw, test_pulse
from oracle import coeffs, norms
from oracle.cpu_port import CpuCheby
sw = assemble_sw(5)
L = kx.EIGS[2] * 900.0
x = test_pulse(sw)
X = np.empty((64, sw.n)); amp = 1e-3 * np.abs(x).max()
for k in range(64):
    rng = np.random.default_rng(1234 + k); X[k] = x
    X[k, sw.n_u:] += amp * rng.standard_normal(sw.n_eta)
    X[k, :sw.n_u] += amp * np.sqrt(sw.g / sw.H) * sw.radius / 32 * rng.standard_normal(sw.n_u)
t, w = kx.sample_times_weights(64, 900.0, 1.0)
cols = [0, 13, 31, 32, 47, 63]
c, n = coeffs.cheby_coeffs(10000, 1e-8, L); a = coeffs.real_recurrence_coeffs(c, n)
cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-14)
ref, _ = cpu.apply_batch(a, n, L, X[cols], -t[cols])
reff, _ = cpu.apply_batch(a, n, L, np.repeat(x[None, :], len(cols), axis=0), t[cols])
spec = None
for eps in ("1e-13", "1e-12", "1e-14", None):
    if eps: os.environ["KEXP_ECON_EPS"] = eps
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-11, max_it=40, spectrum=spec)
    spec = op.spectrum
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L, spectral_radius=kx.EIGS[2] if eps else None)
    bi, bo = op.batch(64), op.batch(64)
    bi.set_order(np.argsort(np.abs(t), kind="stable")); bi.upload(X)
    for _ in range(2): ce.apply_dev(bi, bo, -t)
    op.synchronize(); op.stats_reset(); op.timer_start()
    for _ in range(3): ce.apply_dev(bi, bo, -t)
    ms = op.timer_stop() / 3
    st = op.stats()
    Y = bo.download()
    err = max(max(norms.calculate_norm(sw, Y[k], ref[i])) for i, k in enumerate(cols))
    F = ce.apply_batch(x, t, shared=True)
    errf = max(max(norms.calculate_norm(sw, F[k], reff[i])) for i, k in enumerate(cols))
    print(f"econ_eps={eps}: bwd {64 / (ms * 1e-3):.1f} apps/s, operator applications per bwd pass {st['op_applications'] / 3:.0f}, "
          f"its/solve {st['cg_iterations'] / st['cg_solves']:.1f}, parity err bwd {err:.2e}, fwd(shared basis) {errf:.2e}", flush=True)
    op.close()
