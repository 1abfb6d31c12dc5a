"""Parity error of the default product path (ref 5, 64 samples) against the CPU port at CG rtol 1e-14,
in the calculate_norm.py norms, for a few CG/Chebyshev tolerances, with and without term relaxation."""
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
cols = [0, 31, 63]
c, n = coeffs.cheby_coeffs(10000, 1e-8, L); a = coeffs.real_recurrence_coeffs(c, n)
cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-14)
ref, _ = cpu.apply_batch(a, n, L, X[cols], -t[cols])
reff, _ = cpu.apply_batch(a, n, L, np.repeat(x[None, :], 3, axis=0), t[cols])
spec = None
for relax in ("1", "0"):
    os.environ["KEXP_RELAX"] = relax
    for rtol in (1e-12, 1e-11, 1e-10):
        op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=rtol, max_it=40, spectrum=spec)
        spec = op.spectrum
        ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
        bi, bo = op.batch(64), op.batch(64); bi.upload(X)
        for _ in range(2): ce.apply_dev(bi, bo, -t)
        op.synchronize(); op.stats_reset(); op.timer_start()
        for _ in range(3): ce.apply_dev(bi, bo, -t)
        ms = op.timer_stop() / 3
        st = op.stats()
        Y = bo.download()
        err = max(max(norms.calculate_norm(sw, Y[k], ref[i])) for i, k in enumerate(cols))
        F = ce.apply_batch(x, t, shared=True)
        errf = max(max(norms.calculate_norm(sw, F[k], reff[i])) for i, k in enumerate(cols))
        print(f"relax={relax} rtol={rtol:g}: bwd {64 / (ms * 1e-3):.1f} apps/s, its/solve {st['cg_iterations'] / st['cg_solves']:.1f}, "
              f"parity err bwd {err:.2e}, fwd(shared basis) {errf:.2e}", flush=True)
        op.close()
