"""BASELINE.json configs[2] (M4): krylov_exp Arnoldi exp(-tL)v at ref 5, kdim 40 (test_krylov.py:89-90) and 50,
t = 2 h, gamma = t/10 (test_krylov.py:27-30), against cheby_exp on the same input (apply(x, y, -t))."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import krylov_exp_b200 as kx
from krylov_exp_b200.assembly import assemble_sw, test_pulse
sw = assemble_sw(5)
t = 7200.0
gamma = t / 10.0
op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=60)
sh = kx.ShiftedOperator(op, sw.K, sw.Mu, sw.Meta_blocks, gamma, rtol=1e-12)
x = test_pulse(sw)
M = sw.M()
rel = lambda a, b: float(np.sqrt(((a - b) @ (M @ (a - b))) / (b @ (M @ b))))
ce = kx.cheby_exp(op, None, None, 10000, 1e-8, kx.EIGS[2] * t)
yc = np.empty(sw.n)
ce.apply(x, yc, -t)                                                     # warm-up
t0 = time.perf_counter(); ce.apply(x, yc, -t); tc = time.perf_counter() - t0
print(f"cheby_exp: ncheb {ce.ncheb}, {tc * 1e3:.1f} ms")
nu = sw.n_u
E = lambda z: sw.H * (z[:nu] @ (sw.Mu @ z[:nu])) + sw.g * (z[nu:] @ (sw.Meta() @ z[nu:]))
for kdim in (40, 50):
    ke = kx.krylov_exp(sh, None, None, gamma, kdim, op, None, None)
    y = np.empty(sw.n)
    ke.apply(x, y, t, residuals=False)                                  # warm-up (allocations, first launches)
    op.stats_reset()
    t0 = time.perf_counter(); ke.apply(x, y, t, residuals=False); tk = time.perf_counter() - t0
    st = op.stats()
    print(f"krylov_exp kdim {kdim}: {tk * 1e3:.0f} ms, BiCGStab iterations {st['bicg_iterations']} ({st['bicg_iterations'] / kdim:.0f}/solve), "
          f"launches {st['kernel_launches']}, vs cheby rel M-norm {rel(y, yc):.2e}, energy ratio {E(y) / E(x):.8f}")
ke = kx.krylov_exp(sh, None, None, gamma, 40, op, None, None)
y = np.empty(sw.n)
ke.apply(x, y, t, residuals=True)
print("reference-style residual estimates (krylov_exp.py:95-101), every 5th:", np.array2string(ke.residuals[::5], precision=2))
