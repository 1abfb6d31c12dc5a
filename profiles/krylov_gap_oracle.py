"""CPU only (no GPU, no libkexp): the Krylov-vs-Chebyshev gap of BASELINE configs[2] at ref 5, oracle against oracle.

    python profiles/krylov_gap_oracle.py [ref_level] > profiles/r02_krylov_gap_oracle.txt

krylov_exp.apply restated in oracle/krylov_ref.py (krylov_exp.py:50-109) with EXACT sparse-LU shifted solves, against
cheby_exp.apply(x, y, -t) on the C port (oracle/cpu_cheby.c, CG 1e-13), t = 7200 s, gamma = t/10 (test_krylov.py:27-30),
kdim 40 (test_krylov.py:89-90) and 50 (BASELINE.json).  DESIGN.md (c) quotes these numbers; tests/test_gpu_parity_r2.py
checks that the GPU Arnoldi shows the same gap.
"""
import os
import sys
import time

import numpy as np
from scipy.sparse.linalg import splu

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import coeffs, krylov_ref                      # noqa: E402
from oracle.assemble_sw import assemble_sw, test_pulse     # noqa: E402
from oracle.cpu_port import CpuCheby                       # noqa: E402
from krylov_exp_b200.samples import EIGS                   # noqa: E402

ref = int(sys.argv[1]) if len(sys.argv) > 1 else 5
sw = assemble_sw(ref)
t = 7200.0
gamma = t / 10.0
M = sw.M()
t0 = time.time()
lu_s = splu((M + gamma * sw.K).tocsc())
lu_m = splu(M.tocsc())
print(f"ref {ref}: n = {sw.n}, sparse LU of M + gamma K and M in {time.time() - t0:.1f} s")
x = test_pulse(sw)
L = EIGS[ref - 3] * t
c, n = coeffs.cheby_coeffs(10000, 1e-8, L)
cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-13)
yc, _ = cpu.apply_batch(coeffs.real_recurrence_coeffs(c, n), n, L, x[None, :], np.array([-t]))
yc = yc[0]
rel = lambda a, b: float(np.sqrt(((a - b) @ (M @ (a - b))) / (b @ (M @ b))))
print(f"Chebyshev: L = {L:.3f}, {n} terms")
for kdim in (20, 40, 50):
    yo, H, u, _ = krylov_ref.krylov_apply(lambda v: lu_s.solve(M @ v), lambda v: lu_m.solve(sw.K @ v),
                                          lambda a, b: a @ (M @ b), x, t, gamma, kdim, residuals=False)
    Hk = H[:kdim, :kdim]
    print(f"kdim {kdim}: ||y_krylov - y_cheby||_M / ||y_cheby||_M = {rel(yo, yc):.4e}   cond(H) = {np.linalg.cond(Hk):.3e}")
