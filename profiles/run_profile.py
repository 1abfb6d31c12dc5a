"""Small driver for ncu: the backward batch of the bench step (bench.py settings: Chebyshev mass solve at rtol 1e-11 with
per-term relaxation, economised series, panels sorted by |t_k|) through the C ABI.
Usage: python profiles/run_profile.py [ref_level] [B] [reps] [--full-series]
Set KEXP_NO_GRAPH=1 so that every kernel is a separate launch for ncu."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krylov_exp_b200 as kx
from krylov_exp_b200.assembly import assemble_sw, test_pulse

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
ref = int(argv[0]) if len(argv) > 0 else 5
B = int(argv[1]) if len(argv) > 1 else 64
reps = int(argv[2]) if len(argv) > 2 else 1
sw = assemble_sw(ref)
op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-11, max_it=40)
eig = kx.EIGS[ref - 3] if ref <= 5 else kx.EIGS[2] * 2.0 ** (ref - 5) * 1.03
L = eig * 900.0
ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L, spectral_radius=None if "--full-series" in sys.argv else eig)
t, w = kx.sample_times_weights(max(B, 2), 900.0, 1.0)
t = t[:B]
x = test_pulse(sw)
X = np.repeat(x[None, :], B, axis=0) * (1.0 + 0.01 * np.arange(B))[:, None]
bi, bo = op.batch(B), op.batch(B)
bi.set_order(np.argsort(np.abs(t), kind="stable"))
bi.upload(X)
for _ in range(reps):
    ce.apply_dev(bi, bo, -t)
op.synchronize()
st = op.stats()
print("ms", op.last_device_ms(), "its/solve", st["cg_iterations"] / st["cg_solves"], "launches", st["kernel_launches"],
      "op applications", st["op_applications"], "path", st["solve_path"])
