"""Small driver for ncu: one batched cheby apply (ref 5, 64 samples) through the C ABI.
Usage: python profiles/run_profile.py [ref_level] [B] [nterms_cap]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krylov_exp_b200 as kx
from krylov_exp_b200.assembly import assemble_sw, test_pulse

ref = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sw = assemble_sw(ref)
op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-12, max_it=40)
L = kx.EIGS[min(ref, 5) - 3] * 900.0
ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
t, w = kx.sample_times_weights(B, 900.0, 1.0)
x = test_pulse(sw)
X = np.repeat(x[None, :], B, axis=0) * (1.0 + 0.01 * np.arange(B))[:, None]
bi, bo = op.batch(B), op.batch(B)
bi.upload(X)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for _ in range(reps):
    ce.apply_dev(bi, bo, -t)
op.synchronize()
st = op.stats()
print("ms", op.last_device_ms(), "its/solve", st["cg_iterations"] / st["cg_solves"], "launches", st["kernel_launches"])
