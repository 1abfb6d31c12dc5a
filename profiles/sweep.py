"""BASELINE.json configs[4]: synthetic sweep, ref_level x averaging samples -- batched SpMM (Y = K X) alone
and the full Chebyshev apply, one GPU (run under torchrun for more: every rank does the same local work).
    python profiles/sweep.py [max_ref] > gpurun_out/sweep.json
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krylov_exp_b200 as kx
from krylov_exp_b200.assembly import assemble_sw, test_pulse

PEAK = 6544.3
max_ref = int(sys.argv[1]) if len(sys.argv) > 1 else 6
rows = []
for ref in range(4, max_ref + 1):
    sw = assemble_sw(ref)
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=1e-11, max_it=40)
    eig = kx.EIGS[ref - 3] if ref <= 5 else kx.EIGS[2] * 2.06 ** (ref - 5)
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, eig * 900.0)
    ce_e = kx.cheby_exp(op, None, None, 10000, 1e-8, eig * 900.0, spectral_radius=eig)     # economised per-column series
    x = test_pulse(sw)
    for B in (8, 16, 32, 64, 128, 256):
        if sw.n * B * 8 * 12 > 60e9:
            continue
        X = np.repeat(x[None, :], B, axis=0) * (1.0 + 0.01 * np.arange(B))[:, None]
        t, w = kx.sample_times_weights(B, 900.0, 1.0)
        bi, bo = op.batch(B), op.batch(B)
        bi.upload(X)
        lib = op.lib
        for _ in range(2):
            op._check(lib.kexp_spmm_K_dev(op.ctx, bi.handle, bo.handle), "spmm")
        op.synchronize()
        op.timer_start()
        reps = 10
        for _ in range(reps):
            op._check(lib.kexp_spmm_K_dev(op.ctx, bi.handle, bo.handle), "spmm")
        ms_spmm = op.timer_stop() / reps
        bytes_spmm = 16.0 * sw.n * B + 80.0 * (op.nnz["Kt"] / 9.0) * bi.num_panels
        for _ in range(2):
            ce.apply_dev(bi, bo, -t)
        op.synchronize()
        op.timer_start()
        for _ in range(3):
            ce.apply_dev(bi, bo, -t)
        ms_apply = op.timer_stop() / 3
        bi.set_order(np.argsort(np.abs(t), kind="stable"))
        bi.upload(X)
        for _ in range(2):
            ce_e.apply_dev(bi, bo, -t)
        op.synchronize()
        op.timer_start()
        for _ in range(3):
            ce_e.apply_dev(bi, bo, -t)
        ms_econ = op.timer_stop() / 3
        rows.append({"ref": ref, "N": sw.n, "B": B, "spmm_ms": ms_spmm, "spmm_gbs": bytes_spmm / ms_spmm / 1e6,
                     "spmm_frac": bytes_spmm / ms_spmm / 1e6 / PEAK, "spmm_gflops": 2.0 * op.nnz["Kt"] * B / ms_spmm / 1e6,
                     "apply_ms": ms_apply, "applications_per_s_full_series": B / ms_apply * 1e3,
                     "applications_per_s": B / ms_econ * 1e3, "ncheb": int(ce.ncheb)})
        print(json.dumps(rows[-1]), flush=True)
        bi.close(); bo.close()
    op.close()
