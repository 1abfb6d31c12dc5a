"""BASELINE.json configs[4]: synthetic sweep ref_level 4-7 x averaging samples 8-256 -- batched SpMM (Y = K~ X) alone and the
full Chebyshev apply, at 1/2/4/8 GPUs.

    python profiles/sweep.py [--refs 4,5,6,7] [--samples 8,16,32,64,128,256] > profiles/r02_sweep_n1.jsonl
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 profiles/sweep.py ...

With N GPUs the B samples of a row are sharded over the ranks (B/N columns each, as the averaging window is); every figure is
the whole-job aggregate: columns of all ranks / max-over-ranks device time.  Bytes of the SpMM: 16 N B + matK~ per panel.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krylov_exp_b200 as kx
from krylov_exp_b200.assembly import assemble_sw, test_pulse

ap = argparse.ArgumentParser()
ap.add_argument("--refs", default="4,5,6,7")
ap.add_argument("--samples", default="8,16,32,64,128,256")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def max_over_ranks(v):
    if world == 1:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6544.3
for ref in [int(r) for r in args.refs.split(",")]:
    sw = assemble_sw(ref)
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, device=local, rtol=1e-11, max_it=40)
    eig = kx.EIGS[ref - 3] if ref <= 5 else kx.EIGS[2] * 2.0 ** (ref - 5) * 1.03
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, eig * 900.0)
    ce_e = kx.cheby_exp(op, None, None, 10000, 1e-8, eig * 900.0, spectral_radius=eig)     # economised per-column series
    x = test_pulse(sw)
    for Btot in [int(b) for b in args.samples.split(",")]:
        if Btot % world:
            continue
        B = Btot // world
        t_all, w_all = kx.sample_times_weights(max(Btot, 2), 900.0, 1.0)
        t = np.ascontiguousarray(t_all[:Btot][rank::world])
        X = np.repeat(x[None, :], B, axis=0) * (1.0 + 0.01 * (rank + world * np.arange(B)))[:, None]
        bi, bo = op.batch(B), op.batch(B)
        bi.upload(X)
        lib = op.lib

        def timed(fn, reps, warm=2):
            for _ in range(warm):
                fn()
            op.synchronize()
            if dist is not None:
                dist.barrier()
            op.timer_start()
            for _ in range(reps):
                fn()
            return max_over_ranks(op.timer_stop() / reps)
        ms_spmm = timed(lambda: op._check(lib.kexp_spmm_K_dev(op.ctx, bi.handle, bo.handle), "spmm"), 10)
        bytes_spmm = (16.0 * sw.n * B + 80.0 * (op.nnz["Kt"] / 9.0) * bi.num_panels) * world
        ms_apply = timed(lambda: ce.apply_dev(bi, bo, -t), 3)
        bi.set_order(np.argsort(np.abs(t), kind="stable"))
        bi.upload(X)
        ms_econ = timed(lambda: ce_e.apply_dev(bi, bo, -t), 3)
        row = {"ref": ref, "N": sw.n, "samples": Btot, "n_gpus": world, "columns_per_gpu": B, "spmm_ms": ms_spmm,
               "spmm_gbs": bytes_spmm / ms_spmm / 1e6, "spmm_frac_of_hbm_per_gpu": bytes_spmm / ms_spmm / 1e6 / PEAK / world,
               "spmm_gflops": 2.0 * op.nnz["Kt"] * Btot / ms_spmm / 1e6,
               "apply_ms": ms_apply, "applications_per_s_full_series": Btot / ms_apply * 1e3,
               "apply_econ_ms": ms_econ, "applications_per_s": Btot / ms_econ * 1e3, "ncheb": int(ce.ncheb),
               "solve_path": int(op.stats()["solve_path"])}
        if rank == 0:
            print(json.dumps(row), flush=True)
        bi.close(); bo.close()
    op.close()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
