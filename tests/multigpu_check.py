"""Launched under torchrun (one process per GPU): the sample-sharded averaged apply equals the
single-GPU result, both with torch.distributed's NCCL and with NCCL inside libkexp.so.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import krylov_exp_b200 as kx
    from krylov_exp_b200.assembly import assemble_sw, test_pulse
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sw = assemble_sw(3)
    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, device=local, rtol=1e-13, max_it=80)
    dt, rho, Mbar = 3600.0, 1.0, 18
    L = kx.EIGS[0] * dt * rho
    ce = kx.cheby_exp(op, None, None, 10000, 1e-8, L)
    t, w = kx.sample_times_weights(Mbar, dt, rho)
    x = test_pulse(sw)
    X = np.stack([x * (1.0 + 0.05 * k) for k in range(Mbar)])
    ens = kx.SampleEnsemble(Mbar)
    sl = ens.local_slice
    tl, wl = ens.local_times_weights(dt, rho)
    # (1) torch NCCL
    v1 = ens.averaged_apply(ce, X[sl], -tl, wl)
    # (2) NCCL inside the library
    ens.init_library_comm(op)
    v2 = ens.averaged_apply(ce, X[sl], -tl, wl)
    # reference: everything on this rank's GPU
    ref = ce.apply_average(X, -t, w)
    M = sw.M()
    rel = lambda a, b: float(np.sqrt(((a - b) @ (M @ (a - b))) / (b @ (M @ b))))
    e1, e2 = rel(v1, ref), rel(v2, ref)
    print(f"rank {rank}/{world}: local samples {ens.n_local}, torch-nccl {e1:.2e}, lib-nccl {e2:.2e}", flush=True)
    assert e1 < 1e-12 and e2 < 1e-12
    # (3) cyclic partition, economised series, fused Clenshaw average with the library's all-reduce
    ce.set_spectral_radius(kx.EIGS[0])
    cyc = kx.SampleEnsemble(Mbar, partition="cyclic")
    cyc._lib_comm = ens._lib_comm
    tc, wc = cyc.local_times_weights(dt, rho)
    v3 = cyc.averaged_apply(ce, X[cyc.local_slice], -tc, wc, fused=True)
    v4 = cyc.averaged_apply(ce, X[cyc.local_slice], -tc, wc, fused=False)
    e3, e4 = rel(v3, ref), rel(v4, ref)
    print(f"rank {rank}/{world}: cyclic/economised fused {e3:.2e}, separate {e4:.2e}", flush=True)
    assert e3 < 1e-11 and e4 < 1e-11
    dist.barrier()
    op.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
