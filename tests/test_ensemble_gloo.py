"""CPU, world_size 2 (gloo): the sample-sharded weighted average + all-reduce equals the
single-process weighted sum (`DU *= wt; ensemble.allreduce(DU, V)`, timestepping_methods.py:10-11)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, Mbar, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import krylov_exp_b200 as kx
    ens = kx.SampleEnsemble(Mbar)
    assert ens.rank == rank and ens.world == world
    n = 37
    rng = np.random.default_rng(11)
    DU = rng.standard_normal((Mbar, n))                # what exp(-t_k L) N(exp(t_k L) U) would produce
    t, w = kx.sample_times_weights(Mbar, 900.0, 1.0)
    tl, wl = ens.local_times_weights(900.0, 1.0)
    assert np.array_equal(tl, t[ens.local_slice])
    v = wl @ DU[ens.local_slice]                       # local weighted column sum (the kernel's job on the GPU)
    ens.allreduce(v)                                   # the one collective of the path
    ref = w @ DU
    q.put((rank, float(np.abs(v - ref).max()), ens.n_local))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_average_world2():
    world, Mbar = 2, 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Mbar, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(o[0] for o in out) == [0, 1]
    assert all(o[1] < 1e-14 for o in out)
    assert sum(o[2] for o in out) == Mbar


def _worker_balanced(rank, world, port, Mbar, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import krylov_exp_b200 as kx
    t, w = kx.sample_times_weights(Mbar, 900.0, 1.0)
    terms = 8 + np.round(17 * np.abs(t) / np.abs(t).max()).astype(int)      # series length grows with |t_k|
    ens = kx.SampleEnsemble(Mbar, partition="balanced", terms=terms)
    n = 23
    rng = np.random.default_rng(5)
    DU = rng.standard_normal((Mbar, n))
    tl, wl = ens.local_times_weights(900.0, 1.0)
    assert np.array_equal(tl, t[ens.local_slice]) and len(tl) == ens.n_local
    v = wl @ DU[ens.local_slice]
    ens.allreduce(v)
    q.put((rank, float(np.abs(v - w @ DU).max()), ens.n_local, int(terms[ens.local_slice].max())))
    dist.barrier()
    dist.destroy_process_group()


def test_balanced_shares_world2():
    """Unequal cost-balanced shares (SampleEnsemble(partition="balanced")): the all-reduced weighted sum is the single-process
    one, every sample is held by exactly one rank, and the rank with the longest series does not hold more samples."""
    world, Mbar = 2, 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_balanced, args=(r, world, port, Mbar, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(o[1] < 1e-14 for o in out)
    assert sum(o[2] for o in out) == Mbar
    long_rank = max(out, key=lambda o: o[3])
    assert long_rank[2] <= min(o[2] for o in out)
