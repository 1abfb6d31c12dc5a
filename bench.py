#!/usr/bin/env python
"""bench.py -- exp(sL)v applications per second on the averaged-model hot path.

Workload M1 (BASELINE.json configs[1]): icosahedral sphere ref_level 5,
degree-3 geometry, BDM2 x DG1 (N = 215 040), dt = 0.25 h, rho = 1, tol = 1e-8
(L = 13.4595, 29 Chebyshev terms), an averaging window of 64 samples.  With N GPUs the 64 samples
are sharded over the ranks (strong scaling: sample k on rank k mod N, averaged_sw_explicit.py:37,
107,233-235 -- one ensemble member per sample); `--weak` gives every GPU 64 samples of a 64 N window
instead, and a short weak-scaling measurement is always reported beside the strong one.

One "step" = one averaged nonlinear evaluation's worth of exponentials
(timestepping_methods.py:7-11):
    forward   exp(+t_k L) U         one shared input, 64 sample times
    backward  exp(-t_k L) X_k       64 distinct inputs
    average   V = sum_k w_k (.)     weighted column reduction (+ NCCL all-reduce when N > 1)
= 128 applications of exp(sL) per step (whole job).

  value  applications/s with inputs resident in HBM (device batches)
  e2e    same through the host-pointer C-ABI calls (pinned host buffers, H2D/D2H inside)
  roofline  dominant kernel: algorithmic bytes / CUDA-event launch time vs measured HBM peak
  cpu_baseline  the oracle's C/OpenMP port of the reference algorithm on the host cores

`--impl reference` times the CPU port alone (the reference itself needs Firedrake/PETSc,
which cannot be installed here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = 900.0            # 0.25 h (averaged_sw_explicit.py:18,80)
RHO = 1.0
TOL = 1e-8            # averaged_sw_explicit.py:166-172
NCHEB0 = 10000
METRIC = "exp(sL)v applications/sec at icosa ref5 x 64 samples"


def workload(ref_level, samples, seed=1234):
    """Operator + M1 inputs.  Assembly is host-side input generation (krylov_exp_b200.assembly)."""
    from krylov_exp_b200.assembly import assemble_sw, test_pulse
    from krylov_exp_b200.samples import EIGS, sample_times_weights
    sw = assemble_sw(ref_level)
    eig = EIGS[ref_level - 3] if 3 <= ref_level <= 5 else EIGS[2] * 2.0 ** (ref_level - 5) * 1.03
    L = eig * DT * RHO
    t, w = sample_times_weights(samples, DT, RHO)
    x = test_pulse(sw)
    workload.eig = eig                     # the bound L is built from (averaged_sw_explicit.py:77-82,168)
    return sw, L, t, w, x


def noisy_inputs(sw, x, cols, out, seed=1234):
    """M1 backward inputs: column k of the window = pulse + noise from default_rng(seed + k).
    `cols`: the GLOBAL sample indices held in the rows of `out` (a rank's share of the window)."""
    amp = 1e-3 * np.abs(x).max()
    for i, k in enumerate(cols):
        rng = np.random.default_rng(seed + int(k))
        out[i] = x
        out[i, sw.n_u:] += amp * rng.standard_normal(sw.n_eta)
        out[i, :sw.n_u] += amp * np.sqrt(sw.g / sw.H) * sw.radius / 2 ** sw.ref_level * rng.standard_normal(sw.n_u)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_reference_run(sw, L, t, x, ncols, steps, warmup, rtol=1e-5, literal=True, nthreads=0):
    """The reference's CPU algorithm (literal recurrence cheby_exp.py:110-165, CG at PETSc's default
    rtol 1e-5, one sample time per core) through the oracle's C/OpenMP port.  Returns per-step seconds."""
    from oracle import coeffs
    from oracle.cpu_port import CpuCheby
    cpu = CpuCheby(sw.K, sw.Mu, sw.Meta_inv_blocks, rtol=rtol, max_it=500)
    c, n = coeffs.cheby_coeffs(NCHEB0, TOL, L)
    coef = c if literal else coeffs.real_recurrence_coeffs(c, n)
    X = noisy_inputs(sw, x, range(ncols), np.empty((ncols, sw.n)))
    tk = np.resize(-t, ncols)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        cpu.apply_batch(coef, n, L, X, tk, literal=literal, nthreads=nthreads)   # explicit: torchrun exports OMP_NUM_THREADS=1
        times.append(time.perf_counter() - t0)
    return times[warmup:], (nthreads if nthreads > 0 else cpu.num_threads()), n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kexp", choices=["kexp", "reference"])
    ap.add_argument("--ref-level", type=int, default=5)
    ap.add_argument("--samples", type=int, default=64, help="averaging samples of the window (whole job); per GPU with --weak")
    ap.add_argument("--weak", action="store_true", help="weak scaling: --samples per GPU, window of samples*N")
    ap.add_argument("--partition", default="balanced", choices=["balanced", "cyclic"],
                    help="how the window is split over N > 1 GPUs: cost-balanced unequal shares sorted by |t_k| "
                         "(ensemble.balanced_partition) or sample k on rank k mod N")
    ap.add_argument("--cg-rtol", type=float, default=1e-11,
                    help="relative residual of the batched mass solve; 1e-11 gives <= 3.2e-12 parity error at ref 5 "
                         "(tests/test_gpu_parity_r2.py::test_M1_exact_bench_settings_all_columns)")
    ap.add_argument("--max-it", type=int, default=40)
    ap.add_argument("--solver", default="chebyshev", choices=["chebyshev", "cg"],
                    help="mass solve: Chebyshev semi-iteration (no inner products) or batched PCG")
    ap.add_argument("--no-condense", action="store_true")
    ap.add_argument("--full-series", action="store_true",
                    help="every column runs all ncheb terms (no per-column economised re-expansion)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the variant / weak-scaling side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0))
    Btot = args.samples * (world if args.weak else 1)          # samples of the averaging window (whole job)
    workload_name = (f"M1: icosahedral ref{args.ref_level} BDM2xDG1, dt=0.25h rho=1 tol=1e-8, window of {Btot} samples"
                     + (f" ({args.samples}/GPU, weak)" if args.weak else f" sharded over {world} GPU(s)"))

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        sw, L, t, w, x = workload(args.ref_level, args.samples)
        ncols = cores
        times, nthreads, ncheb = cpu_reference_run(sw, L, t, x, ncols, args.steps, max(args.warmup, 1), nthreads=cores)
        per = float(np.mean(times))
        val = ncols / per
        sample = (f"{ncols} of the {args.samples} backward applications per step (one per host thread), literal "
                  f"2*ncheb-1={2 * ncheb - 1} operator solves each, block-Jacobi CG rtol 1e-5 (PETSc default)")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": "applications/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name, "ncheb": ncheb, "L": L},
            "cpu_baseline": {"value": val, "unit": "applications/s", "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "applications/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = Python on Firedrake/PETSc (not installable here); timed: oracle/cpu_cheby.c, the C/OpenMP "
                    "port of its algorithm, without Firedrake's per-solve RHS re-assembly"}))
        return

    # ------------------------------------------------------------------ GPU arm
    import torch
    import torch.distributed as dist
    import krylov_exp_b200 as kx

    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sw, L, _, _, x = workload(args.ref_level, Btot)
    n = sw.n
    t_all, w_all = kx.sample_times_weights(Btot, DT, RHO)
    if Btot % world:
        raise SystemExit(f"--samples {Btot} is not a multiple of the {world} GPUs")

    op = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, device=local_rank, condense=not args.no_condense,
                     rtol=args.cg_rtol, max_it=args.max_it, solver=args.solver)
    ce = kx.cheby_exp(op, None, None, NCHEB0, TOL, L, spectral_radius=None if args.full_series else workload.eig)
    ens = kx.SampleEnsemble(Btot, rank=rank, world=world, partition="cyclic")
    lib_comm = False
    if world > 1:
        try:                                   # NCCL inside libkexp.so: the all-reduce runs on the library's stream
            ens.init_library_comm(op)
            lib_comm = True
        except Exception as e:                 # torch.distributed's NCCL on torch's stream (needs a host-side join)
            if rank == 0:
                print(f"library NCCL unavailable ({e}); using torch.distributed", file=sys.stderr)

    def barrier():
        op.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(val):
        if world == 1:
            return float(val)
        tt = torch.tensor([val], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    class Shard:
        """One rank's share of a window, pinned host buffers for the e2e path, device batches for `value`.  Partition: "cyclic"
        (sample k on rank k mod N: every rank holds the same mix of |t_k|, i.e. of economised series lengths) or "balanced"
        (ensemble.balanced_partition: |t_k|-sorted groups of unequal size -- the step of a rank is decided by its longest
        series, so the rank that holds the longest ones gets fewer samples)."""

        def __init__(self, Bwin, partition="cyclic"):
            self.Bwin = Bwin
            tw, ww = kx.sample_times_weights(Bwin, DT, RHO)
            if partition == "balanced" and world > 1:
                self.groups = kx.balanced_partition(ce.series_terms(tw), world)
            else:
                self.groups = [np.arange(r, Bwin, world) for r in range(world)]
            self.cols = self.groups[rank]
            self.B = len(self.cols)
            self.t = np.ascontiguousarray(tw[self.cols])
            self.w = np.ascontiguousarray(ww[self.cols])
            self.hX = torch.empty((self.B, n), dtype=torch.float64).pin_memory().numpy()
            self.hY = torch.empty((self.B, n), dtype=torch.float64).pin_memory().numpy()
            self.hx = torch.empty((n,), dtype=torch.float64).pin_memory().numpy()
            self.hx[:] = x
            noisy_inputs(sw, x, self.cols, self.hX)
            self.f_in, self.f_out, self.b_in, self.b_out = (op.batch(self.B) for _ in range(4))
            self.f_in.upload(self.hx, shared=True)
            self.b_in.set_order(np.argsort(np.abs(self.t), kind="stable"))     # panels of similar |t_k|: similar series length
            self.b_in.upload(self.hX)
            self.v = torch.zeros(n, dtype=torch.float64, device=dev)

        def step_dev(self):
            ce.apply_dev(self.f_in, self.f_out, self.t)
            ce.apply_dev(self.b_in, self.b_out, -self.t)
            self.b_out.average_into(self.w, self.v.data_ptr())
            if world > 1:
                if lib_comm:
                    op._check(op.lib.kexp_allreduce_sum_dev(op.ctx, self.v.data_ptr(), n), "kexp_allreduce_sum_dev")
                else:
                    op.synchronize()
                    dist.all_reduce(self.v)

        def step_host(self):
            """The same step through the host-pointer C-ABI calls: H2D of 1 + B vectors, D2H of B + 1 vectors.  The D2H of
            the forward results stays in flight under the backward half; the all-reduce runs on the device."""
            ce.apply_batch(self.hx, self.t, shared=True, out=self.hY, wait=False)
            vh = ce.apply_average(self.hX, -self.t, self.w, allreduce=lib_comm)
            op.wait_transfers()
            if world > 1 and not lib_comm:
                ens.allreduce(vh)
            return vh

        def timed(self, fn, steps):
            """K steps bracketed by barrier + synchronize on both sides; device time on the library's stream (CUDA events),
            host time around it; max over ranks."""
            barrier()
            h0 = time.perf_counter()
            op.timer_start()
            for _ in range(steps):
                fn()
            ms_dev = op.timer_stop()
            barrier()
            ms_host = (time.perf_counter() - h0) * 1e3
            return max_over_ranks(ms_dev), max_over_ranks(ms_host)

        def close(self):
            for b_ in (self.f_in, self.f_out, self.b_in, self.b_out):
                b_.close()

    sh = Shard(Btot, args.partition)
    B = sh.B
    for _ in range(max(args.warmup, 3)):
        sh.step_dev()
    barrier()
    op.set_async(True)               # tuned now: applies queue back to back, convergence checked at the closing synchronize
    sh.step_dev()
    barrier()
    op.stats_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ms_dev, ms_host = sh.timed(sh.step_dev, args.steps)
    clocks = sampler.stop()
    st = op.stats()
    # with the library communicator everything (all-reduce included) is on the library's stream: CUDA-event time;
    # with torch's all-reduce (other stream, host join per step) the host-bracketed time is the one that contains it
    ms_total = ms_dev if (world == 1 or lib_comm) else ms_host
    apps_per_step = 2 * Btot
    ms_per_step = ms_total / args.steps
    value = apps_per_step / (ms_per_step * 1e-3)
    cg_its_per_solve = st["cg_iterations"] / max(st["cg_solves"], 1)
    terms_per_step = st["op_applications"] / args.steps
    solve_path = {0: "batched PCG", 1: "Chebyshev, one kernel per iteration", 2: "Chebyshev, one cooperative kernel per solve",
                  3: "persistent narrow-panel apply kernel"}.get(int(st["solve_path"]), "?")

    # ------------------------------------------------------------ where the step goes: forward / backward / average alone
    def _timed_ms(fn, reps=3):
        fn()
        op.synchronize()
        op.timer_start()
        for _ in range(reps):
            fn()
        return op.timer_stop() / reps
    ce.apply_dev(sh.f_in, sh.f_out, sh.t)
    fwd_path = int(op.stats()["solve_path"])
    breakdown = {"forward_ms": _timed_ms(lambda: ce.apply_dev(sh.f_in, sh.f_out, sh.t)),
                 "backward_ms": _timed_ms(lambda: ce.apply_dev(sh.b_in, sh.b_out, -sh.t)),
                 "average_ms": _timed_ms(lambda: sh.b_out.average_into(sh.w, sh.v.data_ptr()))}

    # ------------------------------------------------------------ multi-GPU numerics: the all-reduced average against
    # a single-GPU recomputation of the whole window on rank 0
    parity = None
    if world > 1:
        sh.step_dev()                          # the all-reduced average of a complete step (the breakdown above left a local sum in v)
        barrier()
        v_multi = sh.v.cpu().numpy().copy()
        if rank == 0:
            Xall = noisy_inputs(sw, x, range(Btot), np.empty((Btot, n)))
            v_single = ce.apply_average(Xall, -t_all, w_all)
            M = sw.M()
            d = v_multi - v_single
            parity = float(np.sqrt((d @ (M @ d)) / (v_single @ (M @ v_single))))
            del Xall
        barrier()

    # ------------------------------------------------------------ the same step with one of the exact rewritings
    # switched off (a second context), for transparency; N = 1 only
    def plain_variant(shared_basis, economise):
        if not shared_basis:
            os.environ["KEXP_SHARED_BASIS"] = "0"
        op2 = kx.Operator(sw.K, sw.Mu, sw.Meta_inv_blocks, device=local_rank, condense=not args.no_condense,
                          rtol=args.cg_rtol, max_it=args.max_it, solver=args.solver, spectrum=op.spectrum)
        os.environ.pop("KEXP_SHARED_BASIS", None)
        ce2 = kx.cheby_exp(op2, None, None, NCHEB0, TOL, L, spectral_radius=workload.eig if economise else None)
        f_in, f_out, g_in, g_out = op2.batch(B), op2.batch(B), op2.batch(B), op2.batch(B)
        f_in.upload(sh.hx, shared=True)
        g_in.set_order(np.argsort(np.abs(sh.t), kind="stable"))
        g_in.upload(sh.hX)
        for _ in range(3):
            ce2.apply_dev(f_in, f_out, sh.t); ce2.apply_dev(g_in, g_out, -sh.t)
        op2.synchronize()
        op2.timer_start()
        for _ in range(3):
            ce2.apply_dev(f_in, f_out, sh.t); ce2.apply_dev(g_in, g_out, -sh.t)
            g_out.average_into(sh.w, sh.v.data_ptr())
        ms2 = op2.timer_stop() / 3
        for b_ in (f_in, f_out, g_in, g_out):
            b_.close()
        op2.close()
        return 2 * B / (ms2 * 1e-3)

    extras = not args.no_extras and not args.no_roofline
    # ------------------------------------------------------------ the averaged evaluation with the fused backward half:
    # forward batch as above, then sum_k w_k exp(-t_k L) N_k by ONE Clenshaw recurrence (kexp_cheby_average_fused_dev) --
    # what timestepping_methods.py:7-11 needs; the backward results are never formed, so it is reported beside `value`
    fused = None
    if extras:
        def step_fused():
            ce.apply_dev(sh.f_in, sh.f_out, sh.t)
            ce.average_dev(sh.b_in, -sh.t, sh.w, sh.v.data_ptr())
            if world > 1 and lib_comm:
                op._check(op.lib.kexp_allreduce_sum_dev(op.ctx, sh.v.data_ptr(), n), "kexp_allreduce_sum_dev")
        if world == 1 or lib_comm:
            for _ in range(3):
                step_fused()
            op.synchronize()
            v_f = sh.v.clone()
            sh.step_dev()
            op.synchronize()
            dv = float(torch.linalg.norm(v_f - sh.v) / torch.linalg.norm(sh.v))
            ks = max(args.steps, 5)
            ms_f, _ = sh.timed(step_fused, ks)
            ms_f /= ks
            fused = {"ms_per_averaged_evaluation": ms_f, "equivalent_applications_per_s": apps_per_step / (ms_f * 1e-3),
                     "rel_l2_vs_separate_applications": dv,
                     "what": "forward shared-input batch (all results delivered) + fused Clenshaw weighted average of the "
                             "backward applications (only their sum is formed)"}

    value_indep = value_full = None
    if extras and rank == 0 and world == 1:
        value_indep = plain_variant(shared_basis=False, economise=not args.full_series)
        value_full = plain_variant(shared_basis=False, economise=False)

    # ------------------------------------------------------------ e2e through host pointers
    e2e = None
    if not args.no_e2e:
        sh.step_host()
        sh.step_host()
        ksteps = max(2, min(args.steps, 5))
        _, e_host = sh.timed(sh.step_host, ksteps)
        e_ms = e_host / ksteps
        e2e = {"value": apps_per_step / (e_ms * 1e-3), "unit": "applications/s",
               "h2d_bytes_per_step": int(8 * n * (Btot + world)), "d2h_bytes_per_step": int(8 * n * (Btot + world)),
               "ms_per_step": e_ms,
               "timing": "host clock around the host-pointer C-ABI calls (pinned buffers; the D2H of the forward results overlaps "
                         "the backward half; all-reduce on the device), barrier + synchronize on both sides, max over ranks"}

    # ------------------------------------------------------------ weak-scaling side measurement (N > 1, strong run)
    weak = None
    if extras and world > 1 and not args.weak:
        shw = Shard(args.samples * world)
        for _ in range(2):
            shw.step_dev()
        kw = max(2, min(args.steps, 3))
        msw_dev, msw_host = shw.timed(shw.step_dev, kw)
        msw = (msw_dev if lib_comm else msw_host) / kw
        weak = {"value": 2 * args.samples * world / (msw * 1e-3), "unit": "applications/s", "samples_per_gpu": args.samples,
                "ms_per_step": msw, "scaling": "weak"}
        shw.close()

    # ------------------------------------------------------------ roofline of the dominant kernel (rank 0)
    roofline, kernels = None, None
    path_t, solve_label = -1, {}
    if not args.no_roofline and rank == 0:
        peak, peak_src = measured_peak_gbs()
        op.set_async(False)
        op.set_kernel_timing(True)
        op.stats_reset()
        ce.apply_dev(sh.b_in, sh.b_out, -sh.t)
        op.synchronize()
        st_t = op.stats()
        path_t = int(st_t["solve_path"])
        kernels = {}
        total_ms = 0.0
        names = ("spmm_K", "cg_init", "spmm_cg", "cheb_update", "misc") if args.solver == "chebyshev" else \
                ("spmm_K", "cg_init", "spmm_cg", "cg_update_xr", "cheb_update", "misc")
        Wp = B / sh.b_in.num_panels
        for name in names:
            cs = op.class_stats(name)
            executed = cs["launches"]
            if name in ("spmm_cg", "cg_update_xr", "cg_update_p"):
                executed = st_t["cg_iterations"]      # PCG: launches after convergence return at once
                if args.solver == "chebyshev":        # K iterations per solve = 1 start kernel + K-1 fused kernels
                    executed = (st_t["cg_iterations"] - st_t["cg_solves"]) if name == "spmm_cg" else 0
            bytes_launch = cs["bytes_per_col"] * Wp + cs["matrix_bytes"]
            if name == "cheb_update":                 # the row-range launches of one term form one pass
                executed = st_t["op_applications"]
            if name == "spmm_cg" and path_t == 2 and cs["launches"] > 0:
                # cooperative kernel: ONE launch per mass solve = start + (K-1) iterations, matrix staged once
                executed = cs["launches"]
                K = st_t["cg_iterations"] / st_t["cg_solves"]
                bytes_launch = ((K - 1) * 32.0 + 24.0) * op.n_solve * Wp + cs["matrix_bytes"]
            if name == "spmm_cg" and path_t == 3 and cs["launches"] > 0:
                # persistent narrow-panel kernel: ONE launch per apply = all terms (K~ product, mass solve, updates)
                executed = cs["launches"]
                K = st_t["cg_iterations"] / st_t["cg_solves"]
                terms = st_t["op_applications"] / cs["launches"]
                mk = op.class_stats("spmm_K")["matrix_bytes"]
                mr = op.class_stats("cheb_update")["matrix_bytes"]
                bytes_launch = terms * ((16.0 * n + 40.0 * n) * Wp + mk + mr + ((K - 1) * 32.0 + 24.0) * op.n_solve * Wp)
            if executed <= 0 or cs["ms"] <= 0:
                continue
            avg_ms = cs["ms"] / executed
            kernels[name] = {"ms_total": cs["ms"], "launches": int(executed), "avg_us": avg_ms * 1e3,
                             "algorithmic_bytes": bytes_launch, "gbs": bytes_launch / (avg_ms * 1e-3) / 1e9 if bytes_launch > 0 else None}
            total_ms += cs["ms"]
        for k in kernels.values():
            k["share"] = k["ms_total"] / total_ms if total_ms > 0 else None
        solve_label = {0: "bsr3_spmm_kernel<W,1> (q = S p + fused dots)", 1: "bsr3_cheb_iter_kernel (one Chebyshev iteration of the mass solve per launch)",
                       2: "cheb_solve_coop_kernel (one whole mass solve per launch)", 3: "cheb_apply_narrow_kernel (one whole apply per launch)"}
        label = {"spmm_K": "bsr3_spmm_kernel<W,0> (T = K~ S)", "cg_init": "cheb_solve_init_kernel / cg_init_kernel",
                 "spmm_cg": solve_label.get(path_t, "?"), "cg_update_xr": "cg_update_kernel",
                 "cheb_update": "cheb_update_kernel (recurrence + accumulate)", "misc": "pack / combine / column sums"}
        for k in kernels:
            kernels[k]["kernel"] = label[k]
        cand = [k for k in kernels if kernels[k]["gbs"]]
        dom = max(cand, key=lambda k: kernels[k]["ms_total"])
        kd = kernels[dom]
        roofline = {"bound": "hbm", "kernel": kd["kernel"], "achieved": kd["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kd["gbs"] / peak, "traffic": None, "peak_source": peak_src,
                    "avg_launch_us": kd["avg_us"], "algorithmic_bytes_per_launch": kd["algorithmic_bytes"],
                    "share_of_step": kd["share"], "panel_columns": Wp}
        # whole-apply effective bandwidth from the byte model (all kernels)
        model_bytes = sum(k["algorithmic_bytes"] * k["launches"] for k in kernels.values())
        roofline["whole_apply_gbs"] = model_bytes / (total_ms * 1e-3) / 1e9 if total_ms > 0 else None
        roofline["whole_apply_frac"] = roofline["whole_apply_gbs"] / peak if total_ms > 0 else None
        op.set_kernel_timing(False)
        # measured DRAM traffic of the dominant kernel and the measured L2 bandwidth (profiles/, regenerated per round by
        # profiles/make_traffic_json.py from the ncu --set full capture and profiles/micro/l2_bandwidth.cu)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
            ent = tr.get(dom) if path_t == tr.get("solve_path", 2) and args.ref_level == tr.get("ref_level", 5) else None
            if isinstance(ent, dict):
                roofline["traffic"] = ent.get("dram_bytes_per_launch")
                roofline["traffic_source"] = ent.get("source")
                if ent.get("l2_bytes_per_launch") and tr.get("l2_peak_gbs"):
                    l2_gbs = ent["l2_bytes_per_launch"] / (kd["avg_us"] * 1e-6) / 1e9
                    roofline["l2_achieved_gbs"] = l2_gbs
                    roofline["l2_peak_gbs"] = tr["l2_peak_gbs"]
                    roofline["l2_frac"] = l2_gbs / tr["l2_peak_gbs"]
                roofline["note"] = ("the panel of the mass solve is L2 resident by design, so `achieved`/`frac` are EFFECTIVE (algorithmic) "
                                    "bandwidth against the HBM copy peak, `traffic` is the measured DRAM traffic per launch (ncu) and "
                                    "`l2_frac` the measured L2 traffic rate against the measured L2 bandwidth")
        except Exception:
            pass

    # ------------------------------------------------------------ CPU baseline (rank 0, N = 1)
    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        times, nthreads, ncheb = cpu_reference_run(sw, L, sh.t, x, cores, 1, 1, nthreads=cores)
        cpu = {"value": cores / float(np.mean(times)), "unit": "applications/s", "cores": nthreads, "kind": "port",
               "sample": f"{cores} backward applications (one per host thread), literal 2*ncheb-1={2 * ncheb - 1} solves, "
                         f"block-Jacobi CG rtol 1e-5; oracle/cpu_cheby.c"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "applications/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3) + 1, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if args.weak else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name, "ncheb": int(ce.ncheb), "L": L, "N": int(n), "applications_per_step": apps_per_step,
                       "window_samples": Btot, "samples_per_gpu": [int(len(g)) for g in sh.groups] if world > 1 else B,
                       "partition": (args.partition if world > 1 else "single GPU"),
                       "l2": "inputs/outputs of a step ~0.9 GB at 64 samples per GPU (>> 126 MB L2), no explicit flush; inside one mass "
                             "solve the 32-column panel (3 x 23.6 MB + matrix) is deliberately kept L2 resident, so the roofline "
                             "figure is effective (algorithmic) bandwidth",
                       "mass_solve": f"{'condensed ' if op.condensed else ''}block-Jacobi {'Chebyshev iteration' if args.solver == 'chebyshev' else 'PCG'} rtol {args.cg_rtol:g}, "
                                     f"{cg_its_per_solve:.1f} iterations/solve, n_solve={op.n_solve}; backward batch: "
                                     f"{solve_label.get(path_t, solve_path) if roofline else solve_path}; forward (one shared column): "
                                     f"{'cheb_apply_narrow_kernel (persistent, one launch per apply)' if fwd_path == 3 else 'same kernels, 2-column panel'}",
                       "parallelism": f"window of {Btot} samples sharded cyclically (sample k on rank k mod N), {B}/GPU; one all-reduce of N "
                                      f"doubles per step ({'NCCL on the library stream' if lib_comm else 'torch.distributed'})",
                       "panels_per_batch": sh.b_in.num_panels,
                       "forward_batch": "shared input: recurrence for ONE column on the common operator + dense re-expansion to the "
                                        "sample times (exact, DESIGN.md); value_independent_columns = same step without it",
                       "series": ("full reference series for every column" if args.full_series else
                                  "per-column economised re-expansion of the reference series (same polynomial, dropped tail "
                                  "<= 1e-13; t_valid = L/eig = dt*rho); value_full_series = all 29 terms for all 128 columns, "
                                  "no shared-input basis"),
                       "operator_applications_per_step": terms_per_step},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(st["kernel_launches"]),
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "parity_vs_single_gpu": parity,
            "value_independent_columns": value_indep, "value_full_series": value_full,
            "averaged_evaluation_fused": fused, "weak_scaling": weak,
            "timing": {"device_ms": ms_dev, "host_ms": ms_host, "rank0_breakdown": breakdown},
        }
        print(json.dumps(out))
    sh.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
