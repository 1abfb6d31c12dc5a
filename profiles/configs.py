"""BASELINE.json configs at their own sizes on one B200 (timings; parity is what tests/ check at oracle-feasible sizes).
    python profiles/configs.py > gpurun_out/configs.txt
M1 is bench.py's default workload.  M2: ref 5, rho = 7.45 (natural Mbar = 64, 130 terms).  M3: test_cheby.py as written
(ref 4, t = 6 h, tol 1e-6, B = 1).  M4: profiles/krylov_m4.py.  M5: ref 6, degree-1 mesh, H = 8000
(averaged_sw_williamson6.py:42-47), spectral bound from the device operator (the reference's table stops at ref 5),
64 samples on one GPU and the 8 samples one of 8 GPUs would hold."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import krylov_exp_b200 as kx
from krylov_exp_b200.assembly import assemble_sw, test_pulse


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def averaged_step(sw, op, eig, dt, rho, B, tol, filt=False):
    L = eig * dt * rho
    out = {}
    for name, sr in (("full series", None), ("economised", eig)):
        ce = kx.cheby_exp(op, None, None, 10000, tol, L, filter=filt, spectral_radius=sr)
        t, w = kx.sample_times_weights(B, dt, rho)
        x = test_pulse(sw)
        X = np.repeat(x[None, :], B, axis=0) * (1.0 + 0.01 * np.arange(B))[:, None]
        fi, fo, bi, bo = op.batch(B), op.batch(B), op.batch(B), op.batch(B)
        fi.upload(x, shared=True)
        bi.set_order(np.argsort(np.abs(t), kind="stable")); bi.upload(X)

        def step():
            ce.apply_dev(fi, fo, t); ce.apply_dev(bi, bo, -t); op.synchronize()
        op.stats_reset()
        ms = timed(step)
        st = op.stats()
        out[name] = (ms, 2 * B / ms * 1e3, ce.ncheb, st["op_applications"] / 5)
        for b in (fi, fo, bi, bo):
            b.close()
    return out


sw5 = assemble_sw(5)
op5 = kx.Operator(sw5.K, sw5.Mu, sw5.Meta_inv_blocks, rtol=1e-11, max_it=40)
for label, rho in (("M1 ref 5, dt 0.25 h, rho 1, 64 samples", 1.0), ("M2 ref 5, dt 0.25 h, rho 7.45, 64 samples", 7.45)):
    r = averaged_step(sw5, op5, kx.EIGS[2], 900.0, rho, 64, 1e-8)
    for k, (ms, aps, n, apps) in r.items():
        print(f"{label}: {k}: ncheb {n}, {apps:.0f} operator applications/step, {ms:.1f} ms/step, {aps:.0f} applications/s", flush=True)
op5.close()

sw4 = assemble_sw(4, c_scaled=True)                     # test_cheby.py:30,38-42 symmetric scaling
op4 = kx.Operator(sw4.K, sw4.Mu, sw4.Meta_inv_blocks, rtol=1e-11, max_it=40)
t3 = 6 * 3600.0
ce = kx.cheby_exp(op4, None, None, 10000, 1e-6, kx.EIGS[1] * t3)
x = test_pulse(sw4); y = np.empty_like(x)
ms = timed(lambda: ce.apply(x, y, t3))
print(f"M3 test_cheby.py (ref 4, t = 6 h, tol 1e-6): ncheb {ce.ncheb}, one apply(x, y, t) through host arrays {ms:.1f} ms", flush=True)
op4.close()

t0 = time.perf_counter()
sw6 = assemble_sw(6, degree=1, H=8000.0)
print(f"M5 host assembly ref 6 (degree-1 mesh): {time.perf_counter() - t0:.0f} s, n = {sw6.n}", flush=True)
op6 = kx.Operator(sw6.K, sw6.Mu, sw6.Meta_inv_blocks, rtol=1e-11, max_it=40)
t0 = time.perf_counter()
eig6 = op6.largest_imag_eig(tol=1e-8)
print(f"M5 spectral bound from the device operator: {eig6:.6e} ({time.perf_counter() - t0:.1f} s); x1.01 used", flush=True)
for B in (64, 8):
    r = averaged_step(sw6, op6, 1.01 * eig6, 900.0, 1.0, B, 1e-8, filt=True)
    for k, (ms, aps, n, apps) in r.items():
        print(f"M5 ref 6, H = 8000, filtered coefficients, {B} samples: {k}: ncheb {n}, {apps:.0f} operator applications/step, "
              f"{ms:.1f} ms/step, {aps:.0f} applications/s", flush=True)
op6.close()
