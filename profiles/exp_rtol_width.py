import sys, os, time, json
sys.path.insert(0,'/root/repo')
import numpy as np
import krylov_exp_b200 as kx
from krylov_exp_b200.assembly import assemble_sw, test_pulse
from oracle import coeffs, norms
from oracle.cpu_port import CpuCheby
sw=assemble_sw(5)
L=kx.EIGS[2]*900.0
x=test_pulse(sw)
def inputs(B):
    X=np.empty((B,sw.n)); amp=1e-3*np.abs(x).max()
    for k in range(B):
        rng=np.random.default_rng(1234+k); X[k]=x
        X[k,sw.n_u:]+=amp*rng.standard_normal(sw.n_eta)
        X[k,:sw.n_u]+=amp*np.sqrt(sw.g/sw.H)*sw.radius/32*rng.standard_normal(sw.n_u)
    return X
# reference for parity: CPU port at 1e-14 on 3 columns
t64,w64=kx.sample_times_weights(64,900.0,1.0)
X64=inputs(64)
cols=[0,31,63]
c,n=coeffs.cheby_coeffs(10000,1e-8,L); a=coeffs.real_recurrence_coeffs(c,n)
cpu=CpuCheby(sw.K,sw.Mu,sw.Meta_inv_blocks,rtol=1e-14)
ref,_=cpu.apply_batch(a,n,L,X64[cols],-t64[cols])
for rtol in (1e-12,1e-11,1e-10,1e-9):
    op=kx.Operator(sw.K,sw.Mu,sw.Meta_inv_blocks,rtol=rtol,max_it=40)
    ce=kx.cheby_exp(op,None,None,10000,1e-8,L)
    bi,bo=op.batch(64),op.batch(64); bi.upload(X64)
    for _ in range(3): ce.apply_dev(bi,bo,-t64)
    op.synchronize(); op.stats_reset()
    op.timer_start()
    for _ in range(3): ce.apply_dev(bi,bo,-t64)
    ms=op.timer_stop()/3
    st=op.stats()
    Y=bo.download()
    err=max(max(norms.calculate_norm(sw,Y[k],ref[i])) for i,k in enumerate(cols))
    print(f"rtol {rtol:g}: {64/(ms*1e-3):.1f} apps/s, its/solve {st['cg_iterations']/st['cg_solves']:.1f}, parity err {err:.2e}", flush=True)
    op.close()
# batch width sweep at rtol 1e-12
for B in (64,32,16,8):
    op=kx.Operator(sw.K,sw.Mu,sw.Meta_inv_blocks,rtol=1e-12,max_it=40)
    ce=kx.cheby_exp(op,None,None,10000,1e-8,L)
    t,w=kx.sample_times_weights(B,900.0,1.0)
    bi,bo=op.batch(B),op.batch(B); bi.upload(X64[:B])
    for _ in range(3): ce.apply_dev(bi,bo,-t)
    op.synchronize()
    op.timer_start()
    for _ in range(4): ce.apply_dev(bi,bo,-t)
    ms=op.timer_stop()/4
    print(f"B={B}: {ms:.2f} ms/apply, {B/(ms*1e-3):.1f} apps/s", flush=True)
    op.close()
