import sys, os
sys.path.insert(0,'/root/repo')
import numpy as np
import krylov_exp_b200 as kx
from krylov_exp_b200.assembly import assemble_sw, test_pulse
sw=assemble_sw(5); L=kx.EIGS[2]*900.0
op=kx.Operator(sw.K,sw.Mu,sw.Meta_inv_blocks,rtol=1e-11,max_it=40)
ce=kx.cheby_exp(op,None,None,10000,1e-8,L)
x=test_pulse(sw); t,w=kx.sample_times_weights(64,900.0,1.0)
bi,bo=op.batch(64),op.batch(64); bi.upload(x,shared=True)
for _ in range(3): ce.apply_dev(bi,bo,t)
op.synchronize(); op.timer_start()
for _ in range(5): ce.apply_dev(bi,bo,t)
print('fwd shared ms', op.timer_stop()/5)
op.set_kernel_timing(True); op.stats_reset()
ce.apply_dev(bi,bo,t); op.synchronize(); st=op.stats()
for name in ("spmm_K","cg_init","spmm_cg","cheb_update","misc"):
    cs=op.class_stats(name); print(name, cs['launches'], 'total ms', round(cs['ms'],3), 'avg us', round(1e3*cs['ms']/max(cs['launches'],1),1))
print(st)
