cd /root/repo
timeout 600 python -m pytest tests/test_gpu_cheby.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
for rtol in 1e-12 1e-11; do
timeout 300 python bench.py --steps 3 --warmup 3 --cg-rtol $rtol --no-cpu-baseline --no-e2e > gpurun_out/b_$rtol.json 2>gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b_$rtol.json'))
print('$rtol', round(d['value'],1), d['config']['mass_solve'], d['gpu_launches'])
for k,v in d['kernels'].items(): print('   ',k, round(v['avg_us'],1), round(v['gbs'] or 0), round(v['share'] or 0,3), v['launches'])
print(d['roofline'])
PY
done
