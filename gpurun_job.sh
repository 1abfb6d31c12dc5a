cd /root/repo
timeout 900 python -m pytest tests/test_gpu_cheby.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -4
for coop in 0 1; do
KEXP_COOP=$coop timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/b.json 2>gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b.json'))
print('coop=$coop', round(d['value'],1), 'indep', round(d['value_independent_columns'],1), d['config']['mass_solve'], d['gpu_launches'])
for k,v in d['kernels'].items(): print('   ',k, round(v['avg_us'],1), round(v['gbs'] or 0), round(v['share'] or 0,3), v['launches'])
PY
done
