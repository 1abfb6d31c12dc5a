cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b.json 2>gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b.json'))
print(round(d['value'],1), 'indep', d['value_independent_columns'], 'e2e', d['e2e']['value'], d['config']['mass_solve'], d['gpu_launches'])
for k,v in d['kernels'].items(): print('   ',k, round(v['avg_us'],1), round(v['gbs'] or 0), round(v['share'] or 0,3), v['launches'])
PY
