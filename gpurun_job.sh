set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_cheby.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo rc=$?; tail -5 gpurun_out/bench6.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench6.json'))
print(d['value'], d['config']['mass_solve'], d['gpu_launches'])
for k,v in d['kernels'].items(): print(k, round(v['avg_us'],1), round(v['gbs'] or 0), round(v['share'],3), v['launches'])
PY
