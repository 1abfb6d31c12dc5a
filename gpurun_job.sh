cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
KEXP_PDL=1 timeout 600 python -m pytest tests/test_gpu_cheby.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
for pdl in 0 1; do
KEXP_PDL=$pdl timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline > gpurun_out/b.json 2>gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b.json'))
print('pdl=$pdl', round(d['value'],1), d['config']['mass_solve'], d['gpu_launches'])
PY
done
