cd /root/repo
for cfg in "KEXP_SCALAR_LANES=0" "KEXP_SCALAR_LANES=1" "KEXP_SCALAR_LANES=1 KEXP_PANEL_CAP=64"; do
env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.json 2>gpurun_out/b.err || tail -5 gpurun_out/b.err
python - <<PY
import json
d=json.load(open('gpurun_out/b.json'))
print('[$cfg]', round(d['value'],1), 'iter kernel us', round(d['kernels']['spmm_cg']['avg_us'],1), 'gbs', round(d['kernels']['spmm_cg']['gbs']), 'panels', d['config']['panels_per_batch'])
PY
done
KEXP_SCALAR_LANES=1 timeout 600 python -m pytest tests/test_gpu_cheby.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
