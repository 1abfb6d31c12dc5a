cd /root/repo
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -5
timeout 1200 python profiles/sweep.py 6 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; tail -3 gpurun_out/sweep.err; cat gpurun_out/sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['ref'], d['B'], 'spmm', round(d['spmm_ms'],3), 'ms', round(d['spmm_gbs']), 'GB/s', 'apply', round(d['apply_ms'],1), 'ms', round(d['applications_per_s'],1), 'apps/s')
"
