cd /root/repo
python profiles/exp_fwd_breakdown.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_cheby.py tests/test_gpu_edges.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-roofline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])"
