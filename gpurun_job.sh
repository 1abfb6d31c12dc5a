cd /root/repo
timeout 600 python -m pytest tests/test_gpu_krylov.py -x -q 2>&1 | tail -8
echo "== smem"; timeout 300 python profiles/krylov_m4.py 2>&1 | grep -v "^$" | tail -4
echo "== l2"; KEXP_BICG_SMEM=0 timeout 300 python profiles/krylov_m4.py 2>&1 | grep -v "^$" | tail -4
