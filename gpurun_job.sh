set -x
cd /root/repo
export KEXP_NO_GRAPH=1
python profiles/run_profile.py 5 64 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r1.csv python profiles/run_profile.py 5 64 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/plain.log
python profiles/run_profile.py 5 64 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'cg_update_xr|bsr3_spmm_kernel|cg_update_p' -s 300 -c 6 -o gpurun_out/prof_r1 python profiles/run_profile.py 5 64 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out/
