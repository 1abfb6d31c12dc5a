set -x
cd /root/repo
export KEXP_NO_GRAPH=1
python profiles/run_profile.py 5 64 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_r1c.csv python profiles/run_profile.py 5 64 > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/plain.log
python profiles/run_profile.py 5 64 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bsr3_cheb_iter_kernel|bsr3_spmm_kernel|cheb_update_kernel' -s 40 -c 8 -o gpurun_out/prof_r1c python profiles/run_profile.py 5 64 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
