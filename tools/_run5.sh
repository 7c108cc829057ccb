cd /root/repo
timeout 600 python -m pytest tests/test_texthead.py -m gpu -x -q 2>&1 | tail -4
timeout 100 python tools/time_texthead.py 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -o gpurun_out/prof_all2 python tools/profile_step.py > gpurun_out/ncu_all2.log 2>&1; tail -2 gpurun_out/ncu_all2.log
ncu -i gpurun_out/prof_all2.ncu-rep --page raw --csv > gpurun_out/prof_all2_raw.csv; wc -l gpurun_out/prof_all2_raw.csv
rm -f gpurun_out/prof_all2.ncu-rep
