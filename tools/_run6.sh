cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; head -c 300 gpurun_out/bench.json; echo
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
