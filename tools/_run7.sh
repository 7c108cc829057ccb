cd /root/repo
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_multirank.log; cat gpurun_out/pytest_multirank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -2 gpurun_out/bench_n2.err; head -c 400 gpurun_out/bench_n2.json; echo
