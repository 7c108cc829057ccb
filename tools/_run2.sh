cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "arc or mag or head or focal or margin or cos_logits" 2>&1 | tail -6
timeout 120 python tools/time_head.py 2>&1 | tail -2
