cd /root/repo
timeout 600 python -m pytest tests/test_texthead.py -m gpu -x -q 2>&1 | tail -5
for m in mixed tc fp32; do echo "mode $m"; TGFR_TEXTHEAD_PRECISION=$m timeout 100 python tools/time_texthead.py 2>&1 | tail -2; done
