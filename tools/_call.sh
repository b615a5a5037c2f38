cd /root/repo
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4
cd tools
timeout 120 python stage_bench.py 2>&1 | head -3
B200TRK_TC_WIDE=0 timeout 120 python stage_bench.py 2>&1 | head -1
cd ..
timeout 400 python bench.py --steps 100 --warmup 10 2>/dev/null | tail -1
