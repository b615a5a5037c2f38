cd /root/repo
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py --steps 100 --warmup 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline'])"
