timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --config c3 --steps 5 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']); print(d['value'], d['cpu_baseline']['value'])"
