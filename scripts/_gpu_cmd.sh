for pair in 1 0; do
  echo "== PAIR=$pair"
  ORI_TC_PAIR=$pair python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], [(k['kernel'], round(k['ms'],2), round(k['achieved'])) for k in d['roofline']['kernels']], d['clocks'])"
done
