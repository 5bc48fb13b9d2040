for r in 1 2 3 4; do
  for tree in . _old; do
    (cd $tree && python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ctc 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tree', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])")
  done
done
