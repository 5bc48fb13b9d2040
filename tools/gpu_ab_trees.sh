# Interleaved bench of two trees on ONE box: the current one and an older one checked out and built under ./_old
# (git worktree add -f _old <commit> && (cd _old && python -c "import __graft_entry__ as g; g.build()"); remove the worktree afterwards).
# Output: tree, ms_per_step, e2e ms_per_step, median SM MHz per run (profiles/r02_tree_ab_round_start_vs_final.txt).
for r in 1 2 3 4; do
  for tree in . _old; do
    (cd $tree && python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ctc 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tree', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])")
  done
done
