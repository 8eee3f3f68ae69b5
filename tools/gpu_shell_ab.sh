#!/bin/bash
# A/B of the small-set shell search (csrc/shell.cu) against the general engine's list mode on the C5 legs.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -k "shell or residence or pair_list or hydration or cluster" 2>&1 | tail -3
for mode in 0 1; do
  export MDP_SHELL_GRID=$mode
  timeout 600 python bench.py --steps 2 --warmup 1 --frames 16 --skip-msd --skip-gk --skip-cpu --skip-triclinic > $OUT/bench_shell_$mode.json 2> $OUT/bench_shell_$mode.err; echo "bench shell=$mode rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_shell_$mode.json'))
r=d['residence']; print('residence', {k:r[k] for k in ('ms_per_step','search_ms','correlation_ms','neighbour_entries','cnt_sha256')})
c=d['clusters_hydration']; print('c5', {k:c[k] for k in ('ms_per_step','hydration_search_ms','hydration_epilogue_ms','cluster_search_ms','cluster_epilogue_ms','hydration_entries','cluster_entries','cluster_member_molecules','oriented_waters')})
PY
done
if [ "${1:-}" = "ncu" ]; then
  export MDP_SHELL_GRID=1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shell_grid -s 1 -c 1 -f -o $OUT/prof_shell \
      python bench.py --steps 1 --warmup 1 --frames 16 --res-frames 1000 --skip-msd --skip-gk --skip-cpu --skip-triclinic --skip-clusters > $OUT/ncu_shell.log 2>&1
  echo "ncu shell rc=$?"
fi
