#!/bin/bash
# Shell search (csrc/shell.cu) and FFT correlation (csrc/fftcorr.cu): parity tests, then the residence / Green-Kubo legs.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -k "shell or residence or pair_list or hydration or cluster or xcorr or fft or conduct or visc" 2>&1 | tail -4
timeout 600 python bench.py --steps 2 --warmup 1 --frames 16 --skip-msd --skip-cpu --skip-triclinic > $OUT/bench_sf.json 2> $OUT/bench_sf.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_sf.json'))
r=d['residence']; print('residence', {k:r[k] for k in ('ms_per_step','search_ms','correlation_ms','neighbour_entries','cnt_sha256')}, r['roofline']['frac'])
c=d['clusters_hydration']; print('c5', {k:c[k] for k in ('ms_per_step','hydration_search_ms','hydration_epilogue_ms','cluster_search_ms','cluster_epilogue_ms','hydration_entries','cluster_entries')})
g=d['green_kubo']; print('gk', g['ms_per_step'], g['method'], g['roofline']['frac'], g['roofline']['note'][:200])
PY
if [ "${1:-}" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shell_grid -s 1 -c 1 -f -o $OUT/prof_shell \
      python bench.py --steps 1 --warmup 1 --frames 16 --res-frames 1000 --skip-msd --skip-gk --skip-cpu --skip-triclinic --skip-clusters > $OUT/ncu_shell.log 2>&1
  echo "ncu shell rc=$?"
fi
