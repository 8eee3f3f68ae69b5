#!/bin/bash
# Pair-kernel iteration on the GPU box: pair parity tests, a short pair-only bench and one ncu --set full capture of k_pair.
# Usage: gpurun --timeout 1200 -- 'bash tools/gpu_pair_iter.sh [all|pair]'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
WHAT=${1:-pair}
if [ "$WHAT" = "all" ]; then
  timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
else
  timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 600 -k "pair or rdf or cn or bin" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
fi
tail -8 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 6 --warmup 3 --skip-msd --skip-cpu --skip-gk --skip-residence --skip-clusters > $OUT/bench_pair.json 2> $OUT/bench_pair.err; echo "bench rc=$?"
tail -3 $OUT/bench_pair.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_pair.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','evaluated_pair_evals_per_step','kernel_share','gpu_launches')})
    print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
    print('tricl', d.get('rdf_triclinic',{}).get('pair_kernel_ms_per_step'))
except Exception as e:
    print('no bench json', e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 3 -c 1 -f -o $OUT/prof_pair \
    python bench.py --steps 1 --warmup 1 --frames-per-step 16 --skip-msd --skip-cpu --skip-gk --skip-residence --skip-clusters --skip-triclinic > $OUT/ncu_pair.log 2>&1
echo "ncu pair rc=$?"
