#!/bin/bash
# Quick GPU iteration: parity tests + a short bench of the pair path (no ncu).  gpurun --timeout 900 -- 'bash tools/gpu_quick.sh [pytest-k-expr]'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
K=${1:-}
if [ -n "$K" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 600 -k "$K" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
else
  timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
fi
tail -15 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-msd --skip-cpu > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench rc=$?"
tail -3 $OUT/bench_quick.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_quick.json'))
    print({k:d[k] for k in ('value','ms_per_step','evaluated_pair_evals_per_step','kernel_share','gpu_launches')})
    print(d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'])
except Exception as e:
    print('no bench json', e)
PY
