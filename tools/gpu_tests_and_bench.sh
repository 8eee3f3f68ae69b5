#!/bin/bash
# Whole GPU test-suite, then the driver's bench invocation.  gpurun --timeout 2400 -- 'bash tools/gpu_tests_and_bench.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 $OUT/pytest_gpu.log
bash tools/gpu_bench_full.sh
