#!/bin/bash
# First hardware run of the four opt-in paths written without a GPU (device dump parser, run-based survival correlation,
# small-set shell search, FFT correlation): their gated parity tests, then the residence and file-based bench legs with and without them.
# Usage: gpurun --timeout 1200 -- 'bash tools/gpu_experimental.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
MDP_TEST_DEVICE_PARSE=1 MDP_TEST_SURVIVAL_RUNS=1 MDP_TEST_SHELL_GRID=1 MDP_TEST_XCORR_FFT=1 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider \
    --timeout 600 -k "device_dump_parser or survival_runs or shell_grid or xcorr_fft" > $OUT/pytest_experimental.log 2>&1
echo "pytest experimental rc=$?"; tail -15 $OUT/pytest_experimental.log
for mode in off on; do
  if [ $mode = on ]; then export MDP_SURVIVAL_RUNS=1 MDP_SHELL_GRID=1 MDP_DEVICE_PARSE=1 MDP_XCORR_FFT=1; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-msd --skip-triclinic > $OUT/bench_exp_$mode.json 2> $OUT/bench_exp_$mode.err
  echo "bench $mode rc=$?"; tail -2 $OUT/bench_exp_$mode.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_exp_$mode.json"))
    r = d["residence"]; print("$mode residence", {k: r[k] for k in ("ms_per_step", "search_ms", "exchange_ms", "correlation_ms", "cnt0")})
    g = d.get("green_kubo"); print("$mode acf", g and g["ms_per_step"])
    f = d.get("rdf_from_files"); print("$mode files", f and {k: f[k] for k in ("ms_per_frame", "text_MB_per_s")})
except Exception as e:
    print("no json", e)
PY
done
