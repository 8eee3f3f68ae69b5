#!/bin/bash
# One ncu --set full capture of a named kernel inside a reduced bench run.
# Usage: bash tools/gpu_ncu_kernel.sh <kernel-regex> <out-name> [extra bench.py flags...]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K=$1; O=$2; shift 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/$O \
    python bench.py --steps 1 --warmup 1 --frames-per-step 8 --skip-cpu "$@" > gpurun_out/ncu_$O.log 2>&1
echo "ncu $O rc=$?"
tail -2 gpurun_out/ncu_$O.log
