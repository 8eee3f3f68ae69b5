#!/bin/bash
# One ncu --set full capture of the pair kernel on the bench workload (16 frames).  Output: gpurun_out/prof_pair.ncu-rep
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 3 -c 1 -f -o gpurun_out/prof_pair \
    python bench.py --steps 1 --warmup 1 --frames-per-step ${1:-16} --skip-msd --skip-cpu --skip-gk --skip-residence > gpurun_out/ncu_pair.log 2>&1
echo "ncu pair rc=$?"
tail -2 gpurun_out/ncu_pair.log
