#!/bin/bash
# One GPU-box visit: peaks, parity tests, smoke, bench, ncu launch list + one full capture of the pair kernel.
# Usage (from the build container):  gpurun --timeout 1700 -- 'bash tools/gpu_round.sh [quick]'
# Everything worth keeping is written under gpurun_out/.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
MODE=${1:-full}
echo "== $(date -u) mode=$MODE" | tee $OUT/round.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv | tee -a $OUT/round.log
lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' | tee -a $OUT/round.log

echo "== peaks" | tee -a $OUT/round.log
timeout 300 tools/peaks > $OUT/peaks.json 2> $OUT/peaks.err; echo "peaks rc=$?" | tee -a $OUT/round.log
cat $OUT/peaks.json | tee -a $OUT/round.log

echo "== smoke" | tee -a $OUT/round.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/round.log
tail -5 $OUT/smoke.log | tee -a $OUT/round.log

echo "== pytest -m gpu" | tee -a $OUT/round.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/round.log
tail -40 $OUT/pytest_gpu.log | tee -a $OUT/round.log

echo "== bench" | tee -a $OUT/round.log
timeout 900 python bench.py --steps 6 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/round.log
cat $OUT/bench.json | tee -a $OUT/round.log
tail -5 $OUT/bench.err | tee -a $OUT/round.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
cat $OUT/bench_reference.json | tee -a $OUT/round.log

if [ "$MODE" != "quick" ]; then
  echo "== ncu launch list" | tee -a $OUT/round.log
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 1 --frames-per-step 16 --msd-frames 32 --skip-cpu --gk-steps 20000 --gk-flux-frames 256 --res-frames 500 > $OUT/ncu_launch_bench.log 2>&1
  echo "launch list rc=$?" | tee -a $OUT/round.log
  echo "== ncu full capture of k_pair and k_msd_single" | tee -a $OUT/round.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 3 -c 1 -f -o $OUT/prof_pair \
      python bench.py --steps 1 --warmup 1 --frames-per-step 16 --skip-msd --skip-cpu --skip-gk --skip-residence > $OUT/ncu_pair.log 2>&1
  echo "ncu pair rc=$?" | tee -a $OUT/round.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msd_single -s 3 -c 1 -f -o $OUT/prof_msd \
      python bench.py --steps 1 --warmup 1 --frames-per-step 4 --msd-frames 64 --skip-cpu --skip-gk --skip-residence --skip-triclinic > $OUT/ncu_msd.log 2>&1
  echo "ncu msd rc=$?" | tee -a $OUT/round.log
  bash tools/gpu_ncu_kernel.sh k_msd_window prof_msdw --skip-triclinic --skip-residence --skip-gk --msd-frames 64 | tee -a $OUT/round.log
  bash tools/gpu_ncu_kernel.sh k_xcorr prof_xcorr --skip-msd --skip-triclinic --skip-residence --gk-steps 50000 | tee -a $OUT/round.log
  bash tools/gpu_ncu_kernel.sh k_charge_flux prof_flux --skip-msd --skip-triclinic --skip-residence --gk-steps 20000 --gk-flux-frames 1024 | tee -a $OUT/round.log
  bash tools/gpu_ncu_kernel.sh k_bitmask_autocorr prof_bitmask --skip-msd --skip-triclinic --skip-gk --res-frames 2000 | tee -a $OUT/round.log
fi
echo "== done $(date -u)" | tee -a $OUT/round.log
