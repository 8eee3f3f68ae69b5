#!/bin/bash
# Round-2 evidence run: peaks, whole GPU suite, default bench + reference arm, ncu launch list, full captures of the main kernels.
# gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv | tee $OUT/round2.log
lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' | tee -a $OUT/round2.log
timeout 600 tools/peaks > $OUT/peaks.json 2> $OUT/peaks.err; echo "peaks rc=$?" | tee -a $OUT/round2.log; cat $OUT/peaks.json
timeout 1400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/round2.log
tail -4 $OUT/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
bash tools/gpu_bench_full.sh
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --frames 64 --msd-atoms 250000 --msd-frames 512 --skip-cpu --gk-steps 20000 --gk-flux-frames 1024 --res-frames 500 --c5-frames 100 > $OUT/ncu_launch_bench.log 2>&1
echo "launch list rc=$?"
echo "== ncu full captures"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair_fast -s 3 -c 1 -f -o $OUT/prof_pair_fast \
    python bench.py --steps 1 --warmup 1 --frames 16 --skip-msd --skip-cpu --skip-gk --skip-residence --skip-clusters --skip-triclinic > $OUT/ncu_pair_fast.log 2>&1
echo "ncu pair_fast rc=$?"
MDP_PAIR_F64=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 3 -c 1 -f -o $OUT/prof_pair \
    python bench.py --steps 1 --warmup 1 --frames 16 --skip-msd --skip-cpu --skip-gk --skip-residence --skip-clusters --skip-triclinic > $OUT/ncu_pair.log 2>&1
echo "ncu pair(f64) rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msd_single -s 3 -c 1 -f -o $OUT/prof_msd \
    python bench.py --steps 1 --warmup 1 --frames 4 --msd-atoms 1000000 --msd-frames 64 --skip-cpu --skip-gk --skip-residence --skip-clusters --skip-triclinic --skip-msd-window > $OUT/ncu_msd.log 2>&1
echo "ncu msd rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shell_grid -s 1 -c 1 -f -o $OUT/prof_shell \
    python bench.py --steps 1 --warmup 1 --frames 4 --res-frames 1000 --skip-msd --skip-gk --skip-cpu --skip-triclinic --skip-clusters > $OUT/ncu_shell.log 2>&1
echo "ncu shell rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fft_pass -s 3 -c 1 -f -o $OUT/prof_fft \
    python bench.py --steps 1 --warmup 1 --frames 4 --skip-msd --skip-cpu --skip-residence --skip-triclinic --skip-clusters --gk-flux-frames 256 > $OUT/ncu_fft.log 2>&1
echo "ncu fft rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_charge_flux -s 3 -c 1 -f -o $OUT/prof_flux \
    python bench.py --steps 1 --warmup 1 --frames 4 --skip-msd --skip-cpu --skip-residence --skip-triclinic --skip-clusters --gk-flux-frames 2048 --gk-steps 20000 > $OUT/ncu_flux.log 2>&1
echo "ncu flux rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dump_rows -s 2 -c 1 -f -o $OUT/prof_dump \
    python bench.py --steps 1 --warmup 1 --frames 4 --skip-msd --skip-gk --skip-residence --skip-triclinic --skip-clusters > $OUT/ncu_dump.log 2>&1
echo "ncu dump rows rc=$?"
bash tools/gpu_files_trace.sh > $OUT/trace.log 2>&1; echo "files timeline rc=$?"; head -12 $OUT/files_timeline.txt
echo "== done $(date -u)"
