#!/bin/bash
# N-rank == 1-rank under NCCL + a short strong-scaling run.  gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_nrank_check.sh 2'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
N=${1:-2}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 1000 -k "nrank" > $OUT/pytest_nrank.log 2>&1; echo "pytest nrank rc=$?"
tail -8 $OUT/pytest_nrank.log
for n in 1 $N; do
  if [ $n = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 4 --warmup 3 --skip-cpu > $OUT/bench_scale_$n.json 2> $OUT/bench_scale_$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $n --steps 4 --warmup 3 --skip-cpu > $OUT/bench_scale_$n.json 2> $OUT/bench_scale_$n.err
  fi
  echo "bench N=$n rc=$?"; tail -2 $OUT/bench_scale_$n.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_scale_$n.json'))
    print('N=$n', d['value'], d['ms_per_step'], d['hist_sha256'][:12], d.get('nrank_equals_1rank'), 'e2e', d['e2e']['value'])
    print('   msd', d['msd']['value'], d['msd']['msd_last_frame'], 'flux', d['green_kubo']['charge_flux']['value'], 'acf ms', d['green_kubo']['ms_per_step'])
    r=d['residence']; print('   residence', r['ms_per_step'], r['search_ms'], r['exchange_ms'], r['correlation_ms'], r['cnt_sha256'][:12])
except Exception as e:
    print('no json', e)
PY
done
