#!/bin/bash
# The driver's own invocation of the bench (N = 1, default sizes) plus the reference arm.  gpurun --timeout 1500 -- 'bash tools/gpu_bench_full.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$(date +%s)
timeout 1200 python bench.py --gpus 1 --steps 8 --warmup 3 > $OUT/bench_full.json 2> $OUT/bench_full.err; echo "bench rc=$? in $(( $(date +%s) - S )) s"
tail -5 $OUT/bench_full.err
S=$(date +%s)
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/bench_full_reference.json 2> $OUT/bench_full_reference.err; echo "reference rc=$? in $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
for k in ('value','ms_per_step','scaling','parity_frame0','parity_frame0_in_step_output','parity_frame0_triclinic','hist_sha256','gpu_launches','kernel_share'):
    print(k, d.get(k))
print('roofline', d['roofline']['achieved'], d['roofline']['frac']); print('e2e', d['e2e']); print('cpu', d['cpu_baseline'])
for leg in ('rdf_triclinic','msd','green_kubo','residence','clusters_hydration','dump_parse','rdf_from_files','c1'):
    v=d.get(leg)
    if not v: print(leg, None); continue
    print(leg, {k:(v[k] if not isinstance(v[k],dict) else {kk:vv for kk,vv in v[k].items() if kk in ('achieved','frac','bound','value')}) for k in v if k not in ('config','note','api')})
r=json.load(open('gpurun_out/bench_full_reference.json')); print('reference', r['value'], r['cpu_baseline'])
PY
