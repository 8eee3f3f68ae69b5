#!/bin/bash
# Why is the rdf_from_files leg slower inside the full bench.py than alone?  Timeline of the leg's passes, with the CPU legs.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
free -g | head -2
for mode in cpu passive; do
  if [ $mode = passive ]; then export OMP_WAIT_POLICY=PASSIVE KMP_BLOCKTIME=0 GOMP_SPINCOUNT=0; fi
  rm -f gpurun_out/bench_files_trace_$mode.json
  ARGS="--frames 64 --skip-msd --skip-gk --skip-residence --skip-clusters --skip-triclinic"
  MDP_PIPELINE_TRACE=gpurun_out/bench_files_trace_$mode.json timeout 900 python bench.py --steps 2 --warmup 3 $ARGS > gpurun_out/bench_files_dbg_$mode.json 2> gpurun_out/bench_files_dbg_$mode.err
  echo "== $mode rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_files_dbg_$mode.json"))
print("leg:", {k: d["rdf_from_files"][k] for k in ("ms_per_frame", "value", "parser")})
for line in open("gpurun_out/bench_files_trace_$mode.json"):
    tl = json.loads(line)
    kinds = {}
    for s in tl["spans"]:
        kinds.setdefault(s["what"], []).append((s["t0_ms"], s["t1_ms"]))
    if not kinds or len(tl["spans"]) < 20: continue
    print("pass total %.1f ms:" % tl["total_ms"], "; ".join("%s n=%d sum=%.1f last=%.1f" % (k.split(" (")[0][:28], len(v), sum(b - a for a, b in v), max(b for _, b in v)) for k, v in kinds.items()))
PY
done
