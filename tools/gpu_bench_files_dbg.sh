#!/bin/bash
# Timeline of the rdf_from_files leg inside the FULL default bench.py run (all legs before it), against the leg alone.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mode in all alone; do
  rm -f gpurun_out/bench_files_trace_$mode.json
  if [ $mode = alone ]; then ARGS="--frames 64 --skip-msd --skip-gk --skip-residence --skip-clusters --skip-triclinic --skip-cpu --files-leg"; else ARGS=""; fi
  MDP_PIPELINE_TRACE=gpurun_out/bench_files_trace_$mode.json timeout 900 python bench.py --steps 8 --warmup 3 $ARGS > gpurun_out/bench_files_dbg_$mode.json 2> gpurun_out/bench_files_dbg_$mode.err
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
