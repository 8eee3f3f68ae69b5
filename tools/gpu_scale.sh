#!/bin/bash
# Strong-scaling run of the bench on one box: N = 1, 2, 4, 8 (as many as the box has).  gpurun --gpus 8 --timeout 1800 -- 'bash tools/gpu_scale.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -le $NG ] || continue
  S=$(date +%s)
  if [ $n = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 6 --warmup 3 --skip-cpu > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 6 --warmup 3 --skip-cpu > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  fi
  echo "N=$n rc=$? in $(( $(date +%s) - S )) s"; tail -2 $OUT/scale_$n.err | cut -c1-200
done
python - <<'PY'
import json, os
base = None
for n in (1, 2, 4, 8):
    f = f"gpurun_out/scale_{n}.json"
    if not os.path.exists(f) or os.path.getsize(f) == 0: continue
    d = json.load(open(f))
    row = {"rdf": d["value"], "rdf_e2e": d["e2e"]["value"], "rdf_tricl": d["rdf_triclinic"]["value"], "msd": d["msd"]["value"],
           "flux": d["green_kubo"]["charge_flux"]["value"], "acf_ms": d["green_kubo"]["ms_per_step"],
           "res_ms": d["residence"]["ms_per_step"], "c5_ms": d["clusters_hydration"]["ms_per_step"]}
    if base is None: base = row
    eff = {k: (row[k] / base[k] / n if not k.endswith("_ms") else base[k] / row[k] / n) for k in row}
    print(f"N={n}", {k: f"{v:.4g}" for k, v in row.items()})
    print("   efficiency", {k: f"{v:.3f}" for k, v in eff.items()}, "sha", d["hist_sha256"][:12], d.get("nrank_equals_1rank"),
          "res sha", d["residence"]["cnt_sha256"][:12], "msd_last", d["msd"]["msd_last_frame"])
    r = d["residence"]; print("   residence", {k: round(r[k], 3) for k in ("search_ms", "exchange_ms", "correlation_ms")})
PY
