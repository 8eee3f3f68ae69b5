#!/bin/bash
# Timeline of calc_atomic_rdf(filename=<256 C2-sized dump files>) through the text pipeline (MDP_PIPELINE_TRACE) + a
# cProfile of the calling thread.  Writes gpurun_out/files_timeline.json and gpurun_out/files_timeline.txt.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/files_timeline.json
python - <<'PY'
import os, sys, time, shutil, tempfile, json, cProfile, pstats, io
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from mdproptools_b200.structural import rdf_cn
torch.cuda.set_device(0)
frames = bench.make_frames(32, bench.SEED, "cuda")
d = tempfile.mkdtemp(prefix="mdp_trace_")
rng = np.random.default_rng(1)
host = frames.cpu().numpy()
for f in range(32):
    ids = rng.permutation(bench.N_ATOMS) + 1
    x, y, z = host[f][:, ids - 1]
    body = "\n".join(["%d 1 %g %g %g" % t for t in zip(ids.tolist(), x.tolist(), y.tolist(), z.tolist())])
    open(os.path.join(d, f"dump.c2.{f}.dump"), "w").write(
        f"ITEM: TIMESTEP\n{f}\nITEM: NUMBER OF ATOMS\n{bench.N_ATOMS}\nITEM: BOX BOUNDS pp pp pp\n0.0 167.19\n0.0 167.19\n0.0 167.19\nITEM: ATOMS id type x y z\n" + body + "\n")
for c in range(1, 8):
    for f in range(32):
        shutil.copy(os.path.join(d, f"dump.c2.{f}.dump"), os.path.join(d, f"dump.c2.{c * 32 + f}.dump"))
pat = os.path.join(d, "dump.c2.*.dump")
def run():
    t = time.perf_counter()
    rdf_cn.calc_atomic_rdf(20, 0.05, 1, [39.9], [[1], [1]], pat, save_mode=False)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3
run(); run()
os.environ["MDP_PIPELINE_TRACE"] = "gpurun_out/files_timeline.json"
ms = run()
del os.environ["MDP_PIPELINE_TRACE"]
tl = json.loads(open("gpurun_out/files_timeline.json").read().strip().splitlines()[-1])
lines = [f"calc_atomic_rdf on 256 files of {bench.N_ATOMS} atoms: {ms:.1f} ms wall ({ms / 256:.3f} ms/frame); pipeline pass {tl['total_ms']:.1f} ms", ""]
def union(iv):
    iv = sorted(iv); tot, cur_a, cur_b = 0.0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None: tot += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    return tot + (cur_b - cur_a if cur_b is not None else 0.0)
kinds = {}
for s in tl["spans"]:
    kinds.setdefault((s["where"], s["what"]), []).append((s["t0_ms"], s["t1_ms"]))
lines.append(f"{'where':5s} {'what':52s} {'n':>4s} {'sum ms':>9s} {'busy ms':>9s} {'first':>8s} {'last':>8s}")
for (w, k), iv in sorted(kinds.items(), key=lambda kv: min(a for a, _ in kv[1])):
    lines.append(f"{w:5s} {k:52s} {len(iv):4d} {sum(b - a for a, b in iv):9.2f} {union(iv):9.2f} {min(a for a, _ in iv):8.2f} {max(b for _, b in iv):8.2f}")
lines.append("")
lines.append("per batch (ms from the start of the pass):")
byb = {}
for s in tl["spans"]:
    byb.setdefault(s["batch"], []).append(s)
for b in sorted(byb):
    lines.append(f"  batch {b}: " + "; ".join(f"{s['what'].split(' (')[0]} {s['t0_ms']:.1f}-{s['t1_ms']:.1f}" for s in sorted(byb[b], key=lambda s: s['t0_ms'])))
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
buf = io.StringIO(); pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(14)
lines.append(""); lines.append("cProfile of the calling thread (one more pass):"); lines.append(buf.getvalue())
open("gpurun_out/files_timeline.txt", "w").write("\n".join(lines))
print("\n".join(lines))
shutil.rmtree(d)
PY
