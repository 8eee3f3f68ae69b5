#!/bin/bash
# file-based RDF entry point: device parser (text pipeline) against the host parser, batch size / reader thread sweep
set -u
cd "$(dirname "$0")/.."
python - <<'PY'
import os, sys, time, shutil, tempfile
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from mdproptools_b200.structural import rdf_cn
torch.cuda.set_device(0)
frames = bench.make_frames(32, bench.SEED, "cuda")
d = tempfile.mkdtemp(prefix="mdp_sweep_")
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
out = {}
def run():
    t = time.perf_counter()
    df = rdf_cn.calc_atomic_rdf(20, 0.05, 1, [39.9], [[1], [1]], pat, save_mode=False)
    torch.cuda.synchronize()
    out["df"] = df
    return (time.perf_counter() - t) / 256 * 1e3
os.environ["MDP_DEVICE_PARSE"] = "0"
run()
print(f"host parser (default batch/readers): {min(run(), run()):.3f} ms/frame", flush=True)
ref = out["df"].values.copy()
os.environ["MDP_DEVICE_PARSE"] = "1"
run()
print("device parser result identical to host parser result:", bool(np.array_equal(ref, out["df"].values)))
print(f"device parser (default batch/readers): {min(run(), run()):.3f} ms/frame", flush=True)
for mb in (32, 64, 128, 256):
    for rd in (8, 12, 16):
        os.environ["MDP_TEXT_GROUP_MB"] = str(mb); os.environ["MDP_READERS"] = str(rd)
        run()
        print(f"device parser, text groups of {mb} MB, readers {rd}: {min(run(), run(), run()):.3f} ms/frame", flush=True)
shutil.rmtree(d)
PY
