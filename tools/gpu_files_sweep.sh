#!/bin/bash
# file-based RDF entry point: batch size / reader thread sweep + a cProfile of one pass
set -u
cd "$(dirname "$0")/.."
python - <<'PY'
import os, sys, time, shutil, tempfile, cProfile, pstats
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
def run():
    t = time.perf_counter()
    rdf_cn.calc_atomic_rdf(20, 0.05, 1, [39.9], [[1], [1]], pat, save_mode=False)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / 256 * 1e3
run()
for mb in (64, 128, 256):
    for rd in (4, 8, 12):
        os.environ["MDP_BATCH_MB"] = str(mb); os.environ["MDP_READERS"] = str(rd)
        print(f"batch {mb} MB readers {rd}: {min(run(), run()):.3f} ms/frame", flush=True)
os.environ["MDP_BATCH_MB"] = "128"; os.environ["MDP_READERS"] = "8"
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
shutil.rmtree(d)
PY
