"""Host-side profile of the file-based RDF entry point on C2-sized dumps (where the per-frame time goes)."""
import cProfile, pstats, sys, os, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mdproptools_b200.structural import rdf_cn

torch.cuda.set_device(0)
frames = bench.make_frames(16, bench.SEED, "cuda")
r = bench.bench_rdf_from_files(torch, frames, bench.N_ATOMS * (bench.N_ATOMS - 1) / 2, nfiles=16)
print({k: v for k, v in r.items() if k not in ("api", "note")})
# profile one more pass
import shutil
d = tempfile.mkdtemp()
rng = np.random.default_rng(1)
host = frames.cpu().numpy()
for f in range(16):
    ids = rng.permutation(bench.N_ATOMS) + 1
    x, y, z = host[f][:, ids - 1]
    body = "\n".join(["%d 1 %g %g %g" % t for t in zip(ids.tolist(), x.tolist(), y.tolist(), z.tolist())])
    open(os.path.join(d, f"dump.c2.{f}.dump"), "w").write(
        f"ITEM: TIMESTEP\n{f}\nITEM: NUMBER OF ATOMS\n{bench.N_ATOMS}\nITEM: BOX BOUNDS pp pp pp\n0.0 167.19\n0.0 167.19\n0.0 167.19\nITEM: ATOMS id type x y z\n" + body + "\n")
pat = os.path.join(d, "dump.c2.*.dump")
rdf_cn.calc_atomic_rdf(20, 0.05, 1, [39.9], [[1], [1]], pat, save_mode=False)
pr = cProfile.Profile(); pr.enable()
rdf_cn.calc_atomic_rdf(20, 0.05, 1, [39.9], [[1], [1]], pat, save_mode=False)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
shutil.rmtree(d)
