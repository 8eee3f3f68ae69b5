"""Host-side profile of the RDF array front end (where the e2e step time goes beyond the kernels)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mdproptools_b200.structural import rdf_cn

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
F = 64
frames = bench.make_frames(F, bench.SEED, "cuda")
host = torch.empty((F, 3, bench.N_ATOMS), dtype=torch.float64, pin_memory=True)
host.copy_(frames)
types = np.ones(bench.N_ATOMS)
L = bench.lattice_lengths()
for _ in range(2):
    rdf_cn.calc_atomic_rdf_from_arrays(host, types, L, bench.R_CUT, bench.BIN, [[1], [1]], batch_frames=16)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    rdf_cn.calc_atomic_rdf_from_arrays(host, types, L, bench.R_CUT, bench.BIN, [[1], [1]], batch_frames=16)
torch.cuda.synchronize()
print("ms per call", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    rdf_cn.calc_atomic_rdf_from_arrays(host, types, L, bench.R_CUT, bench.BIN, [[1], [1]], batch_frames=16)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
