#!/bin/bash
# Device dump parser (csrc/dump_device.cu): kernel throughput on C2-sized text already in HBM, and the raw costs of the
# stages a text pipeline would have (page cache -> pinned, pinned -> device).
set -u
cd "$(dirname "$0")/.."
python - <<'PY'
import os, sys, time, tempfile, shutil, threading
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from mdproptools_b200 import ops
torch.cuda.set_device(0)
F = 16
frames = bench.make_frames(F, bench.SEED, "cuda")
host = frames.cpu().numpy()
rng = np.random.default_rng(1)
N = bench.N_ATOMS
texts = []
for f in range(F):
    ids = rng.permutation(N) + 1
    x, y, z = host[f][:, ids - 1]
    texts.append(("\n".join(["%d 1 %g %g %g" % t for t in zip(ids.tolist(), x.tolist(), y.tolist(), z.tolist())]) + "\n").encode())
begin, end, off = [], [], 0
for t in texts:
    begin.append(off); off += len(t); end.append(off)
blob = b"".join(texts)
print("text bytes per frame", len(blob) / F)
pin = torch.empty((len(blob),), dtype=torch.uint8, pin_memory=True)
pin.numpy()[:] = np.frombuffer(blob, dtype=np.uint8)
text_d = pin.cuda()
begin_d, end_d = torch.tensor(begin).cuda(), torch.tensor(end).cuda()
cols = ["id", "type", "x", "y", "z"]
want = ["id", "type", "x", "y", "z"]
colsel = [want.index(c) for c in cols]
out = torch.empty((F, len(want), N), dtype=torch.float64, device="cuda")
seen = torch.empty((F, (N + 31) // 32), dtype=torch.int32, device="cuda")
status = torch.empty((F, 2), dtype=torch.int64, device="cuda")
longest = max(e - b for b, e in zip(begin, end))
def parse():
    ops.dump_parse_device(text_d, begin_d, end_d, longest, N, len(cols), colsel, 0, out, seen, status)
parse(); torch.cuda.synchronize()
print("status", status[:2].tolist())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): parse()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"k_dump_rows (+memsets): {ms:.3f} ms per {F} frames = {ms / F * 1e3:.1f} us/frame, {len(blob) / ms / 1e6:.1f} GB/s of text")
# check against the frames
perm_ok = torch.allclose(out[:, 2:5, :], frames[:, :, :].to(torch.float64), rtol=1e-5, atol=1e-4)
print("values agree with the source frames to %g precision:", bool(perm_ok))
# pinned -> device
e0.record()
for _ in range(10): text_d.copy_(pin, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"H2D of the text: {ms:.3f} ms per {F} frames, {len(blob) / ms / 1e6:.1f} GB/s")
# page cache -> pinned with reader threads
d = tempfile.mkdtemp(prefix="mdp_dp_")
paths = []
for f in range(F):
    p = os.path.join(d, f"t{f}.txt"); open(p, "wb").write(texts[f]); paths.append(p)
pv = memoryview(pin.numpy())
def rd(k, nthr):
    for f in range(k, F, nthr):
        with open(paths[f], "rb", buffering=0) as fh:
            fh.readinto(pv[begin[f]:end[f]])
for nthr in (1, 4, 8, 16):
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        th = [threading.Thread(target=rd, args=(k, nthr)) for k in range(nthr)]
        [x.start() for x in th]; [x.join() for x in th]
        best = min(best, time.perf_counter() - t)
    print(f"page cache -> pinned, {nthr} threads: {best * 1e3:.2f} ms per {F} frames, {len(blob) / best / 1e9:.1f} GB/s")
shutil.rmtree(d)
PY
