"""Frame pipeline: dump text -> pinned SoA host buffers -> HBM on a side stream, overlapped with compute.

Subsystem (a) of the north star.  A reader thread parses frames with the native parser straight into a
ring of pinned staging tensors ``[F, C, N]`` (ctypes releases the GIL while the C++ parser runs), the copy
to the device is issued on a dedicated copy stream and the consumer only waits on the copy's event, so
parsing, PCIe transfer and kernels of consecutive batches overlap.
"""
from __future__ import annotations

import queue
import threading
from dataclasses import dataclass

import numpy as np
import torch

from . import dump as _dump


@dataclass
class FrameMeta:
    index: int               # position in the trajectory (pymatgen order)
    timestep: int
    natoms: int
    box: _dump.Box


@dataclass
class Batch:
    metas: list              # FrameMeta per frame
    columns: list            # column names, order of axis 1
    host: torch.Tensor       # [F, C, N] float64 (pinned when CUDA is available)
    dev: torch.Tensor | None # [F, C, N] on the device (None when to_device=False)
    ready: object | None     # torch.cuda.Event recorded after the H2D copy

    def wait(self):
        if self.ready is not None:
            cur = torch.cuda.current_stream()
            cur.wait_event(self.ready)
            if self.dev is not None:
                # the batch was allocated under the copy stream: tell the caching allocator that this stream reads it too,
                # so that the block is not handed to a later copy while kernels queued here are still using it
                self.dev.record_stream(cur)
        return self.dev

    def col(self, name):
        return self.columns.index(name)


class FrameBatches:
    """Iterate over a trajectory in batches of frames with equal atom count.

    frame_slice: optional (start, stop, step) applied to the global frame index (used for frame sharding
    across ranks and for ``get_clusters(frame=...)``); frames outside it are skipped without being parsed.
    """

    def __init__(self, pattern, columns, max_batch_bytes=64 << 20, max_batch_frames=256, to_device=True, device=None,
                 frame_select=None, nthreads=0, prefetch=2):
        self.pattern = pattern
        self.columns = list(columns)
        self.max_batch_bytes = int(max_batch_bytes)
        self.max_batch_frames = int(max_batch_frames)
        self.cuda = torch.cuda.is_available()
        self.to_device = to_device and self.cuda
        self.device = device
        self.frame_select = frame_select
        self.nthreads = nthreads
        self.prefetch = prefetch
        self.total_frames = None   # known once iteration has finished

    def _produce(self, q: "queue.Queue"):
        try:
            import ctypes

            from .. import _lib

            copy_stream = torch.cuda.Stream(device=self.device) if self.to_device else None
            C = len(self.columns)
            pending = []                     # (trajectory index, frame text) of the batch being collected
            cur_n, cap = None, 0
            # pinned staging buffers are reused round-robin: a buffer comes up again only after prefetch + 2 newer batches
            # have been handed over, i.e. after the consumer has asked for the batch that followed it
            ring, ring_k = [None] * (self.prefetch + 3), 0
            ring_ev = [None] * len(ring)     # H2D copy that last read each buffer: complete before the buffer is refilled

            def flush():
                nonlocal pending, ring_k
                if not pending:
                    return
                F = len(pending)
                slot = ring_k % len(ring)
                ring_k += 1
                buf = ring[slot]
                if ring_ev[slot] is not None:
                    ring_ev[slot].synchronize()
                    ring_ev[slot] = None
                if buf is None or buf.shape[1:] != (C, cur_n) or buf.shape[0] < F:
                    buf = ring[slot] = torch.empty((max(F, cap), C, cur_n), dtype=torch.float64, pin_memory=self.cuda)
                h = buf[:F]
                frames = _dump.parse_frames([b for _, b in pending], self.columns, h.numpy(), self.nthreads)
                metas = [FrameMeta(idx, fr.timestep, fr.natoms, fr.box) for (idx, _), fr in zip(pending, frames)]
                dev, ev = None, None
                if self.to_device:
                    with torch.cuda.stream(copy_stream):
                        dev = h.to(self.device or "cuda", non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                    ring_ev[slot] = ev
                q.put(Batch(metas, self.columns, h, dev, ev))
                pending = []

            idx = -1
            hdr = (ctypes.c_double * 16)()
            for buf in _dump.iter_frame_buffers(self.pattern):
                idx += 1
                if self.frame_select is not None and not self.frame_select(idx):
                    continue
                # peek at natoms: a batch holds frames of one size
                _lib.check(_lib.lib().mdp_dump_header(buf, len(buf), hdr, None, 0), "mdp_dump_header")
                n = int(hdr[1])
                if pending and (n != cur_n or len(pending) >= cap):
                    flush()
                if not pending:
                    cur_n = n
                    cap = max(1, min(self.max_batch_frames, self.max_batch_bytes // max(1, C * n * 8)))
                pending.append((idx, buf))
            self.total_frames = idx + 1
            flush()
            q.put(None)
        except BaseException as exc:  # noqa: BLE001 - forwarded to the consumer
            q.put(exc)

    def __iter__(self):
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        th = threading.Thread(target=self._produce, args=(q,), daemon=True)
        th.start()
        while True:
            item = q.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            yield item
        th.join()


class ArrayBatches:
    """Stream an in-memory host trajectory ``pos`` [T, C, N] (pinned torch tensor for truly asynchronous copies) to the
    device in batches of frames through TWO persistent device buffers: the copy of batch k+1 runs on a dedicated copy
    stream while the kernels of batch k run on the current stream, and a buffer is refilled only after the kernels that
    read it have finished (event recorded by ``done``).  No allocation happens inside the loop.

        for f0, f1, x in (ab := ArrayBatches(pos, batch_frames)):
            ... launch kernels reading x on the current stream ...
            ab.done()
    """

    def __init__(self, pos: torch.Tensor, batch_frames: int, device=None):
        self.pos = pos
        self.T = pos.shape[0]
        self.nb = max(1, min(int(batch_frames), self.T))
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        shape = (self.nb,) + tuple(pos.shape[1:])
        self.bufs = [torch.empty(shape, dtype=pos.dtype, device=self.device) for _ in range(2 if self.T > self.nb else 1)]
        self.consumed = [None] * len(self.bufs)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.main = torch.cuda.current_stream(self.device)
        self._k = None

    def _stage(self, f0, k):
        f1 = min(self.T, f0 + self.nb)
        k %= len(self.bufs)
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[k] is not None:
                self.copy_stream.wait_event(self.consumed[k])
            else:
                self.copy_stream.wait_stream(self.main)      # the buffers were allocated on the current stream
            x = self.bufs[k][: f1 - f0]
            x.copy_(self.pos[f0:f1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return f0, f1, x, ev, k

    def __iter__(self):
        nxt = self._stage(0, 0)
        while nxt is not None:
            f0, f1, x, ev, k = nxt
            nxt = self._stage(f1, k + 1) if f1 < self.T else None
            self.main.wait_event(ev)
            self._k = k
            yield f0, f1, x
            if self._k is not None:
                self.done()
        self.main.wait_stream(self.copy_stream)

    def done(self):
        """Call after the last kernel reading the current batch has been launched."""
        if self._k is not None:
            self.consumed[self._k] = torch.cuda.Event()
            self.consumed[self._k].record(self.main)
            self._k = None
