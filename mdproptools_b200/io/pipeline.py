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


class _ProducerStopped(BaseException):
    """Raised inside the producer thread when the consumer has gone away."""


class _StoppableQueue:
    """queue.Queue whose put() gives up (raises _ProducerStopped in the producer) once stop() was called."""

    def __init__(self, maxsize):
        self._q = queue.Queue(maxsize=maxsize)
        self._stop = threading.Event()

    def put(self, item):
        while True:
            if self._stop.is_set():
                raise _ProducerStopped()
            try:
                self._q.put(item, timeout=0.05)
                return
            except queue.Full:
                continue

    def get(self):
        return self._q.get()

    def stop(self):
        self._stop.set()
        try:
            while True:                       # drain: a producer blocked in put() gets its slot and then sees the flag
                self._q.get_nowait()
        except queue.Empty:
            pass


def _guard_multi(source, owner):
    """Pass (index, text) pairs through; a MultiFrameFile from a sharded read ends the stream and is recorded."""
    try:
        yield from source
    except _dump.MultiFrameFile:
        owner.multi_frame_seen = True


class FrameBatches:
    """Iterate over a trajectory in batches of frames with equal atom count.

    frame_slice: optional (start, stop, step) applied to the global frame index (used for frame sharding
    across ranks and for ``get_clusters(frame=...)``); frames outside it are skipped without being parsed.
    """

    def __init__(self, pattern, columns, max_batch_bytes=256 << 20, max_batch_frames=256, to_device=True, device=None,
                 frame_select=None, nthreads=0, prefetch=2, device_parse=None, file_shard=None):
        import os as _os
        self.pattern = pattern
        self.columns = list(columns)
        if _os.environ.get("MDP_BATCH_MB"):                    # tuning knob for measurements
            max_batch_bytes = int(_os.environ["MDP_BATCH_MB"]) << 20
        self.max_batch_bytes = int(max_batch_bytes)
        self.max_batch_frames = int(max_batch_frames)
        self.cuda = torch.cuda.is_available()
        self.to_device = to_device and self.cuda
        self.device = device
        self.frame_select = frame_select
        # file_shard=(rank, world): this rank READS only every world-th file (one frame per file assumed and checked: a file
        # with another number of frames sets multi_frame_seen and ends the iteration; the caller then falls back to
        # frame_select, where every rank reads everything).  total_frames = number of files.
        self.file_shard = file_shard
        self.multi_frame_seen = False
        self.nthreads = nthreads
        self.prefetch = prefetch
        self.total_frames = None   # known once iteration has finished
        # EXPERIMENTAL, off by default: ship the TEXT to the device and parse it there (csrc/dump_device.cu); frames the
        # device parser refuses (numbers off the exact fast path, irregular ids) are re-parsed by the host parser
        if device_parse is None:
            import os
            device_parse = os.environ.get("MDP_DEVICE_PARSE", "0") not in ("", "0")
        self.device_parse = bool(device_parse) and self.to_device
        self.device_parsed_frames = 0
        self.host_reparsed_frames = 0

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
                if self.device_parse:
                    got = self._flush_device([b for _, b in pending], [i for i, _ in pending], h, copy_stream)
                    if got is not None:
                        metas, dev, ev = got
                        ring_ev[slot] = ev
                        q.put(Batch(metas, self.columns, h, dev, ev))
                        pending = []
                        return
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
            if self.file_shard is not None:
                source = _dump.iter_frame_buffers(self.pattern, file_shard=self.file_shard)
            else:
                source = ((None, b) for b in _dump.iter_frame_buffers(self.pattern))
            sharded_total = len(_dump.dump_files(self.pattern)) if self.file_shard is not None else None
            for fidx, buf in _guard_multi(source, self):
                idx = idx + 1 if fidx is None else fidx
                if fidx is None and self.frame_select is not None and not self.frame_select(idx):
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
            self.total_frames = idx + 1 if sharded_total is None else sharded_total
            flush()
            q.put(None)
        except _ProducerStopped:
            return                        # the consumer left; nothing to forward
        except BaseException as exc:  # noqa: BLE001 - forwarded to the consumer
            try:
                q.put(exc)
            except _ProducerStopped:
                pass

    def _flush_device(self, bufs, indices, h, copy_stream):
        """Device parse of one batch (EXPERIMENTAL): text -> pinned bytes -> H2D -> k_dump_rows -> D2H of the parsed SoA
        into ``h`` (the consumers read types and ids on the host).  Returns (metas, dev, event), or None when the batch
        does not qualify (no id column, too many columns) and the host parser should take it."""
        import ctypes

        from .. import _lib, ops

        cols = _dump.frame_columns(bufs[0])
        want = self.columns
        if "id" not in cols or any(w not in cols for w in want) or len(want) > 16:
            return None
        colsel = [want.index(c) if c in want else -1 for c in cols]
        if max([cols.index("id")] + [k for k, v in enumerate(colsel) if v >= 0]) >= 64:
            return None
        F, C, n = h.shape
        hdr = (ctypes.c_double * 16)()
        metas, begin, end, off = [], [], [], 0
        for idx, b in zip(indices, bufs):
            _lib.check(_lib.lib().mdp_dump_header(b, len(b), hdr, None, 0), "mdp_dump_header")
            tric = hdr[11] != 0.0
            box = _dump.Box([[hdr[2], hdr[3]], [hdr[4], hdr[5]], [hdr[6], hdr[7]]], [hdr[8], hdr[9], hdr[10]] if tric else None)
            metas.append(FrameMeta(idx, int(hdr[0]), int(hdr[1]), box))
            a = b.find(b"ITEM: ATOMS")
            begin.append(off + b.find(b"\n", a) + 1)
            end.append(off + len(b))
            off += len(b)
        text = torch.empty((off,), dtype=torch.uint8, pin_memory=True)
        tv = text.numpy()
        pos = 0
        for b in bufs:
            tv[pos:pos + len(b)] = np.frombuffer(b, dtype=np.uint8)
            pos += len(b)
        device = self.device or torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.stream(copy_stream):
            text_d = text.to(device, non_blocking=True)
            begin_d = torch.tensor(begin, dtype=torch.int64).to(device)
            end_d = torch.tensor(end, dtype=torch.int64).to(device)
            dev = torch.empty((F, C, n), dtype=torch.float64, device=device)
            seen = torch.empty((F, (n + 31) // 32), dtype=torch.int32, device=device)
            status = torch.empty((F, 2), dtype=torch.int64, device=device)
            ops.dump_parse_device(text_d, begin_d, end_d, max(e - b0 for b0, e in zip(begin, end)), n, len(cols), colsel,
                                  cols.index("id"), dev, seen, status, stream=copy_stream)
            st = status.cpu()                                   # synchronises the copy stream: the parse has finished
            redo = [f for f in range(F) if int(st[f, 0]) != n or int(st[f, 1]) != 0]
            good = [f for f in range(F) if f not in set(redo)]
            if not redo:
                h.copy_(dev, non_blocking=True)
            else:
                for f in good:
                    h[f].copy_(dev[f], non_blocking=True)
            for f in redo:                                      # nothing is approximated: the host parser decides
                _dump.parse_frame(bufs[f], want, self.nthreads, out=h[f].numpy())
                dev[f].copy_(h[f], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        copy_stream.synchronize()                               # h is complete before the batch is handed over
        self.device_parsed_frames += len(good)
        self.host_reparsed_frames += len(redo)
        return metas, dev, ev

    def __iter__(self):
        q: "queue.Queue" = _StoppableQueue(self.prefetch)
        th = threading.Thread(target=self._produce, args=(q,), daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            # the consumer may leave early (an exception in its loop body, a generator that is dropped): tell the producer to
            # stop, unblock its pending put, and wait for it -- otherwise the thread would sit in q.put() for ever, holding
            # the pinned staging ring, the device batches and the read-ahead pool
            q.stop()
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
