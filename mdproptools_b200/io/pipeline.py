"""Frame pipeline: dump text -> HBM, overlapped with compute.  Subsystem (a) of the north star.

With a GPU (default): reader threads read the files straight into pinned TEXT buffers, the text is copied to the device
on a copy stream and parsed there (csrc/dump_device.cu, the same exact number conversion as the host parser); the parsed
``[F, C, N]`` columns are copied back into a pinned host tensor on a third stream for the consumers that look at ids and
types on the host.  Reads of the next group of files, PCIe transfers, the parse kernel and the consumer's kernels overlap.

Without a GPU, for .gz files, dumps without an id column, or with MDP_DEVICE_PARSE=0: a producer thread parses frames
with the native host parser straight into a ring of pinned staging tensors ``[F, C, N]`` (ctypes releases the GIL while
the C++ parser runs) and the copy to the device is issued on a dedicated copy stream.
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
        tr = getattr(self, "_trace", None)
        if tr is not None:                                  # consumer timeline: its stream between two wait() calls
            e = tr.event()
            if tr.consumer_last is not None:
                tr.gpu_span("consumer kernels", tr.consumer_last[0], tr.consumer_last[1], e)
            tr.consumer_last = (self._trace_idx, e)
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


class _Trace:
    """Timeline of one pass of the text pipeline (MDP_PIPELINE_TRACE=<file>: one JSON line per pass is appended).  Host
    intervals are perf_counter times, device intervals pairs of CUDA events on the stream that did the work; everything is
    reported in milliseconds from the start of the pass.  Measurement aid only: nothing reads it back."""

    def __init__(self, path):
        import time
        self.path, self.clock = path, time.perf_counter
        self.t0 = self.clock()
        self.base = torch.cuda.Event(enable_timing=True)
        self.base.record()
        self.host, self.gpu, self.lock = [], [], threading.Lock()

    def span(self, name, idx, t_start):
        with self.lock:
            self.host.append((name, idx, (t_start - self.t0) * 1e3, (self.clock() - self.t0) * 1e3))

    def event(self, stream=None):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream if stream is not None else torch.cuda.current_stream())
        return e

    def gpu_span(self, name, idx, e0, e1):
        with self.lock:
            self.gpu.append((name, idx, e0, e1))

    def dump(self):
        import json
        torch.cuda.synchronize()
        rows = [{"where": "host", "what": n, "batch": i, "t0_ms": round(a, 3), "t1_ms": round(b, 3)} for n, i, a, b in self.host]
        rows += [{"where": "gpu", "what": n, "batch": i, "t0_ms": round(self.base.elapsed_time(a), 3),
                  "t1_ms": round(self.base.elapsed_time(b), 3)} for n, i, a, b in self.gpu]
        rows.sort(key=lambda r: r["t0_ms"])
        with open(self.path, "a") as f:
            f.write(json.dumps({"total_ms": round((self.clock() - self.t0) * 1e3, 3), "spans": rows}) + "\n")


class _PinnedPool:
    """Process-wide pool of pinned staging buffers (flat uint8 tensors).  cudaHostAlloc costs ~0.2 ms per MB: a pass over a
    trajectory that allocated its text ring and host batches afresh spent more time pinning memory than reading files
    (measured: passes of 0.5 s next to passes of 0.08 s, depending on when Python had freed the previous pass's buffers).
    Buffers come back when a pass ends / the consumer moves on; at most ``cap`` bytes are kept."""

    GRAIN = 16 << 20

    def __init__(self, cap=6 << 30):
        self.cap, self.free, self.lock = cap, [], threading.Lock()

    def take(self, nbytes):
        need = max(1, -(-int(nbytes) // self.GRAIN)) * self.GRAIN
        with self.lock:
            best = None
            for k, b in enumerate(self.free):
                if b.numel() >= need and (best is None or b.numel() < self.free[best].numel()):
                    best = k
            if best is not None and self.free[best].numel() <= 4 * need:
                return self.free.pop(best)
        return torch.empty((need,), dtype=torch.uint8, pin_memory=True)

    def give(self, buf):
        with self.lock:
            if sum(b.numel() for b in self.free) + buf.numel() <= self.cap:
                self.free.append(buf)


_POOL = _PinnedPool()


def _file_groups(sizes, max_files, max_bytes):
    """Consecutive files in groups of at most ``max_files`` files and ``max_bytes`` bytes (a file larger than that is a
    group of its own): lists of indices into ``sizes``, in order, covering every file once."""
    groups, cur, cur_b = [], [], 0
    for k, sz in enumerate(sizes):
        if cur and (len(cur) >= max_files or cur_b + sz > max_bytes):
            groups.append(cur)
            cur, cur_b = [], 0
        cur.append(k)
        cur_b += sz
    if cur:
        groups.append(cur)
    return groups


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
        # the device is fixed HERE, in the caller's thread: torch's current device is per thread, and the producer thread
        # would otherwise start on device 0 whatever the caller (rank r of a multi-GPU job) has selected
        self.device = torch.device(device) if device is not None else None
        if self.to_device and (self.device is None or (self.device.type == "cuda" and self.device.index is None)):
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.frame_select = frame_select
        # file_shard=(rank, world): this rank READS only every world-th file (one frame per file assumed and checked: a file
        # with another number of frames sets multi_frame_seen and ends the iteration; the caller then falls back to
        # frame_select, where every rank reads everything).  total_frames = number of files.
        self.file_shard = file_shard
        self.multi_frame_seen = False
        self.nthreads = nthreads
        self.prefetch = prefetch
        self.total_frames = None   # known once iteration has finished
        # Device parse (the default with a GPU, since round 2): the TEXT goes to the device and is parsed there
        # (csrc/dump_device.cu) -- _produce_text below; frames the device parser refuses (numbers off the exact fast path,
        # irregular ids) are re-parsed by the host parser, and trajectories it cannot take at all (no id column, .gz files)
        # use the host pipeline.  MDP_DEVICE_PARSE=0 / device_parse=False select the host parser.
        if device_parse is None:
            device_parse = _os.environ.get("MDP_DEVICE_PARSE", "1") not in ("", "0")
        self.device_parse = bool(device_parse) and self.to_device
        self.device_parsed_frames = 0
        self.host_reparsed_frames = 0
        self._trace = None

    def _produce(self, q: "queue.Queue"):
        try:
            import ctypes

            from .. import _lib

            if self.to_device:
                torch.cuda.set_device(self.device)
            if self.device_parse:
                plan = self._text_plan()
                if plan is not None:
                    self._produce_text(q, *plan)
                    return
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

    # ---------------------------------------------------------------------------------------------------------------
    # text pipeline: file -> pinned text (reader threads) -> HBM -> k_dump_rows -> [F, C, N] on the device (+ a copy on the
    # host for the consumers that look at ids / types there)
    # ---------------------------------------------------------------------------------------------------------------
    def _text_plan(self):
        """(files, file indices) of this rank when the device parser can take the trajectory, else None (host parser):
        plain files whose first frame has an id column and every wanted column within the parser's limits."""
        files = _dump.dump_files(self.pattern)
        if not files or any(f.endswith(".gz") for f in files):
            return None
        import os
        if max(os.path.getsize(f) for f in files) > (1 << 30):      # a file is staged (and shipped) whole
            return None
        with open(files[0], "rb") as fh:
            head = fh.read(4096)
        a = head.find(b"ITEM: ATOMS")
        nl = head.find(b"\n", a) if a >= 0 else -1
        if nl < 0 or self._colsel(head[a + 11:nl].decode().split()) is None:
            return None
        findex = list(range(len(files)))
        if self.file_shard is not None:
            findex = findex[self.file_shard[0]::self.file_shard[1]]
        return [files[i] for i in findex], findex, len(files)

    def _colsel(self, cols):
        """Output slot of every file column (-1: not wanted), or None when the device parser does not apply."""
        want = self.columns
        if "id" not in cols or any(w not in cols for w in want) or len(want) > 16:
            return None
        colsel = [want.index(c) if c in want else -1 for c in cols]
        if max([cols.index("id")] + [k for k, v in enumerate(colsel) if v >= 0]) >= 64:
            return None
        return colsel

    @staticmethod
    def _host_buffer(F, C, n):
        """(pool buffer, pinned [F, C, n] float64 view of it); the buffer goes back to the pool when the consumer asks for
        the next batch."""
        raw = _POOL.take(F * C * n * 8)
        return raw, raw[: F * C * n * 8].view(torch.float64).view(F, C, n)

    def _produce_text(self, q, files, findex, nfiles_all):
        import os
        from concurrent.futures import ThreadPoolExecutor

        from .. import ops

        device = torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        copy_stream, d2h_stream = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        want, C = self.columns, len(self.columns)
        sizes = [os.path.getsize(f) for f in files]
        readers = int(os.environ.get("MDP_READERS", max(4, min(16, (3 * (os.cpu_count() or 4)) // 4))))
        # groups of consecutive files: one pinned text buffer, one H2D copy, (normally) one batch each.  Small groups keep
        # the stages busy together: the first batch is ready after one group's reads, and while group k is parsed groups
        # k+1 and k+2 are being read (64 MB of text = 20 C2 frames; MDP_TEXT_GROUP_MB for measurements)
        group_bytes = min(self.max_batch_bytes, int(os.environ.get("MDP_TEXT_GROUP_MB", "64")) << 20)
        groups = _file_groups(sizes, self.max_batch_frames, group_bytes)
        ahead = 2                                        # groups whose reads are in flight beyond the one being parsed
        nring = self.prefetch + 4 + ahead
        tring, tev = [None] * nring, [None] * nring      # pinned text buffers; the H2D copy that last read each of them
        tr = self._trace
        nbatch = [0]
        sharded = self.file_shard is not None
        state = {"next_index": 0}

        with ThreadPoolExecutor(max_workers=max(1, readers)) as ex:

            def submit(gi):
                """Start reading the files of group gi into its text buffer."""
                slot = gi % nring
                if tev[slot] is not None:
                    tev[slot].synchronize()
                    tev[slot] = None
                need = sum(sizes[k] for k in groups[gi])
                t_sub = tr.clock() if tr is not None else 0.0
                buf = tring[slot]
                if buf is None or buf.numel() < need:
                    if buf is not None:
                        _POOL.give(buf)
                    buf = tring[slot] = _POOL.take(need)
                view, addr, off, futs = memoryview(buf.numpy()), buf.data_ptr(), 0, []
                for k in groups[gi]:
                    futs.append(ex.submit(_dump.read_text_frames, files[k], view, addr, off, sizes[k]))
                    off += sizes[k]
                return slot, buf, view, futs, need, t_sub

            def enqueue_run(run, cols, text_d, view):
                """H2D is under way; parse one run of frames (equal atom count and columns) on the copy stream and start the
                copy of the parsed columns back to the host."""
                F, n = len(run), run[0][1].natoms
                missing = [w for w in want if w not in cols]
                if missing:
                    raise KeyError(f"column(s) {missing} not in dump file (has {cols})")
                h_full, h = self._host_buffer(F, C, n)
                colsel = self._colsel(cols)
                if colsel is None or n <= 0:                   # this run is not for the device parser: host parse, plain H2D
                    _dump.parse_frames([bytes(view[fr.begin:fr.end]) for _, fr in run], want, h.numpy(), self.nthreads)
                    with torch.cuda.stream(copy_stream):
                        dev = h.to(device, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                    self.host_reparsed_frames += F
                    return {"run": run, "h_full": h_full, "h": h, "dev": dev, "ev": ev, "st": None, "view": view}
                begin = [fr.rows for _, fr in run]
                end = [fr.end for _, fr in run]
                st_h = torch.empty((F, 2), dtype=torch.int64, pin_memory=True)
                with torch.cuda.stream(copy_stream):
                    begin_d = torch.tensor(begin, dtype=torch.int64).to(device)
                    end_d = torch.tensor(end, dtype=torch.int64).to(device)
                    dev = torch.empty((F, C, n), dtype=torch.float64, device=device)
                    seen = torch.empty((F, (n + 31) // 32), dtype=torch.int32, device=device)
                    status = torch.empty((F, 2), dtype=torch.int64, device=device)
                    e0 = tr.event(copy_stream) if tr is not None else None
                    ops.dump_parse_device(text_d, begin_d, end_d, max(e - b for b, e in zip(begin, end)), n, len(cols), colsel,
                                          cols.index("id"), dev, seen, status, stream=copy_stream)
                    parsed = torch.cuda.Event(enable_timing=tr is not None)
                    parsed.record(copy_stream)
                    if tr is not None:
                        tr.gpu_span("k_dump_rows", nbatch[0], e0, parsed)
                with torch.cuda.stream(d2h_stream):
                    d2h_stream.wait_event(parsed)
                    e0 = tr.event(d2h_stream) if tr is not None else None
                    st_h.copy_(status, non_blocking=True)
                    h.copy_(dev, non_blocking=True)
                    dev.record_stream(d2h_stream)
                    status.record_stream(d2h_stream)
                    ev = torch.cuda.Event(enable_timing=tr is not None)
                    ev.record(d2h_stream)
                    if tr is not None:
                        tr.gpu_span("D2H parsed columns", nbatch[0], e0, ev)
                nbatch[0] += 1
                return {"run": run, "h_full": h_full, "h": h, "dev": dev, "ev": ev, "st": st_h, "view": view, "k": nbatch[0] - 1}

            def enqueue(gi, sub):
                slot, buf, view, futs, need, t_sub = sub
                frames = []
                for k, fut in zip(groups[gi], futs):
                    _, frs = fut.result()
                    if tr is not None and k == groups[gi][-1]:
                        tr.span("file reads + header scans (reader threads)", gi, t_sub)
                    if sharded:
                        if len(frs) != 1:
                            raise _dump.MultiFrameFile(files[k])
                        frames.append((findex[k], frs[0]))
                        continue
                    for fr in frs:
                        idx = state["next_index"]
                        state["next_index"] += 1
                        if self.frame_select is None or self.frame_select(idx):
                            frames.append((idx, fr))
                if not frames:
                    return []
                with torch.cuda.stream(copy_stream):
                    e0 = tr.event(copy_stream) if tr is not None else None
                    text_d = buf[:need].to(device, non_blocking=True)
                    ev = torch.cuda.Event(enable_timing=tr is not None)
                    ev.record(copy_stream)
                    if tr is not None:
                        tr.gpu_span("H2D text", gi, e0, ev)
                tev[slot] = ev
                recs, i = [], 0
                while i < len(frames):
                    n, cols = frames[i][1].natoms, frames[i][1].columns
                    cap = max(1, min(self.max_batch_frames, self.max_batch_bytes // max(1, C * n * 8)))
                    j = i
                    while j < len(frames) and j - i < cap and frames[j][1].natoms == n and frames[j][1].columns == cols:
                        j += 1
                    recs.append(enqueue_run(frames[i:j], cols, text_d, view))
                    i = j
                return recs

            def finalize(rec):
                """Wait for the parse (this thread only), let the host parser redo what the device refused, hand over."""
                run, h, dev, ev = rec["run"], rec["h"], rec["dev"], rec["ev"]
                t_fin = tr.clock() if tr is not None else 0.0
                if rec["st"] is not None:
                    ev.synchronize()
                    st = rec["st"].numpy()
                    n = run[0][1].natoms
                    redo = [f for f in range(len(run)) if int(st[f, 0]) != n or int(st[f, 1]) != 0]
                    for f in redo:                               # nothing is approximated: the host parser decides
                        fr = run[f][1]
                        _dump.parse_frame(bytes(rec["view"][fr.begin:fr.end]), want, self.nthreads, out=h[f].numpy())
                    if redo:
                        with torch.cuda.stream(copy_stream):
                            for f in redo:
                                dev[f].copy_(h[f], non_blocking=True)
                            ev = torch.cuda.Event()
                            ev.record(copy_stream)
                        ev.synchronize()
                    self.device_parsed_frames += len(run) - len(redo)
                    self.host_reparsed_frames += len(redo)
                metas = [FrameMeta(idx, fr.timestep, fr.natoms, fr.box) for idx, fr in run]
                batch = Batch(metas, self.columns, h, dev, ev)
                batch._pool_buffer = rec["h_full"]
                if tr is not None:
                    batch._trace, batch._trace_idx = tr, rec.get("k", -1)
                    tr.span("producer: wait for the parse, check status", rec.get("k", -1), t_fin)
                    t_put = tr.clock()
                q.put(batch)
                if tr is not None:
                    tr.span("producer: blocked handing over (consumer busy)", rec.get("k", -1), t_put)

            subs = {}

            def run_groups():
                nsub, prev = 0, []
                try:
                    for gi in range(len(groups)):
                        while nsub < len(groups) and nsub <= gi + ahead:     # reads of the next groups run while this one is parsed
                            subs[nsub] = submit(nsub)
                            nsub += 1
                        cur = enqueue(gi, subs.pop(gi))                      # H2D + parse + D2H of this group: asynchronous
                        for rec in prev:                                     # ... while the previous group is checked and handed over
                            finalize(rec)
                        prev = cur
                except _dump.MultiFrameFile:
                    self.multi_frame_seen = True
                for rec in prev:
                    finalize(rec)

            try:
                run_groups()
            finally:
                # the text buffers go back to the pool once nothing reads them any more (reads in flight, H2D copies)
                for sub in subs.values():
                    for fut in sub[3]:
                        fut.cancel()
                ex.shutdown(wait=True)
                copy_stream.synchronize()
                d2h_stream.synchronize()
                for k, buf in enumerate(tring):
                    if buf is not None:
                        _POOL.give(buf)
                        tring[k] = None
        self.total_frames = nfiles_all if sharded else state["next_index"]
        q.put(None)

    def __iter__(self):
        q: "queue.Queue" = _StoppableQueue(self.prefetch)
        import os
        if self.device_parse and os.environ.get("MDP_PIPELINE_TRACE"):
            self._trace = _Trace(os.environ["MDP_PIPELINE_TRACE"])
            self._trace.consumer_last = None
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
                buf = getattr(item, "_pool_buffer", None)     # the consumer is done with this batch's host copy
                if buf is not None:
                    item._pool_buffer = None
                    _POOL.give(buf)
        finally:
            # the consumer may leave early (an exception in its loop body, a generator that is dropped): tell the producer to
            # stop, unblock its pending put, and wait for it -- otherwise the thread would sit in q.put() for ever, holding
            # the pinned staging ring, the device batches and the read-ahead pool
            q.stop()
            th.join()
            if self._trace is not None:
                tr, self._trace = self._trace, None
                if tr.consumer_last is not None:
                    tr.gpu_span("consumer kernels", tr.consumer_last[0], tr.consumer_last[1], tr.event())
                tr.dump()


class ArrayBatches:
    """Stream an in-memory host trajectory ``pos`` [T, C, N] (pinned torch tensor for truly asynchronous copies) to the
    device in batches of frames through TWO persistent device buffers: the copy of batch k+1 runs on a dedicated copy
    stream while the kernels of batch k run on the current stream, and a buffer is refilled only after the kernels that
    read it have finished (event recorded by ``done``).  No allocation happens inside the loop.

        for f0, f1, x in (ab := ArrayBatches(pos, batch_frames)):
            ... launch kernels reading x on the current stream ...
            ab.done()
    """

    def __init__(self, pos: torch.Tensor, batch_frames: int, device=None, ramp: bool = True):
        self.pos = pos
        self.T = pos.shape[0]
        self.nb = max(1, min(int(batch_frames), self.T))
        # ramp: the first batch is an eighth of batch_frames and the size doubles from batch to batch -- the kernels start
        # after a short first copy, and the bulk of the trajectory still goes through in large batches (few launches)
        self.first = max(1, int(batch_frames) // 8) if ramp else self.nb
        self.first = min(self.first, self.nb)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        shape = (self.nb,) + tuple(pos.shape[1:])
        self.bufs = [torch.empty(shape, dtype=pos.dtype, device=self.device) for _ in range(2 if self.T > self.first else 1)]
        self.consumed = [None] * len(self.bufs)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.main = torch.cuda.current_stream(self.device)
        self._k = None

    def _stage(self, f0, i):
        """Start the copy of batch number i (frames from f0 on) into buffer i % 2."""
        f1 = min(self.T, f0 + min(self.nb, self.first << min(i, 30)))
        k = i % len(self.bufs)
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[k] is not None:
                self.copy_stream.wait_event(self.consumed[k])
            else:
                self.copy_stream.wait_stream(self.main)      # the buffers were allocated on the current stream
            x = self.bufs[k][: f1 - f0]
            x.copy_(self.pos[f0:f1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return f0, f1, x, ev, k, i

    def __iter__(self):
        nxt = self._stage(0, 0)
        while nxt is not None:
            f0, f1, x, ev, k, i = nxt
            nxt = self._stage(f1, i + 1) if f1 < self.T else None
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
