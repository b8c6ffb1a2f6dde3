"""Many blocks in flight from ONE host thread.

`to_instance_seg` (vote_instances.py:150-452 in the reference) is a chain of device
stages separated by points where the host needs a number from the device (how many
rows, how many patches were selected, ...).  Run block after block, every such point
drains the GPU; run from several host threads, the threads fight over the interpreter.
Here the chain of one block is a GENERATOR that enqueues its kernels on its own CUDA
stream and yields an event (or a host future) whenever it needs a result; the scheduler
resumes whichever block has its result ready.  The GPU always has the kernels of many
blocks queued, the host never blocks while another block could make progress, and the
sync points per block drop from eight to four (sizes travel together through one pinned
buffer per block).

Same C-ABI calls, same order, same results as BlockAssembler driven by to_instance_seg
(tests/test_sharded.py compares them).
"""
import collections
import concurrent.futures

import numpy as np

from . import cuda_code as cc
from .assembly import BlockAssembler, RowSource


def _torch():
    import torch
    return torch


_PINNED = []
_PINNED_BYTES = {}


def _pinned(n=8):
    """small pinned int64 buffers, recycled (cudaHostAlloc per block would cost more
    than the kernels it serves)."""
    torch = _torch()
    if _PINNED:
        return _PINNED.pop()
    return torch.empty(n, dtype=torch.int64).pin_memory()


class _PinnedBuf:
    """pinned staging memory for a device->host copy that must not block the host
    thread (a pageable destination would); power-of-two size classes, recycled."""

    def __init__(self, dtype, n):
        torch = _torch()
        nbytes = max(256, int(n) * torch.empty(0, dtype=dtype).element_size())
        self.cls = 1 << (nbytes - 1).bit_length()
        free = _PINNED_BYTES.setdefault(self.cls, [])
        self.raw = free.pop() if free else torch.empty(self.cls, dtype=torch.uint8).pin_memory()
        self.t = self.raw[:int(n) * torch.empty(0, dtype=dtype).element_size()].view(dtype)

    def release(self):
        _PINNED_BYTES[self.cls].append(self.raw)
        self.raw = self.t = None


def _take(mask, n):
    """indices of the first n set entries of a bool tensor, in order, without a host
    sync (n is already known on the host)."""
    torch = _torch()
    return torch.nonzero_static(mask, size=int(n)).flatten()


def block_steps(src, fg, mask, numinst, patchshape, pool, **kwargs):
    """generator form of to_instance_seg(..., return_intermediates=True) for a compact
    row block.  Yields torch.cuda.Event / concurrent.futures.Future objects to wait
    for; returns (pairs u32 [n,6] block coordinates, aff f32 [n]) or None through
    StopIteration.value."""
    torch = _torch()
    assert isinstance(src, RowSource)
    ps = np.asarray(patchshape)
    rad = ps // 2
    shape = src.shape
    stream = torch.cuda.current_stream()
    overlap = (numinst > 1).to(torch.uint8)
    mask = mask.clone()
    mask[overlap > 0] = 0                                        # vote_instances.py:226
    radslice = tuple(slice(int(rad[i]), shape[i] - int(rad[i])) for i in range(3))
    asm = BlockAssembler(src, fg, overlap, ps, **kwargs)
    V = asm.V
    dev = asm.dev
    # ---- stage A: gate + compaction; sizes -> host ---------------------------------
    asm.flags = torch.empty(V, dtype=torch.uint8, device=dev)
    cc.call('ppp_gate_rows', cc.ptr(src.patches), cc.ptr(src.vox2row), cc.ptr(asm.overlap),
            cc.ptr(asm.foreground), asm.cfg, cc.ptr(asm.flags), asm.stream)
    asm.fgidx = torch.empty(V, dtype=torch.int32, device=dev)
    rowvox = torch.empty(V, dtype=torch.int32, device=dev)
    nrows = torch.zeros(1, dtype=torch.int64, device=dev)
    scratch = torch.empty(cc.call('ppp_compact_scratch_bytes', V), dtype=torch.uint8, device=dev)
    cc.call('ppp_compact', cc.ptr(asm.flags), V, cc.ptr(asm.fgidx), cc.ptr(rowvox), cc.ptr(nrows),
            cc.ptr(scratch), asm.stream)
    n_mask = torch.count_nonzero(mask[radslice]).reshape(1)
    candm = (asm.flags & 24) == 24                               # CAND | INTERIOR
    sizes = torch.cat([nrows, n_mask, torch.count_nonzero(candm).reshape(1)])
    host = _pinned()
    host[:3].copy_(sizes, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(stream)
    yield ev
    F, nm, ncand = int(host[0]), int(host[1]), int(host[2])
    if nm == 0 or ncand == 0:                                    # :232-245, :291-296
        _PINNED.append(host)
        return None
    # ---- stage B: patches, consensus, rank, sort, cover -------------------------------
    asm.F = F
    asm.rowvox = rowvox[:F]
    rsg = ((int(ps[2]) + 16 + 3) // 4) * 4
    asm.dp = torch.zeros((F, int(ps[0] * ps[1]) * rsg), dtype=torch.float32, device=dev)
    asm.fcmask = torch.empty((F, asm.W), dtype=torch.int32, device=dev)
    asm.rbits = asm.rv = asm.rb16 = None
    asm.small = int(ps[2]) <= 8 and int(ps[0] * ps[1]) <= 64
    if not asm.small:
        asm.rbits = torch.empty((int(ps[0] * ps[1]), F, 2), dtype=torch.int64, device=dev)
    cc.call('ppp_prepare_rows', cc.ptr(src.patches), cc.ptr(src.vox2row), cc.ptr(asm.flags),
            cc.ptr(asm.rowvox), F, asm.cfg, cc.ptr(asm.dp), cc.ptr(asm.fcmask), None,
            cc.ptr(asm.rbits), asm.stream)
    if asm.small:
        asm.received()
    asm._prepared = True
    asm.consensus()
    asm.rank()
    cand = _take(candm, ncand).to(torch.int32)
    order = asm.ranked(cand)
    if isinstance(kwargs.get('score_threshold', False), float):
        raise NotImplementedError("score_threshold: use to_instance_seg")
    n = int(order.numel())
    sparse = bool(kwargs.get('select_patches_for_sparse_data', False))
    if sparse:
        pix, pix_t = [], None
    else:
        pix = [t for t in [500, 100, 50, 10, 0] if t < int(asm.P / 2)]
        pix_t = torch.tensor(pix, dtype=torch.int32, device=dev)
    selected = torch.zeros(n, dtype=torch.uint8, device=dev)
    cscr = torch.empty(cc.call('ppp_cover_scratch_bytes', asm.cfg), dtype=torch.uint8, device=dev)
    asm._latency(lambda st: cc.call(
        'ppp_cover', cc.ptr(mask), cc.ptr(asm.overlap), cc.ptr(order), n, cc.ptr(asm.fgidx),
        cc.ptr(asm.fcmask), asm.cfg, cc.ptr(pix_t), len(pix), cc.ptr(selected), cc.ptr(cscr), st))
    selb = selected != 0
    host[:1].copy_(torch.count_nonzero(selb).reshape(1), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(stream)
    yield ev
    m = int(host[0])
    if m == 0:
        _PINNED.append(host)
        return None
    # ---- stage C: thinning; selection -> host -------------------------------------------
    sel = order[_take(selb, m)].contiguous()
    if not kwargs.get('skipThinCover', False):
        keep = torch.zeros(m, dtype=torch.uint8, device=dev)
        tscr = torch.empty(cc.call('ppp_thin_scratch_bytes', asm.cfg, m), dtype=torch.uint8,
                           device=dev)
        asm._latency(lambda st: cc.call(
            'ppp_thin', cc.ptr(mask), cc.ptr(sel), m, cc.ptr(asm.fgidx), cc.ptr(asm.fcmask),
            asm.cfg, cc.ptr(keep), cc.ptr(tscr), st))
    else:
        keep = torch.ones(m, dtype=torch.uint8, device=dev)
    sel_p, keep_p = _PinnedBuf(torch.int32, m), _PinnedBuf(torch.uint8, m)
    sel_p.t.copy_(sel, non_blocking=True)
    keep_p.t.copy_(keep, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(stream)
    yield ev
    v = sel_p.t.numpy()[keep_p.t.numpy() != 0].astype(np.int64)
    sel_p.release()
    keep_p.release()
    Z, Y, X = shape
    sel_coords = np.stack([v // (Y * X), (v // X) % Y, v % X], axis=1)
    # ---- host: pair enumeration (scipy releases the interpreter lock) -------------------
    fut = pool.submit(asm.patch_pairs, sel_coords)
    yield fut
    pairs = fut.result()
    _PINNED.append(host)
    if pairs is None:
        return None
    # ---- stage D: patch graph ------------------------------------------------------------
    pairs_dev = torch.from_numpy(pairs.view(np.int32)).to(dev, non_blocking=True)
    aff = asm.patch_graph(pairs_dev)
    aff_p = _PinnedBuf(torch.float32, int(aff.numel()))
    aff_p.t.copy_(aff, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(stream)
    yield ev
    out = aff_p.t.numpy().copy()
    aff_p.release()
    return pairs, out


def _ready(token):
    if isinstance(token, concurrent.futures.Future):
        return token.done()
    return token.query()


def _wait(token):
    if isinstance(token, concurrent.futures.Future):
        token.result()
    else:
        token.synchronize()


def run_blocks(jobs, make_steps, max_inflight=16, n_streams=8, host_threads=3, on_done=None):
    """jobs: list of arguments; make_steps(job, pool) -> generator (see block_steps).
    Runs up to `max_inflight` generators side by side, each on one of `n_streams` CUDA
    streams, resuming a block as soon as what it waits for is ready.  Returns the
    generators' return values in job order; on_done(job, value) is called as each one
    finishes."""
    torch = _torch()
    from .vote_instances import _cached_stream
    dev = torch.cuda.current_device()
    streams = [_cached_stream(dev, 'pipe%d' % k) for k in range(n_streams)]
    main = torch.cuda.current_stream()
    ready = torch.cuda.Event()
    ready.record(main)
    for st in streams:
        st.wait_event(ready)
    out = [None] * len(jobs)
    active = collections.deque()        # (job index, generator, stream, token)
    nxt = 0
    with concurrent.futures.ThreadPoolExecutor(max_workers=host_threads) as pool:
        def advance(i, gen, st):
            with torch.cuda.stream(st):
                try:
                    return next(gen)
                except StopIteration as e:
                    out[i] = e.value
                    if on_done is not None:
                        on_done(jobs[i], e.value)
                    return None
        while nxt < len(jobs) or active:
            while nxt < len(jobs) and len(active) < max_inflight:
                st = streams[nxt % n_streams]
                with torch.cuda.stream(st):
                    gen = make_steps(jobs[nxt], pool)
                tok = advance(nxt, gen, st)
                if tok is not None:
                    active.append((nxt, gen, st, tok))
                nxt += 1
            progressed = False
            for _ in range(len(active)):
                i, gen, st, tok = active.popleft()
                if _ready(tok):
                    progressed = True
                    tok = advance(i, gen, st)
                    if tok is None:
                        continue
                active.append((i, gen, st, tok))
            if not progressed and active and not (nxt < len(jobs) and len(active) < max_inflight):
                _wait(active[0][3])
    for st in streams:
        main.wait_stream(st)
    return out
