"""Blockwise instance assembly of ONE volume sharded over the GPUs of a box.

Counterpart of stitch_patch_graph.py:110-399 (stitch_vote_instances) and :553-669
(blockwise_vote_instances) for predictions in the compact row form (a ppp+dec run
decodes the foreground voxels only, decode.py:39-65; every other patch is zero).

Decomposition = the reference's: blocks on the `get_offsets` grid with a
`patchshape // 2` input halo (:575-607), one job per shared face (:188-357), one
global partition of the patch graph (:360-399).  What is new is where the data
lives and who does what:

  * the volume is cut into SLABS along one axis (whole block rows, so the block
    grid is unchanged); rank r holds the patch rows of its slab in HBM
    (`RowShard`), nothing is replicated;
  * exchange (0), NCCL send/recv: the rows within `2*ps + ps//2` of a slab
    border go to the neighbouring rank (block input halo + the face regions that
    reach across the border) -- fg-compacted rows, not dense planes;
  * every rank assembles the blocks of its slab: one host thread keeps many blocks in
    flight on several CUDA streams (pipeline.py);
  * exchange (1), neighbour send/recv: the pair lists of a rank's last block row go to
    the next rank, whose first block row has its lower face neighbours there;
  * face job (block, lower neighbour) runs on the owner of `block`; all face jobs of a
    block row are served by ONE consensus / patch-graph pass over the row's region;
  * exchange (2), ONE tensor all-gather of all block and face edge lists (pairs u32
    [n,6], aff f32 [n], padded to the largest rank);
  * every rank builds the global edge list in the reference's order and runs the
    same deterministic partition (ppp_label_cc on compacted node ids, or the
    mutex watershed), then paints the nodes whose windows touch ITS slab from the
    rows it already holds -- no label exchange, no full-volume reduction.

Labels are identical for every world size (tests/test_sharded.py; bench.py checks
a digest) and equal to the dense single-process driver (stitch_arrays).
"""
import concurrent.futures
import logging
import threading
import time

import numpy as np

from . import cuda_code as cc
from . import vote_instances as vi
from .assembly import RowSource
from .stitch_patch_graph import (get_offsets, face_candidates, face_pairs, _dist)

logger = logging.getLogger(__name__)


def slab_partition(shape, chunksize, world, axis=None, weights=None, bb_offset=None,
                   bb_shape=None):
    """block rows along `axis` dealt to the ranks in contiguous runs.
    weights: cost of every block row along `axis` (e.g. its foreground count); the
    runs then minimise the largest per-rank cost, else they hold equal row counts.
    bb_offset / bb_shape: the block grid covers this box only (`only_bb`,
    stitch_patch_graph.py:745-771); the first / last non-empty slab is extended to the
    volume border (the halos of the outer blocks reach beyond the box).
    Returns (axis, [(lo, hi)] * world) in voxels; empty slabs have lo == hi."""
    full = [int(s) for s in shape]
    org = [0, 0, 0] if bb_offset is None else [int(v) for v in bb_offset]
    shape = full if bb_shape is None else [int(s) for s in bb_shape]
    chunk = [int(min(c, s)) for c, s in zip(chunksize, shape)]
    nrows = [-(-s // c) for s, c in zip(shape, chunk)]
    if axis is None:
        axis = int(np.argmax(nrows))
    n = nrows[axis]
    if weights is None:
        cuts = [(r * n) // world for r in range(world + 1)]
    else:
        w = np.asarray(weights, np.float64)
        assert len(w) == n, "one weight per block row"
        pre = np.concatenate([[0.0], np.cumsum(w)])
        # best[k][i] = smallest possible largest run when rows [0, i) go to k ranks
        best = np.full((world + 1, n + 1), np.inf)
        arg = np.zeros((world + 1, n + 1), np.int64)
        best[0][0] = 0.0
        for k in range(1, world + 1):
            for i in range(n + 1):
                for j in range(i + 1):
                    c = max(best[k - 1][j], pre[i] - pre[j])
                    if c < best[k][i]:
                        best[k][i], arg[k][i] = c, j
        cuts = [n]
        for k in range(world, 0, -1):
            cuts.append(int(arg[k][cuts[-1]]))
        cuts = cuts[::-1]
    out = []
    for r in range(world):
        a, b = cuts[r], cuts[r + 1]
        out.append((org[axis] + min(a * chunk[axis], shape[axis]),
                    org[axis] + min(b * chunk[axis], shape[axis])))
    live = [r for r in range(world) if out[r][1] > out[r][0]]
    if live:
        out[live[0]] = (0, out[live[0]][1])
        out[live[-1]] = (out[live[-1]][0], full[axis])
    return axis, out


class RowShard:
    """the patch rows of one slab of a compact prediction volume, on the device.

    shape    global (Z,Y,X)
    axis, lo, hi   the slab: voxels with lo <= coord[axis] < hi belong to this rank
    coords   i32 [G,3] global (z,y,x) of the stored voxels
    patches  f16 [G,P]
    numinst  u8 [G] instance-count class of the voxel (overlap = numinst > 1), or None
    """

    def __init__(self, shape, axis, lo, hi, coords, patches, numinst=None, fg=None):
        import torch
        self.shape = tuple(int(s) for s in shape)
        self.axis, self.lo, self.hi = int(axis), int(lo), int(hi)
        self.coords = coords.to(torch.int32).contiguous()
        self.patches = patches.contiguous()
        assert patches.dtype == torch.float16
        self.numinst = None if numinst is None else numinst.to(torch.uint8).contiguous()
        # host-side foreground of the stored voxels (fg_key / numinst rule,
        # utilVoteInstances.py:306-322); None: the centre channel decides
        self.fg = None if fg is None else fg.to(torch.uint8).contiguous()
        self.dev = patches.device
        self.ext_lo, self.ext_hi = self.lo, self.hi
        self.vox2row = None
        self.n_own = int(self.coords.shape[0])
        self.halo_bytes = 0

    # -- exchange (0) ------------------------------------------------------
    def exchange_halo(self, slabs, halo):
        """send the rows within `halo` of the slab borders to the ranks that need
        them, receive mine, index everything (vox2row over the extended slab)."""
        import torch
        dist = _dist()
        world = dist.get_world_size() if dist else 1
        rank = dist.get_rank() if dist else 0
        S = self.shape[self.axis]
        ext = [(max(lo - halo, 0), min(hi + halo, S)) if hi > lo else (lo, hi) for lo, hi in slabs]
        self.ext_lo, self.ext_hi = ext[rank]
        coords, patches = self.coords, self.patches
        extras = [self.numinst, self.fg]            # optional u8 [G] columns, same on all ranks
        if world > 1:
            a = coords[:, self.axis]
            send_idx = []
            counts = torch.zeros(world, dtype=torch.int64)
            for d in range(world):
                if d == rank or ext[d][1] <= ext[d][0]:
                    send_idx.append(None)
                    continue
                lo_d, hi_d = ext[d]
                if hi_d <= self.lo or lo_d >= self.hi:
                    send_idx.append(None)
                    continue
                idx = torch.nonzero((a >= lo_d) & (a < hi_d)).flatten()
                send_idx.append(idx if idx.numel() else None)
                counts[d] = idx.numel()
            cdev = self.dev if dist.get_backend() == 'nccl' else torch.device('cpu')
            allc = torch.zeros((world, world), dtype=torch.int64, device=cdev)
            dist.all_gather_into_tensor(allc.view(-1), counts.to(cdev))
            allc = allc.cpu()
            ops, recv, keep = [], [], []
            P = int(patches.shape[1])
            for d in range(world):
                if d == rank:
                    continue
                ns, nr = int(allc[rank, d]), int(allc[d, rank])
                if ns:
                    idx = send_idx[d]
                    bufs = [coords[idx].contiguous(), patches[idx].contiguous()]
                    bufs += [e[idx].contiguous() for e in extras if e is not None]
                    keep.append(bufs)
                    for b in bufs:
                        ops.append(dist.P2POp(dist.isend, b.to(cdev), d))
                        self.halo_bytes += b.numel() * b.element_size()
                if nr:
                    bufs = [torch.empty((nr, 3), dtype=torch.int32, device=cdev),
                            torch.empty((nr, P), dtype=torch.float16, device=cdev)]
                    bufs += [torch.empty(nr, dtype=torch.uint8, device=cdev)
                             for e in extras if e is not None]
                    recv.append(bufs)
                    for b in bufs:
                        ops.append(dist.P2POp(dist.irecv, b, d))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            if recv:
                coords = torch.cat([coords] + [b[0].to(self.dev) for b in recv])
                patches = torch.cat([patches] + [b[1].to(self.dev) for b in recv])
                k = 2
                for i, e in enumerate(extras):
                    if e is not None:
                        extras[i] = torch.cat([e] + [b[k].to(self.dev) for b in recv])
                        k += 1
        self.coords, self.patches = coords, patches
        self.numinst, self.fg = extras
        self._index()

    def _index(self):
        import torch
        box = list(self.shape)
        box[self.axis] = max(self.ext_hi - self.ext_lo, 0)
        self.box = tuple(box)
        V = int(np.prod(box))
        assert V < 2 ** 31, "extended slab too large for 32-bit voxel indices"
        self.vox2row = torch.full((max(V, 1),), -1, dtype=torch.int32, device=self.dev)
        G = int(self.coords.shape[0])
        if G:
            c = self.coords.long()
            loc = [c[:, 0], c[:, 1], c[:, 2]]
            loc[self.axis] = loc[self.axis] - self.ext_lo
            lin = (loc[0] * box[1] + loc[1]) * box[2] + loc[2]
            self.vox2row[lin] = torch.arange(G, dtype=torch.int32, device=self.dev)
        self.vox2row = self.vox2row[:V].reshape(self.box) if V else \
            self.vox2row[:0].reshape(self.box)
        mid = int(self.patches.shape[1]) // 2
        self.mid = self.patches[:, mid].float().contiguous() if G else \
            torch.zeros(0, dtype=torch.float32, device=self.dev)

    # -- block inputs (blockwise_vote_instances :603-637) ---------------------
    def region(self, start, stop, **kwargs):
        """(RowSource, foreground u8, mask u8, numinst u8, start_clipped) of the box
        [start, stop) clipped to the volume; None if it holds no foreground."""
        import torch
        from .utilVoteInstances import getFgThreshold
        s = np.maximum(np.asarray(start, np.int64), 0)
        e = np.minimum(np.asarray(stop, np.int64), self.shape)
        ls, le = s.copy(), e.copy()
        ls[self.axis] -= self.ext_lo
        le[self.axis] -= self.ext_lo
        assert ls[self.axis] >= 0 and le[self.axis] <= self.box[self.axis], \
            "region [%s, %s) leaves the extended slab [%d, %d)" % (s, e, self.ext_lo, self.ext_hi)
        v2r = self.vox2row[int(ls[0]):int(le[0]), int(ls[1]):int(le[1]),
                           int(ls[2]):int(le[2])].contiguous()
        valid = v2r >= 0
        idx = v2r.clamp(min=0).long()
        th = float(np.float32(getFgThreshold(**kwargs)))
        fg = valid & ((self.mid[idx] > th) if self.fg is None else (self.fg[idx] != 0))
        if self.numinst is not None:
            numinst = torch.where(valid, self.numinst[idx], torch.zeros((), dtype=torch.uint8,
                                                                       device=self.dev))
        else:
            numinst = fg.to(torch.uint8)
        fg8 = fg.to(torch.uint8)
        return RowSource(self.patches, v2r), fg8, fg8.clone(), numinst, s


def _block_job(shard, offset, chunksize, ps, kwargs, block_fn=None):
    """blockwise_vote_instances (stitch_patch_graph.py:553-669) on one block of
    the shard.  Returns (pairs u32 [n,6] VOLUME coordinates, aff f32 [n]) or None."""
    margin = ps // 2
    offset = np.asarray(offset)
    src, fg, mask, numinst, start = shard.region(offset - margin, offset + chunksize + margin,
                                                 **kwargs)
    if not bool(fg.any().item()):
        return None
    kw = dict(kwargs)
    kw['return_intermediates'] = True
    pairs, aff = (block_fn or vi.do_block)(src, fg, mask, numinst, **kw)
    if pairs is None:
        return None
    pairs = pairs.astype(np.int64) + np.tile(start, 2)          # :650 and :162
    return pairs.astype(np.uint32), np.asarray(aff, np.float32)


def _face_job(shard, candidates, pa, ps, kwargs, block_fn=None):
    """stitch_patch_graph.py:252-336 on the shard: cross edges of one face."""
    cleaned = candidates[np.unique(pa.reshape(-1))]
    bb_start = np.maximum(np.min(cleaned, axis=0) - ps, 0)
    bb_stop = np.minimum(np.max(cleaned, axis=0) + ps, shard.shape)
    margin = ps // 2
    src, fg, mask, numinst, start = shard.region(
        bb_start - margin, np.maximum(bb_stop, bb_start + 1) + margin, **kwargs)
    overlapping = np.concatenate([candidates[pa[:, 0]], candidates[pa[:, 1]]], axis=1)
    # the reference subtracts the UNclipped region start (:317-321), see
    # stitch_patch_graph.assemble_face
    origin = start if kwargs.get('ppp_fix_face_origin', False) else bb_start - margin
    rel_c = cleaned - origin
    rel_p = overlapping - np.tile(origin, 2)
    kw = dict(kwargs)
    kw.update(skipRanking=True, skipThinCover=True, return_intermediates=True)
    _, aff = (block_fn or vi.do_block)(src, fg, mask, numinst, selected_patches=rel_c,
                                       selected_patch_pairs=rel_p.astype(np.uint32), **kw)
    if aff is None:
        aff = np.zeros(len(overlapping), np.float32)
    return overlapping.astype(np.uint32), np.asarray(aff, np.float32)


def _face_batch(shard, faces, ps, kwargs):
    """the face jobs `faces` = [(job, (candidates, pair index array))] as ONE pass over
    the bounding box of their regions.  Returns {job: (pairs u32 [n,6] volume coordinates,
    aff f32 [n])}, identical to running _face_job on each of them."""
    import torch
    from .assembly import BlockAssembler
    margin = ps // 2
    shape = np.asarray(shard.shape)
    per = []
    for j, (candidates, pa) in faces:
        cleaned = candidates[np.unique(pa.reshape(-1))]
        bb_start = np.maximum(np.min(cleaned, axis=0) - ps, 0)
        bb_stop = np.minimum(np.max(cleaned, axis=0) + ps, shape)
        r_start = np.maximum(bb_start - margin, 0)                    # load_region clips
        r_stop = np.minimum(np.maximum(bb_stop, bb_start + 1) + margin, shape)
        origin = r_start if kwargs.get('ppp_fix_face_origin', False) else bb_start - margin
        overlapping = np.concatenate([candidates[pa[:, 0]], candidates[pa[:, 1]]], axis=1)
        rel_p = overlapping - np.tile(origin, 2)
        inside = np.all((rel_p >= 0) & (rel_p < np.tile(r_stop - r_start, 2)), axis=1)
        eff_p = rel_p + np.tile(r_start, 2)          # where the kernel really looks (:317-321)
        eff_c = cleaned - origin + r_start
        per.append((j, overlapping, inside, eff_p, eff_c, r_start, r_stop))
    b_start = np.min([p[5] for p in per], axis=0)
    b_stop = np.max([p[6] for p in per], axis=0)
    src, fg, mask, numinst, _ = shard.region(b_start, b_stop, **kwargs)
    bshape = tuple(int(v) for v in (b_stop - b_start))
    ckw = {k: v for k, v in kwargs.items() if k != 'patchshape'}
    asm = BlockAssembler(src, fg, (numinst > 1).to(torch.uint8), ps, **ckw)
    asm.prepare(want_dp=False, want_rv=False, want_masks=False)
    dev = shard.dev
    pk = np.concatenate([p[3][p[2]] for p in per]) - np.tile(b_start, 2)
    org = np.concatenate([np.broadcast_to(p[5] - b_start, (int(p[2].sum()), 3)) for p in per])
    cen = np.concatenate([p[4] for p in per]) - b_start
    ok = np.all((cen >= 0) & (cen < np.asarray(bshape)), axis=1)
    cen = torch.from_numpy(np.ascontiguousarray(cen[ok].astype(np.int32))).to(dev)
    need = asm.window_rows(cen, margin)
    # their partners (anything within 2*ps-1 of a wanted row) must have tables too
    asm.received(asm.window_rows(cen, margin + ps - 1))
    asm.consensus(need=need)
    aff_all = np.zeros(0, np.float32)
    if len(pk):
        pairs_dev = torch.from_numpy(np.ascontiguousarray(pk.astype(np.uint32)).view(np.int32)).to(dev)
        org_dev = torch.from_numpy(np.ascontiguousarray(org.astype(np.int32))).to(dev)
        aff_all = asm.patch_graph(pairs_dev, pair_org=org_dev).cpu().numpy()
    out = {}
    o = 0
    for j, overlapping, inside, _, _, _, _ in per:
        n = int(inside.sum())
        aff = np.zeros(len(overlapping), np.float32)
        aff[inside] = aff_all[o:o + n]
        o += n
        out[j] = (overlapping.astype(np.uint32), aff)
    return out


def _run_jobs(jobs, fn, workers):
    """run fn(job) for every job; with workers > 1 from that many host threads, each
    on its own CUDA stream, so that the host part of one job (pair search, sizes
    read back) overlaps the kernels of the others.  Results in job order."""
    import torch
    if workers <= 1 or len(jobs) <= 1:
        return [fn(j) for j in jobs]
    out = [None] * len(jobs)
    nxt = [0]
    lock = threading.Lock()
    err = []
    dev = torch.cuda.current_device()
    main_stream = torch.cuda.current_stream()
    ready = torch.cuda.Event()
    ready.record(main_stream)

    def work(k):
        try:
            torch.cuda.set_device(dev)
            st = vi._cached_stream(dev, 'job%d' % k)
            st.wait_event(ready)
            with torch.cuda.stream(st):
                while True:
                    with lock:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= len(jobs) or err:
                        break
                    out[i] = fn(jobs[i])
                st.synchronize()
        except BaseException as e:           # noqa: BLE001 -- re-raised below
            err.append(e)

    ths = [threading.Thread(target=work, args=(k,)) for k in range(workers)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if err:
        raise err[0]
    return out


def _neighbour_exchange(mine, send_blocks, nxt, recv_blocks, prv, dev):
    """pair lists of `send_blocks` to rank `nxt`, those of `recv_blocks` from rank `prv`
    (None: no such neighbour).  Two messages each way: counts i64 [blocks], pairs i32
    [sum,6].  Returns {block: (pairs u32 [n,6], None)} for the non-empty received blocks."""
    import torch
    dist = _dist()
    if dist is None or (nxt is None and prv is None):
        return {}
    cdev = dev if dist.get_backend() == 'nccl' else torch.device('cpu')
    ops = []
    if nxt is not None:
        cnt = torch.tensor([len(mine[b][1]) if b in mine else 0 for b in send_blocks],
                           dtype=torch.int64)
        ops.append(dist.P2POp(dist.isend, cnt.to(cdev), nxt))
    rc = None
    if prv is not None:
        rc = torch.zeros(len(recv_blocks), dtype=torch.int64, device=cdev)
        ops.append(dist.P2POp(dist.irecv, rc, prv))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    ops = []
    if nxt is not None and int(cnt.sum()):
        data = np.concatenate([mine[b][0] for b in send_blocks if b in mine])
        ops.append(dist.P2POp(dist.isend, torch.from_numpy(
            np.ascontiguousarray(data).view(np.int32)).to(cdev), nxt))
    rd = None
    if prv is not None:
        rc = rc.cpu().numpy()
        if int(rc.sum()):
            rd = torch.empty((int(rc.sum()), 6), dtype=torch.int32, device=cdev)
            ops.append(dist.P2POp(dist.irecv, rd, prv))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    out = {}
    if rd is not None:
        rd = rd.cpu().numpy().view(np.uint32)
        o = 0
        for b, n in zip(recv_blocks, rc):
            if n:
                out[b] = (rd[o:o + int(n)], None)
                o += int(n)
    return out


def _allgather_edges(per_job, n_jobs, owner_of):
    """exchange (1)/(2): every rank contributes {job: (pairs, aff)} for the jobs it
    owns; returns the complete {job: (pairs u32 [n,6], aff f32 [n])} on every rank.
    Tensors only: counts i64 [n_jobs], pairs i32 [max,6], aff f32 [max]."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return dict(per_job)
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' \
        else torch.device('cpu')
    counts = np.zeros(n_jobs, np.int64)
    mine = sorted(per_job)
    for j in mine:
        counts[j] = len(per_job[j][1])
    ct = torch.from_numpy(counts).to(dev)
    dist.all_reduce(ct)
    counts = ct.cpu().numpy()
    per_rank = np.zeros(world, np.int64)
    for j in range(n_jobs):
        per_rank[owner_of(j)] += counts[j]
    mx = int(per_rank.max())
    if mx == 0:
        return {}
    pbuf = np.zeros((mx, 6), np.uint32)
    abuf = np.zeros(mx, np.float32)
    o = 0
    for j in mine:
        n = len(per_job[j][1])
        pbuf[o:o + n] = per_job[j][0]
        abuf[o:o + n] = per_job[j][1]
        o += n
    pt = torch.from_numpy(pbuf.view(np.int32)).to(dev)
    at = torch.from_numpy(abuf).to(dev)
    pall = torch.empty((world, mx, 6), dtype=torch.int32, device=dev)
    aall = torch.empty((world, mx), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(pall.view(-1), pt.view(-1))
    dist.all_gather_into_tensor(aall.view(-1), at)
    pall = pall.cpu().numpy().view(np.uint32)
    aall = aall.cpu().numpy()
    out = {}
    pos = np.zeros(world, np.int64)
    for j in range(n_jobs):
        n = int(counts[j])
        if n == 0:
            continue
        r = owner_of(j)
        out[j] = (pall[r, pos[r]:pos[r] + n], aall[r, pos[r]:pos[r] + n])
        pos[r] += n
    return out


def partition_graph(pairs, aff, shape, dev, **kwargs):
    """the global partition (affGraphToInstances, graph_to_labeling.py:44-54 on the
    graph of setAffgraph, aff_patch_graph.py:31-40) on compacted node ids.
    Returns (node_coords i32 [m,3] device, node_label i32 [m] device, n_labels)."""
    import torch
    n = len(pairs)
    Y, X = int(shape[1]), int(shape[2])
    pd = torch.from_numpy(np.ascontiguousarray(pairs).view(np.int32)).to(dev).long()
    keys = torch.cat([(pd[:, 0] * Y + pd[:, 1]) * X + pd[:, 2],
                      (pd[:, 3] * Y + pd[:, 4]) * X + pd[:, 5]])
    uniq, inv = torch.unique(keys, sorted=True, return_inverse=True)
    m = int(uniq.numel())
    cp = torch.zeros((n, 6), dtype=torch.int32, device=dev)
    cp[:, 2] = inv[:n].to(torch.int32)
    cp[:, 5] = inv[n:].to(torch.int32)
    ckw = {k: v for k, v in kwargs.items() if k != 'patchshape'}
    cfg = cc.make_cfg((1, 1, max(m, 1)), kwargs['patchshape'], **ckw)
    ad = torch.from_numpy(np.ascontiguousarray(aff, np.float32)).to(dev)
    if kwargs.get('mws', False):
        from .assembly import mutex_watershed
        node_vox, node_label, top = mutex_watershed(cp.cpu().numpy().view(np.uint32), aff, cfg)
        label = torch.zeros(m, dtype=torch.int32, device=dev)
        label[torch.from_numpy(node_vox).to(dev).long()] = torch.from_numpy(node_label).to(dev)
        n_labels = top
    else:
        label = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
        ncomp = torch.zeros(1, dtype=torch.int32, device=dev)
        scratch = torch.empty(cc.call('ppp_label_scratch_bytes', max(m, 1), n),
                              dtype=torch.uint8, device=dev)
        cc.call('ppp_label_cc', cc.ptr(cp), cc.ptr(ad), n, cfg, cc.ptr(label), cc.ptr(ncomp),
                cc.ptr(scratch), cc.current_stream_ptr())
        n_labels = int(ncomp.item())
        label = label[:m]
    node_coords = torch.stack([uniq // (Y * X), (uniq // X) % Y, uniq % X], dim=1).to(torch.int32)
    return node_coords, label, n_labels


def stitch_shard(shard, slabs, workers=None, block_fn=None, paint_fn=None, **kwargs):
    """blockwise assembly + stitching of the volume `shard` is a slab of.

    shard   RowShard of this rank (own rows only; the halo is exchanged here)
    slabs   [(lo, hi)] of every rank along shard.axis (slab_partition)
    block_fn / paint_fn: test hooks (the CPU tests drive this host logic with the
    oracle as engine): block_fn(RowSource, fg, mask, numinst, **kw) like do_block;
    paint_fn(shard, pairs, aff, own_box, **kw) -> labels of the own slab.
    Returns (instances i32 device tensor of the OWN slab, info dict)."""
    import torch
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    ps = np.asarray(kwargs['patchshape'])
    shape = shard.shape
    axis = shard.axis
    # only_bb (stitch_patch_graph.py:745-771, 137-138): block grid over the bounding box
    bb_offset = np.asarray(kwargs.pop('bb_offset', np.zeros(3, int)), dtype=int)
    bb_shape = np.asarray(kwargs.pop('bb_shape', shape), dtype=int)
    chunksize = np.minimum(np.asarray(kwargs['chunksize']), bb_shape)
    kwargs = dict(kwargs, chunksize=chunksize)
    if workers is None:
        workers = int(kwargs.get('ppp_block_workers', 6))
    offsets = [o + bb_offset for o in get_offsets(bb_shape, chunksize)]
    nblk = len(offsets)

    def slab_of(coord):
        for r, (lo, hi) in enumerate(slabs):
            if lo <= coord < hi:
                return r
        raise ValueError("coordinate %d outside every slab" % coord)
    owner = [slab_of(int(o[axis])) for o in offsets]

    tm = {}
    t_last = [time.perf_counter()]

    def lap(name):
        if kwargs.get('ppp_phase_sync', True):
            torch.cuda.synchronize() if shard.dev.type == 'cuda' else None
        now = time.perf_counter()
        tm[name] = tm.get(name, 0.0) + (now - t_last[0]) * 1e3
        t_last[0] = now

    # ---- exchange (0): halo rows ---------------------------------------------------
    halo = int(2 * ps[axis] + ps[axis] // 2)
    shard.exchange_halo(slabs, halo)
    lap('halo_exchange')

    # ---- phase 1: the blocks of my slab ----------------------------------------------
    my_blocks = [b for b in range(nblk) if owner[b] == rank]
    mine, box, selected = {}, {}, {}
    key_of = {tuple(int(v) for v in o): i for i, o in enumerate(offsets)}

    def selected_of(b):
        """the selected patches of a block = the nodes of its pair list (:166-171);
        computed where first needed, from any thread (a race only repeats the work)"""
        got = selected.get(b)
        if got is None:
            src = mine[b] if b in mine else box['blocks'][b]
            p = src[0].reshape(-1, 3).astype(np.int64)
            key = (p[:, 0] * shape[1] + p[:, 1]) * shape[2] + p[:, 2]
            _, first = np.unique(key, return_index=True)     # sorted by (z,y,x) like
            got = selected[b] = p[first]                     # np.unique(axis=0)
        return got

    def neighbours(b, sign):
        for dim in range(3):                                     # -z, -y, -x (:125-128)
            nb_off = offsets[b].copy()
            nb_off[dim] += sign * chunksize[dim]
            nb = key_of.get(tuple(int(v) for v in nb_off))
            if nb is not None and (nb < b if sign < 0 else nb > b):
                yield nb, dim

    def face_host(bnd):
        """candidates and cross pairs of the face (block, neighbour, axis) (host: KD
        tree, set order); None if there is nothing to pair"""
        b, nb, dim = bnd
        cur, nbc = face_candidates(selected_of(b), selected_of(nb), offsets[b], dim, ps)
        if len(cur) == 0 or len(nbc) == 0:
            return None
        cands, pa = face_pairs(cur, nbc, ps)
        if len(pa) == 0:
            return None
        return cands, pa

    batch_faces = block_fn is None and kwargs.get('ppp_batch_faces', True)
    early = {}                  # face (block, lower neighbour, axis) -> future of face_host
    finished = set()
    hpool = concurrent.futures.ThreadPoolExecutor(max_workers=3) if batch_faces else None

    def block_done(b, r):
        """the host part of a face between two of my blocks starts as soon as both are
        assembled, on a small thread pool, while the GPU works on the other blocks"""
        finished.add(b)
        if r is None:
            return
        mine[b] = r
        if hpool is None:
            return
        for nb, dim in neighbours(b, -1):
            if nb in mine:
                early[(b, nb, dim)] = hpool.submit(face_host, (b, nb, dim))
        for nb, dim in neighbours(b, +1):
            if nb in mine:
                early[(nb, b, dim)] = hpool.submit(face_host, (nb, b, dim))

    if block_fn is None and shard.dev.type == 'cuda' and kwargs.get('ppp_pipeline', True) and \
            not isinstance(kwargs.get('score_threshold', False), float):
        # one host thread, many blocks in flight (pipeline.py)
        from . import pipeline
        margin = ps // 2
        ckw = {k: v for k, v in kwargs.items() if k != 'patchshape'}

        def steps(b, pool):
            off = np.asarray(offsets[b])
            src, fg, mask, numinst, start = shard.region(off - margin, off + chunksize + margin,
                                                         **kwargs)
            got = yield from pipeline.block_steps(src, fg, mask, numinst, ps, pool, **ckw)
            if got is None:
                return None
            pairs = got[0].astype(np.int64) + np.tile(start, 2)      # :650 and :162
            return pairs.astype(np.uint32), got[1]
        pipeline.run_blocks(my_blocks, steps, max_inflight=int(kwargs.get('ppp_inflight', 16)),
                            n_streams=int(kwargs.get('ppp_streams', 8)), on_done=block_done)
    else:
        res = _run_jobs(my_blocks,
                        lambda b: _block_job(shard, offsets[b], chunksize, ps, kwargs, block_fn),
                        workers)
        for b, r in zip(my_blocks, res):
            block_done(b, r)
    lap('blocks')
    # exchange (1), neighbour to neighbour: a face job reads the selected patches of the
    # lower neighbour block; for the first block row of a slab those belong to the previous
    # rank, which sends the pair lists of its last block row (send/recv on a helper thread
    # while this rank finishes the host part of its interior faces).  Nothing global here:
    # a rank only ever waits for its predecessor.
    def row_of(b):
        return (int(offsets[b][axis]) - int(bb_offset[axis])) // int(chunksize[axis])
    live = [r for r in range(world) if any(o == r for o in owner)]
    my_rows = sorted({row_of(b) for b in my_blocks})
    prv = live[live.index(rank) - 1] if rank in live and live.index(rank) > 0 else None
    nxt = live[live.index(rank) + 1] if rank in live and live.index(rank) + 1 < len(live) else None
    send_blocks = [b for b in my_blocks if row_of(b) == my_rows[-1]] if nxt is not None else []
    recv_blocks = [b for b in range(nblk) if owner[b] == prv and my_rows and
                   row_of(b) == my_rows[0] - 1] if prv is not None else []

    def exchange1():
        if shard.dev.type == 'cuda':
            torch.cuda.set_device(shard.dev)
        box['remote'] = _neighbour_exchange(mine, send_blocks, nxt, recv_blocks, prv, shard.dev)
    ex = threading.Thread(target=exchange1)
    ex.start()
    early_res = {k: f.result() for k, f in early.items()}
    if hpool is not None:
        hpool.shutdown()
    lap('faces_host_early')
    ex.join()
    box['blocks'] = box['remote']            # where selected_of looks for foreign blocks
    lap('neighbour_exchange')

    # ---- phase 2: face jobs, on the owner of the higher block ------------------------
    # (block, lower neighbour, axis) for both non-empty; the first non-empty block of the
    # volume has no non-empty lower neighbour, so :151-156 / :178-181 need no special case
    known = set(mine) | set(box['remote'])
    my_jobs = [(b, nb, dim) for b in my_blocks if b in mine
               for nb, dim in neighbours(b, -1) if nb in known]

    def face(job):
        cp = face_host(job)
        if cp is None:
            return None
        return _face_job(shard, cp[0], cp[1], ps, kwargs, block_fn)
    if block_fn is not None or not kwargs.get('ppp_batch_faces', True):
        res = _run_jobs(my_jobs, face, workers)
    else:
        # all face jobs of one block row share ONE region: the consensus of the slots a
        # face job reads does not depend on the extent of its region (every centre that can
        # vote on them lies inside: the region is padded by patchshape + patchshape//2,
        # :261-278), so one gate/prepare/consensus/patch-graph pass over the row serves them all
        late = [job for job in my_jobs if job not in early_res]
        late_res = dict(zip(late, _run_jobs(late, face_host, workers)))
        host = [early_res[job] if job in early_res else late_res[job] for job in my_jobs]
        lap('faces_host')
        groups = {}
        for job, cp in zip(my_jobs, host):
            if cp is not None:
                groups.setdefault(int(offsets[job[0]][axis]), []).append((job, cp))
        got = {}
        for part in _run_jobs(sorted(groups), lambda k: _face_batch(shard, groups[k], ps, kwargs),
                              min(workers, 2)):
            got.update(part)
        res = [got.get(job) for job in my_jobs]
    lap('faces')
    # exchange (2): ONE all-gather of everything this rank computed, keyed so that sorting
    # the keys gives the reference's order of update_graph calls: block b, then its faces
    # towards -z, -y, -x (stitch_patch_graph.py:178-193)
    contrib = {4 * b: r for b, r in mine.items()}
    for (b, nb, dim), r in zip(my_jobs, res):
        if r is not None:
            contrib[4 * b + 1 + dim] = r
    edges = _allgather_edges(contrib, 4 * nblk, lambda k: owner[k // 4])
    lap('allgather_edges')
    plist = [edges[k][0] for k in sorted(edges)]
    alist = [edges[k][1] for k in sorted(edges)]
    n_faces = sum(1 for k in edges if k % 4)
    info = dict(n_blocks=nblk, n_faces=n_faces, n_edges=int(sum(len(a) for a in alist)),
                my_blocks=len(my_blocks), my_faces=len(my_jobs), halo_bytes=shard.halo_bytes,
                phase_ms=tm,
                rows=int(shard.coords.shape[0]), own_rows=shard.n_own)
    own_box = list(shape)
    own_box[axis] = shard.hi - shard.lo
    inst = torch.zeros(own_box, dtype=torch.int32, device=shard.dev)
    if not plist:
        return inst, info
    pairs = np.concatenate(plist)
    aff = np.concatenate(alist)
    assert pairs.dtype == np.uint32 and aff.dtype == np.float32
    info['pairs'] = pairs
    info['aff'] = aff

    # ---- phase 3: replicated partition, painting of my slab --------------------------
    if paint_fn is not None:
        return paint_fn(shard, pairs, aff, own_box, **kwargs), info
    lap('edge_list')
    node_coords, label, n_labels = partition_graph(pairs, aff, shape, shard.dev, **kwargs)
    info['n_labels'] = n_labels
    lap('partition')
    r_ax = int(ps[axis] // 2)
    a = node_coords[:, axis]
    near = (a >= max(shard.lo - r_ax, shard.ext_lo)) & (a < min(shard.hi + r_ax, shard.ext_hi)) \
        & (label > 0)
    sel = torch.nonzero(near).flatten()
    if sel.numel() and inst.numel():
        nc = node_coords[sel].long()
        loc = [nc[:, 0], nc[:, 1], nc[:, 2]]
        loc[axis] = loc[axis] - shard.ext_lo
        node_row = shard.vox2row[loc[0], loc[1], loc[2]].contiguous()
        zyx = node_coords[sel].clone()
        zyx[:, axis] -= shard.lo
        ckw = {k: v for k, v in kwargs.items() if k != 'patchshape'}
        cfg = cc.make_cfg(own_box, ps, **ckw)
        cc.call('ppp_paint_rows', cc.ptr(shard.patches), cc.ptr(node_row),
                cc.ptr(zyx.contiguous()), cc.ptr(label[sel].contiguous()), int(sel.numel()),
                cfg, cc.ptr(inst), cc.current_stream_ptr())
    lap('paint')
    return inst, info


def gather_slabs(inst, slabs, axis, shape, dst=0):
    """collect the per-rank label slabs on rank `dst` (the volume the reference
    writes, stitch_patch_graph.py:849-869).  Returns the full tensor on `dst`,
    None elsewhere."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return inst
    world, rank = dist.get_world_size(), dist.get_rank()
    cdev = inst.device if dist.get_backend() == 'nccl' else torch.device('cpu')
    if rank == dst:
        full = torch.zeros(tuple(shape), dtype=inst.dtype, device=cdev)
        ops, parts = [], []
        for r, (lo, hi) in enumerate(slabs):
            if hi <= lo:
                continue
            sl = [slice(None)] * 3
            sl[axis] = slice(lo, hi)
            if r == rank:
                full[tuple(sl)] = inst.to(cdev)
                continue
            box = list(shape)
            box[axis] = hi - lo
            buf = torch.empty(box, dtype=inst.dtype, device=cdev)
            parts.append((tuple(sl), buf))
            ops.append(dist.P2POp(dist.irecv, buf, r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for sl, buf in parts:
            full[sl] = buf
        return full
    if inst.numel():
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, inst.to(cdev).contiguous(), dst)]):
            w.wait()
    return None


def rows_from_dense(pred, foreground, numinst, axis, lo, hi, device, th, tile=32):
    """the compact rows of the slab lo <= coord[axis] < hi of a DENSE prediction
    (numpy / zarr-like [P,Z,Y,X]): every voxel whose centre channel passes `th` or that
    the host-side `foreground` (bool [Z,Y,X]) names -- the only voxels whose patches the
    assembly ever reads.  Read `tile` planes at a time, so the dense volume is never held.
    Returns (coords i32 [G,3], patches f16 [G,P], numinst u8 [G], fg u8 [G]) on `device`."""
    import torch
    P = int(pred.shape[0])
    mid = P // 2
    cs, ps_, ns, fs = [], [], [], []
    for a in range(lo, hi, tile):
        b = min(a + tile, hi)
        sl = [slice(None)] * 3
        sl[axis] = slice(a, b)
        blk = np.asarray(pred[(slice(None),) + tuple(sl)])
        fgs = np.asarray(foreground[tuple(sl)]) != 0
        keep = (blk[mid].astype(np.float32) > np.float32(th)) | fgs
        c = np.argwhere(keep)
        if len(c) == 0:
            continue
        ps_.append(np.ascontiguousarray(blk[:, c[:, 0], c[:, 1], c[:, 2]].T.astype(np.float16)))
        fs.append(fgs[c[:, 0], c[:, 1], c[:, 2]].astype(np.uint8))
        ns.append(np.asarray(numinst[tuple(sl)])[c[:, 0], c[:, 1], c[:, 2]].astype(np.uint8))
        c[:, axis] += a
        cs.append(c.astype(np.int32))
    if not cs:
        return (torch.zeros((0, 3), dtype=torch.int32, device=device),
                torch.zeros((0, P), dtype=torch.float16, device=device),
                torch.zeros(0, dtype=torch.uint8, device=device),
                torch.zeros(0, dtype=torch.uint8, device=device))
    # raster order over the slab (tiles are raster-ordered only along `axis` == 0)
    c = np.concatenate(cs)
    order = np.lexsort((c[:, 2], c[:, 1], c[:, 0]))
    cat = lambda parts: torch.from_numpy(np.concatenate(parts)[order]).to(device)
    return cat(cs), cat(ps_), cat(ns), cat(fs)
