"""Device pipeline for ONE block: consensus -> rank -> cover -> thin ->
patch graph -> connected components -> paint, all through the C ABI
(include/ppp_b200.h).  This is the hot path that `to_instance_seg`
(vote_instances.py:150-452 in the reference) drives; every stage keeps its
data on the GPU, only the patch-pair enumeration (scipy cKDTree, exactly as
aff_patch_graph.py:43-110) touches the host.
"""
import ctypes

import numpy as np
import scipy.spatial

from . import cuda_code as cc
from .layout import patch_geometry


def _torch():
    import torch
    return torch


_PYSET_REPLAY = None


def _pyset_replay_ok():
    """ppp_pyset_order replays CPython's set table; verified once per process
    against a real set (another interpreter version may hash or grow differently)."""
    global _PYSET_REPLAY
    if _PYSET_REPLAY is None:
        ok = True
        rng = np.random.default_rng(12345)
        for n in (7, 300, 9000):
            a = np.unique(rng.integers(0, 4 * n, (n, 2)).astype(np.int64), axis=0)
            rng.shuffle(a)
            a = np.ascontiguousarray(a)
            order = np.zeros(len(a), np.int64)
            cc.call('ppp_pyset_order', a.ctypes.data, len(a), order.ctypes.data)
            want = np.array(list(set(map(tuple, a.tolist()))), np.int64).reshape(-1, 2)
            ok = ok and np.array_equal(a[order], want)
        _PYSET_REPLAY = bool(ok)
    return _PYSET_REPLAY


def query_pairs_set_order(tree, r):
    """tree.query_pairs(r, p=1) as an int64 [n,2] array in the order in which
    python iterates over the SET scipy returns (the reference's enumeration order,
    aff_patch_graph.py:57-110) -- without building the set."""
    if _pyset_replay_ok():
        a = np.ascontiguousarray(tree.query_pairs(r, p=1, output_type='ndarray'), np.int64)
        if len(a) == 0:
            return a.reshape(0, 2)
        order = np.zeros(len(a), np.int64)
        cc.call('ppp_pyset_order', a.ctypes.data, len(a), order.ctypes.data)
        return a[order]
    pairs = tree.query_pairs(r, p=1)
    return np.array(list(pairs), dtype=np.int64).reshape(-1, 2)


def query_pairs_filtered(tree, pts, r, thr):
    """query_pairs_set_order followed by the reference's per-axis distance filter
    (aff_patch_graph.py:61-69), in one pass on the host side of the library.
    pts u32 [m,3], thr float [3]; returns int64 [k,2]."""
    if _pyset_replay_ok():
        import ctypes
        a = np.ascontiguousarray(tree.query_pairs(r, p=1, output_type='ndarray'), np.int64)
        if len(a) == 0:
            return a.reshape(0, 2)
        pts = np.ascontiguousarray(pts, np.uint32)
        thr = np.ascontiguousarray(thr, np.float64)
        out = np.empty((len(a), 2), np.int64)
        n_out = ctypes.c_int64(0)
        cc.call('ppp_pyset_pairs', a.ctypes.data, len(a), pts.ctypes.data, thr.ctypes.data,
                out.ctypes.data, ctypes.addressof(n_out))
        return out[:int(n_out.value)]
    pa = query_pairs_set_order(tree, r)
    if len(pa) == 0:
        return pa
    d = np.abs(pts[pa[:, 0]].astype(np.float32) - pts[pa[:, 1]].astype(np.float32))
    return pa[~np.any(d > thr, axis=1)]


def mutex_watershed(pairs, aff, cfg):
    """graph_mws.mws on the graph of setAffgraph (graph_mws.py:7-85,
    aff_patch_graph.py:31-40) through ppp_mws_host: the one serial graph pass of
    the path, host side like the pair search.  pairs u32 [n,6], aff f32 [n]
    (numpy); cfg carries the volume shape.  Returns (node_vox i32 [m],
    node_label i32 [m], number of component ids created); label 0 = node joined nothing."""
    import ctypes
    pairs = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 6)
    aff = np.ascontiguousarray(aff, np.float32)
    n = len(aff)
    assert len(pairs) == n
    node_vox = np.zeros(max(1, 2 * n), np.int32)
    node_label = np.zeros(max(1, 2 * n), np.int32)
    n_nodes = ctypes.c_int64(0)
    top = ctypes.c_int32(0)
    cc.call('ppp_mws_host', pairs.ctypes.data, aff.ctypes.data, n, cfg,
            node_vox.ctypes.data, node_label.ctypes.data,
            ctypes.addressof(n_nodes), ctypes.addressof(top))
    m = int(n_nodes.value)
    return node_vox[:m].copy(), node_label[:m].copy(), int(top.value)


class RowSource:
    """compact prediction of one block (the form a ppp+dec run produces,
    decode.py:39-65): `patches` f16 [G,P] cuda tensor, one row per stored voxel, and
    `vox2row` i32 [Z,Y,X] cuda tensor = row of every block voxel or -1 (an all-zero
    patch).  Stands in for the dense f32 [P,Z,Y,X] block everywhere."""

    def __init__(self, patches, vox2row):
        torch = _torch()
        assert patches.dtype == torch.float16 and patches.is_contiguous()
        assert vox2row.dtype == torch.int32 and vox2row.is_contiguous()
        assert vox2row.dim() == 3
        self.patches = patches
        self.vox2row = vox2row
        self.shape = tuple(int(s) for s in vox2row.shape)
        self.device = patches.device

    def dense(self):
        """the equivalent f32 [P,Z,Y,X] tensor (tests only: this is what the rows
        path avoids)."""
        torch = _torch()
        P = int(self.patches.shape[1])
        out = torch.zeros((P,) + self.shape, dtype=torch.float32, device=self.device)
        m = self.vox2row >= 0
        out[:, m] = self.patches[self.vox2row[m].long()].float().T
        return out


_HP_STREAMS = {}


def _latency_stream(cur):
    """high-priority companion of stream `cur` (created once per stream): the few-CTA,
    latency-bound stages (thinning rounds, cover, the candidate sort) are launched on it,
    so that the block scheduler slots their CTAs in ahead of the large grids that other
    host threads (other blocks) have in flight instead of queueing behind them."""
    torch = _torch()
    key = (cur.device.index, cur.cuda_stream)
    if key not in _HP_STREAMS:
        lo, hi = torch.cuda.Stream.priority_range()
        _HP_STREAMS[key] = torch.cuda.Stream(cur.device, priority=hi)
    return _HP_STREAMS[key]


class BlockAssembler:
    """state of one block on the device.

    pred        f32 [P,Z,Y,X] cuda tensor (the block incl. its halo), or a RowSource
    foreground  host-side foreground (bool/u8 [Z,Y,X] cuda tensor): the
                candidate patch centres (vote_instances.py:276-287)
    overlap     u8 [Z,Y,X] cuda tensor, numinst > 1 (vote_instances.py:211)
    """

    def __init__(self, pred, foreground, overlap, patchshape, **kwargs):
        torch = _torch()
        self.ps, self.P, self.rad, _, self.N, self.K = patch_geometry(patchshape)
        if isinstance(pred, RowSource):
            self.rows = pred
            self.pred = None
            self.shape = pred.shape
            assert pred.patches.shape[1] == self.P, "row length != prod(patchshape)"
        else:
            assert pred.is_cuda and pred.dtype == torch.float32 and pred.is_contiguous()
            self.rows = None
            self.pred = pred
            self.shape = tuple(int(s) for s in pred.shape[1:])
            assert pred.shape[0] == self.P, "channel count != prod(patchshape)"
        self.kwargs = kwargs
        self.cfg = cc.make_cfg(self.shape, self.ps, **kwargs)
        self.V = int(np.prod(self.shape))
        self.dev = pred.device
        self.foreground = foreground.to(torch.uint8).contiguous()
        self.overlap = overlap.to(torch.uint8).contiguous()
        self.stream = cc.current_stream_ptr()
        self._cur = torch.cuda.current_stream()
        self._hp = _latency_stream(self._cur) if kwargs.get('ppp_latency_stream', True) else None
        self.W = (self.P + 31) // 32
        self.cons = None
        self.cnt = None
        self.score = None
        self._prepared = False

    def _latency(self, fn):
        """run fn(stream pointer) on the high-priority companion stream, ordered after
        everything queued on the block's stream and before everything that follows."""
        if self._hp is None:
            return fn(self.stream)
        self._hp.wait_stream(self._cur)
        r = fn(ctypes.c_void_p(self._hp.cuda_stream))
        self._cur.wait_stream(self._hp)
        return r

    # -- step 0 ------------------------------------------------------------
    def prepare(self, want_dp=True, want_rbits=None, want_rv=None, want_masks=True):
        torch = _torch()
        V = self.V
        self.flags = torch.empty(V, dtype=torch.uint8, device=self.dev)
        if self.rows is not None:
            cc.call('ppp_gate_rows', cc.ptr(self.rows.patches), cc.ptr(self.rows.vox2row),
                    cc.ptr(self.overlap), cc.ptr(self.foreground), self.cfg,
                    cc.ptr(self.flags), self.stream)
        else:
            cc.call('ppp_gate', cc.ptr(self.pred), cc.ptr(self.overlap),
                    cc.ptr(self.foreground), self.cfg, cc.ptr(self.flags), self.stream)
        self.fgidx = torch.empty(V, dtype=torch.int32, device=self.dev)
        self.rowvox = torch.empty(V, dtype=torch.int32, device=self.dev)
        nrows = torch.zeros(1, dtype=torch.int64, device=self.dev)
        scratch = torch.empty(cc.call('ppp_compact_scratch_bytes', V), dtype=torch.uint8,
                              device=self.dev)
        cc.call('ppp_compact', cc.ptr(self.flags), V, cc.ptr(self.fgidx),
                cc.ptr(self.rowvox), cc.ptr(nrows), cc.ptr(scratch), self.stream)
        self.F = int(nrows.item())          # the one host sync: sizes the outputs
        self.rowvox = self.rowvox[:self.F]
        F = max(self.F, 1)
        # padded layout [F][psz*psy][rsg] with zero guards (csrc/ppp_common.cuh)
        rsg = ((int(self.ps[2]) + 16 + 3) // 4) * 4
        self.dp = torch.zeros((F, int(self.ps[0] * self.ps[1]) * rsg), dtype=torch.float32,
                              device=self.dev) if want_dp else None
        self.fcmask = torch.empty((F, self.W), dtype=torch.int32, device=self.dev) \
            if want_masks else None
        self.rbits = None
        self.rv = self.rb16 = None
        # small windows (the 7^3 flylight patches): "received" tables instead of rbits
        self.small = (int(self.ps[2]) <= 8 and int(self.ps[0] * self.ps[1]) <= 64 and
                      int(self.kwargs.get('ppp_consensus_impl', 0)) in (0, 4))
        if want_rbits is None:
            want_rbits = not self.small
        if want_rv is None:
            want_rv = self.small
        if want_dp and want_rbits and int(self.ps[2]) <= 64:
            self.rbits = torch.empty((int(self.ps[0] * self.ps[1]), F, 2), dtype=torch.int64,
                                     device=self.dev)
        if want_dp or want_masks or self.rbits is not None:
            if self.rows is not None:
                cc.call('ppp_prepare_rows', cc.ptr(self.rows.patches),
                        cc.ptr(self.rows.vox2row), cc.ptr(self.flags), cc.ptr(self.rowvox),
                        self.F, self.cfg, cc.ptr(self.dp), cc.ptr(self.fcmask), None,
                        cc.ptr(self.rbits), self.stream)
            else:
                cc.call('ppp_prepare_patches', cc.ptr(self.pred), cc.ptr(self.flags),
                        cc.ptr(self.rowvox), self.F, self.cfg, cc.ptr(self.dp),
                        cc.ptr(self.fcmask), None, cc.ptr(self.rbits), self.stream)
        if want_rv and self.small:
            self.received()
        self._prepared = True
        return self.F

    def window_rows(self, centres, half):
        """u8 [F]: 1 for the rows whose voxel lies within `half` (per axis) of one of the
        `centres` (i32 [m,3] device tensor, block coordinates)."""
        torch = _torch()
        need = torch.zeros(max(self.F, 1), dtype=torch.uint8, device=self.dev)
        cc.call('ppp_mark_windows', cc.ptr(centres.contiguous()), int(centres.shape[0]), self.cfg,
                int(half[0]), int(half[1]), int(half[2]), cc.ptr(self.fgidx), cc.ptr(need),
                self.stream)
        return need

    def received(self, need=None):
        """the "received" tables of the small-window consensus (ppp_received); need: u8
        [F], rows with 0 are skipped (must cover the wanted rows and their partners)."""
        torch = _torch()
        F = max(self.F, 1)
        w16 = int(cc.call('ppp_received_row_words', self.cfg))
        self.rv = torch.empty((F, self.P), dtype=torch.float32, device=self.dev)
        self.rb16 = torch.empty((F, w16), dtype=torch.int16, device=self.dev)
        if self.rows is not None:
            cc.call('ppp_received_rows', cc.ptr(self.rows.patches), cc.ptr(self.rows.vox2row),
                    cc.ptr(self.flags), cc.ptr(self.rowvox), cc.ptr(need), self.F, self.cfg,
                    cc.ptr(self.rv), cc.ptr(self.rb16), self.stream)
        else:
            cc.call('ppp_received', cc.ptr(self.pred), cc.ptr(self.flags), cc.ptr(self.rowvox),
                    cc.ptr(need), self.F, self.cfg, cc.ptr(self.rv), cc.ptr(self.rb16),
                    self.stream)

    # -- step 1 ------------------------------------------------------------
    def consensus(self, want_cnt=False, impl=None, need=None):
        """create_consensus_array_cuda (consensus_array.py:71-206).

        impl 0 = automatic (received tables for psx <= 8, bit-guided gather for
        psx < 16, tiled TMA kernel otherwise), 1 = simple gather (cross-check),
        2 / 3 force bit-guided / tiled, 4 = received tables.
        need: u8 [F] device tensor, rows with 0 are skipped (impl 0/4 on small
        windows only; the other kernels compute every row)."""
        torch = _torch()
        if not self._prepared:
            self.prepare()
        if impl is None:
            impl = int(self.kwargs.get('ppp_consensus_impl', 0))
        F = max(self.F, 1)
        self.cons = torch.empty((F, self.K), dtype=torch.float32, device=self.dev)
        self.cnt = torch.empty((F, self.K), dtype=torch.int32, device=self.dev)
        if self.rv is not None and impl in (0, 4):
            if cc.profile_hook is not None:
                cc.profile_hook('consensus_rows',
                                self.F if need is None else int(need.sum().item()))
            cc.call('ppp_consensus_small', cc.ptr(self.rv), cc.ptr(self.rb16),
                    cc.ptr(self.flags), cc.ptr(self.fgidx), cc.ptr(self.rowvox),
                    cc.ptr(need), self.F, self.cfg, cc.ptr(self.cons), cc.ptr(self.cnt),
                    self.stream)
            return self.cons
        Z, Y, X = self.shape
        if self.rbits is None or (impl in (0, 3) and int(self.ps[2]) >= 16 and
                                  (X > 2048 or int(self.ps[0] * self.ps[1]) > 128)) \
                or self.K > 65535:
            impl = 1        # the simple gather has none of the tiled kernels' limits
        scratch = torch.empty(cc.call('ppp_consensus_scratch_bytes', self.cfg),
                              dtype=torch.uint8, device=self.dev)
        cc.call('ppp_consensus', cc.ptr(self.dp), cc.ptr(self.rbits), cc.ptr(self.flags),
                cc.ptr(self.fgidx), cc.ptr(self.rowvox), self.F, self.cfg, cc.ptr(self.cons),
                cc.ptr(self.cnt), impl, cc.ptr(scratch), self.stream)
        return self.cons

    # -- step 2 ------------------------------------------------------------
    def rank(self):
        """rank_patches_cuda (ranked_patches.py:33-74): score volume."""
        torch = _torch()
        self.score = torch.empty(self.shape, dtype=torch.float32, device=self.dev)
        scratch = torch.empty(cc.call('ppp_rank_scratch_bytes', self.cfg, self.F),
                              dtype=torch.uint8, device=self.dev)
        cc.call('ppp_rank', cc.ptr(self.dp), cc.ptr(self.flags), cc.ptr(self.fgidx),
                cc.ptr(self.rowvox), self.F, cc.ptr(self.cons), self.cfg,
                cc.ptr(self.score), cc.ptr(scratch), self.stream)
        return self.score

    def candidates(self):
        """vote_instances.py:276-287: interior foreground voxels, raster order."""
        torch = _torch()
        m = (self.flags & 24) == 24      # CAND | INTERIOR
        return torch.nonzero(m).flatten().to(torch.int32)

    def ranked(self, cand=None, score=None):
        """rank_patches_by_score (ranked_patches.py:21-30): candidate voxel
        indices, best score first, ties in raster order."""
        torch = _torch()
        cand = self.candidates() if cand is None else cand
        score = self.score if score is None else score
        n = int(cand.numel())
        order = torch.empty(max(n, 1), dtype=torch.int32, device=self.dev)
        scratch = torch.empty(cc.call('ppp_rank_sort_scratch_bytes', n), dtype=torch.uint8,
                              device=self.dev)
        self._latency(lambda st: cc.call('ppp_rank_sort', cc.ptr(score), cc.ptr(cand), n,
                                         cc.ptr(order), cc.ptr(scratch), st))
        return order[:n]

    # -- steps 3+4 ---------------------------------------------------------
    def cover(self, mask, order):
        """computeForegroundCover (foreground_cover.py:15-180)."""
        torch = _torch()
        # score_threshold (foreground_cover.py:136-138): the walk stops at the first ranked
        # patch whose score is below it -- scores are sorted, so the list is cut there
        if isinstance(self.kwargs.get('score_threshold', False), float) and order.numel():
            sc = self.score.reshape(-1)[order.long()].double()
            order = order[:int((sc >= self.kwargs['score_threshold']).sum().item())].contiguous()
        n = int(order.numel())
        if n == 0:
            return order
        if self.kwargs.get('select_patches_for_sparse_data', False) and \
                not self.kwargs.get('ppp_cover_serial', False):
            pix, pix_t = [], None            # threshold 0 only: data-parallel form
        else:
            if self.kwargs.get('select_patches_for_sparse_data', False):
                pix = [0]
            else:
                mid = int(self.P / 2)
                pix = [t for t in [500, 100, 50, 10, 0] if t < mid]
            pix_t = torch.tensor(pix, dtype=torch.int32, device=self.dev)
        selected = torch.zeros(n, dtype=torch.uint8, device=self.dev)
        scratch = torch.empty(cc.call('ppp_cover_scratch_bytes', self.cfg),
                              dtype=torch.uint8, device=self.dev)
        mask = mask.to(torch.uint8).contiguous()
        self._latency(lambda st: cc.call(
            'ppp_cover', cc.ptr(mask), cc.ptr(self.overlap), cc.ptr(order), n,
            cc.ptr(self.fgidx), cc.ptr(self.fcmask), self.cfg, cc.ptr(pix_t),
            len(pix), cc.ptr(selected), cc.ptr(scratch), st))
        return order[selected.bool()]

    def thin(self, mask, sel):
        """thinOutForegroundCover (foreground_cover.py:183-256)."""
        torch = _torch()
        m = int(sel.numel())
        if m == 0:
            return sel
        keep = torch.zeros(m, dtype=torch.uint8, device=self.dev)
        scratch = torch.empty(cc.call('ppp_thin_scratch_bytes', self.cfg, m),
                              dtype=torch.uint8, device=self.dev)
        mask = mask.to(torch.uint8).contiguous()
        sel = sel.contiguous()
        self._latency(lambda st: cc.call(
            'ppp_thin', cc.ptr(mask), cc.ptr(sel), m, cc.ptr(self.fgidx),
            cc.ptr(self.fcmask), self.cfg, cc.ptr(keep), cc.ptr(scratch), st))
        return sel[keep.bool()]

    # -- step 4b (host) ------------------------------------------------------
    def coords(self, vox):
        """voxel indices (tensor) -> int64 [n,3] numpy (z,y,x)."""
        v = vox.cpu().numpy().astype(np.int64)
        Z, Y, X = self.shape
        return np.stack([v // (Y * X), (v // X) % Y, v % X], axis=1)

    def patch_pairs(self, sel_coords):
        """computeAndStorePatchPairs (aff_patch_graph.py:43-110).

        `sel_coords` int [m,3] in ranked order.  Returns u32 [n,6] numpy in
        the reference's order (the python-set iteration order of
        cKDTree.query_pairs, then the self pairs) or None."""
        kw = self.kwargs
        ps = self.ps
        m = len(sel_coords)
        if m == 0:
            return None
        ordr = np.argsort(sel_coords[:, 2], kind='stable')       # :45
        pts = sel_coords[ordr].astype(np.uint32)
        tree = scipy.spatial.cKDTree(pts, leafsize=4)
        max_ps = kw.get("max_total_patch_distance_in_ps_multiples", 2)
        pa = query_pairs_filtered(tree, pts, 2 * np.sum(ps),     # :57, :61-69
                                  np.asarray(max_ps * ps, np.float64))
        single = bool(kw.get('includeSinglePatchCCS', False))
        total = len(pa) + (m if single else 0)
        if total == 0:
            return None
        arr = np.zeros((total, 6), np.uint32)
        arr[:len(pa), :3] = pts[pa[:, 0]]
        arr[:len(pa), 3:] = pts[pa[:, 1]]
        if single:
            arr[len(pa):, :3] = pts
            arr[len(pa):, 3:] = pts
        return arr

    # -- step 5 ------------------------------------------------------------
    def patch_graph(self, pairs_dev, fast=None, pair_org=None):
        """computePatchGraph_cuda (aff_patch_graph.py:113-187), matrix form.
        fast=True: parallel double-precision sum instead of the reference's
        serial float order (kwargs ppp_graph_fast)."""
        torch = _torch()
        n = int(pairs_dev.shape[0])
        aff = torch.empty(max(n, 1), dtype=torch.float32, device=self.dev)
        cfg = self.cfg
        if fast is not None:
            cfg = cc.make_cfg(self.shape, self.ps, **dict(self.kwargs, ppp_graph_fast=fast))
        scratch = torch.empty(cc.call('ppp_patch_graph_scratch_bytes', cfg, n),
                              dtype=torch.uint8, device=self.dev)
        if self.rows is not None:
            cc.call('ppp_patch_graph_rows', cc.ptr(self.rows.patches), cc.ptr(self.rows.vox2row),
                    cc.ptr(self.flags), cc.ptr(self.fgidx), cc.ptr(self.cons), cc.ptr(pairs_dev),
                    cc.ptr(pair_org), n, cfg, cc.ptr(aff), cc.ptr(scratch), self.stream)
        else:
            assert pair_org is None, "pair_org needs the compact row form"
            cc.call('ppp_patch_graph', cc.ptr(self.pred), cc.ptr(self.flags),
                    cc.ptr(self.fgidx), cc.ptr(self.cons), cc.ptr(pairs_dev), n, cfg,
                    cc.ptr(aff), cc.ptr(scratch), self.stream)
        return aff[:n]

    # -- step 6 ------------------------------------------------------------
    def label(self, pairs_dev, aff, nodes, pred=None, cfg=None, mws=False, per_channel=False):
        """setAffgraph + affGraphToInstances (aff_patch_graph.py:31-40,
        graph_to_labeling.py:44-84).  Returns (instances i32 [Z,Y,X], n_comp).
        mws: partition by mutex watershed (graph_mws.py) instead of the
        components over aff > 0.  per_channel: one channel per component
        (`one_instance_per_channel`), instances i32 [n_comp,Z,Y,X]."""
        torch = _torch()
        cfg = cfg or self.cfg
        pred = self.pred if pred is None else pred
        n = int(pairs_dev.shape[0])
        V = self.V
        if mws:
            node_vox, node_label, top = mutex_watershed(
                pairs_dev.cpu().numpy(), aff.cpu().numpy(), cfg)
            comp = torch.zeros(V, dtype=torch.int32, device=self.dev)
            nodes = torch.from_numpy(node_vox).to(self.dev)
            comp[nodes.long()] = torch.from_numpy(node_label).to(self.dev)
            return self._paint(pred, nodes, comp, cfg, top, per_channel), top
        comp = torch.empty(V, dtype=torch.int32, device=self.dev)
        ncomp = torch.zeros(1, dtype=torch.int32, device=self.dev)
        scratch = torch.empty(cc.call('ppp_label_scratch_bytes', V, n), dtype=torch.uint8,
                              device=self.dev)
        cc.call('ppp_label_cc', cc.ptr(pairs_dev), cc.ptr(aff), n, cfg, cc.ptr(comp),
                cc.ptr(ncomp), cc.ptr(scratch), self.stream)
        n_comp = int(ncomp.item())
        return self._paint(pred, nodes.contiguous(), comp, cfg, n_comp, per_channel), n_comp

    def _paint(self, pred, nodes, comp, cfg, n_comp, per_channel):
        torch = _torch()
        if pred is None:                 # compact rows
            if per_channel:
                raise NotImplementedError("one_instance_per_channel needs the dense input form")
            inst = torch.zeros(self.shape, dtype=torch.int32, device=self.dev)
            nl = nodes.long()
            Y, X = self.shape[1], self.shape[2]
            zyx = torch.stack([nl // (Y * X), (nl // X) % Y, nl % X], dim=1).to(torch.int32)
            node_row = self.rows.vox2row.reshape(-1)[nl].contiguous()
            node_label = comp[nl].contiguous()
            cc.call('ppp_paint_rows', cc.ptr(self.rows.patches), cc.ptr(node_row),
                    cc.ptr(zyx.contiguous()), cc.ptr(node_label), int(nodes.numel()), cfg,
                    cc.ptr(inst), self.stream)
            return inst
        if per_channel:
            inst = torch.zeros((n_comp,) + self.shape, dtype=torch.int32, device=self.dev)
            if n_comp > 0:
                cc.call('ppp_paint_channels', cc.ptr(pred), cc.ptr(nodes), int(nodes.numel()),
                        cc.ptr(comp), cfg, cc.ptr(inst), self.stream)
            return inst
        inst = torch.zeros(self.shape, dtype=torch.int32, device=self.dev)
        cc.call('ppp_paint', cc.ptr(pred), cc.ptr(nodes), int(nodes.numel()), cc.ptr(comp),
                cfg, cc.ptr(inst), self.stream)
        return inst
