"""Blockwise instance assembly and cross-block stitching — the B200 counterpart
of PatchPerPix/vote_instances/stitch_patch_graph.py.

Same decomposition as the reference:
  * blocks on the `get_offsets` grid (stitch_patch_graph.py:425-440), each read
    with a `patchshape // 2` halo clipped at the volume edge (:575-607) and
    assembled independently up to the patch graph (`do_block` with
    return_intermediates);
  * for every shared face the selected patches of both blocks near the face
    are paired (cKDTree, :210-248), the region around them is re-assembled and
    the cross edges are computed (:261-336);
  * one global connected-components pass + painting (:360-399).

What changes: blocks and face jobs are independent, so they are dealt
round-robin over the ranks of a torch.distributed job (one process per GPU;
the reference runs them serially behind a lock).  The only exchanges are
  (1) all-gather of the per-block selected patches (needed by the face jobs),
  (2) all-gather of the edge lists,
after which every rank runs the same deterministic union-find
(ppp_label_cc) on the full edge list in the reference's order, paints the
patches of its own blocks and the label volumes are max-reduced (NCCL).
Arrays replace the zarr cache as the medium; `main(pred_file, ...)` keeps the
reference's file-level signature and needs the optional zarr package.
"""
import logging
import os

import numpy as np
import scipy.spatial

from . import vote_instances as vi
from .utilVoteInstances import numinst_from_prob

logger = logging.getLogger(__name__)


def get_offsets(total_shape, chunksize):
    """stitch_patch_graph.py:425-440 (raster order)."""
    offs = []
    if len(total_shape) == 2:
        for y in range(0, total_shape[0], chunksize[0]):
            for x in range(0, total_shape[1], chunksize[1]):
                offs.append(np.array([y, x]))
    elif len(total_shape) == 3:
        for z in range(0, total_shape[0], chunksize[0]):
            for y in range(0, total_shape[1], chunksize[1]):
                for x in range(0, total_shape[2], chunksize[2]):
                    offs.append(np.array([z, y, x]))
    else:
        raise NotImplementedError
    return offs


def get_offset_str(offset):
    return "_".join(str(off) for off in offset)


def load_region(arr, start, stop):
    """arr[..., z0:z1, y0:y1, x0:x1] clipped to the volume (load_input with
    padding=False, stitch_patch_graph.py:443-513).  Returns (data, start_clipped)."""
    shape = arr.shape[-3:]
    s = np.maximum(np.asarray(start), 0)
    e = np.minimum(np.asarray(stop), shape)
    sl = tuple(slice(int(a), int(b)) for a, b in zip(s, e))
    return arr[(Ellipsis,) + sl], s


class VolumeInputs:
    """host-side view of one prediction volume.

    pred        [P,Z,Y,X] float16/float32 (numpy, memmap or torch CPU tensor)
    numinst_prob [C,Z,Y,X] or None; fg [Z,Y,X] / [1,Z,Y,X] or None
    """

    def __init__(self, pred, numinst_prob=None, fg=None):
        if len(pred.shape) == 3:                 # 2-D prediction [P,Y,X]: lifted to Z = 1
            pred = pred[:, None]                 # (stitch_patch_graph.py:766-767)
            if numinst_prob is not None and len(numinst_prob.shape) == 3:
                numinst_prob = numinst_prob[:, None]
            if fg is not None and len(fg.shape) == 2:
                fg = fg[None]
        self.pred = pred
        self.numinst_prob = numinst_prob
        self.fg = fg
        self.shape = tuple(int(s) for s in pred.shape[-3:])
        self.device_pred = None

    def to_device(self, max_fraction=0.4):
        """keep the whole prediction volume resident in HBM when it fits (blocks,
        face regions and the patch vectors of the final painting are then sliced
        on the device instead of being gathered with strided host copies).
        Returns True if the volume is resident."""
        import torch
        if self.device_pred is not None:
            return True
        if not torch.cuda.is_available():
            return False
        free, _ = torch.cuda.mem_get_info()
        nbytes = int(np.prod(self.pred.shape)) * np.dtype(self.pred.dtype).itemsize
        if nbytes > max_fraction * free:
            return False
        self.device_pred = torch.from_numpy(np.ascontiguousarray(self.pred)).cuda()
        return True

    def _fg_numinst(self, start, stop, **kwargs):
        """foreground / numinst of a region: fg_key if set, else numinst > 0, else
        the centre channel (returnFg, utilVoteInstances.py:306-322); numinst from
        numinst_key whenever it is set (stitch_patch_graph.py:610-637)."""
        from .utilVoteInstances import resolve_foreground
        numinst = fg = mid = None
        if kwargs.get("numinst_key") is not None and self.numinst_prob is not None:
            prob, _ = load_region(self.numinst_prob, start, stop)
            numinst = numinst_from_prob(np.asarray(prob), **kwargs)
        if kwargs.get("fg_key") is not None and self.fg is not None:
            fg, _ = load_region(self.fg, start, stop)
            fg = np.asarray(fg)
        elif numinst is None:
            m = int(np.prod(kwargs['patchshape'])) // 2
            mid, _ = load_region(self.pred[m:m + 1], start, stop)
            mid = np.asarray(mid)[0]
        foreground = resolve_foreground(fg=fg, numinst=numinst, mid=mid, **kwargs)
        if foreground.ndim == 2:
            foreground = foreground[None]
        if numinst is None:
            numinst = np.copy(foreground)
        return foreground, numinst

    def foreground(self, **kwargs):
        return self._fg_numinst(np.zeros(3, int), np.asarray(self.shape), **kwargs)[0]

    def region(self, start, stop, **kwargs):
        """(block, foreground bool, mask, numinst, start_clipped) of
        blockwise_vote_instances (stitch_patch_graph.py:603-637)."""
        if self.device_pred is not None:
            block, s = load_region(self.device_pred, start, stop)
            block = block.contiguous()
        else:
            block, s = load_region(self.pred, start, stop)
            block = np.ascontiguousarray(block)
        foreground, numinst = self._fg_numinst(start, stop, **kwargs)
        mask = np.copy(foreground)
        return block, foreground, mask, numinst, s


def default_block_fn(block, foreground, mask, numinst, **kwargs):
    """the CUDA path (do_block, vote_instances.py:455)."""
    return vi.do_block(block, foreground, mask, numinst, **kwargs)


def assemble_block(inputs, offset, block_fn=default_block_fn, **kwargs):
    """blockwise_vote_instances (stitch_patch_graph.py:553-669) on arrays.
    Returns (pairs u32 [n,6] in VOLUME coordinates, aff f32 [n]) or (None, None)."""
    ps = np.asarray(kwargs['patchshape'])
    chunksize = np.minimum(np.asarray(kwargs['chunksize']), inputs.shape)
    margin = ps // 2
    offset = np.asarray(offset)
    block, foreground, mask, numinst, start = inputs.region(
        offset - margin, offset + chunksize + margin, **kwargs)
    if not np.any(foreground):
        return None, None
    kw = dict(kwargs)
    kw['return_intermediates'] = True
    pairs, aff = block_fn(block, foreground, mask, numinst, **kw)
    if pairs is None:
        return None, None
    pairs = pairs.astype(np.int64) + np.tile(start, 2)          # :650 and :162
    return pairs.astype(np.uint32), np.asarray(aff, np.float32)


def face_candidates(sel_cur, sel_nb, offset, dim, ps):
    """stitch_patch_graph.py:210-224 for a neighbour on the negative side."""
    cur = sel_cur[sel_cur[:, dim] <= offset[dim] + ps[dim]]
    nb = sel_nb[sel_nb[:, dim] >= offset[dim] - ps[dim]]
    return cur, nb


def face_pairs(cur, nb, ps):
    """stitch_patch_graph.py:236-258: candidate pairs across the face, in the
    reference's (python-set) order.  Returns (candidates, pairs index array)."""
    candidates = np.concatenate([cur, nb])
    tree = scipy.spatial.cKDTree(candidates, leafsize=4)
    from .assembly import query_pairs_set_order
    pa = query_pairs_set_order(tree, 1 * np.sum(ps + 1))
    if len(pa) == 0:
        return candidates, np.zeros((0, 2), np.int64)
    d = np.abs(candidates[pa[:, 0]].astype(np.float32) -
               candidates[pa[:, 1]].astype(np.float32))
    keep = ~np.any(d > ps + 1, axis=1)                           # remove_pairs :75-88
    ncur = len(cur)
    same = (pa[:, 0] < ncur) == (pa[:, 1] < ncur)                # remove_intra_block_pairs
    return candidates, pa[keep & ~same]


def assemble_face(inputs, candidates, pa, block_fn=default_block_fn, **kwargs):
    """stitch_patch_graph.py:252-336: cross edges of one face.
    Returns (pairs u32 [n,6] volume coords, aff f32 [n])."""
    ps = np.asarray(kwargs['patchshape'])
    cleaned = candidates[np.unique(pa.reshape(-1))]
    bb_start = np.maximum(np.min(cleaned, axis=0) - ps, 0)
    bb_stop = np.minimum(np.max(cleaned, axis=0) + ps, inputs.shape)
    margin = ps // 2
    block, foreground, mask, numinst, start = inputs.region(
        bb_start - margin, np.maximum(bb_stop, bb_start + 1) + margin, **kwargs)
    overlapping = np.concatenate([candidates[pa[:, 0]], candidates[pa[:, 1]]], axis=1)
    # The reference makes the coordinates relative to the UNclipped region start
    # (stitch_patch_graph.py:317-321) although the region it loads is clipped at
    # the volume origin, so faces within patchshape//2 of the origin are scored at
    # centres shifted by the clipped amount.  Reproduced by default (the affinities
    # and, with mws, the label ids depend on it); ppp_fix_face_origin=True uses the
    # start of the region actually loaded.
    origin = start if kwargs.get('ppp_fix_face_origin', False) else bb_start - margin
    rel_c = cleaned - origin
    rel_p = overlapping - np.tile(origin, 2)
    kw = dict(kwargs)
    kw.update(skipRanking=True, skipThinCover=True, return_intermediates=True)
    _, aff = block_fn(block, foreground, mask, numinst, selected_patches=rel_c,
                      selected_patch_pairs=rel_p.astype(np.uint32), **kw)
    return overlapping.astype(np.uint32), np.asarray(aff, np.float32)


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def _allgather_obj(obj):
    dist = _dist()
    if dist is None:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def stitch_arrays(inputs, block_fn=default_block_fn, paint_fn=None, **kwargs):
    """blockwise assembly + stitching of one volume (stitch_vote_instances,
    stitch_patch_graph.py:110-399, plus the block loop of main :805-813).

    Returns (instances uint32 [Z,Y,X], foreground bool [Z,Y,X], info dict).
    Under torch.distributed every rank returns the same arrays."""
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    ps = np.asarray(kwargs['patchshape'])
    shape = inputs.shape
    # only_bb (stitch_patch_graph.py:745-764, 137-138): the block grid covers the
    # bounding box of the foreground and starts at its corner
    bb_offset = np.asarray(kwargs.pop('bb_offset', np.zeros(3, int)), dtype=int)
    bb_shape = np.asarray(kwargs.pop('bb_shape', shape), dtype=int)
    chunksize = np.minimum(np.asarray(kwargs['chunksize']), bb_shape)
    kwargs = dict(kwargs, chunksize=chunksize)
    offsets = [o + bb_offset for o in get_offsets(bb_shape, chunksize)]
    nblk = len(offsets)
    if block_fn is default_block_fn and kwargs.get('ppp_device_volume', True):
        inputs.to_device()               # whole volume in HBM when it fits

    # block cache (stitch_patch_graph.py:584-587, 649-669): results of earlier runs are
    # kept under volumes/blocks/<z_y_x>/{patch_pairs, aff_graph_mat} (pairs relative to
    # the block offset) and, per face, under <block>/<neighbour>/ (volume coordinates)
    cache = kwargs.pop('block_cache', None)
    tmp_key = 'volumes/blocks'

    def cached(key):
        if cache is not None and key + '/patch_pairs' in cache:
            return (np.array(cache[key + '/patch_pairs']).astype(np.uint32).reshape(-1, 6),
                    np.array(cache[key + '/aff_graph_mat']).astype(np.float32))
        return None

    def store(key, pairs, aff):
        if cache is not None and pairs is not None:
            cache.create_dataset(key + '/patch_pairs', data=np.asarray(pairs, np.uint32),
                                 overwrite=True)
            cache.create_dataset(key + '/aff_graph_mat', data=np.asarray(aff, np.float32),
                                 overwrite=True)

    # ---- phase 1: blocks, round-robin over the ranks (offsets.py:45) -----------
    mine = {}
    for b in range(rank, nblk, world):
        key = tmp_key + '/' + get_offset_str(offsets[b])
        got = cached(key)
        if got is not None:
            mine[b] = ((got[0].astype(np.int64) + np.tile(offsets[b], 2)).astype(np.uint32),
                       got[1])
            continue
        mine[b] = assemble_block(inputs, offsets[b], block_fn, **kwargs)
        if mine[b][0] is not None:
            store(key, mine[b][0].astype(np.int64) - np.tile(offsets[b], 2), mine[b][1])
    # exchange (1): per-block pairs (selected patches are derived from them)
    blocks = {}
    for part in _allgather_obj(mine):
        blocks.update(part)
    selected = {}
    for b in range(nblk):
        p = blocks[b][0]
        selected[b] = None if p is None else \
            np.unique(p.reshape(-1, 3).astype(np.int64), axis=0)  # :166-171

    # ---- phase 2: face jobs ------------------------------------------------------
    key_of = {tuple(int(v) for v in o): i for i, o in enumerate(offsets)}
    jobs = []      # (block, neighbour) in the reference's visiting order
    first_nonempty = next((b for b in range(nblk) if blocks[b][0] is not None), None)
    for b in range(nblk):
        if blocks[b][0] is None or b == first_nonempty:
            continue                                             # :151-156, :178-181
        for dim in range(3):                                     # -z, -y, -x (:125-128)
            nb_off = offsets[b].copy()
            nb_off[dim] -= chunksize[dim]
            nb = key_of.get(tuple(int(v) for v in nb_off))
            if nb is None or nb >= b or selected[nb] is None:
                continue
            jobs.append((b, nb, dim))
    my_faces = {}
    for j in range(rank, len(jobs), world):
        b, nb, dim = jobs[j]
        fkey = tmp_key + '/' + get_offset_str(offsets[b]) + '/' + get_offset_str(offsets[nb])
        got = cached(fkey)
        if got is not None:
            my_faces[j] = got
            continue
        cur, nbc = face_candidates(selected[b], selected[nb], offsets[b], dim, ps)
        if len(cur) == 0 or len(nbc) == 0:
            continue
        cands, pa = face_pairs(cur, nbc, ps)
        if len(pa) == 0:
            continue
        my_faces[j] = assemble_face(inputs, cands, pa, block_fn, **kwargs)
        store(fkey, *my_faces[j])
    faces = {}
    for part in _allgather_obj(my_faces):                        # exchange (2)
        faces.update(part)

    # ---- global edge list in the reference's order (update_graph calls) ----------
    plist, alist = [], []
    jidx = {}
    for j, (b, nb, dim) in enumerate(jobs):
        jidx.setdefault(b, []).append(j)
    for b in range(nblk):
        if blocks[b][0] is None:
            continue
        plist.append(blocks[b][0])
        alist.append(blocks[b][1])
        for j in jidx.get(b, []):
            if j in faces:
                plist.append(faces[j][0])
                alist.append(faces[j][1])
    info = dict(n_blocks=nblk, n_faces=len(jobs),
                n_edges=int(sum(len(a) for a in alist)))
    foreground = inputs.foreground(**kwargs) if kwargs.get('want_foreground', True) else None
    if not plist:
        return np.zeros(shape, np.uint32), foreground, info
    pairs = np.concatenate(plist).astype(np.uint32)
    aff = np.concatenate(alist).astype(np.float32)
    info['pairs'] = pairs
    info['aff'] = aff
    # ---- phase 3: replicated union-find + painting --------------------------------
    paint = paint_fn or paint_global
    inst = paint(inputs, pairs, aff, rank, world, **kwargs)
    return inst, foreground, info


def paint_global(inputs, pairs, aff, rank=0, world=1, **kwargs):
    """connected components (or, with kwargs mws, the mutex watershed) of the
    global graph and painting of the member patches (affGraphToInstances with
    sparse_labels, stitch_patch_graph.py:380-396).  Each rank paints the nodes
    i = rank mod world; the volumes are max-reduced."""
    import torch
    from . import cuda_code as cc
    ps = np.asarray(kwargs['patchshape'])
    shape = inputs.shape
    dev = torch.device('cuda', torch.cuda.current_device())
    cfg = cc.make_cfg(shape, ps, **{k: v for k, v in kwargs.items() if k != 'patchshape'})
    V = int(np.prod(shape))
    stream = cc.current_stream_ptr()
    pd = torch.from_numpy(pairs.view(np.int32)).to(dev)
    ad = torch.from_numpy(aff).to(dev)
    n = len(pairs)
    if kwargs.get('mws', False):
        from .assembly import mutex_watershed
        node_vox, node_label, _ = mutex_watershed(pairs, aff, cfg)
        comp = torch.zeros(V, dtype=torch.int32, device=dev)
        comp[torch.from_numpy(node_vox).to(dev).long()] = torch.from_numpy(node_label).to(dev)
    else:
        comp = torch.empty(V, dtype=torch.int32, device=dev)
        ncomp = torch.zeros(1, dtype=torch.int32, device=dev)
        scratch = torch.empty(cc.call('ppp_label_scratch_bytes', V, n), dtype=torch.uint8,
                              device=dev)
        cc.call('ppp_label_cc', cc.ptr(pd), cc.ptr(ad), n, cfg, cc.ptr(comp), cc.ptr(ncomp),
                cc.ptr(scratch), stream)
    Y, X = shape[1], shape[2]
    p64 = pairs.astype(np.int64)
    nodes = np.unique(np.concatenate([(p64[:, 0] * Y + p64[:, 1]) * X + p64[:, 2],
                                      (p64[:, 3] * Y + p64[:, 4]) * X + p64[:, 5]]))
    nodes = nodes[rank::world]
    P = int(np.prod(ps))
    # gather the patch vectors of my nodes from the host volume (:380-385)
    z, y, x = nodes // (Y * X), (nodes // X) % Y, nodes % X
    inst = torch.zeros(shape, dtype=torch.int32, device=dev)
    if len(nodes):
        nd = torch.from_numpy(nodes.astype(np.int32)).to(dev)
        if getattr(inputs, 'device_pred', None) is not None:
            zi, yi, xi = (torch.from_numpy(a).to(dev) for a in (z, y, x))
            pt = inputs.device_pred[:, zi, yi, xi].T.float().contiguous()
        else:
            pt = torch.from_numpy(gather_patches(inputs.pred, z, y, x)).to(dev)
        cc.call('ppp_paint_patches', cc.ptr(pt), cc.ptr(nd), len(nodes), cc.ptr(comp), cfg,
                cc.ptr(inst), stream)
    dist = _dist()
    if dist is not None and world > 1:
        dist.all_reduce(inst, op=dist.ReduceOp.MAX)
    return inst.cpu().numpy().astype(np.uint32)


def bounding_box(inputs, **kwargs):
    """(bb_offset, bb_shape) of stitch_patch_graph.py:745-771; None: no foreground."""
    from .postprocess import foreground_bbox
    if not kwargs.get('only_bb'):
        return np.zeros(3, int), np.asarray(inputs.shape)
    return foreground_bbox(inputs.foreground(**kwargs), **kwargs)


def finish_outputs(instances, foreground, **kwargs):
    """stitch_patch_graph.py:824-894 on the device: optional removal of small
    instances + relabelling, the masked volume, optional dilated volumes.
    instances: int tensor / array [Z,Y,X]; returns {dataset name: uint16 numpy}."""
    import torch
    from . import postprocess as pp
    res_key = kwargs.get('res_key', 'vote_instances')
    dev = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
    inst = torch.as_tensor(np.asarray(instances).astype(np.int64)
                           if not torch.is_tensor(instances) else instances).to(dev)
    fg = torch.as_tensor(np.squeeze(np.asarray(foreground)) != 0).to(dev)
    if fg.dim() == 2:
        fg = fg[None]
    if kwargs.get('remove_small_comps', 0) > 0:
        inst = pp.relabel(pp.remove_small_components(inst, kwargs['remove_small_comps']))

    def u16(t):
        return t.cpu().numpy().astype(np.uint16)
    out = {res_key: u16(inst), 'vote_foreground': u16(fg),
           'vote_instances_masked': u16(torch.where(fg, inst, torch.zeros_like(inst)))}
    if kwargs.get('dilate_instances', False):
        dil = pp.dilate_instances(inst)
        out[res_key + '_dil_1'] = u16(dil)
        out[res_key + '_masked_dil_1'] = u16(torch.where(fg, dil, torch.zeros_like(dil)))
    return out


def gather_patches(pred, z, y, x, tile=64):
    """pred[:, z, y, x].T as float32 [n,P] without loading the volume: the nodes are
    grouped by tile and one box per tile is read (the reference reads node by node when
    the prediction is larger than 20 GB, stitch_patch_graph.py:367-385)."""
    n = len(z)
    out = np.zeros((n, int(pred.shape[0])), np.float32)
    if n == 0:
        return out
    if isinstance(pred, np.ndarray) and not isinstance(pred, np.memmap):
        return np.ascontiguousarray(pred[:, z, y, x].T.astype(np.float32))
    key = (z // tile) * 1000003 + (y // tile) * 1009 + (x // tile)
    order = np.argsort(key, kind='stable')
    bounds = np.flatnonzero(np.diff(key[order])) + 1
    for idx in np.split(order, bounds):
        z0, y0, x0 = z[idx].min(), y[idx].min(), x[idx].min()
        box = np.asarray(pred[:, int(z0):int(z[idx].max()) + 1, int(y0):int(y[idx].max()) + 1,
                              int(x0):int(x[idx].max()) + 1])
        out[idx] = box[:, z[idx] - z0, y[idx] - y0, x[idx] - x0].T
    return out


def _main_rows(inputs, bb, **kwargs):
    """the volume as compact rows, one slab per rank, through sharded.stitch_shard.
    Returns (instances u32 [Z,Y,X] on rank 0 (zeros elsewhere), foreground bool)."""
    import torch
    from . import sharded
    from .utilVoteInstances import getFgThreshold
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    dev = torch.device('cuda', torch.cuda.current_device())
    shape = inputs.shape
    foreground, numinst = inputs._fg_numinst(np.zeros(3, int), np.asarray(shape), **kwargs)
    axis, _ = sharded.slab_partition(shape, kwargs['chunksize'], world, bb_offset=bb[0],
                                     bb_shape=bb[1])
    csz = int(min(kwargs['chunksize'][axis], bb[1][axis]))
    per_plane = foreground.sum(axis=tuple(a for a in range(3) if a != axis))
    o = int(bb[0][axis])
    weights = [int(per_plane[o + i:o + i + csz].sum()) for i in range(0, int(bb[1][axis]), csz)]
    axis, slabs = sharded.slab_partition(shape, kwargs['chunksize'], world, axis=axis,
                                         weights=weights, bb_offset=bb[0], bb_shape=bb[1])
    lo, hi = slabs[rank]
    th = float(np.float32(getFgThreshold(**kwargs)))
    c, p, ni, fgr = sharded.rows_from_dense(inputs.pred, foreground, numinst, axis, lo, hi, dev, th)
    shard = sharded.RowShard(shape, axis, lo, hi, c, p, ni, fgr)
    inst, _ = sharded.stitch_shard(shard, slabs, bb_offset=bb[0], bb_shape=bb[1], **kwargs)
    full = sharded.gather_slabs(inst, slabs, axis, shape)
    if full is None:
        return np.zeros(shape, np.uint32), foreground
    return full.cpu().numpy().astype(np.uint32), foreground


def main(pred_file, result_folder='.', **kwargs):
    """stitch_patch_graph.py:672-896: file-level entry point (zarr in,
    `<sample>.hdf` / `.npz` out).  Under torch.distributed every rank works, rank 0
    writes."""
    from .io_util import open_container, write_result
    if kwargs.get('graphToInst', False) or kwargs.get('blockwise_old_stitch_fn', False):
        raise NotImplementedError("stitch_patch_graph options graphToInst / "
                                  "blockwise_old_stitch_fn are outside the B200 hot path")
    assert os.path.exists(pred_file), \
        'Prediction file {} does not exist. Please check!'.format(pred_file)
    sample = os.path.basename(pred_file).split('.')[0]
    kwargs['result_folder'] = result_folder
    in_f = open_container(pred_file, 'r')
    pred = in_f[kwargs['aff_key']]
    numinst_prob = in_f[kwargs['numinst_key']] if kwargs.get('numinst_key') else None
    fg = in_f[kwargs['fg_key']] if kwargs.get('fg_key') else None
    if numinst_prob is not None:
        assert tuple(pred.shape[1:]) == tuple(numinst_prob.shape[1:]), \
            'Please check: affinity and numinst shape do not match!'
    inputs = VolumeInputs(pred, numinst_prob, fg)
    bb = bounding_box(inputs, **kwargs)
    if bb is None:
        logger.info('Volume has no foreground voxel, returning...')
        return
    dist = _dist()
    if kwargs.get('ppp_rows', dist is not None):
        # compact rows, the volume sharded over the ranks (sharded.py): the default under
        # torch.distributed, `ppp_rows=True` selects it for a single process as well
        instances, foreground = _main_rows(inputs, bb, **kwargs)
    else:
        cache = None
        if kwargs.get('ppp_block_cache', True):
            from .io_util import open_zarr
            cache = open_zarr(os.path.join(result_folder, sample + '.zarr'), 'a')
        instances, foreground, _ = stitch_arrays(inputs, bb_offset=bb[0], bb_shape=bb[1],
                                                 block_cache=cache, **kwargs)
    if dist is None or dist.get_rank() == 0:
        os.makedirs(result_folder, exist_ok=True)
        out = finish_outputs(instances, foreground, **kwargs)
        if kwargs.get('save_mip', False):                        # :815-821, 840-845
            from .postprocess import color, write_png
            res = out[kwargs.get('res_key', 'vote_instances')]
            write_png(os.path.join(result_folder, sample + (
                '_cleaned.png' if kwargs.get('remove_small_comps', 0) > 0 else '.png')),
                color(np.max(res, axis=0)))
        write_result(os.path.join(result_folder, sample), out,
                     kwargs.get('output_format', 'hdf'))
    if dist is not None:
        dist.barrier()
    return instances
