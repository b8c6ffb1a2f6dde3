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
from .utilVoteInstances import returnFg, numinst_from_prob

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
        """foreground / numinst of a region, precedence as in the reference's
        returnFg (utilVoteInstances.py:306-322) and stitch_patch_graph.py:610-637."""
        from .utilVoteInstances import getFgThreshold
        th = getFgThreshold(**kwargs)
        numinst = None
        if kwargs.get("numinst_key") is not None and self.numinst_prob is not None:
            prob, _ = load_region(self.numinst_prob, start, stop)
            numinst = numinst_from_prob(np.asarray(prob), **kwargs)
            foreground = (numinst > 0) > th
        elif kwargs.get("fg_key") is not None and self.fg is not None:
            fg, _ = load_region(self.fg, start, stop)
            foreground = np.squeeze(np.asarray(fg)) > th
            if foreground.ndim == 2:
                foreground = foreground[None]
        else:
            mid = int(np.prod(kwargs['patchshape'])) // 2
            m, _ = load_region(self.pred[mid:mid + 1], start, stop)
            foreground = np.asarray(m)[0] > th
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
    chunksize = np.minimum(np.asarray(kwargs['chunksize']), shape)
    kwargs = dict(kwargs, chunksize=chunksize)
    offsets = get_offsets(shape, chunksize)
    nblk = len(offsets)
    if block_fn is default_block_fn and kwargs.get('ppp_device_volume', True):
        inputs.to_device()               # whole volume in HBM when it fits

    # ---- phase 1: blocks, round-robin over the ranks (offsets.py:45) -----------
    mine = {}
    for b in range(rank, nblk, world):
        mine[b] = assemble_block(inputs, offsets[b], block_fn, **kwargs)
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
        cur, nbc = face_candidates(selected[b], selected[nb], offsets[b], dim, ps)
        if len(cur) == 0 or len(nbc) == 0:
            continue
        cands, pa = face_pairs(cur, nbc, ps)
        if len(pa) == 0:
            continue
        my_faces[j] = assemble_face(inputs, cands, pa, block_fn, **kwargs)
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
            patches = np.ascontiguousarray(np.asarray(inputs.pred)[:, z, y, x].T.astype(np.float32))
            pt = torch.from_numpy(patches).to(dev)
        cc.call('ppp_paint_patches', cc.ptr(pt), cc.ptr(nd), len(nodes), cc.ptr(comp), cfg,
                cc.ptr(inst), stream)
    dist = _dist()
    if dist is not None and world > 1:
        dist.all_reduce(inst, op=dist.ReduceOp.MAX)
    return inst.cpu().numpy().astype(np.uint32)


def main(pred_file, result_folder='.', **kwargs):
    """stitch_patch_graph.py:672-896: file-level entry point (zarr in,
    `<sample>.hdf` / `.npz` out)."""
    from .io_util import open_zarr, write_result
    assert os.path.exists(pred_file), \
        'Prediction file {} does not exist. Please check!'.format(pred_file)
    sample = os.path.basename(pred_file).split('.')[0]
    kwargs['result_folder'] = result_folder
    in_f = open_zarr(pred_file, 'r')
    aff_key = kwargs['aff_key']
    pred = in_f[aff_key]
    numinst_prob = in_f[kwargs['numinst_key']] if kwargs.get('numinst_key') else None
    fg = in_f[kwargs['fg_key']] if (numinst_prob is None and kwargs.get('fg_key')) else None
    if kwargs.get('only_bb'):
        raise NotImplementedError("only_bb (bounding-box crop with skeletonisation, "
                                  "stitch_patch_graph.py:745-764) is outside the hot path")
    inputs = VolumeInputs(pred, numinst_prob, fg)
    instances, foreground, _ = stitch_arrays(inputs, **kwargs)
    res_key = kwargs.get('res_key', 'vote_instances')
    os.makedirs(result_folder, exist_ok=True)
    fg16 = np.squeeze(foreground).astype(np.uint16)
    masked = instances.copy()
    masked[fg16 == 0] = 0
    write_result(os.path.join(result_folder, sample),
                 {res_key: instances.astype(np.uint16), 'vote_foreground': fg16,
                  'vote_instances_masked': masked.astype(np.uint16)},
                 kwargs.get('output_format', 'hdf'))
    return instances
