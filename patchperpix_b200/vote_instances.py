"""Drop-in for PatchPerPix/vote_instances/vote_instances.py on the B200 path.

Same entry points and keyword surface as the reference:
    main(**kwargs)                                   vote_instances.py:557
    do_all(aff_file, patchshape, **kwargs)           vote_instances.py:486
    do_block(block, foreground, mask, numinst, **kw) vote_instances.py:455
    to_instance_seg(pred_affs, foreground, mask_to_cover, numinst,
                    patchshape, **kwargs)            vote_instances.py:150
but steps (1)-(6) run on the GPU through include/ppp_b200.h
(patchperpix_b200/assembly.py).  `cuda=False` is refused: this build has no
CPU path.  Arrays may be numpy (copied to the device here — that copy is part
of the end-to-end cost), pinned torch CPU tensors or CUDA tensors.
"""
import glob
import logging
import os

import numpy as np

from . import cuda_code
from .assembly import BlockAssembler, RowSource
from .utilVoteInstances import loadAffinities, getResKey

logger = logging.getLogger(__name__)

_UNSUPPORTED = ('isbiHack', 'debug', 'graphToInst', 'use_score_oracle',
                'mark_close_neighboorhood', 'select_patches_overlap_neighborhood',
                'thin_cover_use_kd', 'shuffle_patches')


def merge_dicts(sink, source):
    """overlay `source` on `sink` in place, descending into sub-dicts present on both
    sides (the call-site helper of vote_instances.py:49-59, kept for callers that import
    it)."""
    if not (isinstance(sink, dict) and isinstance(source, dict)):
        raise TypeError('Args to merge_dicts should be dicts')
    for key, val in source.items():
        both = isinstance(val, dict) and isinstance(sink.get(key), dict)
        sink[key] = merge_dicts(sink[key], val) if both else val
    return sink


def _to_device(a, dtype=None):
    import torch
    if isinstance(a, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(a))
    else:
        t = a
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    t = t.cuda(non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def pack_no_overlap_channels(stack, min_size=2000):
    """no_overlap_per_channel (graph_to_labeling.py:96-113) on the device: `stack` i32
    [n_comp,Z,Y,X] = one channel per component (ppp_paint_channels).  Components larger
    than `min_size` voxels go to the first channel where their voxels are all free, else
    they open a new channel; smaller ones are written into channel 0 (overwriting)."""
    import torch
    out = []
    for k in range(int(stack.shape[0])):
        cur = stack[k]
        if not out:
            out.append(cur.clone())
            continue
        m = cur > 0
        if int(m.sum().item()) > min_size:
            for ch in out:
                if not bool((ch[m] != 0).any().item()):
                    ch[m] = k + 1
                    break
            else:
                out.append(cur.clone())
        else:
            out[0][m] = k + 1
    if not out:
        return stack[:0]
    return torch.stack(out, dim=0)


def to_instance_seg(pred_affs, foreground, mask_to_cover, numinst, patchshape,
                    **kwargs):
    """vote_instances.py:150-452.  Returns (instances u16, foreground u8), or
    (pairs u32 [n,6], aff f32 [n]) with return_intermediates, all numpy."""
    import torch
    if not kwargs.get('cuda', True):
        raise NotImplementedError(
            "patchperpix_b200 only provides the CUDA path (cuda=True)")
    for k in _UNSUPPORTED:
        if kwargs.get(k, False):
            raise NotImplementedError("vote_instances option %r is outside "
                                      "the B200 hot path" % k)
    if float(kwargs.get('sample', 1.0)) < 1.0:
        raise NotImplementedError("vote_instances option 'sample' < 1.0 (random sub-sampling of "
                                  "the patch sets, get_patch_sets.py:52) is outside the B200 hot "
                                  "path")
    patchshape = np.array(patchshape)
    rad = np.array([p // 2 for p in patchshape])
    ret_inter = kwargs.get('return_intermediates', False)

    if isinstance(pred_affs, RowSource):       # compact rows (ppp+dec form), device resident
        pred = pred_affs
        assert not kwargs.get("pad_with_ps", False), "pad_with_ps needs the dense input form"
    else:
        pred = _to_device(pred_affs, torch.float32)
    fg = _to_device(foreground, torch.uint8)
    mask = _to_device(mask_to_cover, torch.uint8).clone()
    numinst_t = _to_device(numinst)
    if kwargs.get("pad_with_ps", False):                       # :166-190
        assert not kwargs.get('blockwise', False), "can only pad whole volumes"
        pad = (int(rad[2]),) * 2 + (int(rad[1]),) * 2 + (int(rad[0]),) * 2
        pred = torch.nn.functional.pad(pred, pad).contiguous()
        fg = torch.nn.functional.pad(fg, pad).contiguous()
        mask = torch.nn.functional.pad(mask, pad).contiguous()
        numinst_t = torch.nn.functional.pad(numinst_t, pad).contiguous()
    shape = tuple(int(s) for s in fg.shape)
    overlap = (numinst_t > 1).to(torch.uint8)                  # :211
    if not kwargs.get('blockwise', False) and kwargs.get('skeletonize_foreground'):
        raise NotImplementedError("skeletonize_foreground needs skimage "
                                  "(non-blockwise only, vote_instances.py:220-224)")
    mask[overlap > 0] = 0                                      # :226
    radslice = tuple(slice(int(rad[i]), shape[i] - int(rad[i])) for i in range(3))

    def _unpad(t):
        if kwargs.get("pad_with_ps", False):
            return t[radslice]
        return t

    def _empty_result():
        if ret_inter:
            return None, None
        inst = torch.zeros(shape, dtype=torch.int32, device=pred.device)
        return (_unpad(inst).cpu().numpy().astype(np.uint16),
                _unpad(fg).cpu().numpy().astype(np.uint8))

    if int(torch.count_nonzero(mask[radslice]).item()) == 0:   # :232-245
        logger.info("no fg found, returning...")
        return _empty_result()

    asm = BlockAssembler(pred, fg, overlap, patchshape, **kwargs)
    asm.prepare()
    cand = asm.candidates()                                    # :276-287
    if cand.numel() == 0:
        logger.info("no patches found, returning...")
        return _empty_result()

    # (1) consensus
    if not kwargs.get('skipConsensus', False):
        need = None
        if kwargs.get('selected_patches') is not None and kwargs.get('skipRanking', False) \
                and kwargs.get('selected_patch_pairs') is not None and asm.small:
            # stitcher's face job (stitch_patch_graph.py:323-328): only the patch graph of the
            # given pairs reads the consensus, i.e. rows inside the windows of those patches
            sc = np.asarray(kwargs['selected_patches'], dtype=np.int64).reshape(-1, 3)
            sc = sc[np.all((sc >= 0) & (sc < np.asarray(shape)), axis=1)]
            need = asm.window_rows(torch.from_numpy(sc.astype(np.int32)).to(asm.dev), rad)
        asm.consensus(need=need)
    if kwargs.get('save_consensus', False):
        return None, None
    # (2) ranking
    order = None
    if not kwargs.get('skipRanking', False):
        asm.rank()
        order = asm.ranked(cand)
        logger.info("num ranked patches %s", int(order.numel()))

    if kwargs.get('aff_graph') is not None:
        raise NotImplementedError("loading a stored patch graph")
    # (3)/(4) selection
    if kwargs.get('selected_patches') is not None:             # :369-375
        sc = np.asarray(kwargs['selected_patches'], dtype=np.int64).reshape(-1, 3)
        sel_coords = sc
    else:
        if order is None:
            raise NotImplementedError("skipRanking needs selected_patches (a stored ranking "
                                      "file is outside the B200 hot path)")
        if kwargs.get('skipSelection', False):
            sel = order
        else:
            sel = asm.cover(mask, order)
        if not kwargs.get('skipThinCover', False) and sel.numel() > 0:
            sel = asm.thin(mask, sel)
        sel_coords = asm.coords(sel)
    # (4b) pairs
    if kwargs.get('selected_patch_pairs') is not None:         # :400-406
        pairs = np.array(kwargs['selected_patch_pairs'], dtype=np.uint32).reshape(-1, 6)
        if len(pairs) == 0:
            pairs = None
    else:
        pairs = asm.patch_pairs(sel_coords)
    if pairs is None:                                          # :413-417
        return _empty_result()
    if kwargs.get('termAfterThinCover', False):
        return None, None
    # (5) patch graph
    # centres handed in by a caller (stitcher) must lie inside the block: the
    # reference would read out of bounds, here such pairs get affinity 0 (no edge)
    inside = np.all(pairs.astype(np.int64) < np.tile(np.asarray(shape, np.int64), 2), axis=1)
    if not inside.all():
        logger.warning("%d patch pairs outside the block: affinity 0", int((~inside).sum()))
        aff = torch.zeros(len(pairs), dtype=torch.float32, device=pred.device)
        if inside.any():
            sub = torch.from_numpy(np.ascontiguousarray(pairs[inside]).view(np.int32)).to(
                pred.device)
            aff[torch.from_numpy(np.nonzero(inside)[0]).to(pred.device)] = asm.patch_graph(sub)
        pairs_dev = torch.from_numpy(pairs.view(np.int32)).to(pred.device)
    else:
        pairs_dev = torch.from_numpy(pairs.view(np.int32)).to(pred.device)
        aff = asm.patch_graph(pairs_dev)
    if kwargs.get("save_patch_graph", False) or kwargs.get("termAfterPatchGraph", False):
        fn = os.path.splitext(os.path.basename(kwargs.get('affinities', 'block')))[0]
        np.save(os.path.join(kwargs['result_folder'], fn + "_selected_patch_pairs.npy"), pairs)
        np.save(os.path.join(kwargs['result_folder'], fn + "_aff_graph.npy"),
                aff.cpu().numpy())
    if ret_inter:
        return pairs, aff.cpu().numpy()
    if kwargs.get('termAfterPatchGraph', False):
        return None, None
    # (6) labelling
    Y, X = shape[1], shape[2]
    nodes_np = np.unique(np.concatenate([
        (pairs[:, 0].astype(np.int64) * Y + pairs[:, 1]) * X + pairs[:, 2],
        (pairs[:, 3].astype(np.int64) * Y + pairs[:, 4]) * X + pairs[:, 5]]))
    nodes = torch.from_numpy(nodes_np.astype(np.int32)).to(pred.device)
    no_overlap = bool(kwargs.get('no_overlap_per_channel', False))
    per_channel = bool(kwargs.get('one_instance_per_channel', False)) or no_overlap
    inst, ncomp = asm.label(pairs_dev, aff, nodes, mws=kwargs.get('mws', False),
                            per_channel=per_channel)
    if ncomp > 65535:
        logger.warning("%d components do not fit the reference's uint16 labels", ncomp)
    if no_overlap:
        inst = pack_no_overlap_channels(inst)
    if per_channel:                                            # graph_to_labeling.py:143-151
        inst = inst[(slice(None),) + radslice] if kwargs.get("pad_with_ps", False) else inst
    else:
        inst = _unpad(inst)
    return (inst.cpu().numpy().astype(np.uint16),
            _unpad(fg).cpu().numpy().astype(np.uint8))


_STREAM_CACHE = {}


def _cached_stream(dev_index, name):
    """side streams are created once per device and role: torch's caching allocator
    keeps one memory pool per stream, a fresh stream per call would re-allocate
    (cudaMalloc) several GB of working buffers every time (~150 ms on the bench image)."""
    import torch
    key = (dev_index, name)
    if key not in _STREAM_CACHE:
        _STREAM_CACHE[key] = torch.cuda.Stream(torch.device('cuda', dev_index))
    return _STREAM_CACHE[key]


def to_instance_seg_stream(samples, patchshape, workers=1, **kwargs):
    """Assemble a sequence of samples as a pipeline (the reference handles them
    strictly one after the other, run_ppp.py:1111-1190, vote_instances.py:586-605):
      * the host->device copy of a sample runs on a copy stream while earlier
        samples are assembled;
      * `workers` samples are assembled side by side, each on its own CUDA stream
        from its own host thread.  Measured on the bench image: 2 workers are
        SLOWER than 1 (48 vs 43 ms per sample: the kernels of one sample already
        fill the GPU, two of them only time-slice), so the default is 1 and the
        pipeline overlaps just the upload.

    samples: iterable of (pred_affs, foreground, mask_to_cover, numinst) host
    arrays (numpy or pinned torch tensors; pred may be float16 as stored by the
    predict step).  Yields what to_instance_seg returns, in order."""
    import collections
    import concurrent.futures
    import threading
    import torch
    dev_index = torch.cuda.current_device()
    dev = torch.device('cuda', dev_index)
    copy = _cached_stream(dev_index, 'copy')
    workers = max(1, int(workers))
    nslots = workers + 1
    bufs = [None] * nslots             # device copies of the samples in flight
    local = threading.local()

    def as_tensor(a):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.to(torch.uint8) if a.dtype == torch.bool else a

    def upload(i, sample):
        host = [as_tensor(a) for a in sample]
        slot = i % nslots
        with torch.cuda.stream(copy):
            old = bufs[slot]
            devt = []
            for k, h in enumerate(host):
                if old is not None and old[k].shape == h.shape and old[k].dtype == h.dtype:
                    t = old[k]
                else:
                    t = torch.empty(h.shape, dtype=h.dtype, device=dev)
                t.copy_(h, non_blocking=True)
                devt.append(t)
            ready = torch.cuda.Event()
            ready.record(copy)
        bufs[slot] = devt
        return ready

    free_streams = [_cached_stream(dev_index, 'worker%d' % k) for k in range(workers)]
    lock = threading.Lock()

    def assemble(slot, ready):
        torch.cuda.set_device(dev_index)
        if not hasattr(local, 'stream'):
            with lock:
                local.stream = free_streams.pop()
        with torch.cuda.stream(local.stream):
            local.stream.wait_event(ready)
            pred, fg, mask, numinst = bufs[slot]
            pred32 = pred.float() if pred.dtype != torch.float32 else pred
            out = to_instance_seg(pred32, fg, mask, numinst, patchshape, **kwargs)
            local.stream.synchronize()
        return out

    inflight = collections.deque()
    with concurrent.futures.ThreadPoolExecutor(max_workers=workers) as pool:
        for i, sample in enumerate(samples):
            # sample i is uploaded while up to `workers` earlier samples are still being
            # assembled; its slot was last used by sample i - nslots, handed out before
            ready = upload(i, sample)
            inflight.append(pool.submit(assemble, i % nslots, ready))
            while len(inflight) > workers:
                yield inflight.popleft().result()
        while inflight:
            yield inflight.popleft().result()


def do_block(block, foreground, mask, numinst, **kwargs):
    """vote_instances.py:455-483."""
    patchshape = kwargs['patchshape']
    del kwargs['patchshape']
    if type(patchshape) != np.ndarray:
        patchshape = np.array(patchshape)
    res = to_instance_seg(block, foreground, mask, numinst, patchshape, **kwargs)
    if kwargs.get('return_intermediates'):
        return res
    instances, _ = res
    rad = np.array([p // 2 for p in patchshape])
    slices = tuple(slice(r, d - r) for r, d in zip(rad, instances.shape))
    return instances[slices]


def do_all(aff_file, patchshape=np.array([1, 25, 25]), **kwargs):
    """vote_instances.py:486-554: load -> assemble -> write <sample>.hdf/.npz."""
    logger.info("processing %s into %s", aff_file, kwargs['result_folder'])
    if type(patchshape) is not np.ndarray:
        patchshape = np.array(patchshape)
    res_ext = getResKey(**kwargs) if kwargs.get('add_suffix', False) else ''
    loaded = loadAffinities(aff_file, res_ext, patchshape=patchshape, **kwargs)
    if loaded is None:
        return
    affinities, numinst, foreground = loaded
    mask = np.copy(foreground)
    if numinst is None:
        numinst = np.copy(foreground)
    kwargs['aff_file'] = aff_file
    res = to_instance_seg(affinities, foreground, mask, numinst, patchshape, **kwargs)
    res_key = kwargs.get('res_key', 'vote_instances')
    instances, foreground = res
    if instances is None and foreground is None:
        return
    foreground = foreground.astype(np.uint8)
    if kwargs.get('crop_to_foreground', True):                 # :535-540
        if kwargs.get('one_instance_per_channel', False) or \
                kwargs.get('no_overlap_per_channel', False):
            instances[:, foreground == 0] = 0
        else:
            instances[foreground == 0] = 0
    fn = os.path.splitext(os.path.basename(aff_file))[0]
    from .io_util import write_result
    write_result(os.path.join(kwargs['result_folder'], fn),
                 {res_key + res_ext: instances,
                  'vote_foreground' + res_ext: foreground},
                 kwargs.get('output_format', 'hdf'))


# what the reference's argument parser would fill in for keys a caller leaves out
# (vote_instances.py:62-147)
_MAIN_DEFAULTS = dict(
    affinities=None, affinities_key="images/pred_affs", basedir=None, mode=None,
    checkpoint=None, debug=False, patch_threshold=0.9, fc_threshold=0.5, consensus=None,
    scores=None, ranked_patches=None, aff_graph=None, selected_patches=None,
    selected_patch_pairs=None, select_patches_for_sparse_data=False, cuda=False,
    skipLookup=False, skipThinCover=False, skipRanking=False, skipConsensus=False,
    termAfterThinCover=False, graphToInst=False, mws=False, includeSinglePatchCCS=False,
    removeIntersection=False, isbiHack=False, mask_fg_border=False, parallel=False,
    save_no_intermediates=False)


def _input_files(args):
    """the prediction files one call of main() covers (vote_instances.py:581-600): a zarr
    store or a single file, every .hdf / .npy of a directory, or the processed files of
    <basedir>/<mode>/processed/<checkpoint>."""
    src = args['affinities']
    if src is None:
        if args['mode'] is None or args['checkpoint'] is None:
            return []
        return glob.glob(os.path.join(args['basedir'], args['mode'], "processed",
                                      args['checkpoint'], "*.hdf"))
    if src.endswith(".zarr") or os.path.isfile(src):
        return [src]
    if os.path.isdir(src):
        return [f for ext in ("*.hdf", "*.npy") for f in glob.glob(os.path.join(src, ext))]
    raise RuntimeError("affinities (%s) should be file or dir" % src)


def main(**kwargs):
    """vote_instances.py:557-605: keyword arguments over the parser's defaults (sys.argv is
    not re-parsed), one do_all per input file."""
    args = merge_dicts(dict(_MAIN_DEFAULTS), kwargs)
    if 'check_required' in kwargs:
        for key, kinds in (('patchshape', (np.ndarray, tuple, list)), ('result_folder', (str,))):
            assert type(args.get(key)) in kinds, \
                "Please check type of {} {}".format(key, type(args.get(key)))
    if args['parallel']:
        raise NotImplementedError("parallel=True (a process pool over files) is not built")
    if args.get('cuda') and not args.get('graphToInst', False):
        args['context'] = cuda_code.init_cuda()
    os.makedirs(args['result_folder'], exist_ok=True)
    for fl in _input_files(args):
        do_all(fl, **args)
    cuda_code.delete_cuda(args.get('context'))
