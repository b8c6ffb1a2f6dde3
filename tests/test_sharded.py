"""Sharded blockwise driver (patchperpix_b200/sharded.py): slabs, halo exchange of
compact rows, tensor all-gather of edge lists, replicated partition.

CPU tests drive the host logic with the oracle as block engine (world 1 and a
world-size-2 gloo job) and must reproduce the golden recorded from the
reference's own blockwise driver; the GPU tests run the CUDA path (and NCCL with
2 devices when the box has them)."""
import json
import hashlib
import os
import socket

import numpy as np
import pytest

from patchperpix_b200 import synth
from patchperpix_b200 import sharded
from patchperpix_b200 import stitch_patch_graph as spg
from tests import golden_util as gu


def _load():
    g = dict(np.load(os.path.join(gu.GOLD, 'blockwise3d_ps5.npz')))
    kw = json.loads(str(g['kwargs']))
    skw = json.loads(str(g['synth']))
    for k in gu._TUPLES:
        if k in skw:
            skw[k] = tuple(skw[k])
    ps = g['patchshape']
    pred, numinst, _ = synth.make_case(patchshape=ps, **skw)
    assert hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest() == str(g['pred_sha1'])
    return g, kw, pred, numinst


def _load3():
    """3x3x3 blocks, reference run with the mutex watershed and with thresholded CC"""
    g = dict(np.load(os.path.join(gu.GOLD, 'blockwise3d_3x3x3_mws.npz')))
    kw = json.loads(str(g['kwargs']))
    skw = json.loads(str(g['synth']))
    for k in gu._TUPLES:
        if k in skw:
            skw[k] = tuple(skw[k])
    pred, numinst, _ = synth.make_case(patchshape=g['patchshape'], **skw)
    assert hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest() == str(g['pred_sha1'])
    return g, kw, pred, numinst


def rows_of(pred, numinst, th, axis, lo, hi):
    """compact rows of the slab [lo, hi) of a dense prediction: the voxels whose
    centre channel passes the threshold (what a ppp+dec run would have decoded)."""
    import torch
    P = pred.shape[0]
    m = pred[P // 2] > np.float32(th)
    sl = [slice(None)] * 3
    sl[axis] = slice(lo, hi)
    c = np.argwhere(m[tuple(sl)])
    c[:, axis] += lo
    patches = pred[:, c[:, 0], c[:, 1], c[:, 2]].T.astype(np.float16)
    ni = numinst[c[:, 0], c[:, 1], c[:, 2]].astype(np.uint8)
    return (torch.from_numpy(c.astype(np.int32)), torch.from_numpy(np.ascontiguousarray(patches)),
            torch.from_numpy(ni))


def _oracle_hooks(pred):
    from oracle import host_logic

    def block_fn(src, fg, mask, numinst, **kw):
        return host_logic.oracle_block_fn(src.dense().numpy(), fg.numpy() > 0, mask.numpy() > 0,
                                          numinst.numpy(), **kw)

    def paint_fn(shard, pairs, aff, own_box, **kw):
        import torch
        ps = np.array(kw['patchshape'])
        inst, _ = host_logic.label_instances(pairs, aff, pred, ps, ps // 2, pred.shape[1:],
                                             np.float32(kw['patch_threshold']), dtype=np.uint32,
                                             mws=kw.get('mws', False))
        sl = [slice(None)] * 3
        sl[shard.axis] = slice(shard.lo, shard.hi)
        return torch.from_numpy(inst[tuple(sl)].astype(np.int32))
    return block_fn, paint_fn


def _run_rank(pred, numinst, kw, rank, world, hooks=True, device='cpu', workers=1):
    import torch
    shape = pred.shape[1:]
    axis, slabs = sharded.slab_partition(shape, kw['chunksize'], world)
    lo, hi = slabs[rank]
    c, p, ni = rows_of(pred, numinst, kw['patch_threshold'], axis, lo, hi)
    shard = sharded.RowShard(shape, axis, lo, hi, c.to(device), p.to(device), ni.to(device))
    extra = {}
    if hooks:
        extra['block_fn'], extra['paint_fn'] = _oracle_hooks(pred)
    inst, info = sharded.stitch_shard(shard, slabs, workers=workers, **extra, **kw)
    return inst, info, axis, slabs


def test_slab_partition():
    axis, slabs = sharded.slab_partition((256, 1024, 1024), (128, 128, 128), 8)
    assert axis == 1 and slabs == [(i * 128, (i + 1) * 128) for i in range(8)]
    axis, slabs = sharded.slab_partition((24, 44, 44), (12, 22, 22), 3)
    assert axis == 0 and slabs == [(0, 0), (0, 12), (12, 24)]
    axis, slabs = sharded.slab_partition((100, 30, 30), (32, 32, 32), 2)
    assert slabs == [(0, 64), (64, 100)]


def test_slab_partition_with_bounding_box_and_weights():
    # block grid over the box [10, 10+80) along y, chunks of 20: rows start at 10, 30, 50, 70
    axis, slabs = sharded.slab_partition((16, 120, 40), (16, 20, 40), 2, bb_offset=(0, 10, 0),
                                         bb_shape=(16, 80, 40))
    assert axis == 1 and slabs == [(0, 50), (50, 120)]
    # heavy first row: it gets a rank of its own
    axis, slabs = sharded.slab_partition((16, 120, 40), (16, 20, 40), 2, bb_offset=(0, 10, 0),
                                         bb_shape=(16, 80, 40), weights=[10, 1, 1, 1])
    assert slabs == [(0, 30), (30, 120)]
    # more ranks than rows: the spare ones stay empty, the outer ones reach the borders
    axis, slabs = sharded.slab_partition((16, 120, 40), (16, 40, 40), 4, bb_offset=(0, 10, 0),
                                         bb_shape=(16, 80, 40))
    live = [s for s in slabs if s[1] > s[0]]
    assert len(live) == 2 and live[0][0] == 0 and live[-1][1] == 120


def test_rows_from_dense_keeps_what_the_assembly_reads():
    import torch
    rng = np.random.default_rng(2)
    pred = rng.random((27, 6, 10, 12)).astype(np.float16)
    fg = rng.random((6, 10, 12)) < 0.2
    ni = rng.integers(0, 3, (6, 10, 12)).astype(np.uint8)
    for axis, lo, hi in ((0, 0, 6), (1, 3, 9), (2, 4, 12)):
        c, p, n, f = sharded.rows_from_dense(pred, fg, ni, axis, lo, hi, 'cpu', 0.5, tile=4)
        keep = (pred[13].astype(np.float32) > 0.5) | fg
        sl = [slice(None)] * 3
        sl[axis] = slice(lo, hi)
        want = np.argwhere(keep[tuple(sl)])
        want[:, axis] += lo
        cc = c.numpy()
        assert np.array_equal(cc, want)
        assert np.array_equal(p.numpy(), pred[:, cc[:, 0], cc[:, 1], cc[:, 2]].T)
        assert np.array_equal(n.numpy(), ni[cc[:, 0], cc[:, 1], cc[:, 2]])
        assert np.array_equal(f.numpy() != 0, fg[cc[:, 0], cc[:, 1], cc[:, 2]])


def test_sharded_world1_oracle_engine_matches_reference_golden():
    g, kw, pred, numinst = _load()
    inst, info, _, _ = _run_rank(pred, numinst, kw, 0, 1)
    assert np.array_equal(inst.numpy().astype(np.uint16), g['instances'])
    n_ref = sum(len(v) for k, v in g.items()
                if k.startswith('blk/') and k.endswith('aff_graph_mat'))
    assert info['n_edges'] == n_ref


def test_sharded_mws_world1_oracle_engine():
    g, kw, pred, numinst = _load()
    want = np.load(os.path.join(gu.GOLD, 'mws_cases.npz'))['blockwise_inst']
    inst, _, _, _ = _run_rank(pred, numinst, dict(kw, mws=True), 0, 1)
    assert np.array_equal(inst.numpy().astype(np.uint16), want)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g, kw, pred, numinst = _load()
        inst, info, axis, slabs = _run_rank(pred, numinst, kw, rank, world)
        full = sharded.gather_slabs(inst, slabs, axis, pred.shape[1:])
        ok = None
        if rank == 0:
            ok = bool(np.array_equal(full.numpy().astype(np.uint16), g['instances']))
        q.put((rank, ok, info['n_edges'], info['halo_bytes'],
               hashlib.sha1(info['pairs'].tobytes() + info['aff'].tobytes()).hexdigest()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_gloo(world):
    """slabs over 2 / 3 ranks (3: one rank owns no block): the halo rows travel,
    every rank ends with the same global edge list, rank 0 collects the
    reference's labels."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] is True, res
    assert len({r[2] for r in res}) == 1 and len({r[4] for r in res}) == 1, res
    assert sum(r[3] for r in res) > 0, "no halo rows were exchanged"


def test_sharded_3x3x3_blocks_cc_oracle_engine():
    g, kw, pred, numinst = _load3()
    inst, info, _, _ = _run_rank(pred, numinst, dict(kw, mws=False), 0, 1)
    assert info['n_blocks'] == 27
    assert np.array_equal(inst.numpy().astype(np.uint16), g['instances_cc'])


# ---------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('mws', [True, False])
def test_3x3x3_blocks_cuda_matches_reference_run(mws):
    """27 blocks, 54 faces: the sharded rows driver and the dense driver both reproduce
    the labels of the reference's own blockwise run."""
    g, kw, pred, numinst = _load3()
    want = g['instances_mws' if mws else 'instances_cc']
    inst, info, _, _ = _run_rank(pred, numinst, dict(kw, mws=mws), 0, 1, hooks=False,
                                 device='cuda', workers=4)
    assert info['n_blocks'] == 27
    assert np.array_equal(inst.cpu().numpy().astype(np.uint16), want)
    prob = np.stack([(numinst == 0), (numinst == 1), (numinst > 1)]).astype(np.float32)
    dense, _, _ = spg.stitch_arrays(spg.VolumeInputs(pred.astype(np.float16), numinst_prob=prob),
                                    **dict(kw, mws=mws))
    assert np.array_equal(dense.astype(np.uint16), want)

@pytest.mark.gpu
@pytest.mark.parametrize('mws', [False, True])
def test_sharded_cuda_matches_reference_golden(mws):
    g, kw, pred, numinst = _load()
    want = np.load(os.path.join(gu.GOLD, 'mws_cases.npz'))['blockwise_inst'] if mws \
        else g['instances']
    # blocks through the single-thread pipeline (pipeline.py, the default) and through
    # host threads calling to_instance_seg; face jobs batched per block row and one by one
    for workers, extra in ((1, {}), (3, dict(ppp_pipeline=False, ppp_batch_faces=False))):
        inst, info, _, _ = _run_rank(pred, numinst, dict(kw, mws=mws, **extra), 0, 1,
                                     hooks=False, device='cuda', workers=workers)
        assert np.array_equal(inst.cpu().numpy().astype(np.uint16), want)


@pytest.mark.gpu
def test_rows_path_equals_dense_path():
    """to_instance_seg on a RowSource == on the dense array the rows stand for."""
    import torch
    from patchperpix_b200 import vote_instances as vi
    from patchperpix_b200.assembly import RowSource
    ps = np.array([7, 7, 7])
    pred, numinst, labels = synth.make_case('neurites', ps, seed=11, shape=(20, 40, 40), n=5,
                                            seg_len=9.0, n_seg=10)
    c, p, ni = rows_of(pred, numinst, 0.5, 0, 0, 20)
    v2r = torch.full(pred.shape[1:], -1, dtype=torch.int32)
    v2r[c[:, 0].long(), c[:, 1].long(), c[:, 2].long()] = torch.arange(len(c), dtype=torch.int32)
    src = RowSource(p.cuda(), v2r.cuda())
    dense = src.dense()
    fg = (dense[171] > 0.5).to(torch.uint8)
    kw = dict(patch_threshold=0.5, fc_threshold=0.5, cuda=True, blockwise=False,
              select_patches_for_sparse_data=True, includeSinglePatchCCS=True,
              consensus_norm_prob_product=True, consensus_prob_product=True,
              consensus_norm_aff=True, consensus_interleaved_cnt=False,
              vi_bg_use_inv_th=False, vi_bg_use_half_th=False, vi_bg_use_less_than_th=True,
              rank_norm_patch_score=True, rank_int_counter=False, patch_graph_norm_aff=True,
              overlapping_inst=True)
    nit = torch.from_numpy(numinst).cuda()
    for mws in (False, True):
        a, _ = vi.to_instance_seg(src, fg, fg.clone(), nit, ps, mws=mws, **kw)
        b, _ = vi.to_instance_seg(dense, fg, fg.clone(), nit, ps, mws=mws, **kw)
        assert a.max() > 0 and np.array_equal(a, b)
        pa, aa = vi.to_instance_seg(src, fg, fg.clone(), nit, ps, mws=mws,
                                    return_intermediates=True, **kw)
        pb, ab = vi.to_instance_seg(dense, fg, fg.clone(), nit, ps, mws=mws,
                                    return_intermediates=True, **kw)
        assert np.array_equal(pa, pb) and np.array_equal(aa, ab)


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        g, kw, pred, numinst = _load()
        inst, info, axis, slabs = _run_rank(pred, numinst, kw, rank, world, hooks=False,
                                            device='cuda', workers=2)
        full = sharded.gather_slabs(inst, slabs, axis, pred.shape[1:])
        ok = None
        if rank == 0:
            ok = bool(np.array_equal(full.cpu().numpy().astype(np.uint16), g['instances']))
        q.put((rank, ok, info['n_edges'], info['halo_bytes']))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_nccl_two_gpus():
    """the NCCL path proper: 2 ranks, one GPU each."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][1] is True, res
    assert res[0][2] == res[1][2] and res[0][3] + res[1][3] > 0
