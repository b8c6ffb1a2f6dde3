"""Blockwise assembly + stitching (stitch_patch_graph.py) against the golden
recorded from the reference's own blockwise driver (tools/gen_golden.py
blockwise).  The CPU test drives the product's host logic (block grid, halos,
face pairing, edge order) with the oracle as the per-block engine; the GPU test
uses the CUDA path end to end."""
import json
import hashlib
import os

import numpy as np
import pytest

from patchperpix_b200 import synth
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
    prob = np.stack([(numinst == 0), (numinst == 1), (numinst > 1)]).astype(np.float32)
    inputs = spg.VolumeInputs(pred.astype(np.float16), numinst_prob=prob)
    return g, kw, inputs


def _check(g, inst, fg, info, aff_tol=0.0):
    assert np.array_equal(inst.astype(np.uint16), g['instances'])
    assert np.array_equal(np.squeeze(fg).astype(np.uint16), g['foreground'])
    # per-block and per-face intermediates, in volume coordinates
    n_ref = sum(len(v) for k, v in g.items() if k.startswith('blk/') and k.endswith('aff_graph_mat'))
    assert info['n_edges'] == n_ref
    # every stored pair list / affinity vector (blocks are stored block-relative,
    # stitch_patch_graph.py:650, faces in volume coordinates, :338-347)
    mine = {}
    for p, a in zip(info['pairs'], info['aff']):
        mine.setdefault(tuple(int(x) for x in p), []).append(float(a))
    for k in g:
        if not k.endswith('patch_pairs'):
            continue
        pp = g[k].astype(np.int64)
        aa = g[k.replace('patch_pairs', 'aff_graph_mat')]
        name = k.split('/')[1:-1]
        if len(name) == 1:
            pp = pp + np.tile([int(v) for v in name[0].split('_')], 2)
        for p, a in zip(pp, aa):
            got = mine.get(tuple(int(x) for x in p))
            assert got is not None, (k, p)
            assert min(abs(float(a) - v) for v in got) <= aff_tol * max(1.0, abs(float(a))), (k, p, a, got)


def test_blockwise_host_logic_with_oracle_engine():
    from oracle import host_logic
    g, kw, inputs = _load()
    inst, fg, info = spg.stitch_arrays(inputs, block_fn=host_logic.oracle_block_fn,
                                       paint_fn=host_logic.oracle_paint_fn, **kw)
    _check(g, inst, fg, info)
    # block pairs are stored block-relative by the reference (stitch_patch_graph.py:650)
    offs = spg.get_offsets(inputs.shape, kw['chunksize'])
    k0 = 'blk/' + spg.get_offset_str(offs[-1]) + '/patch_pairs'
    if k0 in g:
        ref_global = g[k0].astype(np.int64) + np.tile(offs[-1], 2)
        mine = info['pairs'].astype(np.int64)
        assert (mine[:, None, :] == ref_global[None, :, :]).all(-1).any(0).all()


def test_blockwise_mws_with_oracle_engine():
    """same run with the mutex-watershed partition of the global graph."""
    from oracle import host_logic
    g, kw, inputs = _load()
    want = np.load(os.path.join(gu.GOLD, 'mws_cases.npz'))['blockwise_inst']
    inst, _, _ = spg.stitch_arrays(inputs, block_fn=host_logic.oracle_block_fn,
                                   paint_fn=host_logic.oracle_paint_fn, **dict(kw, mws=True))
    assert np.array_equal(inst.astype(np.uint16), want)


@pytest.mark.gpu
def test_blockwise_mws_cuda():
    g, kw, inputs = _load()
    want = np.load(os.path.join(gu.GOLD, 'mws_cases.npz'))['blockwise_inst']
    inst, _, _ = spg.stitch_arrays(inputs, **dict(kw, mws=True))
    assert np.array_equal(inst.astype(np.uint16), want)


@pytest.mark.gpu
def test_blockwise_cuda_end_to_end():
    g, kw, inputs = _load()
    inst, fg, info = spg.stitch_arrays(inputs, **kw)
    _check(g, inst, fg, info, aff_tol=1e-5)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import host_logic
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g, kw, inputs = _load()
        inst, fg, info = spg.stitch_arrays(inputs, block_fn=host_logic.oracle_block_fn,
                                           paint_fn=host_logic.oracle_paint_fn, **kw)
        q.put((rank, bool(np.array_equal(inst.astype(np.uint16), g['instances'])),
               info['n_edges'], info['n_blocks'], info['n_faces']))
    finally:
        dist.destroy_process_group()


def test_blockwise_two_ranks_gloo():
    """blocks and face jobs dealt over 2 ranks (gloo, CPU): every rank must end
    with the same edge list and the reference's labels."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2:] == res[1][2:]
