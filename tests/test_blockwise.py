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


def _check(g, inst, fg, info):
    assert np.array_equal(inst.astype(np.uint16), g['instances'])
    assert np.array_equal(np.squeeze(fg).astype(np.uint16), g['foreground'])
    # per-block and per-face intermediates, in volume coordinates
    n_ref = sum(len(v) for k, v in g.items() if k.startswith('blk/') and k.endswith('aff_graph_mat'))
    assert info['n_edges'] == n_ref


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


@pytest.mark.gpu
def test_blockwise_cuda_end_to_end():
    g, kw, inputs = _load()
    inst, fg, info = spg.stitch_arrays(inputs, **kw)
    _check(g, inst, fg, info)


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
