"""Parity at BASELINE sizes: the full configs[1] image (520x696, 1x41x41 patches) and
a configs[2]-shaped block (16x128x128, 5x21x21 patches) against goldens recorded from the
UNMODIFIED reference (tools/gen_golden.py BIG_CASES: sampled consensus rows, whole-array
sums, scores, ranked order digest, cover, thinning, pairs, affinities, labels)."""
import hashlib
import os

import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu
CASES = [n for n in ('worms2d_c2_full', 'blobs3d_c3_block')
         if os.path.exists(os.path.join(gu.GOLD, n + '.npz'))]


@pytest.mark.parametrize('name', CASES)
def test_every_stage_against_the_reference_run(name):
    import torch
    from patchperpix_b200.assembly import BlockAssembler
    from patchperpix_b200.consensus_array import ConsensusArray
    from patchperpix_b200 import vote_instances as vi
    g, kw, ps, pred = gu.load(name)
    P = int(np.prod(ps))
    predt = torch.from_numpy(pred).cuda()
    fg = pred[P // 2] > np.float32(kw['patch_threshold'])
    fgt = torch.from_numpy(fg.astype(np.uint8)).cuda()
    ov = torch.from_numpy((g['numinst'] > 1).astype(np.uint8)).cuda()
    asm = BlockAssembler(predt, fgt, ov, ps, **kw)
    asm.prepare()
    asm.consensus(want_cnt=True)
    ca = ConsensusArray(asm)
    assert np.array_equal(ca.gate(), g['gate'])
    rows = g['rows']
    pos, neg = ca.compact('pos'), ca.compact('neg')
    # vote counters: sampled rows bit-exact, and the sum over the WHOLE array
    assert np.array_equal((pos.astype(np.int64) + neg)[rows], g['cnt'])
    assert int(pos.astype(np.int64).sum() + neg.astype(np.int64).sum()) == int(g['cnt_sum'])
    cons = ca.compact('cons')
    assert np.max(np.abs(cons[rows] - g['cons_norm'])) <= 1e-5
    tot = float(cons.astype(np.float64).sum())
    assert abs(tot - float(g['cons_norm_sum'])) <= 1e-6 * max(1.0, abs(float(g['cons_norm_sum'])))
    # scores, ranked order
    score = asm.rank().cpu().numpy()
    assert np.max(np.abs(score - g['score'])) <= 1e-5
    cand = asm.candidates()
    order = asm.ranked(cand, torch.from_numpy(g['score']).cuda())
    oc = asm.coords(order).astype(np.int32)
    assert hashlib.sha1(np.ascontiguousarray(oc).tobytes()).hexdigest() == str(g['ranked_sha1'])
    assert np.array_equal(oc[:len(g['ranked_head'])], g['ranked_head'])
    mask = torch.from_numpy((fg & ~(g['numinst'] > 1)).astype(np.uint8)).cuda()
    sel = asm.cover(mask, order)
    assert np.array_equal(asm.coords(sel), g['cover'])
    thin = asm.thin(mask, sel)
    assert np.array_equal(asm.coords(thin), g['thin'])
    pairs = asm.patch_pairs(asm.coords(thin))
    assert np.array_equal(pairs, g['pairs'])
    pd = torch.from_numpy(pairs.view(np.int32)).cuda()
    aff = asm.patch_graph(pd).cpu().numpy()
    assert np.max(np.abs(aff - g['aff'])) <= 1e-5
    assert np.array_equal(aff > 0, g['aff'] > 0) and np.array_equal(aff != 0, g['aff'] != 0)
    del asm, ca, predt
    torch.cuda.empty_cache()
    # and the whole path through the entry point, own scores and affinities
    inst, _ = vi.to_instance_seg(pred, fg, fg.copy(), g['numinst'], ps, **kw)
    assert np.array_equal(inst, g['instances'])
