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


def test_c1_counters_equal_the_reference_cpu_path():
    """BASELINE configs[0] / SURVEY.md A.8: the reference's PYTHON CPU path
    (fillLookup -> computeFGBGsets -> create_consensus_array, run unmodified by
    tools/gen_c1_cpu.py on the bundled flylight crop) accumulates +1 / -1 votes in int16;
    with plain vote counting, the inverse-threshold background band and no overlap
    handling the CUDA path's counters must give exactly pos - neg."""
    import hashlib
    import torch
    from patchperpix_b200 import synth
    from patchperpix_b200.assembly import BlockAssembler
    from patchperpix_b200.consensus_array import ConsensusArray
    g = np.load(os.path.join(gu.GOLD, 'c1_cpu_consensus.npz'))
    ps = np.array([7, 7, 7])
    pred = synth.crop_case(g['labels'], ps, seed=int(g['seed']))
    assert hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest() == str(g['pred_sha1'])
    th = float(g['th'])
    kw = dict(patch_threshold=th, fc_threshold=0.5, vi_bg_use_inv_th=True,
              consensus_norm_prob_product=False, consensus_prob_product=False,
              consensus_norm_aff=False, consensus_interleaved_cnt=False, overlapping_inst=False)
    predt = torch.from_numpy(pred).cuda()
    fg = (predt[171] > th).to(torch.uint8)
    for impl in (0, 2):                                  # received tables, bit-guided gather
        asm = BlockAssembler(predt, fg, torch.zeros_like(fg), ps, **kw)
        asm.prepare(want_rbits=True)
        asm.consensus(want_cnt=True, impl=impl)
        ca = ConsensusArray(asm)
        votes = ca.compact('pos').astype(np.int64) - ca.compact('neg').astype(np.int64)
        assert np.array_equal(votes[g['rows']], g['cons'].astype(np.int64)), impl
        assert int(votes.sum()) == int(g['cons_sum'])
        assert int(np.abs(votes).sum()) == int(g['cons_abs_sum'])
        # counter mode: the consensus array itself is that difference
        assert np.array_equal(ca.compact('cons')[g['rows']], g['cons'].astype(np.float32))
