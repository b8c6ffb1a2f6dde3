"""GPU parity: the CUDA path (through the C ABI) against the golden vectors
recorded from the unmodified reference and against the CPU oracle.

Bars (BASELINE.json north_star): vote counters and instance labels bit-exact;
normalised consensus / scores / patch affinities within a stated fp32
tolerance (the reference itself sums with float atomics in arbitrary order).
"""
import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu

CONS_TOL = 1e-5      # max abs error, normalised consensus in [-1, 1]
SCORE_TOL = 1e-5
AFF_TOL = 1e-5


def _asm(name):
    import torch
    from patchperpix_b200.assembly import BlockAssembler
    g, kw, ps, pred = gu.load(name)
    predt = torch.from_numpy(pred).cuda()
    mid = int(np.prod(ps)) // 2
    fg = pred[mid] > np.float32(kw['patch_threshold'])
    fgt = torch.from_numpy(fg.astype(np.uint8)).cuda()
    ov = torch.from_numpy((g['numinst'] > 1).astype(np.uint8)).cuda()
    asm = BlockAssembler(predt, fgt, ov, ps, **kw)
    return g, kw, ps, pred, fg, asm


@pytest.mark.parametrize('name', gu.NAMES)
def test_consensus_counts_bit_exact_and_affinities_close(name):
    from patchperpix_b200.consensus_array import ConsensusArray
    from oracle import cpu_oracle
    g, kw, ps, pred, fg, asm = _asm(name)
    asm.prepare()
    asm.consensus(want_cnt=True)
    ca = ConsensusArray(asm)
    assert np.array_equal(ca.gate(), g['gate'])
    rows = g['rows']
    pos, neg = ca.compact('pos'), ca.compact('neg')
    if 'cnt' in g:
        assert np.array_equal((pos.astype(np.int64) + neg)[rows], g['cnt'])
        assert int(pos.astype(np.int64).sum() + neg.sum()) == int(g['cnt_sum'])
    # positive / negative votes separately against the oracle
    var = cpu_oracle.variant_from_kwargs(kw)
    O = cpu_oracle.Oracle(pred, g['numinst'] > 1, ps, var)
    _, op, on = O.consensus(want_cons=False)
    assert np.array_equal(pos, op)
    assert np.array_equal(neg, on)
    cons = ca.compact('cons')
    ref = g['cons_norm'] if 'cons_norm' in g else g['cons_raw']
    if var['prod_mode'] == 0 and not var['norm_aff']:
        assert np.array_equal(cons[rows], ref)           # integer-valued
    else:
        scale = 1.0 if var['norm_aff'] else float(np.prod(ps))
        assert np.max(np.abs(cons[rows] - ref)) <= CONS_TOL * scale
    # dense reference layout round trip
    if pred.size < 2_000_000 and 'cons_norm' in g and len(rows) == int(g['gate'].sum()):
        from patchperpix_b200 import layout
        dense = ca.to_dense()
        assert np.max(np.abs(layout.dense_to_compact(dense, g['gate'], ps) - ref)) \
            <= CONS_TOL


@pytest.mark.parametrize('name', gu.NAMES)
def test_rank_cover_thin_graph_labels(name):
    import torch
    g, kw, ps, pred, fg, asm = _asm(name)
    asm.prepare()
    asm.consensus()
    score = asm.rank().cpu().numpy()
    scale = 1.0 if kw.get('rank_norm_patch_score', True) else float(np.prod(ps)) ** 2
    assert np.max(np.abs(score - g['score'])) <= SCORE_TOL * scale
    # ranking: use the golden scores so that float noise cannot flip ties
    cand = asm.candidates()
    order = asm.ranked(cand, torch.from_numpy(g['score']).cuda())
    assert np.array_equal(asm.coords(order), g['ranked'])
    # with the device scores the order may only differ inside near-ties
    order_dev = asm.coords(asm.ranked(cand))
    sg = g['score'][tuple(order_dev.T)]
    assert np.all(np.diff(sg) <= 2 * SCORE_TOL * scale)
    mask = torch.from_numpy((fg & ~(g['numinst'] > 1)).astype(np.uint8)).cuda()
    sel = asm.cover(mask, order)
    assert np.array_equal(asm.coords(sel), g['cover'])
    # threshold 0 runs in the data-parallel "first coverer" form: the serial walk of
    # the reference (kept for the dense threshold schedule) must select the same
    asm.kwargs['ppp_cover_serial'] = True
    assert torch.equal(asm.cover(mask, order), sel)
    del asm.kwargs['ppp_cover_serial']
    thin = asm.thin(mask, sel) if not kw.get('skipThinCover', False) else sel
    assert np.array_equal(asm.coords(thin), g['thin'])
    if not kw.get('skipThinCover', False):
        # rounds of local maxima (default) == one selection per step (ppp_tune bit 16)
        from patchperpix_b200 import cuda_code as cc
        cfg0 = asm.cfg
        asm.cfg = cc.make_cfg(asm.shape, asm.ps, **dict(asm.kwargs, ppp_tune=0x10000))
        assert torch.equal(asm.thin(mask, sel), thin)
        asm.cfg = cfg0
    pairs = asm.patch_pairs(asm.coords(thin))
    assert np.array_equal(pairs, g['pairs'])
    pd = torch.from_numpy(pairs.view(np.int32)).cuda()
    aff = asm.patch_graph(pd)
    ascale = 1.0 if kw.get('patch_graph_norm_aff', True) else float(np.prod(ps)) ** 2
    affn = aff.cpu().numpy()
    # the reference sums up to P^2 terms in one float, serially: its own result
    # drifts from the exact sum (by ~1e-3 relative at 41x41), and the mutex
    # watershed orders edges by |aff|.  The default kernel adds in the reference's
    # order: tight against the golden and against the oracle's float accumulator
    # (what is left is the 1e-7 noise of the consensus sums).
    from oracle import cpu_oracle
    O = cpu_oracle.Oracle(pred, g['numinst'] > 1, ps, cpu_oracle.variant_from_kwargs(kw))
    O.consensus()
    O.norm()
    assert np.max(np.abs(affn - g['aff'])) <= AFF_TOL * ascale
    assert np.max(np.abs(affn - O.patch_graph(pairs))) <= AFF_TOL * ascale
    assert np.array_equal(affn > 0, g['aff'] > 0) and np.array_equal(affn != 0, g['aff'] != 0)
    # the parallel double-precision path: tight against the double-accumulating
    # oracle, loose against the golden, same signs
    fast = asm.patch_graph(pd, fast=True).cpu().numpy()
    exact = O.patch_graph(pairs, exact_sum=True)
    assert np.max(np.abs(fast - exact)) <= AFF_TOL * ascale
    loose = max(AFF_TOL, 4e-7 * float(np.prod(ps)) ** 1.5) * ascale
    assert np.max(np.abs(fast - g['aff'])) <= loose
    assert np.array_equal(fast > 0, g['aff'] > 0)
    # labels from the golden affinities (decouples CC/paint from float noise)
    nodes = torch.unique(torch.cat([
        (pd[:, 0] * asm.shape[1] + pd[:, 1]) * asm.shape[2] + pd[:, 2],
        (pd[:, 3] * asm.shape[1] + pd[:, 4]) * asm.shape[2] + pd[:, 5]])).int()
    inst, ncomp = asm.label(pd, torch.from_numpy(g['aff']).cuda(), nodes)
    assert np.array_equal(inst.cpu().numpy().astype(np.uint16), g['instances'])
    assert ncomp == len(np.unique(g['instances'])) - (1 if (g['instances'] == 0).any() else 0) \
        or ncomp >= g['instances'].max()


@pytest.mark.parametrize('name', gu.NAMES)
def test_end_to_end_labels_identical(name):
    """numpy in -> numpy out through the reference-facing entry point."""
    from patchperpix_b200 import vote_instances as vi
    g, kw, ps, pred = gu.load(name)
    mid = int(np.prod(ps)) // 2
    fg = pred[mid] > np.float32(kw['patch_threshold'])
    inst, fgo = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), g['numinst'].copy(),
                                   ps.copy(), **kw)
    assert inst.dtype == np.uint16 and fgo.dtype == np.uint8
    assert np.array_equal(fgo, fg.astype(np.uint8))
    assert np.array_equal(inst, g['instances'])
    pairs, aff = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), g['numinst'].copy(),
                                    ps.copy(), **dict(kw, return_intermediates=True))
    assert np.array_equal(pairs, g['pairs'])


@pytest.mark.parametrize('name', gu.NAMES)
def test_tiled_and_simple_consensus_agree(name):
    """the two CUDA implementations of step 1: counters identical, sums close."""
    g, kw, ps, pred, fg, asm = _asm(name)
    asm.prepare(want_rbits=True)
    import torch
    asm.consensus(impl=1)
    c1, n1 = asm.cons.clone(), asm.cnt.clone()
    for impl in (2, 3, 4):                   # bit-guided gather, tiled, received tables
        if impl == 4 and not asm.small:
            continue
        asm.consensus(impl=impl)
        assert torch.equal(n1, asm.cnt), impl
        # same centres, same order, same FMAs: the float sums are bit-identical
        assert torch.equal(c1, asm.cons), impl
