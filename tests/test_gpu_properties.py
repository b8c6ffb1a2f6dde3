"""GPU parity beyond the golden vectors.

1. fresh seeded inputs (not among the goldens) through the CPU oracle and the
   CUDA path: counters bit-exact, every later stage identical / within the
   stated fp32 tolerance, labels identical;
2. edge cases of the reference's entry point: no foreground, foreground only
   in the border, a single voxel, a volume exactly one patch large, empty pair
   lists;
3. the bench workload at FULL size (configs[1]: 520x696, patches 1x41x41),
   where the oracle would take hours: size-independent properties — vote
   conservation (a checksum of checksums against an independent torch count),
   determinism, sortedness of the ranking, the greedy-cover invariant, set-cover
   thinning keeps the coverage, painted labels stay inside selected patches.
"""
import numpy as np
import pytest

from patchperpix_b200 import synth

pytestmark = pytest.mark.gpu

TOL = 1e-5           # normalised consensus / scores / affinities (fp32 sums)

FLY = dict(patch_threshold=0.5, fc_threshold=0.5, cuda=True, overlapping_inst=True,
           vi_bg_use_inv_th=False, vi_bg_use_half_th=False, vi_bg_use_less_than_th=True,
           consensus_norm_prob_product=True, consensus_prob_product=True,
           consensus_norm_aff=True, consensus_interleaved_cnt=False,
           rank_norm_patch_score=True, rank_int_counter=False, patch_graph_norm_aff=True,
           select_patches_for_sparse_data=True, includeSinglePatchCCS=True, mws=False,
           skipThinCover=False, max_total_patch_distance_in_ps_multiples=2,
           return_intermediates=False, pad_with_ps=False)

SEEDED = [
    ('worms', dict(seed=101, shape=(56, 72), n_worms=4, width=(4, 7), length=(30, 60),
                   hard_frac=0.1), (1, 11, 11), {}),
    ('worms', dict(seed=102, shape=(48, 48), n_worms=3, width=(4, 6), length=(25, 45),
                   hard_frac=0.2), (1, 9, 9), dict(patch_threshold=0.7, fc_threshold=0.6,
                                                   vi_bg_use_inv_th=True,
                                                   vi_bg_use_less_than_th=False)),
    ('neurites', dict(seed=103, shape=(14, 30, 30), n=3, radius=(1.5, 2.5), seg_len=8.0,
                      n_seg=5), (5, 5, 5), {}),
    ('blobs', dict(seed=104, shape=(9, 26, 26), n=5, rad_xy=(3, 6), rad_z=(1.5, 3),
                   hard_frac=0.15), (3, 7, 7), dict(mws=True)),
]


@pytest.mark.parametrize('case', range(len(SEEDED)))
def test_fresh_seeded_inputs_against_oracle(case):
    import torch
    from oracle import cpu_oracle, host_logic
    from patchperpix_b200 import vote_instances as vi
    from patchperpix_b200.assembly import BlockAssembler
    from patchperpix_b200.consensus_array import ConsensusArray
    kind, skw, ps, over = SEEDED[case]
    ps = np.array(ps)
    kw = dict(FLY, **over)
    pred, numinst, _ = synth.make_case(kind=kind, patchshape=ps, **skw)
    mid = int(np.prod(ps)) // 2
    fg = pred[mid] > np.float32(kw['patch_threshold'])
    assert fg.sum() > 50
    # --- oracle, stage by stage ------------------------------------------------
    O = cpu_oracle.Oracle(pred, numinst > 1, ps, cpu_oracle.variant_from_kwargs(kw))
    want = host_logic.assemble(pred, fg, numinst, ps, kw, O)
    # --- CUDA -------------------------------------------------------------------
    asm = BlockAssembler(torch.from_numpy(pred).cuda(),
                         torch.from_numpy(fg.astype(np.uint8)).cuda(),
                         torch.from_numpy((numinst > 1).astype(np.uint8)).cuda(), ps, **kw)
    asm.prepare()
    asm.consensus(want_cnt=True)
    ca = ConsensusArray(asm)
    assert np.array_equal(ca.compact('pos'), O.cnt_pos)          # bit-exact counters
    assert np.array_equal(ca.compact('neg'), O.cnt_neg)
    assert np.max(np.abs(ca.compact('cons') - O.cons)) <= TOL
    score = asm.rank().cpu().numpy()
    assert np.max(np.abs(score - want['score'])) <= TOL
    inst, fgo = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), numinst.copy(),
                                   ps.copy(), **kw)
    assert np.array_equal(inst, want['instances'])               # identical labels
    pairs, aff = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), numinst.copy(),
                                    ps.copy(), **dict(kw, return_intermediates=True))
    assert np.array_equal(pairs, want['pairs'])
    assert np.array_equal(aff > 0, want['aff'] > 0)


def _call(pred, fg, ps, **over):
    from patchperpix_b200 import vote_instances as vi
    kw = dict(FLY, **over)
    numinst = fg.astype(np.uint8)
    return vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), numinst, np.array(ps), **kw)


def test_edge_no_foreground():
    ps = (1, 5, 5)
    pred = np.full((25, 1, 20, 20), 0.05, np.float32)
    inst, fg = _call(pred, np.zeros((1, 20, 20), bool), ps)
    assert inst.shape == (1, 20, 20) and not inst.any() and not fg.any()
    pairs, aff = _call(pred, np.zeros((1, 20, 20), bool), ps, return_intermediates=True)
    assert pairs is None and aff is None                         # vote_instances.py:232-245


def test_edge_foreground_only_in_border():
    """no interior patch centre (vote_instances.py:276-296): nothing is labelled."""
    ps = (1, 5, 5)
    pred = np.full((25, 1, 12, 12), 0.05, np.float32)
    fg = np.zeros((1, 12, 12), bool)
    fg[0, 0:2, :] = True
    pred[:, fg] = 0.95
    inst, _ = _call(pred, fg, ps)
    assert not inst.any()


def test_edge_single_voxel_and_single_patch_volume():
    ps = (1, 5, 5)
    # one foreground voxel: its patch marks only itself, the self pair has no
    # other pixel to agree with -> affinity 0 -> no edge -> nothing is painted
    # (aff_patch_graph.py:36, computePatchGraph.cu:131-135)
    pred = np.full((25, 1, 11, 11), 0.05, np.float32)
    fg = np.zeros((1, 11, 11), bool)
    fg[0, 5, 5] = True
    pred[12, 0, 5, 5] = 0.95
    inst, _ = _call(pred, fg, ps)
    assert not inst.any()
    # a 3x3 blob: one selected patch, kept as its own component through the self
    # pair (includeSinglePatchCCS), dropped without it
    labels = np.zeros((1, 11, 11), np.int32)
    labels[0, 4:7, 4:7] = 1
    pred = synth.patches_from_labels(labels, ps, seed=3, noise=0.0)
    fg = labels > 0
    inst, _ = _call(pred, fg, ps)
    assert np.array_equal(inst > 0, fg) and inst.max() == 1
    inst, _ = _call(pred, fg, ps, includeSinglePatchCCS=False)
    assert not inst.any()                                        # no pair, no component
    # a volume exactly one patch large, all foreground: one centre, one instance
    pred = np.full((25, 1, 5, 5), 0.95, np.float32)
    fg = np.ones((1, 5, 5), bool)
    inst, _ = _call(pred, fg, ps)
    assert np.array_equal(inst, np.ones((1, 5, 5), np.uint16))


def test_edge_even_patchshape_rejected():
    pred = np.zeros((16, 1, 8, 8), np.float32)
    with pytest.raises(AssertionError):
        _call(pred, np.ones((1, 8, 8), bool), (1, 4, 4))


# ---------------------------------------------------------------------------
# full size
# ---------------------------------------------------------------------------
@pytest.fixture(scope='module')
def full():
    import torch
    import bench
    from patchperpix_b200 import synth
    from patchperpix_b200.assembly import BlockAssembler
    dev = torch.device('cuda', 0)
    # BASELINE configs[1] at its full size: 2-D worms 696x520, patchshape 1x41x41
    ps = np.array([1, 41, 41])
    labels, numinst = synth.worms_2d((520, 696), n_worms=40, seed=2)
    pred = synth.patches_from_labels(labels, ps, seed=2, device=dev)
    P = int(np.prod(ps))
    fg = (pred[P // 2] > 0.5).to(torch.uint8)
    overlap = torch.from_numpy((numinst > 1).astype(np.uint8)).to(dev)
    mask = fg.clone()
    mask[overlap > 0] = 0
    kw = dict(bench.KW, blockwise=False)
    asm = BlockAssembler(pred, fg, overlap, ps, **kw)
    asm.prepare()
    asm.consensus(want_cnt=True)
    asm.rank()
    return dict(pred=pred, fg=fg, overlap=overlap, mask=mask, asm=asm, ps=ps, kw=kw)


def _shift(t, dy, dx):
    """t[y + dy, x + dx] with zero fill (2-D volumes [1,Y,X])."""
    import torch
    out = torch.zeros_like(t)
    Y, X = t.shape[-2:]
    ys, ye = max(0, -dy), min(Y, Y - dy)
    xs, xe = max(0, -dx), min(X, X - dx)
    if ys < ye and xs < xe:
        out[..., ys:ye, xs:xe] = t[..., ys + dy:ye + dy, xs + dx:xe + dx]
    return out


def test_full_size_vote_conservation(full):
    """sum over all slots of the positive (negative) counters == sum over the
    centres of C(h,2) (h*l), h / l = gated high / background pixels of the
    centre's patch — counted here with plain torch ops on the raw prediction."""
    import torch
    asm, pred, ps = full['asm'], full['pred'], full['ps']
    th = np.float32(0.5)
    gate = (pred[int(np.prod(ps)) // 2] > th) & (full['overlap'] == 0)
    ry, rx = int(ps[1]) // 2, int(ps[2]) // 2
    Y, X = gate.shape[-2:]
    centre = torch.zeros_like(gate)
    centre[..., ry:Y - ry, rx:X - rx] = True
    centre &= pred[int(np.prod(ps)) // 2] > th
    h = torch.zeros(gate.shape, dtype=torch.int64, device=pred.device)
    l = torch.zeros_like(h)
    for po in range(pred.shape[0]):
        dy, dx = po // int(ps[2]) - ry, po % int(ps[2]) - rx
        gsh = _shift(gate, dy, dx)
        h += (gsh & (pred[po] > th)).long()
        l += (gsh & (pred[po] < th)).long()          # vi_bg_use_less_than_th
    h = h * centre
    l = l * centre
    pos_want = int((h * (h - 1) // 2).sum().item())
    neg_want = int((h * l).sum().item())
    cnt = asm.cnt.view(torch.int32).long()
    pos = int((cnt & 0xffff).sum().item())
    neg = int((cnt >> 16).sum().item())
    assert pos == pos_want and neg == neg_want
    assert pos > 10 ** 9                              # this really is the full workload
    # nothing is stored for rows that are not gated
    assert float(asm.cons.abs().max()) <= 1.0 + TOL


def test_full_size_deterministic(full):
    import torch
    from patchperpix_b200.assembly import BlockAssembler
    a = full['asm']
    b = BlockAssembler(full['pred'], full['fg'], full['overlap'], full['ps'], **full['kw'])
    b.prepare()
    b.consensus(want_cnt=True)
    b.rank()
    assert torch.equal(a.cnt, b.cnt)
    assert torch.equal(a.cons, b.cons)               # single writer per slot, fixed order
    assert torch.equal(a.score, b.score)
    # the cross-check kernel (one CTA per voxel) agrees bit for bit at full size too
    b.consensus(want_cnt=True, impl=1)
    assert torch.equal(a.cnt, b.cnt) and torch.equal(a.cons, b.cons)


def test_full_size_ranking_cover_thin_paint(full):
    import torch
    asm, pred, mask, ps = full['asm'], full['pred'], full['mask'], full['ps']
    shape = asm.shape
    cand = asm.candidates()
    order = asm.ranked(cand)
    s = asm.score.flatten()[order.long()]
    assert bool((s[1:] <= s[:-1]).all())                          # sorted, descending
    tie = s[1:] == s[:-1]
    assert bool((order[1:][tie] > order[:-1][tie]).all())         # ties keep raster order
    assert torch.equal(torch.sort(order)[0], torch.sort(cand)[0])  # a permutation
    sel = asm.cover(mask, order)
    asm.kwargs['ppp_cover_serial'] = True                          # the serial walk agrees
    assert torch.equal(asm.cover(mask, order), sel)
    del asm.kwargs['ppp_cover_serial']
    thin = asm.thin(mask, sel)
    # thinning runs in rounds of local maxima; the one-selection-per-step form of the
    # reference (ppp_tune bit 16) must keep exactly the same patches
    from patchperpix_b200 import cuda_code as cc
    cfg0 = asm.cfg
    asm.cfg = cc.make_cfg(asm.shape, asm.ps, **dict(asm.kwargs, ppp_tune=0x10000))
    assert torch.equal(asm.thin(mask, sel), thin)
    asm.cfg = cfg0
    assert set(thin.tolist()) <= set(sel.tolist()) <= set(order.tolist())
    fc = np.float32(full['kw']['fc_threshold'])
    ry, rx = int(ps[1]) // 2, int(ps[2]) // 2

    def covered(centres):
        """voxels marked (> fc) by the patches of `centres`."""
        c = torch.zeros(int(np.prod(shape)), dtype=torch.bool, device=pred.device)
        c[centres.long()] = True
        c = c.view(shape)
        out = torch.zeros(shape, dtype=torch.bool, device=pred.device)
        for po in range(pred.shape[0]):
            dy, dx = po // int(ps[2]) - ry, po % int(ps[2]) - rx
            out |= _shift(c & (pred[po] > fc), -dy, -dx)
        return out
    # greedy cover with pixel threshold 0 (foreground_cover.py:35-39, sparse data):
    # whatever any candidate could cover is covered by the selection
    # (both loops only look at the interior `radslice`, :128-129 and :196-199)
    interior = torch.zeros(shape, dtype=torch.bool, device=pred.device)
    interior[:, ry:shape[1] - ry, rx:shape[2] - rx] = True
    skip = full['overlap'].flatten()[order.long()] > 0            # :141-145
    reachable = covered(order[~skip]) & (mask > 0)
    got = covered(sel) & (mask > 0)
    assert not bool((reachable & interior & ~got).any())
    assert not bool((got & ~reachable).any())
    # set-cover thinning keeps the coverage (foreground_cover.py:183-256)
    assert not bool((got & interior & ~covered(thin)).any())
    assert thin.numel() < sel.numel()
    # painted labels lie inside the patches of the thinned selection
    pairs = asm.patch_pairs(asm.coords(thin))
    pd = torch.from_numpy(pairs.view(np.int32)).to(pred.device)
    aff = asm.patch_graph(pd)
    inst, ncomp = asm.label(pd, aff, thin)
    pt = np.float32(full['kw']['patch_threshold'])
    inside = torch.zeros(shape, dtype=torch.bool, device=pred.device)
    c = torch.zeros(int(np.prod(shape)), dtype=torch.bool, device=pred.device)
    c[thin.long()] = True
    c = c.view(shape)
    for po in range(pred.shape[0]):
        dy, dx = po // int(ps[2]) - ry, po % int(ps[2]) - rx
        inside |= _shift(c & (pred[po] > pt), -dy, -dx)
    assert bool(((inst > 0) <= inside).all())
    assert 0 < ncomp <= thin.numel() and int(inst.max().item()) == ncomp
    # mutex watershed on the same graph: same painted support or smaller
    inst2, top = asm.label(pd, aff, thin, mws=True)
    assert bool(((inst2 > 0) <= inside).all()) and top >= 1


def test_streamed_samples_equal_single_calls():
    """to_instance_seg_stream (copy stream overlapping the next upload) returns
    exactly what one to_instance_seg call per sample returns, in order, for
    float16 and float32 inputs of different shapes."""
    from patchperpix_b200 import vote_instances as vi
    ps = np.array([1, 9, 9])
    samples, want = [], []
    for seed, shape, dt in ((201, (40, 56), np.float16), (202, (48, 48), np.float32),
                            (203, (40, 56), np.float16)):
        pred, numinst, _ = synth.make_case(kind='worms', patchshape=ps, seed=seed, shape=shape,
                                           n_worms=3, width=(4, 6), length=(20, 40))
        fg = pred[40] > np.float32(0.5)
        samples.append((pred.astype(dt), fg, fg.copy(), numinst))
        want.append(vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), numinst.copy(),
                                       ps.copy(), **FLY)[0])
    got = [r[0] for r in vi.to_instance_seg_stream(iter(samples), ps, **FLY)]
    assert len(got) == 3
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    assert list(vi.to_instance_seg_stream(iter([]), ps, **FLY)) == []


def test_file_level_entry_point(tmp_path):
    """vote_instances.main (vote_instances.py:557-605) on `.npy` predictions, single file
    and directory form: same labels as the array call, foreground-cropped
    (:535-540), written as `<sample>.npz` (h5py is optional in this image)."""
    from patchperpix_b200 import vote_instances as vi
    ps = np.array([1, 9, 9])
    want = {}
    d = tmp_path / 'pred'
    d.mkdir()
    for seed in (301, 302):
        pred, _, _ = synth.make_case(kind='worms', patchshape=ps, seed=seed, shape=(40, 56),
                                     n_worms=3, width=(4, 6), length=(20, 40))
        fg = pred[40] > np.float32(0.5)
        np.save(d / ('s%d.npy' % seed), pred[:, 0])              # legacy 2-D layout [P,Y,X]
        inst, fgo = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), fg.astype(np.uint8),
                                       ps.copy(), **FLY)
        inst[fgo == 0] = 0
        want[seed] = inst
    kw = {k: v for k, v in FLY.items() if k not in ('return_intermediates', 'pad_with_ps')}
    out1 = tmp_path / 'one'
    vi.main(affinities=str(d / 's301.npy'), patchshape=[1, 9, 9], result_folder=str(out1),
            output_format='npz', check_required=False, **kw)
    got = np.load(out1 / 's301.npz')
    assert np.array_equal(got['vote_instances'], want[301])
    assert got['vote_foreground'].dtype == np.uint8
    out2 = tmp_path / 'all'
    vi.main(affinities=str(d), patchshape=[1, 9, 9], result_folder=str(out2),
            output_format='npz', check_required=False, **kw)
    for seed in (301, 302):
        assert np.array_equal(np.load(out2 / ('s%d.npz' % seed))['vote_instances'], want[seed])
