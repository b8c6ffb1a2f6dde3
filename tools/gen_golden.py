#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference here.

Runs only where /root/reference exists (this container).  For each case it
feeds seeded synthetic predictions (patchperpix_b200/synth.py) to the
reference's `to_instance_seg` (vote_instances.py:150) with `cuda=True`, the
kernels compiled for the host through oracle/ref_shim (SURVEY.md §8c Route 2b),
and records every intermediate of the path: un-normalised consensus sums, vote
counters, normalised consensus, rank scores, ranked order, selected patches
after cover / after thinning, patch pairs, patch affinities, instance labels.
The vectors are stored compactly (gated fg rows x positive offsets, see
patchperpix_b200/layout.py); big cases keep a seeded sample of rows.

usage: python tools/gen_golden.py [case ...]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import ref_runner            # noqa: E402
from patchperpix_b200 import synth, layout  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

# name -> (synth kwargs, patchshape, reference kwargs overrides, max_rows)
CASES = {
    # the survey's two-disc probe, flylight flags
    'discs2d_ps9': (dict(kind='discs', seed=1), (1, 9, 9), {}, None),
    # crossing worms with overlap voxels, heavy noise, flylight flags
    'worms2d_ps9_hard': (dict(kind='worms', seed=5, shape=(72, 96), n_worms=5,
                              width=(5, 8), length=(40, 80), hard_frac=0.15),
                         (1, 9, 9), {}, None),
    # the C2 patch size (1x41x41) on a small image
    'worms2d_ps41': (dict(kind='worms', seed=7, shape=(96, 128), n_worms=4,
                          width=(8, 11), length=(60, 110)),
                     (1, 41, 41), {}, 64),
    # flylight-shaped 3-D block, 7^3 patches
    'neurites3d_ps7': (dict(kind='neurites', seed=11, shape=(20, 36, 36), n=4,
                            radius=(1.5, 2.5), seg_len=10.0, n_seg=6),
                       (7, 7, 7), {}, None),
    # anisotropic 3-D patches, th 0.7 with the 1-th background band, plain
    # probability product, interleaved counter, integer rank counter
    'blobs3d_ps355_inv': (dict(kind='blobs', seed=13, shape=(10, 28, 28), n=6,
                               rad_xy=(4, 7), rad_z=(1.5, 3), hard_frac=0.2),
                          (3, 5, 5),
                          dict(patch_threshold=0.7, fc_threshold=0.6,
                               vi_bg_use_inv_th=True,
                               vi_bg_use_less_than_th=False,
                               consensus_norm_prob_product=False,
                               consensus_prob_product=True,
                               consensus_interleaved_cnt=True,
                               rank_int_counter=True), None),
    # plain vote counter (no probability product), th/2 background, no
    # normalisation anywhere, no overlap handling, dense cover thresholds
    'worms2d_ps7_counter': (dict(kind='worms', seed=17, shape=(64, 64),
                                 n_worms=4, width=(5, 8), length=(30, 60),
                                 hard_frac=0.1),
                            (1, 7, 7),
                            dict(vi_bg_use_inv_th=False, vi_bg_use_half_th=True,
                                 vi_bg_use_less_than_th=False,
                                 consensus_norm_prob_product=False,
                                 consensus_prob_product=False,
                                 consensus_norm_aff=False,
                                 consensus_interleaved_cnt=False,
                                 rank_norm_patch_score=False,
                                 patch_graph_norm_aff=False,
                                 overlapping_inst=False,
                                 select_patches_for_sparse_data=False), None),
}


def run_case(name, S, rec):
    skw, ps, over, max_rows = CASES[name]
    ps = np.array(ps)
    pred, numinst, labels = synth.make_case(patchshape=ps, **skw)
    kw = S.default_kwargs(**over)
    th = kw['patch_threshold']
    mid = int(np.prod(ps)) // 2
    fg = pred[mid] > th
    rec.clear()
    stages = {}

    vi = S.vi
    orig_cover = S.mods['foreground_cover'].computeForegroundCover
    orig_thin = S.mods['foreground_cover'].thinOutForegroundCover
    orig_rank = S.mods['ranked_patches'].rank_patches_by_score

    def cover(*a, **k):
        res = orig_cover(*a, **k)
        stages['cover'] = np.array([p[0] for p in res[0]], np.int32).reshape(-1, 3)
        return res

    def thin(*a, **k):
        res = orig_thin(*a, **k)
        stages['thin'] = np.array([p[0] for p in res[0]], np.int32).reshape(-1, 3)
        return res

    def rank(*a, **k):
        res = orig_rank(*a, **k)
        stages['ranked'] = np.array([p[0] for p in res], np.int32).reshape(-1, 3)
        return res

    vi.computeForegroundCover = cover
    vi.thinOutForegroundCover = thin
    S.mods['ranked_patches'].rank_patches_by_score = rank
    t0 = time.time()
    try:
        inst, fgo = vi.to_instance_seg(
            pred.copy(), fg.copy(), fg.copy(), numinst.copy(), ps.copy(), **kw)
    finally:
        vi.computeForegroundCover = orig_cover
        vi.thinOutForegroundCover = orig_thin
        S.mods['ranked_patches'].rank_patches_by_score = orig_rank
    dt = time.time() - t0

    gate = fg.copy()
    if kw['overlapping_inst']:
        gate &= ~(numinst > 1)
    import hashlib
    p16 = pred.astype(np.float16)
    out = dict(
        pred_sha1=hashlib.sha1(p16.tobytes()).hexdigest(), numinst=numinst,
        patchshape=ps.astype(np.int32),
        kwargs=json.dumps({k: v for k, v in kw.items()
                           if isinstance(v, (bool, int, float, str))}),
        synth=json.dumps({k: (list(v) if isinstance(v, tuple) else v)
                          for k, v in skw.items()}),
        instances=inst, gate=gate,
    )
    if p16.nbytes <= (1 << 20):     # big inputs are regenerated from `synth`
        out['pred_f16'] = p16
    rows = np.arange(int(gate.sum()))
    if max_rows is not None and len(rows) > max_rows:
        rows = np.sort(np.random.default_rng(0).choice(
            len(rows), max_rows, replace=False))
    out['rows'] = rows.astype(np.int32)
    sn = rec.result()
    if sn.get('cons_raw') is not None:
        out['cons_raw'] = layout.dense_to_compact(sn['cons_raw'], gate, ps)[rows]
    if sn.get('cnt') is not None:
        c = layout.dense_to_compact(sn['cnt'], gate, ps)[rows]
        assert np.all(c == np.round(c)) and c.max() < 65536
        out['cnt'] = c.astype(np.uint16)
    if sn.get('cons_norm') is not None:
        out['cons_norm'] = layout.dense_to_compact(sn['cons_norm'], gate, ps)[rows]
    # whole-array checksums (double) so that sampled cases still pin the total
    for key in ('cons_raw', 'cnt', 'cons_norm'):
        if sn.get(key) is not None:
            out[key + '_sum'] = np.float64(sn[key].astype(np.float64).sum())
            # nothing may be written outside gated rows / positive offsets
            full = layout.dense_to_compact(sn[key], gate, ps)
            assert np.isclose(full.astype(np.float64).sum(), out[key + '_sum'],
                              rtol=1e-9, atol=1e-6), key
    out['score'] = sn['score']
    out['pairs'] = sn['pairs']
    out['aff'] = np.array(sn['aff'])
    for k in ('ranked', 'cover', 'thin'):
        if k in stages:
            out[k] = stages[k]
    os.makedirs(GOLD, exist_ok=True)
    fn = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(fn, **out)
    print('%-24s %6.1fs  fg=%d gated=%d rows=%d pairs=%d inst=%d  %.2f MB' % (
        name, dt, int(fg.sum()), int(gate.sum()), len(rows),
        len(sn['pairs']), len(np.unique(inst)) - 1,
        os.path.getsize(fn) / 1e6))


class Recorder:
    """snapshots kernel outputs in launch order (see ref_runner._Kernel)."""

    def __init__(self):
        self.clear()

    def clear(self):
        self.d = {}

    def result(self):
        return self.d

    def __call__(self, kind, options, args):
        arrs = [a for a in args]
        if kind == 'K_FILL':
            if '-DOUTPUT_BOTH' in options:
                self.d['cons_raw'] = np.array(arrs[-2])
                self.d['cnt'] = np.array(arrs[-1])
            elif '-DOUTPUT_CNT' in options:
                self.d['cnt'] = np.array(arrs[-1])
            else:
                self.d['cons_raw'] = np.array(arrs[-1])
        elif kind == 'K_NORM':
            self.d['cons_norm'] = np.array(arrs[1])
        elif kind == 'K_RANK':
            self.d['score'] = np.array(arrs[-1])
        elif kind == 'K_GRAPH':
            self.d['aff'] = np.asarray(arrs[2])     # filled slice by slice
            self.d['pairs'] = np.array(arrs[3])


# BASELINE-size cases (parity was pinned at toy sizes only).  Too big for
# the dense snapshots of run_case: the recorder keeps a row sample + whole-array sums.
BIG_CASES = {
    # configs[1] at its full size: the 2-D worms image, 1x41x41 patches
    'worms2d_c2_full': (dict(kind='worms', seed=2, shape=(520, 696), n_worms=40),
                        (1, 41, 41), {}, 48),
    # configs[2]-shaped block: nuclei-like blobs, 5x21x21 patches, 16x128x128
    'blobs3d_c3_block': (dict(kind='blobs', seed=3, shape=(16, 128, 128), n=20,
                              rad_xy=(6, 14), rad_z=(2, 4)),
                         (5, 21, 21), {}, 32),
}


class BigRecorder:
    """like Recorder, but keeps only `rows` of the compact consensus arrays and
    double-precision sums (the dense arrays are tens of GB)."""

    def __init__(self, gate, ps, rows):
        self.gate, self.ps, self.rows = gate, ps, rows
        self.d = {}

    def _keep(self, key, arr):
        self.d[key + '_sum'] = np.float64(np.asarray(arr).astype(np.float64).sum())
        full = layout.dense_to_compact(np.asarray(arr), self.gate, self.ps)
        assert np.isclose(full.astype(np.float64).sum(), self.d[key + '_sum'],
                          rtol=1e-9, atol=1e-6), key
        self.d[key] = full[self.rows].copy()

    def __call__(self, kind, options, args):
        arrs = [a for a in args]
        if kind == 'K_FILL':
            if '-DOUTPUT_CNT' in options:
                self._keep('cnt', arrs[-1])
            else:
                self._keep('cons_raw', arrs[-1])
        elif kind == 'K_NORM':
            self._keep('cons_norm', arrs[1])
        elif kind == 'K_RANK':
            self.d['score'] = np.array(arrs[-1])
        elif kind == 'K_GRAPH':
            self.d['aff'] = np.asarray(arrs[2])
            self.d['pairs'] = np.array(arrs[3])


def run_big(name):
    import hashlib
    skw, ps, over, max_rows = BIG_CASES[name]
    ps = np.array(ps)
    pred, numinst, labels = synth.make_case(patchshape=ps, **skw)
    mid = int(np.prod(ps)) // 2
    gate0 = (pred[mid] > 0.5) & ~(numinst > 1)
    rows = np.sort(np.random.default_rng(0).choice(int(gate0.sum()), max_rows, replace=False))
    rec = BigRecorder(gate0, ps, rows)
    S = ref_runner.RefSession(recorder=rec)
    kw = S.default_kwargs(**over)
    fg = pred[mid] > kw['patch_threshold']
    stages = {}
    vi = S.vi
    orig_cover = S.mods['foreground_cover'].computeForegroundCover
    orig_thin = S.mods['foreground_cover'].thinOutForegroundCover
    orig_rank = S.mods['ranked_patches'].rank_patches_by_score

    def cover(*a, **k):
        res = orig_cover(*a, **k)
        stages['cover'] = np.array([p[0] for p in res[0]], np.int32).reshape(-1, 3)
        return res

    def thin(*a, **k):
        res = orig_thin(*a, **k)
        stages['thin'] = np.array([p[0] for p in res[0]], np.int32).reshape(-1, 3)
        return res

    def rank(*a, **k):
        res = orig_rank(*a, **k)
        stages['ranked'] = np.array([p[0] for p in res], np.int32).reshape(-1, 3)
        return res
    vi.computeForegroundCover = cover
    vi.thinOutForegroundCover = thin
    S.mods['ranked_patches'].rank_patches_by_score = rank
    t0 = time.time()
    inst, fgo = vi.to_instance_seg(pred, fg.copy(), fg.copy(), numinst.copy(), ps.copy(), **kw)
    dt = time.time() - t0
    sn = rec.d
    out = dict(
        pred_sha1=hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest(), numinst=numinst,
        patchshape=ps.astype(np.int32),
        kwargs=json.dumps({k: v for k, v in kw.items()
                           if isinstance(v, (bool, int, float, str))}),
        synth=json.dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in skw.items()}),
        instances=inst, gate=gate0, rows=rows.astype(np.int32),
        cons_raw=sn['cons_raw'], cons_raw_sum=sn['cons_raw_sum'],
        cnt=sn['cnt'].astype(np.uint16), cnt_sum=sn['cnt_sum'],
        cons_norm=sn['cons_norm'], cons_norm_sum=sn['cons_norm_sum'],
        score=sn['score'], pairs=sn['pairs'], aff=np.array(sn['aff']),
        ranked_sha1=hashlib.sha1(stages['ranked'].tobytes()).hexdigest(),
        ranked_head=stages['ranked'][:4096], cover=stages['cover'], thin=stages['thin'])
    fn = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(fn, **out)
    print('%-24s %6.1fs  fg=%d rows=%d pairs=%d inst=%d  %.2f MB' % (
        name, dt, int(fg.sum()), len(rows), len(sn['pairs']), len(np.unique(inst)) - 1,
        os.path.getsize(fn) / 1e6), flush=True)


def run_blockwise(S, mws=False):
    """the reference's blockwise driver (stitch_patch_graph.main) on an
    in-memory zarr stand-in (oracle/ref_runner.FakeGroup).  mws: only return
    the label volume of a run with the mutex-watershed partition."""
    import shutil
    sp = ref_runner.load_stitch_module(S)
    ps = np.array([5, 5, 5])
    skw = dict(kind='neurites', seed=21, shape=(24, 44, 44), n=6, radius=(1.5, 2.5),
               seg_len=9.0, n_seg=8)
    pred, numinst, labels = synth.make_case(patchshape=ps, **skw)
    root = '/tmp/ppp_gold_blk'
    shutil.rmtree(root, ignore_errors=True)
    pred_path = os.path.join(root, 'sample.zarr')
    store = ref_runner.FakeGroup.open(pred_path, 'w')
    store['volumes/pred_affs'] = pred.astype(np.float16)
    prob = np.stack([(numinst == 0), (numinst == 1), (numinst > 1)]).astype(np.float32)
    store['volumes/pred_numinst'] = prob
    kw = S.default_kwargs(
        blockwise=True, chunksize=[12, 22, 22], patchshape=[5, 5, 5],
        aff_key='volumes/pred_affs', numinst_key='volumes/pred_numinst', fg_key=None,
        numinst_threshs=[0.9, 0.1], only_bb=False, output_format='hdf',
        num_parallel_blocks=1, ignore_small_comps=0, skeletonize_foreground=False,
        remove_small_comps=0, res_key='vote_instances', mws=mws)
    del kw['result_folder']
    t0 = time.time()
    sp.main(pred_path, result_folder=os.path.join(root, 'out'), **kw)
    dt = time.time() - t0
    out = ref_runner.FakeGroup.open(os.path.join(root, 'out', 'sample.hdf'))
    if mws:
        return np.asarray(out['vote_instances'])
    blk = ref_runner.FakeGroup.open(os.path.join(root, 'out', 'sample.zarr'))
    import hashlib
    res = dict(
        pred_sha1=hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest(),
        numinst=numinst, patchshape=ps.astype(np.int32),
        kwargs=json.dumps({k: v for k, v in kw.items()
                           if isinstance(v, (bool, int, float, str, list)) or v is None}),
        synth=json.dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in skw.items()}),
        instances=np.asarray(out['vote_instances']),
        foreground=np.asarray(out['vote_foreground']),
    )
    for k, v in blk.d.items():
        res['blk/' + k.replace('volumes/blocks/', '')] = np.asarray(v)
    fn = os.path.join(GOLD, 'blockwise3d_ps5.npz')
    np.savez_compressed(fn, **res)
    print('%-24s %6.1fs  blocks+faces=%d inst=%d  %.2f MB' % (
        'blockwise3d_ps5', dt, len(blk.d) // 2, len(np.unique(res['instances'])) - 1,
        os.path.getsize(fn) / 1e6))


def run_blockwise_big(S):
    """the reference's blockwise driver on a 3x3x3 block grid with the mutex-watershed
    partition (the flylight default): labels only."""
    import hashlib
    import shutil
    sp = ref_runner.load_stitch_module(S)
    ps = np.array([5, 5, 5])
    skw = dict(kind='neurites', seed=23, shape=(36, 66, 66), n=14, radius=(1.5, 2.5),
               seg_len=9.0, n_seg=10)
    pred, numinst, labels = synth.make_case(patchshape=ps, **skw)
    root = '/tmp/ppp_gold_blk3'
    shutil.rmtree(root, ignore_errors=True)
    pred_path = os.path.join(root, 'sample.zarr')
    store = ref_runner.FakeGroup.open(pred_path, 'w')
    store['volumes/pred_affs'] = pred.astype(np.float16)
    prob = np.stack([(numinst == 0), (numinst == 1), (numinst > 1)]).astype(np.float32)
    store['volumes/pred_numinst'] = prob
    res = {}
    for mws in (True, False):
        kw = S.default_kwargs(
            blockwise=True, chunksize=[12, 22, 22], patchshape=[5, 5, 5],
            aff_key='volumes/pred_affs', numinst_key='volumes/pred_numinst', fg_key=None,
            numinst_threshs=[0.9, 0.1], only_bb=False, output_format='hdf',
            num_parallel_blocks=1, ignore_small_comps=0, skeletonize_foreground=False,
            remove_small_comps=0, res_key='vote_instances', mws=mws)
        del kw['result_folder']
        out_dir = os.path.join(root, 'out_mws' if mws else 'out_cc')
        t0 = time.time()
        sp.main(pred_path, result_folder=out_dir, **kw)
        dt = time.time() - t0
        out = ref_runner.FakeGroup.open(os.path.join(out_dir, 'sample.hdf'))
        res['instances_mws' if mws else 'instances_cc'] = np.asarray(out['vote_instances'])
        print('blockwise 3x3x3 mws=%s %.1fs inst=%d' % (
            mws, dt, len(np.unique(np.asarray(out['vote_instances']))) - 1), flush=True)
    res.update(
        pred_sha1=hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest(),
        numinst=numinst, patchshape=ps.astype(np.int32),
        kwargs=json.dumps({k: v for k, v in kw.items()
                           if isinstance(v, (bool, int, float, str, list)) or v is None}),
        synth=json.dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in skw.items()}))
    fn = os.path.join(GOLD, 'blockwise3d_3x3x3_mws.npz')
    np.savez_compressed(fn, **res)
    print('blockwise3d_3x3x3_mws %.2f MB' % (os.path.getsize(fn) / 1e6))


from tests.golden_util import bb_case  # noqa: E402


def run_blockwise_bb(S):
    """the reference's blockwise driver with the flylight pre-crop and post side:
    only_bb + ignore_small_comps (clean_mask), remove_small_comps + relabel, dilated and
    masked outputs (stitch_patch_graph.py:745-764, 824-894), mws partition."""
    import hashlib
    import importlib.util
    import shutil
    sp = ref_runner.load_stitch_module(S)
    spec = importlib.util.spec_from_file_location(
        'ppp_ref_postprocess', os.path.join(ref_runner.REF_ROOT, 'PatchPerPix', 'util',
                                            'postprocess.py'))
    post = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(post)
    sp.remove_small_components = post.remove_small_components     # the real ones
    sp.relabel = post.relabel
    sp.io.imsave = lambda *a, **k: None
    ps, skw, pred, numinst = bb_case()
    root = '/tmp/ppp_gold_bb'
    shutil.rmtree(root, ignore_errors=True)
    pred_path = os.path.join(root, 'sample.zarr')
    store = ref_runner.FakeGroup.open(pred_path, 'w')
    store['volumes/pred_affs'] = pred.astype(np.float16)
    prob = np.stack([(numinst == 0), (numinst == 1), (numinst > 1)]).astype(np.float32)
    store['volumes/pred_numinst'] = prob
    kw = S.default_kwargs(
        blockwise=True, chunksize=[12, 22, 22], patchshape=[5, 5, 5],
        aff_key='volumes/pred_affs', numinst_key='volumes/pred_numinst', fg_key=None,
        numinst_threshs=[0.9, 0.1], only_bb=True, output_format='hdf',
        num_parallel_blocks=1, ignore_small_comps=30, skeletonize_foreground=False,
        remove_small_comps=60, dilate_instances=True, res_key='vote_instances', mws=True)
    del kw['result_folder']
    t0 = time.time()
    sp.main(pred_path, result_folder=os.path.join(root, 'out'), **kw)
    out = ref_runner.FakeGroup.open(os.path.join(root, 'out', 'sample.hdf'))
    res = {k: np.asarray(out[k]) for k in ('vote_instances', 'vote_foreground',
                                           'vote_instances_masked', 'vote_instances_dil_1',
                                           'vote_instances_masked_dil_1')}
    blk = ref_runner.FakeGroup.open(os.path.join(root, 'out', 'sample.zarr'))
    res['block_keys'] = np.array(sorted(k for k in blk.d if k.endswith('patch_pairs')))
    res.update(
        pred_sha1=hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest(),
        patchshape=ps.astype(np.int32),
        kwargs=json.dumps({k: v for k, v in kw.items()
                           if isinstance(v, (bool, int, float, str, list)) or v is None}))
    fn = os.path.join(GOLD, 'blockwise3d_bb_post.npz')
    np.savez_compressed(fn, **res)
    print('blockwise3d_bb_post %.1fs inst=%d blocks+faces=%d  %.2f MB' % (
        time.time() - t0, len(np.unique(res['vote_instances'])) - 1, len(res['block_keys']),
        os.path.getsize(fn) / 1e6))
    print(res['block_keys'][:6])


def run_mws(S):
    """mutex-watershed labelling (kwargs mws=True, the flylight default): the
    reference's setAffgraph + affGraphToInstances on the recorded pairs / aff of
    every case, plus graph_mws.mws on seeded random graphs (ties, self loops,
    repeated pairs) that exercise the id bookkeeping."""
    from tests import golden_util
    apg = S.mods['aff_patch_graph']
    g2l = S.mods['graph_to_labeling']
    res = {}
    for name in golden_util.NAMES:
        g, kw, ps, pred = golden_util.load(name)
        graph = apg.setAffgraph(g['aff'], g['pairs'])
        inst = np.zeros(pred.shape[1:], np.uint16)
        inst, _ = g2l.affGraphToInstances(
            graph, pred, ps, ps // 2, None, None, inst, g['gate'], mws=True,
            patch_threshold=kw['patch_threshold'], debug=False)
        res['inst/' + name] = inst
        print('mws %-24s inst=%d (cc: %d)' % (name, len(np.unique(inst)) - 1,
                                             len(np.unique(g['instances'])) - 1))
        # one_instance_per_channel (graph_to_labeling.py:57-95), both partitions; the
        # stack is stored as (channel of every painted voxel, voxel index) pairs
        if pred.shape[1] * pred.shape[2] * pred.shape[3] <= 20000:
            for tag, use_mws in (('cc', False), ('mws', True)):
                graph = apg.setAffgraph(g['aff'], g['pairs'])
                z = np.zeros(pred.shape[1:], np.uint16)
                stack, _ = g2l.affGraphToInstances(
                    graph, pred, ps, ps // 2, None, None, z, g['gate'], mws=use_mws,
                    one_instance_per_channel=True, patch_threshold=kw['patch_threshold'],
                    debug=False)
                res['opc_%s/%s' % (tag, name)] = np.packbits(stack > 0)
                res['opc_%s_shape/%s' % (tag, name)] = np.array(stack.shape, np.int32)
                vals = np.array([np.unique(c[c > 0]).tolist() or [0] for c in stack], np.int32)
                res['opc_%s_vals/%s' % (tag, name)] = vals.reshape(-1)
    rng = np.random.default_rng(99)
    n_graphs = 24
    for gi in range(n_graphs):
        nn = int(rng.integers(6, 160))
        coords = np.stack([rng.integers(0, 4, nn), rng.integers(0, 40, nn),
                           rng.integers(0, 40, nn)], 1).astype(np.uint32)
        coords = np.unique(coords, axis=0)
        rng.shuffle(coords)
        nn = len(coords)
        ne = int(nn * rng.uniform(1.0, 4.0))
        a = rng.integers(0, nn, ne)
        b = rng.integers(0, nn, ne)
        w = rng.normal(0.15, 0.5, ne).astype(np.float32)
        if gi % 3 == 0:
            w = np.round(w * 4) / 4                  # many ties and exact zeros
        if gi % 4 == 1:
            w = np.abs(w) * np.where(rng.random(ne) < 0.15, -1, 1).astype(np.float32)
        pairs = np.concatenate([coords[a], coords[b]], 1).astype(np.uint32)
        graph = apg.setAffgraph(w, pairs)
        ccs = S.mods['graph_mws'].mws(graph)
        lab = {}
        for k, cc in enumerate(ccs):
            for nd in cc:
                lab[tuple(int(v) for v in nd)] = k + 1
        nodes = np.array([list(nd) for nd in graph.nodes()], np.int32).reshape(-1, 3)
        labels = np.array([lab.get(tuple(int(v) for v in nd), 0) for nd in graph.nodes()],
                          np.int32)
        res['rnd/%02d/pairs' % gi] = pairs
        res['rnd/%02d/aff' % gi] = w
        res['rnd/%02d/nodes' % gi] = nodes
        res['rnd/%02d/labels' % gi] = labels
        print('mws random %02d nodes=%d edges=%d comps=%d max=%d' % (
            gi, len(nodes), graph.number_of_edges(), len(set(labels[labels > 0])),
            labels.max() if len(labels) else 0))
    res['n_random'] = np.int32(n_graphs)
    res['blockwise_inst'] = run_blockwise(ref_runner.RefSession(), mws=True)
    print('mws blockwise inst=%d' % (len(np.unique(res['blockwise_inst'])) - 1))
    fn = os.path.join(GOLD, 'mws_cases.npz')
    np.savez_compressed(fn, **res)
    print('mws_cases %.2f MB' % (os.path.getsize(fn) / 1e6))


from no_overlap_case import no_overlap_case  # noqa: E402


def run_no_overlap(S):
    """no_overlap_per_channel (graph_to_labeling.py:96-113): instances larger than 2000
    voxels go to the first channel they do not overlap, smaller ones into channel 0.
    Four big overlapping discs and a small one through the reference's to_instance_seg."""
    import hashlib
    ps = np.array([1, 9, 9])
    pred, numinst, labels = no_overlap_case()
    kw = S.default_kwargs(no_overlap_per_channel=True)
    fg = pred[40] > kw['patch_threshold']
    res = {}
    for tag, use_mws in (('cc', False), ('mws', True)):
        k2 = dict(kw, mws=use_mws)
        stack, _ = S.vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), numinst.copy(),
                                        ps.copy(), **k2)
        stack = np.asarray(stack)
        res['stack_' + tag] = stack.astype(np.uint16)
        print('no_overlap %s: channels %d, labels %s' % (tag, stack.shape[0],
                                                        [np.unique(c).tolist() for c in stack]))
    res['pred_sha1'] = hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest()
    res['numinst'] = numinst
    res['kwargs'] = json.dumps({k: v for k, v in kw.items()
                                if isinstance(v, (bool, int, float, str))})
    fn = os.path.join(GOLD, 'chan_nooverlap2d_ps9.npz')
    np.savez_compressed(fn, **res)
    print('chan_nooverlap2d_ps9 %.2f MB' % (os.path.getsize(fn) / 1e6))


def main():
    if sys.argv[1:] == ['nooverlap']:
        run_no_overlap(ref_runner.RefSession())
        return 0
    if sys.argv[1:] == ['mws']:
        run_mws(ref_runner.RefSession())
        return 0
    if sys.argv[1:] == ['blockwise_bb']:
        run_blockwise_bb(ref_runner.RefSession())
        return 0
    if sys.argv[1:] == ['blockwise3']:
        run_blockwise_big(ref_runner.RefSession())
        return 0
    if sys.argv[1:] == ['blockwise']:
        S = ref_runner.RefSession()
        run_blockwise(S)
        return 0
    if sys.argv[1:] and all(n in BIG_CASES for n in sys.argv[1:]):
        for n in sys.argv[1:]:
            run_big(n)
        return 0
    names = sys.argv[1:] or list(CASES)
    rec = Recorder()
    S = ref_runner.RefSession(recorder=rec)
    for n in names:
        run_case(n, S, rec)
    return 0


if __name__ == '__main__':
    sys.exit(main())
