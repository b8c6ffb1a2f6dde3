"""The reference-facing boundary of the stage: containers, foreground rule,
bounding-box pre-crop, post-processing, and the shipped flylight configuration
passed VERBATIM to both entry points (run_ppp.py:1163-1190).

CPU tests cover the host logic against numpy restatements of the reference's
functions (file:line in each test); the GPU tests run the entry points."""
import json
import os
import zipfile
import zlib

import numpy as np
import pytest

from patchperpix_b200 import io_util, synth
from patchperpix_b200 import postprocess as pp
from patchperpix_b200 import stitch_patch_graph as spg
from patchperpix_b200 import utilVoteInstances as uvi

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, 'golden', 'flylight_default_kwargs.json')
REF_TOML = '/root/reference/experiments/flylight/setups/setup01/default.toml'
REF_ZIP = '/root/reference/experiments/flylight/JRC_SS05008-20160318_24_B2_crop.zip'


def flylight_kwargs(blockwise=True):
    """what run_ppp.py:1169-1190 passes for the shipped flylight setup."""
    d = json.load(open(FIX))
    kw = dict(d['vote_instances'])
    kw.update(d['model'])
    p = d['prediction']
    if blockwise:
        kw.update(d['visualize'])
        kw.update(aff_key=p['aff_key'], numinst_key=p['numinst_key'], fg_key=p['fg_key'],
                  fg_folder=p['fg_folder'], fg_thresh=p['fg_thresh'])
    else:
        kw.update(aff_key=p['aff_key'], numinst_key=p['numinst_key'], fg_key=p['fg_key'])
    return kw


# ---------------------------------------------------------------------------
# containers
# ---------------------------------------------------------------------------
def test_zarrlite_roundtrip(tmp_path):
    g = io_util.ZarrLiteGroup(str(tmp_path / 'a.zarr'), 'w')
    a = np.random.default_rng(0).random((5, 13, 17, 9)).astype(np.float16)
    g.create_dataset('volumes/pred_affs', data=a, chunks=(5, 4, 8, 4))
    g.create_dataset('volumes/gz', data=a, chunks=(2, 13, 5, 9), compressor={'id': 'gzip'})
    g.create_dataset('volumes/raw', data=a, compressor=None)
    r = io_util.open_zarr(str(tmp_path / 'a.zarr'))
    for k in ('volumes/pred_affs', 'volumes/gz', 'volumes/raw'):
        z = r[k]
        assert z.shape == a.shape and z.dtype == a.dtype
        assert np.array_equal(np.array(z), a)
        assert np.array_equal(z[2], a[2])
        assert np.array_equal(z[:, 3:11, 2:9, 1:8], a[:, 3:11, 2:9, 1:8])
        assert np.array_equal(z[171 % 5:4, -1], a[1:4, -1])
        assert np.array_equal(z[..., 3:5], a[..., 3:5])
    assert 'volumes/gz' in r and 'volumes' in r and 'volumes/nope' not in r
    with pytest.raises(RuntimeError, match='lzma'):
        io_util._decompress(b'', {'id': 'lzma'}, 'chunk')


def test_zstd_chunks_through_pyarrow():
    pa = pytest.importorskip('pyarrow')
    raw = np.arange(5000, dtype=np.uint16).tobytes()
    assert io_util._zstd_decompress(pa.Codec('zstd').compress(raw, asbytes=True)) == raw


def _bitshuffle_by_stages(block, ts):
    """bitshuffle's scalar route, stage by stage (byte transpose of the elements, 8x8 bit
    transposes of consecutive bytes into eight bit rows, regrouping of the bit rows per
    element byte), written with plain loops: the forward direction, independent of the
    closed form io_util._bit_unshuffle inverts."""
    size = len(block) // ts
    if size % 8 or size == 0:
        return bytes(block)                      # c-blosc 1.x copies such a block
    n = size * ts
    a = bytearray(n)
    for i in range(size):                        # stage 1: [i][j] -> [j][i]
        for j in range(ts):
            a[j * size + i] = block[i * ts + j]
    rowlen = n // 8
    b = bytearray(n)
    for ii in range(rowlen):                     # stage 2: bit k of bytes 8ii..8ii+7 -> row k
        for k in range(8):
            v = 0
            for q in range(8):
                v |= ((a[8 * ii + q] >> k) & 1) << q
            b[k * rowlen + ii] = v
    c = bytearray(n)
    per = size // 8
    for k in range(8):                           # stage 3: [k][j][per] -> [j][k][per]
        for j in range(ts):
            c[(j * 8 + k) * per:(j * 8 + k + 1) * per] = b[(k * ts + j) * per:(k * ts + j + 1) * per]
    return bytes(c) + bytes(block[n:])


def _blosc_frame(raw, ts, blocksize, mode, codec, split=True):
    """a c-blosc 1.x chunk built from the format description (header, block starts, per-block
    streams with i32 lengths, raw storage when compression does not shrink a stream)."""
    import pyarrow as pa
    comp = {4: lambda x: pa.Codec('zstd').compress(x, asbytes=True),
            3: lambda x: zlib.compress(x, 1),
            1: lambda x: pa.Codec('lz4_raw').compress(x, asbytes=True)}[codec]
    nblocks = (len(raw) + blocksize - 1) // blocksize
    flags = (codec << 5) | {0: 0, 1: 0x1, 2: 0x4}[mode] | (0 if split else 0x10)
    body, starts = b'', []
    for b in range(nblocks):
        blk = raw[b * blocksize:(b + 1) * blocksize]
        if mode == 1 and ts > 1:
            q = len(blk) // ts
            blk = b''.join(bytes(blk[j:q * ts:ts]) for j in range(ts)) + blk[q * ts:]
        elif mode == 2 and len(blk) >= ts:
            blk = _bitshuffle_by_stages(blk, ts)
        do_split = split and ts <= 16 and blocksize // ts >= 128 and len(blk) == blocksize
        ns = ts if do_split else 1
        ne = len(blk) // ns
        starts.append(16 + 4 * nblocks + len(body))
        for j in range(ns):
            part = blk[j * ne:(j + 1) * ne]
            c = comp(part)
            if len(c) >= ne:
                c = part
            body += len(c).to_bytes(4, 'little') + c
    head = bytes([2, 1, flags, ts]) + len(raw).to_bytes(4, 'little') + blocksize.to_bytes(4, 'little')
    total = 16 + 4 * nblocks + len(body)
    return head + total.to_bytes(4, 'little') + b''.join(x.to_bytes(4, 'little') for x in starts) + body


def test_blosc_chunks(tmp_path):
    """Blosc frames as predict_no_gp.py:243-257 writes them (zstd, bit-shuffle) and the other
    shapes a frame can take: byte shuffle, no shuffle, split and unsplit blocks, a short
    last block, streams stored raw, the whole chunk stored raw.  The encoder above is the
    test's own (no Blosc library exists in this image)."""
    pytest.importorskip('pyarrow')
    rng = np.random.default_rng(5)
    smooth = (np.cumsum(rng.integers(0, 3, 6000)) % 2048).astype(np.float16)
    noise = rng.integers(0, 65536, 3000).astype(np.uint16)
    for arr in (smooth, noise, smooth.astype(np.float32), smooth[:1003].astype(np.uint8)):
        raw, ts = arr.tobytes(), arr.dtype.itemsize
        for blocksize in (512, 2048, 4096 + 8 * ts, len(raw), 2 * len(raw)):
            for mode in (0, 1, 2):
                for codec, split in ((4, True), (4, False), (3, True), (1, False)):
                    fr = _blosc_frame(raw, ts, blocksize, mode, codec, split)
                    assert io_util.blosc_decode(fr) == raw, (arr.dtype, blocksize, mode, codec, split)
    # stored uncompressed (flag 0x2): the data follows the header
    fr = bytes([2, 1, 0x2 | 0x4 | (4 << 5), 2]) + (10).to_bytes(4, 'little') + \
        (10).to_bytes(4, 'little') + (26).to_bytes(4, 'little') + bytes(range(10))
    assert io_util.blosc_decode(fr) == bytes(range(10))
    with pytest.raises(RuntimeError, match='header says'):
        io_util.blosc_decode(fr + b'x')
    with pytest.raises(RuntimeError, match='blosclz'):
        io_util.blosc_decode(_blosc_frame(b'ab' * 600, 2, 2400, 0, 4)[:2] + bytes([0]) +
                             _blosc_frame(b'ab' * 600, 2, 2400, 0, 4)[3:])
    # a store whose .zarray names the blosc compressor, read through ZarrLite
    a = (rng.random((3, 6, 10, 8)) * 0.9).astype(np.float16)
    d = tmp_path / 'b.zarr' / 'volumes' / 'pred_affs'
    os.makedirs(d)
    chunks = (3, 4, 10, 8)
    with open(d / '.zarray', 'w') as f:
        json.dump(dict(zarr_format=2, shape=a.shape, chunks=chunks, dtype='<f2', order='C',
                       fill_value=0, filters=None,
                       compressor=dict(id='blosc', cname='zstd', clevel=3, shuffle=2, blocksize=0)), f)
    with open(tmp_path / 'b.zarr' / '.zgroup', 'w') as f:
        json.dump(dict(zarr_format=2), f)
    with open(tmp_path / 'b.zarr' / 'volumes' / '.zgroup', 'w') as f:
        json.dump(dict(zarr_format=2), f)
    for c in range(2):
        blk = np.zeros(chunks, np.float16)
        part = a[:, 4 * c:4 * c + 4]
        blk[:, :part.shape[1]] = part
        with open(d / ('0.%d.0.0' % c), 'wb') as f:
            f.write(_blosc_frame(blk.tobytes(), 2, 1024, 2, 4))
    z = io_util.ZarrLiteGroup(str(tmp_path / 'b.zarr'))['volumes/pred_affs']
    assert np.array_equal(np.array(z), a)
    assert np.array_equal(z[:, 3:6, 2:5], a[:, 3:6, 2:5])


@pytest.mark.skipif(not os.path.exists(REF_ZIP), reason="reference tree absent")
def test_bundled_flylight_sample_is_readable(tmp_path):
    """BASELINE configs[0] input: gzip-compressed zarr chunks, read without zarr."""
    with zipfile.ZipFile(REF_ZIP) as z:
        z.extractall(tmp_path)
    f = io_util.open_container(str(tmp_path / 'JRC_SS05008-20160318_24_B2_crop.zarr'))
    raw, gt = f['volumes/raw'], f['volumes/gt_instances']
    assert raw.shape == (3, 50, 50, 50) and gt.shape == (3, 50, 50, 50)
    assert int((np.array(gt) > 0).sum()) == 24204


# ---------------------------------------------------------------------------
# foreground rule, bounding box
# ---------------------------------------------------------------------------
def test_foreground_precedence():
    """utilVoteInstances.py:275-322: fg_key, else numinst > 0, else the centre channel."""
    rng = np.random.default_rng(1)
    affs = rng.random((27, 4, 5, 6)).astype(np.float32)
    fg = rng.random((1, 4, 5, 6)).astype(np.float32)
    numinst = rng.integers(0, 3, (4, 5, 6)).astype(np.uint8)
    kw = dict(patchshape=[3, 3, 3], patch_threshold=0.5, fg_thresh_vi=-1)
    assert np.array_equal(uvi.returnFg(affs, numinst, fg, fg_key='f', numinst_key='n', **kw),
                          fg[0] > 0.5)
    assert np.array_equal(uvi.returnFg(affs, numinst, fg, fg_key=None, numinst_key='n', **kw),
                          (numinst > 0) > 0.5)
    assert np.array_equal(uvi.returnFg(affs, None, None, fg_key=None, numinst_key=None, **kw),
                          affs[13] > 0.5)
    kw['fg_thresh_vi'] = 0.8
    assert np.array_equal(uvi.returnFg(affs, None, None, fg_key=None, numinst_key=None, **kw),
                          affs[13] > 0.8)
    # block-level view of a volume follows the same rule, numinst still from numinst_key
    prob = np.stack([(numinst == 0), (numinst == 1), (numinst > 1)]).astype(np.float32)
    vol = spg.VolumeInputs(affs, numinst_prob=prob, fg=fg)
    f, n = vol._fg_numinst(np.zeros(3, int), np.array(vol.shape), fg_key='f', numinst_key='n',
                           numinst_threshs=[0.9, 0.1], **kw)
    assert np.array_equal(f, fg[0] > 0.8) and np.array_equal(n, numinst)


def test_two_dimensional_prediction_is_lifted():
    vol = spg.VolumeInputs(np.zeros((9, 20, 30), np.float16))
    assert vol.shape == (1, 20, 30) and vol.pred.shape == (9, 1, 20, 30)


def test_clean_mask_and_bbox():
    """stitch_patch_graph.py:46-57, 745-764."""
    from scipy import ndimage
    rng = np.random.default_rng(0)
    m = rng.random((10, 30, 30)) < 0.08
    m[2:6, 5:20, 8:12] = True
    labeled = ndimage.label(m, np.ones((3, 3, 3)))[0]
    labels, counts = np.unique(labeled, return_counts=True)
    ref = np.isin(labeled, labels[counts > 5]) & (labeled > 0)
    assert np.array_equal(pp.clean_mask(m, np.ones((3, 3, 3)), 5), ref)
    off, shape = pp.foreground_bbox(m, ignore_small_comps=5)
    nz = np.argwhere(ref)
    assert np.array_equal(off, nz.min(0)) and np.array_equal(shape, nz.max(0) - nz.min(0) + 1)
    assert pp.foreground_bbox(np.zeros((3, 4, 5), bool)) is None


def test_bbox_without_skeleton_contains_reference_box():
    """DOCUMENTED DEVIATION: with skeletonize_foreground the reference takes the box of
    the 3-D skeleton (skimage, optional here).  Without skimage the box of the cleaned
    mask is used; any subset of the mask -- the skeleton is one -- has its box inside."""
    m = np.zeros((12, 40, 40), bool)
    m[3:9, 5:35, 18:23] = True
    off, shape = pp.foreground_bbox(m, ignore_small_comps=10, skeletonize_foreground=True)
    try:
        import skimage  # noqa: F401
        return                      # real skeleton available: nothing to document
    except ImportError:
        pass
    assert np.array_equal(off, [3, 5, 18]) and np.array_equal(shape, [6, 30, 5])
    sub = np.zeros_like(m)
    sub[5:7, 8:30, 20] = True       # a thinner curve inside the mask
    o2, s2 = pp.foreground_bbox(sub)
    assert np.all(o2 >= off) and np.all(o2 + s2 <= off + shape)


# ---------------------------------------------------------------------------
# post side
# ---------------------------------------------------------------------------
def test_post_ops_match_reference_semantics():
    """util/postprocess.py:24-52, stitch_patch_graph.py:873-881 restated in numpy."""
    import torch
    from scipy import ndimage
    rng = np.random.default_rng(0)
    a = (rng.integers(0, 6, (12, 20, 20)) * (rng.random((12, 20, 20)) < 0.3)).astype(np.int32)
    a[2:5, 3:9, 3:9] = 7
    a[0, 0, 0] = 9

    def ref_remove(array, compsize):
        labels, counts = np.unique(array, return_counts=True)
        out = array.copy()
        out[np.isin(array, labels[counts <= compsize])] = 0
        return out

    def ref_relabel(array):
        out = np.zeros_like(array)
        c = 1
        for l in np.unique(array):
            if l != 0:
                out[array == l] = c
                c += 1
        return out

    def ref_dilate(inst):
        d = inst.copy()
        for l in np.unique(inst):
            if l != 0:
                d[ndimage.binary_dilation(d == l, iterations=1)] = l
        return d
    t = torch.from_numpy(a)
    assert np.array_equal(pp.remove_small_components(t, 300).numpy(), ref_remove(a, 300))
    assert np.array_equal(pp.relabel(pp.remove_small_components(t, 300)).numpy(),
                          ref_relabel(ref_remove(a, 300)))
    assert np.array_equal(pp.dilate_instances(t).numpy(), ref_dilate(a))
    fg = a > 0
    out = spg.finish_outputs(a, fg, remove_small_comps=300, dilate_instances=True)
    want = ref_relabel(ref_remove(a, 300))
    assert np.array_equal(out['vote_instances'], want.astype(np.uint16))
    assert np.array_equal(out['vote_instances_masked'], np.where(fg, want, 0))
    assert np.array_equal(out['vote_instances_dil_1'], ref_dilate(want))
    assert set(out) == {'vote_instances', 'vote_foreground', 'vote_instances_masked',
                        'vote_instances_dil_1', 'vote_instances_masked_dil_1'}


@pytest.mark.skipif(not os.path.exists(REF_TOML), reason="reference tree absent")
def test_fixture_is_the_shipped_toml():
    import tomllib
    with open(REF_TOML, 'rb') as f:
        cfg = tomllib.load(f)
    d = json.load(open(FIX))
    assert d['vote_instances'] == cfg['vote_instances'] and d['model'] == cfg['model']
    assert d['visualize'] == cfg.get('visualize', {})


def test_unsupported_keys_raise_by_name():
    from patchperpix_b200 import vote_instances as vi
    pred = np.zeros((27, 4, 8, 8), np.float32)
    fg = np.zeros((4, 8, 8), bool)
    for key, val in (('isbiHack', True), ('sample', 0.5), ('thin_cover_use_kd', True),
                     ('mark_close_neighboorhood', True), ('cuda', False)):
        kw = dict(patch_threshold=0.5, cuda=True)
        kw[key] = val
        with pytest.raises(NotImplementedError):
            vi.to_instance_seg(pred, fg, fg, fg, [3, 3, 3], **kw)


# ---------------------------------------------------------------------------
# GPU: the entry points with the shipped configuration
# ---------------------------------------------------------------------------
def _volume(tmp_path, shape=(24, 48, 48)):
    ps = np.array([7, 7, 7])
    pred, numinst, labels = synth.make_case('neurites', ps, seed=21, shape=shape, n=5,
                                            radius=(2.0, 3.0), seg_len=10.0, n_seg=10)
    pred[:, :, :6, :] = 0            # an empty margin, so that only_bb really crops
    numinst[:, :6, :] = 0
    prob = np.stack([(numinst == 0), (numinst == 1), (numinst > 1)]).astype(np.float32)
    path = str(tmp_path / 'sample.zarr')
    g = io_util.ZarrLiteGroup(path, 'w')
    g.create_dataset('volumes/pred_affs', data=pred.astype(np.float16),
                     chunks=(343, 12, 24, 24))
    g.create_dataset('volumes/pred_numinst', data=prob, chunks=(3, 12, 24, 24))
    return path, pred, numinst, prob


@pytest.mark.gpu
def test_blockwise_entry_point_with_flylight_toml(tmp_path):
    """stitch_patch_graph.main(pred_file, **[vote_instances], **[model], **[visualize],
    aff_key=..., ...) exactly as run_ppp.py:1169-1179 calls it."""
    from oracle import host_logic
    path, pred, numinst, prob = _volume(tmp_path)
    kw = flylight_kwargs(blockwise=True)
    kw['chunksize'] = [16, 24, 24]           # several blocks on the small test volume
    out_dir = str(tmp_path / 'out')
    inst = spg.main(path, result_folder=out_dir, **dict(kw))
    assert inst is not None and inst.max() > 0
    res = np.load(os.path.join(out_dir, 'sample.npz')) if not os.path.exists(
        os.path.join(out_dir, 'sample.hdf')) else None
    if res is not None:
        assert set(res.files) >= {'vote_instances', 'vote_foreground', 'vote_instances_masked'}
        assert np.array_equal(res['vote_instances'], inst.astype(np.uint16))
    assert os.path.exists(os.path.join(out_dir, 'sample.png'))          # save_mip
    # the same through the host logic with the CPU oracle as block engine
    inputs = spg.VolumeInputs(pred.astype(np.float16), numinst_prob=prob)
    bb = spg.bounding_box(inputs, **kw)
    assert bb[0][1] >= 6, "only_bb did not crop the empty margin"
    ref, _, _ = spg.stitch_arrays(inputs, block_fn=host_logic.oracle_block_fn,
                                  paint_fn=host_logic.oracle_paint_fn, bb_offset=bb[0],
                                  bb_shape=bb[1], **dict(kw))
    assert np.array_equal(inst.astype(np.uint16), ref.astype(np.uint16))
    # the same call on the compact-row engine (sharded.py; the default under torchrun)
    rows = spg.main(path, result_folder=str(tmp_path / 'out_rows'), **dict(kw, ppp_rows=True))
    assert np.array_equal(rows.astype(np.uint16), inst.astype(np.uint16))
    # second call: every block and face comes from the cache (skip-if-exists, :584-587)
    cache = io_util.open_zarr(os.path.join(out_dir, 'sample.zarr'))
    assert len(cache['volumes/blocks'].keys()) > 1
    from patchperpix_b200 import vote_instances as vi
    calls = []
    orig = vi.do_block
    vi.do_block = lambda *a, **k: calls.append(1) or orig(*a, **k)
    try:
        again = spg.main(path, result_folder=out_dir, **dict(kw))
    finally:
        vi.do_block = orig
    assert not calls and np.array_equal(again, inst)


@pytest.mark.gpu
def test_single_block_entry_point_with_flylight_toml(tmp_path):
    """vote_instances.main(**[vote_instances], **[model], numinst_key=..., aff_key=...,
    fg_key=...) as run_ppp.py:1184-1190 calls it (blockwise switched off)."""
    from patchperpix_b200 import vote_instances as vi
    from oracle import cpu_oracle, host_logic
    path, pred, numinst, prob = _volume(tmp_path, shape=(20, 40, 40))
    kw = flylight_kwargs(blockwise=False)
    kw.update(blockwise=False, return_intermediates=False, affinities=path,
              result_folder=str(tmp_path / 'out1'), check_required=False)
    # DOCUMENTED DEVIATION: outside the blockwise driver the reference thins the mask to
    # cover to its 3-D skeleton (vote_instances.py:220-224, skimage); that key is refused
    # by name here, the blockwise driver (the flylight default) never reads it
    with pytest.raises(NotImplementedError, match='skeletonize_foreground'):
        vi.main(**kw)
    kw['skeletonize_foreground'] = False
    vi.main(**kw)
    res = np.load(os.path.join(kw['result_folder'], 'sample.npz'))
    p16 = pred.astype(np.float16).astype(np.float32)
    ni = uvi.numinst_from_prob(prob, **kw)
    fg = (ni > 0) > 0.5
    O = cpu_oracle.Oracle(p16, ni > 1, np.array([7, 7, 7]), cpu_oracle.variant_from_kwargs(kw))
    ref = host_logic.assemble(p16, fg, ni, np.array([7, 7, 7]), kw, O)
    want = ref['instances'].copy()
    want[~fg] = 0                                        # crop_to_foreground, :535-540
    assert np.array_equal(res['vote_instances'], want)


@pytest.mark.gpu
def test_score_threshold_cuts_the_cover():
    """foreground_cover.py:136-138 (a float score_threshold stops the walk)."""
    from patchperpix_b200 import vote_instances as vi
    from oracle import cpu_oracle, host_logic
    ps = np.array([1, 9, 9])
    pred, numinst, _ = synth.make_case('worms', ps, seed=5, shape=(64, 80), n_worms=4,
                                       width=(5, 8), length=(30, 70), hard_frac=0.05)
    kw = dict(json.load(open(FIX))['vote_instances'], blockwise=False, mws=False,
              return_intermediates=False, skeletonize_foreground=False, overlapping_inst=True)
    fg = pred[40] > np.float32(0.5)
    O = cpu_oracle.Oracle(pred, numinst > 1, ps, cpu_oracle.variant_from_kwargs(kw))
    base = host_logic.assemble(pred, fg, numinst, ps, kw, O)['instances']
    hit = False
    for thr in (0.35, 0.6):
        k2 = dict(kw, score_threshold=thr)
        inst, _ = vi.to_instance_seg(pred, fg, fg.copy(), numinst, ps, **k2)
        ref = host_logic.assemble(pred, fg, numinst, ps, k2, O)['instances']
        assert np.array_equal(inst, ref), thr
        hit = hit or not np.array_equal(ref, base)
    assert hit, "thresholds too low to change anything: the test would be vacuous"


def test_gather_patches_reads_boxes_not_the_volume(tmp_path):
    """stitch_patch_graph.py:367-385: node patches of a prediction that stays on disk."""
    rng = np.random.default_rng(3)
    a = rng.random((27, 20, 70, 90)).astype(np.float16)
    g = io_util.ZarrLiteGroup(str(tmp_path / 'p.zarr'), 'w')
    g.create_dataset('volumes/pred_affs', data=a, chunks=(27, 8, 32, 32))
    arr = io_util.open_zarr(str(tmp_path / 'p.zarr'))['volumes/pred_affs']
    z, y, x = rng.integers(0, 20, 300), rng.integers(0, 70, 300), rng.integers(0, 90, 300)
    got = spg.gather_patches(arr, z, y, x, tile=16)
    assert np.array_equal(got, a[:, z, y, x].T.astype(np.float32))
    assert np.array_equal(spg.gather_patches(a, z, y, x), got)


def test_pair_order_fallback_without_the_set_replay(monkeypatch):
    """aff_patch_graph.py:57-110 enumerates a python SET; ppp_pyset_order replays CPython's
    table, and if its self-check ever fails (another interpreter) the plain python path
    must give the same pairs in the same order."""
    import scipy.spatial
    from patchperpix_b200 import assembly
    rng = np.random.default_rng(5)
    pts = np.unique(rng.integers(0, 40, (400, 3)), axis=0).astype(np.uint32)
    tree = scipy.spatial.cKDTree(pts, leafsize=4)
    thr = np.array([14.0, 14.0, 14.0])
    assert assembly._pyset_replay_ok()
    fast = assembly.query_pairs_filtered(tree, pts, 42, thr)
    fast2 = assembly.query_pairs_set_order(tree, 42)
    monkeypatch.setattr(assembly, '_PYSET_REPLAY', False)
    slow = assembly.query_pairs_filtered(tree, pts, 42, thr)
    slow2 = assembly.query_pairs_set_order(tree, 42)
    assert len(fast) > 100 and np.array_equal(fast, slow) and np.array_equal(fast2, slow2)


# ---------------------------------------------------------------------------
# the pre-crop and the post side against a run of the reference's own blockwise driver
# (tools/gen_golden.py blockwise_bb: only_bb + ignore_small_comps, remove_small_comps +
# relabel with the reference's util functions, dilated and masked outputs, mws)
# ---------------------------------------------------------------------------
def _bb_golden():
    from tests.golden_util import bb_case
    g = dict(np.load(os.path.join(HERE, 'golden', 'blockwise3d_bb_post.npz')))
    kw = json.loads(str(g['kwargs']))
    ps, skw, pred, numinst = bb_case()
    prob = np.stack([(numinst == 0), (numinst == 1), (numinst > 1)]).astype(np.float32)
    return g, kw, pred, numinst, prob


_BB_KEYS = ('vote_instances', 'vote_foreground', 'vote_instances_masked', 'vote_instances_dil_1',
            'vote_instances_masked_dil_1')


def test_bounding_box_grid_and_post_side_match_the_reference_run():
    """host logic (bbox, block grid at the box corner, post-processing) with the oracle as
    block engine: the five output volumes of the reference's run, value for value."""
    from oracle import host_logic
    g, kw, pred, numinst, prob = _bb_golden()
    inputs = spg.VolumeInputs(pred.astype(np.float16), numinst_prob=prob)
    bb = spg.bounding_box(inputs, **kw)
    # the block keys the reference wrote start at the box corner
    blocks = {str(k).split('/')[2] for k in g['block_keys'] if str(k).count('/') == 3}
    offs = {spg.get_offset_str(o + bb[0]) for o in spg.get_offsets(bb[1], np.minimum(
        kw['chunksize'], bb[1]))}
    assert spg.get_offset_str(bb[0]) in blocks and blocks <= offs
    inst, fg, _ = spg.stitch_arrays(inputs, block_fn=host_logic.oracle_block_fn,
                                    paint_fn=host_logic.oracle_paint_fn, bb_offset=bb[0],
                                    bb_shape=bb[1], **kw)
    out = spg.finish_outputs(inst, fg, **kw)
    for k in _BB_KEYS:
        assert np.array_equal(out[k], g[k].astype(np.uint16)), k


@pytest.mark.gpu
@pytest.mark.parametrize('rows', [False, True])
def test_file_level_entry_matches_the_reference_run(tmp_path, rows):
    """stitch_patch_graph.main on a zarr store, dense engine and compact-row engine."""
    g, kw, pred, numinst, prob = _bb_golden()
    path = str(tmp_path / 'sample.zarr')
    z = io_util.ZarrLiteGroup(path, 'w')
    z.create_dataset('volumes/pred_affs', data=pred.astype(np.float16), chunks=(125, 10, 30, 30))
    z.create_dataset('volumes/pred_numinst', data=prob, chunks=(3, 10, 30, 30))
    out_dir = str(tmp_path / 'out')
    spg.main(path, result_folder=out_dir, **dict(kw, ppp_rows=rows))
    fn = os.path.join(out_dir, 'sample.npz')
    if not os.path.exists(fn):
        pytest.skip("h5py present: .hdf written, checked elsewhere")
    res = np.load(fn)
    for k in _BB_KEYS:
        assert np.array_equal(res[k], g[k].astype(np.uint16)), k


# ---------------------------------------------------------------------------
# loadAffinities (utilVoteInstances.py:136-251) on real containers
# ---------------------------------------------------------------------------
def _store(tmp_path, name, arrays):
    path = str(tmp_path / name)
    g = io_util.ZarrLiteGroup(path, 'w')
    for k, v in arrays.items():
        g.create_dataset(k, data=v)
    return path


def test_load_affinities_zarr_3d_with_crops_and_numinst(tmp_path):
    rng = np.random.default_rng(7)
    pred = rng.random((27, 6, 8, 10)).astype(np.float16)
    prob = rng.random((3, 6, 8, 10)).astype(np.float32)
    path = _store(tmp_path, 's.zarr', {'volumes/pred_affs': pred, 'volumes/pred_numinst': prob})
    kw = dict(aff_key='volumes/pred_affs', numinst_key='volumes/pred_numinst', fg_key=None,
              patch_threshold=0.5, numinst_threshs=[0.9, 0.1])
    aff, numinst, fg = uvi.loadAffinities(path, '', patchshape=np.array([3, 3, 3]), **kw)
    assert aff.shape == (27, 6, 8, 10) and np.array_equal(aff, pred)
    want = np.zeros((6, 8, 10), np.uint8)
    want[prob[1] > 0.9] = 1
    want[prob[2] > 0.1] = 2
    assert np.array_equal(np.squeeze(numinst), want)
    assert np.array_equal(np.squeeze(fg), want > 0)
    # crop_*_s / crop_*_e (:171-200)
    aff_c, _, _ = uvi.loadAffinities(path, '', patchshape=np.array([3, 3, 3]),
                                     **dict(kw, crop_z_s=1, crop_z_e=5, crop_y_s=2, crop_x_e=7))
    assert np.array_equal(aff_c, pred[:, 1:5, 2:, :7])


def test_load_affinities_channel_last_2d_and_logits(tmp_path):
    """2-D predictions are lifted to Z = 1, a channel-last array is rotated (:158-164),
    stored logits are squashed with expit (:249-250)."""
    rng = np.random.default_rng(8)
    logits = (rng.normal(0, 3, (12, 14, 25))).astype(np.float16)      # [Y,X,P], P = 5x5
    fgp = rng.random((1, 12, 14)).astype(np.float32)
    path = _store(tmp_path, 'l.zarr', {'volumes/pred_affs': logits, 'volumes/pred_fgbg': fgp})
    aff, numinst, fg = uvi.loadAffinities(path, '', patchshape=np.array([1, 5, 5]),
                                          aff_key='volumes/pred_affs', numinst_key=None,
                                          fg_key='volumes/pred_fgbg', patch_threshold=0.5)
    assert aff.shape == (25, 1, 12, 14)
    want = 1.0 / (1.0 + np.exp(-np.moveaxis(logits, -1, 0).astype(np.float32)))
    assert np.allclose(aff[:, 0], want, atol=1e-6)
    assert numinst is None
    # the foreground comes from the container as stored (loadFg, :275-303), here fg_key
    assert np.array_equal(np.squeeze(fg), fgp[0] > 0.5)


def test_load_affinities_skips_finished_samples(tmp_path):
    """utilVoteInstances.py:146-149: a container that already holds the result is skipped."""
    pred = np.zeros((9, 1, 4, 4), np.float16)
    path = _store(tmp_path, 'd.zarr', {'volumes/pred_affs': pred,
                                       'vote_instances_05_tfgc': np.zeros((1, 4, 4), np.uint16)})
    assert uvi.loadAffinities(path, '_05_tfgc', patchshape=np.array([1, 3, 3]),
                              aff_key='volumes/pred_affs', patch_threshold=0.5) is None


def test_slab_partition_is_optimal_on_small_cases():
    """the dynamic programme against brute force over all contiguous splits."""
    import itertools
    from patchperpix_b200 import sharded
    rng = np.random.default_rng(4)
    for _ in range(20):
        n, world = int(rng.integers(3, 9)), int(rng.integers(2, 5))
        w = rng.integers(1, 20, n)
        _, slabs = sharded.slab_partition((8, n * 10, 8), (8, 10, 8), world, axis=1, weights=w)
        got = max(int(w[lo // 10:hi // 10].sum()) for lo, hi in slabs)
        best = min(max(int(w[a:b].sum()) for a, b in zip((0,) + cuts, cuts + (n,)))
                   for cuts in itertools.combinations_with_replacement(range(n + 1), world - 1))
        assert got == best, (w, slabs)
        assert slabs[0][0] == 0 and max(hi for lo, hi in slabs) == n * 10
