"""CPU-side checks of the boundary: the C-ABI library loads and exports every
symbol include/ppp_b200.h declares (no compute calls), the config mapping
follows the reference's flag table, and the host pair enumeration matches the
oracle / golden order."""
import ctypes
import os
import re

import numpy as np
import pytest

from patchperpix_b200 import cuda_code, layout
from tests import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'ppp_b200.h')).read()
    declared = set(re.findall(r'\b(ppp_[a-z_0-9]+)\s*\(', hdr))
    assert declared == set(cuda_code.exported_symbols())
    lib = cuda_code.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ppp_version() >= 100
    assert ctypes.sizeof(cuda_code.PppCfg) == 80


def test_threshold_literals_behave_like_double_compares():
    # kernels compare float values against double literals
    for th in (0.5, 0.7, 0.8, 0.9, 0.1, 0.3):
        f = np.float32(cuda_code._f32_floor(th))
        c = np.float32(cuda_code._f32_ceil(th))
        vals = np.array([np.nextafter(np.float32(th), np.float32(-1)), np.float32(th),
                         np.nextafter(np.float32(th), np.float32(2))], np.float32)
        for v in vals:
            assert (float(v) > th) == bool(v > f)
            assert (float(v) < th) == bool(v < c)


def test_cfg_flag_mapping():
    kw = dict(patch_threshold=0.9, fc_threshold=0.5)
    c = cuda_code.make_cfg((1, 8, 8), (1, 3, 3), **kw)       # inv_th default True
    assert abs(c.bg_lt - (1.0 - 0.9)) < 1e-6 and c.prod_mode == 2 and c.norm_aff == 1
    c = cuda_code.make_cfg((1, 8, 8), (1, 3, 3), **dict(kw, patch_threshold=0.4))
    assert abs(c.bg_lt - 0.4) < 1e-6                         # falls back to < th
    c = cuda_code.make_cfg((1, 8, 8), (1, 3, 3), vi_bg_use_inv_th=False,
                           vi_bg_use_half_th=True, consensus_norm_prob_product=False,
                           rank_int_counter=True, patch_graph_norm_aff=False, **kw)
    assert abs(c.bg_lt - 0.45) < 1e-6 and c.prod_mode == 1
    assert c.rank_flags == 3 and c.graph_flags == 0
    with pytest.raises(RuntimeError):
        cuda_code.make_cfg((1, 8, 8), (1, 3, 3), vi_bg_use_inv_th=False, **kw)


def test_layout_roundtrip():
    ps = (3, 5, 5)
    _, P, r, n, N, K = layout.patch_geometry(ps)
    offs = layout.offsets_of_k(ps)
    assert len(offs) == K
    assert np.array_equal(layout.k_of_offset(ps, offs[:, 0], offs[:, 1], offs[:, 2]),
                          np.arange(K))
    rng = np.random.default_rng(0)
    gate = rng.random((4, 6, 7)) > 0.5
    comp = rng.random((int(gate.sum()), K)).astype(np.float32)
    dense = layout.compact_to_dense(comp, gate, ps)
    assert np.array_equal(layout.dense_to_compact(dense, gate, ps), comp)


@pytest.mark.parametrize('name', gu.NAMES)
def test_host_pair_enumeration_matches_reference_order(name):
    from patchperpix_b200.assembly import BlockAssembler
    g, kw, ps, pred = gu.load(name)
    asm = BlockAssembler.__new__(BlockAssembler)
    asm.kwargs = kw
    asm.ps = layout.patch_geometry(ps)[0]
    pairs = asm.patch_pairs(g['thin'].astype(np.int64))
    assert np.array_equal(pairs, g['pairs'])


def test_no_cpu_fallback():
    from patchperpix_b200 import vote_instances as vi
    with pytest.raises(NotImplementedError):
        vi.to_instance_seg(np.zeros((9, 1, 8, 8), np.float32), np.zeros((1, 8, 8), bool),
                           np.zeros((1, 8, 8), bool), np.zeros((1, 8, 8), np.uint8),
                           np.array([1, 3, 3]), cuda=False)


@pytest.mark.parametrize('n', [0, 1, 5, 6, 40, 1000, 30000, 120000])
def test_pyset_order_replays_cpython_sets(n):
    """ppp_pyset_order (host helper of the pair enumeration) against a real set,
    across every table size up to the x2 growth regime (> 50000 entries)."""
    from patchperpix_b200 import cuda_code as cc
    rng = np.random.default_rng(n)
    a = np.unique(rng.integers(0, max(4, 3 * n), (n, 2)).astype(np.int64), axis=0)
    rng.shuffle(a)
    a = np.ascontiguousarray(a)
    order = np.zeros(len(a), np.int64)
    cc.call('ppp_pyset_order', a.ctypes.data if len(a) else None, len(a),
            order.ctypes.data if len(a) else None)
    want = np.array(list(set(map(tuple, a.tolist()))), np.int64).reshape(-1, 2)
    assert np.array_equal(a[order], want)


def test_query_pairs_set_order_equals_python_set():
    import scipy.spatial
    from patchperpix_b200.assembly import query_pairs_set_order
    rng = np.random.default_rng(3)
    pts = rng.integers(0, 120, (700, 3)).astype(np.uint32)
    tree = scipy.spatial.cKDTree(pts, leafsize=4)
    for r in (4, 21, 42):
        want = np.array(list(tree.query_pairs(r, p=1)), np.int64).reshape(-1, 2)
        assert np.array_equal(query_pairs_set_order(tree, r), want)


def test_query_pairs_filtered_equals_reference_loop():
    """set order + per-axis distance filter in one library call == the reference's
    remove-from-the-set loop (aff_patch_graph.py:57-69)."""
    import scipy.spatial
    from patchperpix_b200.assembly import query_pairs_filtered
    rng = np.random.default_rng(5)
    ps = np.array([3, 7, 5])
    pts = rng.integers(0, 60, (900, 3)).astype(np.uint32)
    pts = pts[np.argsort(pts[:, 2], kind='stable')]
    tree = scipy.spatial.cKDTree(pts, leafsize=4)
    pairs = tree.query_pairs(2 * np.sum(ps), p=1)
    for p in list(pairs):
        if np.any(np.abs(pts[p[0]].astype(np.float32) - pts[p[1]].astype(np.float32)) > 2 * ps):
            pairs.remove(p)
    want = np.array(list(pairs), np.int64).reshape(-1, 2)
    got = query_pairs_filtered(tree, pts, 2 * np.sum(ps), np.asarray(2 * ps, np.float64))
    assert np.array_equal(got, want)
