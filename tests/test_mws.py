"""mutex-watershed labelling (kwargs mws=True, the flylight default) against
vectors recorded from the unmodified reference (tools/gen_golden.py mws):
graph_mws.mws on seeded random graphs and affGraphToInstances(mws=True) on the
recorded pairs / affinities of every golden case."""
import os

import numpy as np
import pytest

from tests import golden_util
from patchperpix_b200 import cuda_code as cc
from patchperpix_b200.assembly import mutex_watershed

MWS = dict(np.load(os.path.join(golden_util.GOLD, 'mws_cases.npz')))
N_RANDOM = int(MWS['n_random'])


def _graph(gi):
    k = 'rnd/%02d/' % gi
    return MWS[k + 'pairs'], MWS[k + 'aff'], MWS[k + 'nodes'], MWS[k + 'labels']


@pytest.mark.parametrize('gi', range(N_RANDOM))
def test_oracle_mws_random_graphs(gi):
    import networkx as nx
    from oracle import host_logic
    pairs, aff, nodes, labels = _graph(gi)
    g = nx.Graph()
    for i, a in enumerate(aff):
        if a != 0:
            g.add_edge(tuple(int(v) for v in pairs[i, :3]),
                       tuple(int(v) for v in pairs[i, 3:6]), aff=a)
    assert [list(n) for n in g.nodes()] == nodes.tolist()
    lab = {}
    for k, comp in enumerate(host_logic.mutex_watershed(g)):
        for n in comp:
            lab[n] = k + 1
    got = np.array([lab.get(tuple(int(v) for v in n), 0) for n in nodes], np.int32)
    assert np.array_equal(got, labels)


@pytest.mark.parametrize('gi', range(N_RANDOM))
def test_cabi_mws_random_graphs(gi):
    """ppp_mws_host is host code: it runs without a GPU."""
    pairs, aff, nodes, labels = _graph(gi)
    Z, Y, X = 4, 40, 40
    cfg = cc.make_cfg((Z, Y, X), (1, 3, 3), patch_threshold=0.5, fc_threshold=0.5)
    vox, lab, top = mutex_watershed(pairs, aff, cfg)
    want_vox = (nodes[:, 0].astype(np.int64) * Y + nodes[:, 1]) * X + nodes[:, 2]
    assert np.array_equal(vox, want_vox)          # node insertion order
    assert np.array_equal(lab, labels)            # same ids, gaps included
    assert top >= (labels.max() if len(labels) else 0)   # ids ever created


@pytest.mark.parametrize('seed,rep,big_space', [(1, 0.1, False), (2, 0.5, True), (3, 0.9, False)])
def test_cabi_mws_stress_against_oracle(seed, rep, big_space):
    """larger seeded graphs than the recorded ones -- repeated pairs in both orientations
    (the later affinity wins, the place in the neighbour list stays), self pairs, tied
    |aff|, many exclusions, sparse and compact voxel spaces -- against the oracle
    restatement (itself pinned on the reference's graphs above)."""
    import networkx as nx
    from oracle import host_logic
    rng = np.random.default_rng(seed)
    Z, Y, X = (40, 300, 300) if big_space else (2, 30, 30)
    m, n = 700, 6000
    pts = np.unique(np.stack([rng.integers(0, Z, m), rng.integers(0, Y, m),
                              rng.integers(0, X, m)], 1), axis=0)
    pts = pts[rng.permutation(len(pts))]
    i, j = rng.integers(0, len(pts), n), rng.integers(0, len(pts), n)
    j[:200] = i[:200]                                           # self pairs
    i[200:700], j[200:700] = j[5200:5700], i[5200:5700]         # repeats, other orientation
    i[700:900], j[700:900] = i[5700:5900], j[5700:5900]         # repeats, same orientation
    aff = np.round(rng.random(n) * 0.95 + 0.02, 2).astype(np.float32)   # many ties
    aff[rng.random(n) < rep] *= -1
    aff[:200] = np.abs(aff[:200])
    aff[rng.random(n) < 0.02] = 0
    pairs = np.concatenate([pts[i], pts[j]], 1).astype(np.uint32)
    g = nx.Graph()
    for k, a in enumerate(aff):
        if a != 0:
            g.add_edge(tuple(int(v) for v in pairs[k, :3]), tuple(int(v) for v in pairs[k, 3:]),
                       aff=a)
    lab = {}
    for k, comp in enumerate(host_logic.mutex_watershed(g)):
        for nd in comp:
            lab[nd] = k + 1
    nodes = np.array(list(g.nodes()), np.int64)
    want = np.array([lab.get(tuple(int(v) for v in nd), 0) for nd in nodes], np.int32)
    cfg = cc.make_cfg((Z, Y, X), (1, 3, 3), patch_threshold=0.5, fc_threshold=0.5)
    vox, got, top = mutex_watershed(pairs, aff, cfg)
    assert np.array_equal(vox, (nodes[:, 0] * Y + nodes[:, 1]) * X + nodes[:, 2])
    assert np.array_equal(got, want)
    assert len(np.unique(got)) > 3


def test_cabi_mws_empty():
    cfg = cc.make_cfg((1, 8, 8), (1, 3, 3), patch_threshold=0.5, fc_threshold=0.5)
    vox, lab, top = mutex_watershed(np.zeros((0, 6), np.uint32), np.zeros(0, np.float32), cfg)
    assert len(vox) == 0 and len(lab) == 0 and top == 0
    # only zero affinities: no edge enters the graph (aff_patch_graph.py:36)
    vox, lab, top = mutex_watershed(np.ones((3, 6), np.uint32), np.zeros(3, np.float32), cfg)
    assert len(vox) == 0 and top == 0


@pytest.mark.parametrize('name', golden_util.NAMES)
def test_oracle_mws_instances(name):
    from oracle import host_logic
    g, kw, ps, pred = golden_util.load(name)
    inst, _ = host_logic.label_instances(
        g['pairs'], g['aff'], pred, ps, ps // 2, pred.shape[1:],
        np.float32(kw['patch_threshold']), mws=True)
    assert np.array_equal(inst, MWS['inst/' + name])


@pytest.mark.gpu
@pytest.mark.parametrize('name', golden_util.NAMES)
def test_gpu_mws_instances(name):
    """to_instance_seg(mws=True) end to end on the device."""
    from patchperpix_b200 import vote_instances as vi
    g, kw, ps, pred = golden_util.load(name)
    kw = dict(kw, mws=True)
    mid = int(np.prod(ps)) // 2
    fg = pred[mid] > kw['patch_threshold']
    inst, _ = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), g['numinst'].copy(),
                                 ps.copy(), **kw)
    assert np.array_equal(inst, MWS['inst/' + name])


# ---------------------------------------------------------------------------
# one_instance_per_channel (graph_to_labeling.py:57-95): one volume per component
# ---------------------------------------------------------------------------
OPC_NAMES = sorted(k.split('/', 1)[1] for k in MWS if k.startswith('opc_cc/'))


def _opc_expected(tag, name):
    shape = tuple(int(v) for v in MWS['opc_%s_shape/%s' % (tag, name)])
    bits = np.unpackbits(MWS['opc_%s/%s' % (tag, name)])[:int(np.prod(shape))].reshape(shape)
    vals = MWS['opc_%s_vals/%s' % (tag, name)].astype(np.uint16)
    return bits.astype(np.uint16) * vals.reshape((-1,) + (1,) * (len(shape) - 1))


@pytest.mark.parametrize('tag', ['cc', 'mws'])
@pytest.mark.parametrize('name', OPC_NAMES)
def test_oracle_one_instance_per_channel(name, tag):
    from oracle import host_logic
    g, kw, ps, pred = golden_util.load(name)
    stack, _ = host_logic.label_instances(
        g['pairs'], g['aff'], pred, ps, ps // 2, pred.shape[1:],
        np.float32(kw['patch_threshold']), mws=(tag == 'mws'), per_channel=True)
    assert np.array_equal(stack, _opc_expected(tag, name))


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['cc', 'mws'])
@pytest.mark.parametrize('name', OPC_NAMES)
def test_gpu_one_instance_per_channel(name, tag):
    from patchperpix_b200 import vote_instances as vi
    g, kw, ps, pred = golden_util.load(name)
    kw = dict(kw, mws=(tag == 'mws'), one_instance_per_channel=True)
    mid = int(np.prod(ps)) // 2
    fg = pred[mid] > kw['patch_threshold']
    stack, _ = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), g['numinst'].copy(),
                                  ps.copy(), **kw)
    want = _opc_expected(tag, name)
    assert stack.dtype == np.uint16 and stack.shape == want.shape
    assert np.array_equal(stack, want)


# ---------------------------------------------------------------------------
# no_overlap_per_channel (graph_to_labeling.py:96-113)
# ---------------------------------------------------------------------------
def _no_overlap_case():
    import hashlib
    import json
    sys_path_tools = os.path.join(os.path.dirname(golden_util.GOLD), '..', 'tools')
    g = dict(np.load(os.path.join(golden_util.GOLD, 'chan_nooverlap2d_ps9.npz')))
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        'gen_golden_case', os.path.join(os.path.abspath(sys_path_tools), 'no_overlap_case.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pred, numinst, _ = mod.no_overlap_case()
    assert hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest() == str(g['pred_sha1'])
    assert np.array_equal(numinst, g['numinst'])
    return g, json.loads(str(g['kwargs'])), pred, numinst


@pytest.mark.parametrize('tag', ['cc', 'mws'])
def test_oracle_no_overlap_per_channel(tag):
    from oracle import cpu_oracle, host_logic
    g, kw, pred, numinst = _no_overlap_case()
    ps = np.array([1, 9, 9])
    kw = dict(kw, mws=(tag == 'mws'))
    fg = pred[40] > np.float32(kw['patch_threshold'])
    O = cpu_oracle.Oracle(pred, numinst > 1, ps, cpu_oracle.variant_from_kwargs(kw))
    out = host_logic.assemble(pred, fg, numinst, ps, kw, O)
    assert out['instances'].shape[0] == 2
    assert np.array_equal(out['instances'], g['stack_' + tag])


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['cc', 'mws'])
def test_gpu_no_overlap_per_channel(tag):
    from patchperpix_b200 import vote_instances as vi
    g, kw, pred, numinst = _no_overlap_case()
    ps = np.array([1, 9, 9])
    kw = dict(kw, mws=(tag == 'mws'))
    fg = pred[40] > np.float32(kw['patch_threshold'])
    stack, _ = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), numinst.copy(), ps.copy(), **kw)
    assert stack.dtype == np.uint16
    assert np.array_equal(stack, g['stack_' + tag])


def test_channel_packing_matches_oracle_on_random_stacks():
    """vote_instances.pack_no_overlap_channels (torch ops, device-agnostic) against the
    oracle's restatement on random instance stacks with a small size threshold."""
    import torch
    from oracle import host_logic
    from patchperpix_b200.vote_instances import pack_no_overlap_channels
    rng = np.random.default_rng(8)
    for trial in range(20):
        n = int(rng.integers(1, 9))
        stack = np.zeros((n, 1, 24, 24), np.int32)
        for k in range(n):
            if rng.random() < 0.15:
                continue                                  # empty component (mws id gap)
            y, x = rng.integers(0, 16, 2)
            h, w = rng.integers(2, 9, 2)
            stack[k, 0, y:y + h, x:x + w] = k + 1
        want = host_logic.pack_no_overlap([c.copy() for c in stack], stack.shape[1:], np.int32,
                                          min_size=20)
        got = pack_no_overlap_channels(torch.from_numpy(stack), min_size=20).numpy()
        assert np.array_equal(got, want), trial
