"""TEST INFRASTRUCTURE ONLY — numpy/python restatement of the reference's HOST
side of instance assembly (the serial steps between the kernels).

Each function cites the reference file:line it follows.  Written for clarity,
not speed; pinned against tests/golden/*.npz (ranked order, cover, thinning,
pairs, labels recorded from the unmodified reference).  Only tests/,
__graft_entry__.smoke() and bench.py's baseline legs may import this.
"""
import networkx as nx
import numpy as np
import scipy.spatial


def interior_patches(foreground, rad):
    """vote_instances.py:276-287: fg coords in raster order, borders dropped."""
    allp = np.transpose(np.where(foreground))
    shp = np.array(foreground.shape)
    keep = np.all(allp >= rad, axis=1) & np.all(allp < shp - rad, axis=1)
    return allp[keep]


def rank_by_score(all_patches, score):
    """ranked_patches.py:21-30: stable sort, descending."""
    s = score[tuple(all_patches.T)]
    items = [(all_patches[i], s[i]) for i in range(len(all_patches))]
    items = sorted(items, key=lambda x: x[1], reverse=True)
    return items


def _window(idx, rad):
    return tuple(slice(int(idx[i] - rad[i]), int(idx[i] + rad[i] + 1))
                 for i in range(3))


def foreground_cover(overlap_mask, mask_to_cover, patchshape, ranked, rad,
                     pred, fc_threshold, sparse=True, score_threshold=None):
    """foreground_cover.py:15-180 (without the optional neighbourhood passes).

    `ranked` = list of (coord, score).  Returns the selected sub-list."""
    mask = mask_to_cover.copy()
    radslice = tuple(slice(int(rad[i]), mask.shape[i] - int(rad[i]))
                     for i in range(3))
    selected = np.zeros(len(ranked), bool)
    mid = int(np.prod(patchshape) / 2)
    pix_ths = [0] if sparse else [t for t in [500, 100, 50, 10, 0] if t < mid]
    for pix_th in pix_ths:
        rpidx = 0                      # by value in the reference (:31,43)
        while np.max(mask[radslice]) > 0 and rpidx < len(ranked):
            rpidx += 1
            r = rpidx - 1
            if selected[r]:
                continue
            if isinstance(score_threshold, float) and \
                    ranked[r][1] < score_threshold:
                break
            idx = ranked[r][0]
            if overlap_mask[tuple(idx)] > 0:
                continue
            patch = pred[(slice(None),) + tuple(idx)].reshape(patchshape)
            sl = _window(idx, rad)
            hit = patch > fc_threshold
            if np.count_nonzero(mask[sl][hit]) > pix_th:
                selected[r] = True
                mask[sl][hit] = 0
        if np.sum(mask[radslice]) < 1:
            break
    return [rp for i, rp in enumerate(ranked) if selected[i]]


def _fg_set(idx, pred, mask, patchshape, rad, th):
    """get_patch_sets.py:32-54 (sample == 1)."""
    start = idx - rad
    stop = idx + rad + 1
    if not (np.all(start >= 0) and np.all(stop <= mask.shape)):
        return set()
    patch = pred[(slice(None),) + tuple(idx)].reshape(patchshape)
    sl = _window(idx, rad)
    pf = start + np.argwhere(np.logical_and(patch > th, mask[sl]))
    return set(map(tuple, pf))


def thin_cover(mask_to_cover, selected_list, patchshape, rad, pred,
               fc_threshold):
    """foreground_cover.py:183-256 (thin_cover_use_kd False).

    Terminates when the best remaining set is empty (the reference would
    spin forever there)."""
    mask = mask_to_cover.copy()
    radslice = tuple(slice(int(rad[i]), mask.shape[i] - int(rad[i]))
                     for i in range(3))
    sel = np.zeros(len(selected_list), bool)
    sets = [_fg_set(rp[0], pred, mask_to_cover, patchshape, rad, fc_threshold)
            for rp in selected_list]
    while np.max(mask[radslice]) > 0:
        best = int(np.argmax([len(s) for s in sets]))
        if len(sets[best]) == 0 and sel[best]:
            break
        sel[best] = True
        best_fg = _fg_set(selected_list[best][0], pred, mask, patchshape, rad,
                          fc_threshold)
        if len(best_fg) == 0:
            break
        mask[tuple(zip(*list(best_fg)))] = 0
        sets = [s - best_fg for s in sets]
    return [rp for i, rp in enumerate(selected_list) if sel[i]]


def patch_pairs(selected_list, patchshape, include_single=True, max_ps_dist=2):
    """aff_patch_graph.py:43-110.  Returns u32 [n,6] in the reference's order
    (python-set iteration order of cKDTree.query_pairs) or None."""
    sel = sorted(selected_list, key=lambda p: p[0][2])
    n = len(sel)
    pts = np.zeros((n, 3), np.uint32)
    for i, p in enumerate(sel):
        pts[i] = p[0]
    pairs = set()
    if n > 0:
        tree = scipy.spatial.cKDTree(pts, leafsize=4)
        pairs = tree.query_pairs(2 * np.sum(patchshape), p=1)
    for p in list(pairs):
        if np.any(np.abs(pts[p[0]].astype(np.float32) -
                         pts[p[1]].astype(np.float32)) >
                  max_ps_dist * np.asarray(patchshape)):
            pairs.remove(p)
    total = len(pairs) + (n if include_single else 0)
    if total == 0:
        return None
    arr = np.zeros((total, 6), np.uint32)
    for i, p in enumerate(pairs):
        arr[i, :3] = pts[p[0]]
        arr[i, 3:] = pts[p[1]]
    if include_single:
        for i, p in enumerate(sel):
            arr[len(pairs) + i, :3] = p[0]
            arr[len(pairs) + i, 3:] = p[0]
    return arr


def mutex_watershed(g):
    """graph_mws.py:7-85 restated: greedy pass over the edges by decreasing
    |aff|; an attractive edge joins two clusters unless a repulsive edge seen
    earlier connects them.  Returns the reference's list of components (some
    empty: ids of merged-away components stay in the numbering)."""
    ids = {n: i for i, n in enumerate(g.nodes())}
    names = list(g.nodes())
    edges = []
    for u, v, a in g.edges.data('aff'):
        edges.append((ids[u], ids[v], a, True) if a > 0 else (ids[u], ids[v], -a, False))
    edges.sort(key=lambda e: e[2], reverse=True)          # stable
    comp = {i: 0 for i in ids.values()}                   # node -> component id, 0 = none
    member = {0: set(ids.values())}                       # component id -> nodes
    repulsive = []

    def blocked(test):
        return any(test(e, f) or test(f, e) for e, f in repulsive)

    for u, v, _, attractive in edges:
        if not attractive:
            repulsive.append((u, v))
            continue
        cu, cv = comp[u], comp[v]
        if cu == 0 and cv == 0:
            new = max(comp.values()) + 1
            member[new] = {u, v}
            member[0] -= {u, v}
            comp[u] = comp[v] = new
        elif cu == 0 or cv == 0:
            c = max(cu, cv)
            lone = u if cu == 0 else v
            if not blocked(lambda e, f: comp[e] == c and f == lone):
                member[c] |= {u, v}
                member[0] -= {u, v}
                comp[u] = comp[v] = c
        elif cu != cv:
            if not blocked(lambda e, f: comp[e] == cu and comp[f] == cv):
                keep, gone = min(cu, cv), max(cu, cv)
                member[keep] = member[cu] | member[cv]
                for e in member[gone]:
                    comp[e] = keep
                member[gone] = set()
    return [[names[i] for i in member[c]] for c in member if c > 0]


def pack_no_overlap(channels, shape, dtype, min_size=2000):
    """graph_to_labeling.py:96-113 (no_overlap_per_channel): instances larger than
    `min_size` voxels go to the first channel where their voxels are all free, else
    to a new channel; smaller ones are written into channel 0 (overwriting)."""
    out = []
    for k, cur in enumerate(channels):
        if not out:
            out.append(cur.copy())
            continue
        m = cur > 0
        if np.sum(m) > min_size:
            for ch in out:
                if np.all(ch[m] == 0):
                    ch[m] = k + 1
                    break
            else:
                out.append(cur.copy())
        else:
            out[0][m] = k + 1
    return np.stack(out, axis=0) if out else np.zeros((0,) + tuple(shape), dtype)


def label_instances(pairs, aff, pred, patchshape, rad, shape, patch_threshold,
                    dtype=np.uint16, mws=False, per_channel=False, no_overlap=False):
    """aff_patch_graph.py:31-40 + graph_to_labeling.py:44-84; per_channel =
    one_instance_per_channel (:57-95, 114-115): every component painted into its
    own volume, stacked."""
    g = nx.Graph()
    for i, a in enumerate(aff):
        if a != 0:
            g.add_edge(tuple(int(v) for v in pairs[i, :3]),
                       tuple(int(v) for v in pairs[i, 3:6]), aff=a)
    pos = nx.Graph()
    for e0, e1, a in g.edges.data('aff'):
        if a > 0:
            pos.add_edge(e0, e1, weight=a)
    inst = np.zeros(shape, dtype)
    comps, channels = [], []
    ccs = mutex_watershed(g) if mws else nx.connected_components(pos)
    for k, cc in enumerate(ccs):
        comps.append(sorted(cc))
        target = np.zeros(shape, dtype) if (per_channel or no_overlap) else inst
        for idx in cc:
            idx = np.array(idx)
            patch = pred[(slice(None),) + tuple(idx)].reshape(patchshape)
            sl = _window(idx, rad)
            target[sl][patch > patch_threshold] = k + 1
        if per_channel or no_overlap:
            channels.append(target)
    if no_overlap:
        inst = pack_no_overlap(channels, shape, dtype)
    elif per_channel:
        inst = np.stack(channels, axis=0) if channels else np.zeros((0,) + tuple(shape), dtype)
    return inst, comps


def assemble(pred, foreground, numinst, patchshape, kw, kern):
    """to_instance_seg (vote_instances.py:150-452) with `kern` = an
    oracle.cpu_oracle.Oracle supplying the kernel steps."""
    ps = np.array(patchshape)
    rad = ps // 2
    mask = foreground.copy()
    overlap_mask = 1 * (numinst > 1)
    mask[overlap_mask > 0] = 0
    out = {}
    allp = interior_patches(foreground, rad)
    kern.consensus()
    cons = kern.norm()
    score = kern.rank()
    ranked = rank_by_score(allp, score)
    out['score'] = score
    out['ranked'] = np.array([p[0] for p in ranked], np.int32).reshape(-1, 3)
    fc = np.float32(kw['fc_threshold'])
    sel = foreground_cover(overlap_mask, mask, ps, ranked, rad, pred, fc,
                           sparse=kw['select_patches_for_sparse_data'],
                           score_threshold=kw.get('score_threshold'))
    out['cover'] = np.array([p[0] for p in sel], np.int32).reshape(-1, 3)
    if not kw.get('skipThinCover', False) and len(sel) > 0:
        sel = thin_cover(mask, sel, ps, rad, pred, fc)
    out['thin'] = np.array([p[0] for p in sel], np.int32).reshape(-1, 3)
    pairs = patch_pairs(sel, ps, kw['includeSinglePatchCCS'],
                        kw.get('max_total_patch_distance_in_ps_multiples', 2))
    out['pairs'] = pairs
    if pairs is None:
        out['instances'] = np.zeros(foreground.shape, np.uint16)
        return out
    aff = kern.patch_graph(pairs)
    out['aff'] = aff
    inst, comps = label_instances(pairs, aff, pred, ps, rad, foreground.shape,
                                  np.float32(kw['patch_threshold']),
                                  mws=kw.get('mws', False),
                                  per_channel=kw.get('one_instance_per_channel', False),
                                  no_overlap=kw.get('no_overlap_per_channel', False))
    out['instances'] = inst
    out['components'] = comps
    return out


def oracle_block_fn(block, foreground, mask, numinst, **kw):
    """do_block(..., return_intermediates=True) (vote_instances.py:455-476) on
    the CPU oracle; honours selected_patches / selected_patch_pairs the way
    stitch_vote_instances uses them (stitch_patch_graph.py:323-328)."""
    from . import cpu_oracle
    ps = np.array(kw['patchshape'])
    rad = ps // 2
    pred = np.ascontiguousarray(block, np.float32)
    overlap_mask = 1 * (numinst > 1)
    mask = mask.copy()
    mask[overlap_mask > 0] = 0
    radslice = tuple(slice(int(rad[i]), mask.shape[i] - int(rad[i])) for i in range(3))
    if np.count_nonzero(mask[radslice]) == 0:
        return None, None
    allp = interior_patches(foreground, rad)
    if len(allp) == 0:
        return None, None
    O = cpu_oracle.Oracle(pred, numinst > 1, ps, cpu_oracle.variant_from_kwargs(kw))
    O.consensus()
    O.norm()
    if kw.get('selected_patch_pairs') is not None:
        pairs = np.array(kw['selected_patch_pairs'], dtype=np.uint32).reshape(-1, 6)
    else:
        score = O.rank()
        ranked = rank_by_score(allp, score)
        fc = np.float32(kw['fc_threshold'])
        sel = foreground_cover(overlap_mask, mask, ps, ranked, rad, pred, fc,
                               sparse=kw['select_patches_for_sparse_data'])
        if not kw.get('skipThinCover', False) and len(sel) > 0:
            sel = thin_cover(mask, sel, ps, rad, pred, fc)
        pairs = patch_pairs(sel, ps, kw['includeSinglePatchCCS'],
                            kw.get('max_total_patch_distance_in_ps_multiples', 2))
    if pairs is None or len(pairs) == 0:
        return None, None
    assert np.all(pairs.astype(np.int64) < np.tile(np.array(pred.shape[1:]), 2)), \
        "pair centre outside the block (the reference would read out of bounds)"
    return pairs, O.patch_graph(pairs)


def oracle_paint_fn(inputs, pairs, aff, rank=0, world=1, **kw):
    """global labelling (stitch_patch_graph.py:360-399) with networkx."""
    ps = np.array(kw['patchshape'])
    pred = np.asarray(inputs.pred).astype(np.float32)
    inst, _ = label_instances(pairs, aff, pred, ps, ps // 2, inputs.shape,
                              np.float32(kw['patch_threshold']), dtype=np.uint32,
                              mws=kw.get('mws', False))
    return inst
