"""Seeded synthetic patch predictions for the instance-assembly hot path.

The reference ships no predictions (the bundled flylight crop
only holds raw + GT), so every test, golden vector and bench input is made
here: draw a label volume, derive the ideal shape patch of every voxel
(``pred[po][c] = [label(c+off(po)) == label(c) and label(c) > 0]``), squash it
to 0.05/0.95, add hash noise, round through float16 (the dtype the reference's
predict/decode stages store, experiments/flylight/setups/setup01/decode.py:102-109)
and widen to float32 (what ``do_block`` receives, stitch_patch_graph.py:640-646).

The noise is an integer hash of (element index, seed) so that the numpy/CPU
and the torch/CUDA evaluation give bit-identical arrays.
"""
import numpy as np

_M1 = np.int64(-7046029254386353131)   # 0x9E3779B97F4A7C15 as int64
_M2 = np.int64(-4658895280553007687)   # 0xBF58476D1CE4E5B9
_M3 = np.int64(-7723592293110705685)   # 0x94D049BB133111EB


def _hash01_np(idx, seed):
    """splitmix64-style hash of int64 indices -> float32 uniform in [0,1)."""
    with np.errstate(over='ignore'):
        z = (idx + np.int64(seed)) * _M1
        z = (z ^ ((z >> np.int64(30)) & np.int64(0x3FFFFFFFF))) * _M2
        z = (z ^ ((z >> np.int64(27)) & np.int64(0x1FFFFFFFFF))) * _M3
        z = z ^ ((z >> np.int64(31)) & np.int64(0x1FFFFFFFF))
    return ((z >> np.int64(40)) & np.int64(0xFFFFFF)).astype(np.float32) \
        * np.float32(1.0 / 16777216.0)


def _hash01_torch(idx, seed):
    import torch
    z = (idx + seed) * int(_M1)
    z = (z ^ ((z >> 30) & 0x3FFFFFFFF)) * int(_M2)
    z = (z ^ ((z >> 27) & 0x1FFFFFFFFF)) * int(_M3)
    z = z ^ ((z >> 31) & 0x1FFFFFFFF)
    return ((z >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)


# ----------------------------------------------------------------------------
# label volumes
# ----------------------------------------------------------------------------
def _draw_capsule(labels, numinst, p0, p1, radius, value, origin=None):
    """paint all voxels within `radius` of segment p0-p1 (float zyx coords).
    `origin`: global coordinates of labels[0,0,0] (labels is a window of the volume)."""
    shape = labels.shape
    org = np.zeros(3, int) if origin is None else np.asarray(origin, int)
    lo = np.floor(np.minimum(p0, p1) - radius).astype(int)
    hi = np.ceil(np.maximum(p0, p1) + radius).astype(int) + 1
    lo = np.maximum(lo, org)
    hi = np.minimum(hi, org + np.asarray(shape))
    if np.any(hi <= lo):
        return None
    zz, yy, xx = np.meshgrid(*[np.arange(lo[i], hi[i]) for i in range(3)],
                             indexing='ij')
    lo = lo - org
    hi = hi - org
    pts = np.stack([zz, yy, xx], axis=-1).astype(np.float32)
    d = (p1 - p0).astype(np.float32)
    dd = float(np.dot(d, d))
    if dd < 1e-12:
        t = np.zeros(pts.shape[:-1], np.float32)
    else:
        t = np.clip(((pts - p0) @ d) / dd, 0.0, 1.0)
    closest = p0 + t[..., None] * d
    dist2 = np.sum((pts - closest) ** 2, axis=-1)
    m = dist2 <= radius * radius
    return (slice(lo[0], hi[0]), slice(lo[1], hi[1]), slice(lo[2], hi[2])), m


def _paint_polyline(labels, numinst, pts, radius, value, origin=None):
    """union of capsules = one instance (counts once per voxel in numinst)."""
    for a, b in zip(pts[:-1], pts[1:]):
        res = _draw_capsule(labels, numinst, np.asarray(a, np.float32),
                            np.asarray(b, np.float32), radius, value, origin)
        if res is None:
            continue
        sl, m = res
        new = m & (labels[sl] != value)
        numinst[sl][new] += 1
        labels[sl][m] = value


def worms_2d(shape_yx=(520, 696), n_worms=60, seed=2,
             width=(8, 12), length=(150, 250)):
    """BBBC010-style curved capsules; crossings allowed.

    Returns labels int32 [1,Y,X] (top-most instance wins) and numinst uint8."""
    rng = np.random.default_rng(seed)
    Y, X = shape_yx
    labels = np.zeros((1, Y, X), np.int32)
    numinst = np.zeros((1, Y, X), np.uint8)
    for i in range(n_worms):
        L = rng.uniform(*length)
        w = rng.uniform(*width)
        nseg = 12
        ang = rng.uniform(0, 2 * np.pi)
        curv = rng.uniform(-0.25, 0.25)
        p = np.array([0.0, rng.uniform(0, Y), rng.uniform(0, X)])
        pts = [p.copy()]
        for s in range(nseg):
            ang += curv + rng.normal(0, 0.08)
            p = p + (L / nseg) * np.array([0.0, np.sin(ang), np.cos(ang)])
            pts.append(p.copy())
        _paint_polyline(labels, numinst, pts, w / 2.0, i + 1)
    return labels, numinst


def blobs_3d(shape=(128, 512, 512), n=2500, seed=3,
             rad_xy=(6, 14), rad_z=(2, 4)):
    """nuclei-style ellipsoids, touching allowed."""
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    labels = np.zeros(shape, np.int32)
    numinst = np.zeros(shape, np.uint8)
    for i in range(n):
        c = np.array([rng.uniform(0, Z), rng.uniform(0, Y), rng.uniform(0, X)])
        r = np.array([rng.uniform(*rad_z), rng.uniform(*rad_xy),
                      rng.uniform(*rad_xy)])
        lo = np.maximum(np.floor(c - r).astype(int), 0)
        hi = np.minimum(np.ceil(c + r).astype(int) + 1, shape)
        if np.any(hi <= lo):
            continue
        zz, yy, xx = np.meshgrid(*[np.arange(lo[k], hi[k]) for k in range(3)],
                                 indexing='ij')
        m = (((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 +
             ((xx - c[2]) / r[2]) ** 2) <= 1.0
        sl = (slice(lo[0], hi[0]), slice(lo[1], hi[1]), slice(lo[2], hi[2]))
        numinst[sl][m] += 1
        labels[sl][m] = i + 1
    return labels, numinst


def neurites_3d(shape=(256, 1024, 1024), n=300, seed=4, radius=(2, 3),
                seg_len=24.0, n_seg=40, window=None):
    """FlyLight-style thin 3-D polylines.
    window = (axis, lo, hi): only that slab of the volume is allocated and drawn
    (same random sequence, so the slabs of different ranks fit together)."""
    rng = np.random.default_rng(seed)
    shape = tuple(int(s) for s in shape)
    origin = None
    wshape = shape
    if window is not None:
        ax, wlo, whi = window
        origin = np.zeros(3, int)
        origin[ax] = wlo
        wshape = list(shape)
        wshape[ax] = whi - wlo
        wshape = tuple(wshape)
    labels = np.zeros(wshape, np.int32)
    numinst = np.zeros(wshape, np.uint8)
    for i in range(n):
        p = np.array([rng.uniform(0, shape[0]), rng.uniform(0, shape[1]),
                      rng.uniform(0, shape[2])])
        d = rng.normal(size=3)
        d[0] *= 0.4
        d /= np.linalg.norm(d)
        r = rng.uniform(*radius)
        pts = [p.copy()]
        for s in range(n_seg):
            d = d + 0.35 * rng.normal(size=3)
            d[0] *= 0.7
            d /= np.linalg.norm(d)
            p = p + seg_len * d
            pts.append(p.copy())
            if np.any(p < -seg_len) or np.any(p > np.array(shape) + seg_len):
                break
        _paint_polyline(labels, numinst, pts, r, i + 1, origin)
    return labels, numinst


def neurite_rows(shape, patchshape, axis=0, lo=0, hi=None, seed=4, noise=0.04, device='cuda',
                 chunk_rows=1 << 17, box=None, volume=None, **kw):
    """the compact row form of make_case('neurites', ...) for the slab
    lo <= coord[axis] < hi: (coords i32 [G,3] global, patches f16 [G,P], numinst u8 [G])
    as torch tensors on `device`, rows in raster order.  Values are bit-identical to
    patches_from_labels on the whole volume at the stored voxels (labels > 0); every
    other voxel of a ppp+dec prediction is zero (decode.py:39-65)."""
    import torch
    ps = [int(p) for p in patchshape]
    r = [p // 2 for p in ps]
    shape = tuple(int(s) for s in shape)
    hi = shape[axis] if hi is None else hi
    wlo, whi = max(lo - r[axis], 0), min(hi + r[axis], shape[axis])
    if volume is not None:           # (labels, numinst) of the WHOLE volume, already drawn
        wsl = [slice(None)] * 3
        wsl[axis] = slice(wlo, whi)
        labels, numinst = volume[0][tuple(wsl)], volume[1][tuple(wsl)]
    else:
        labels, numinst = neurites_3d(shape, seed=seed, window=(axis, wlo, whi), **kw)
    dev = torch.device(device)
    lab_t = torch.as_tensor(labels, device=dev)
    ni_t = torch.as_tensor(numinst, device=dev)
    # -1 outside the VOLUME; inside the window margin the real labels are present
    pad = [r[2], r[2], r[1], r[1], r[0], r[0]]
    pad_lo = [r[0], r[1], r[2]]
    pad[2 * (2 - axis)] = r[axis] - (lo - wlo)           # window margin already there
    pad[2 * (2 - axis) + 1] = r[axis] - (whi - hi)
    pad_lo[axis] = r[axis] - (lo - wlo)
    lab = torch.nn.functional.pad(lab_t, pad, value=-1)
    own = [slice(None)] * 3
    own[axis] = slice(lo - wlo, hi - wlo)
    c = torch.nonzero(lab_t[tuple(own)] > 0)             # raster order, slab-local
    if box is not None:                                  # only the rows inside a 3-D box
        b0 = torch.as_tensor(np.asarray(box[0]), device=dev).clone()
        b1 = torch.as_tensor(np.asarray(box[1]), device=dev).clone()
        b0[axis] -= lo
        b1[axis] -= lo
        c = c[((c >= b0) & (c < b1)).all(dim=1)]
    G = int(c.shape[0])
    P = ps[0] * ps[1] * ps[2]
    V = shape[0] * shape[1] * shape[2]
    gc = c.clone()
    gc[:, axis] += lo
    vglob = (gc[:, 0] * shape[1] + gc[:, 1]) * shape[2] + gc[:, 2]
    wc = c.clone()
    wc[:, axis] += lo - wlo
    lab_c = lab_t[wc[:, 0], wc[:, 1], wc[:, 2]]
    ni_c = ni_t[wc[:, 0], wc[:, 1], wc[:, 2]]
    offs = torch.tensor([(dz, dy, dx) for dz in range(ps[0]) for dy in range(ps[1])
                         for dx in range(ps[2])], device=dev)          # [P,3], padded coords
    po = torch.arange(P, device=dev, dtype=torch.int64)
    patches = torch.empty((G, P), dtype=torch.float16, device=dev)
    # position of centre c in the padded window: wc + pad_lo - r  (offset adds 0..ps-1)
    basec = wc + torch.tensor([pad_lo[i] - r[i] for i in range(3)], device=dev)
    for s0 in range(0, G, chunk_rows):
        s1 = min(s0 + chunk_rows, G)
        q = basec[s0:s1, None, :] + offs[None, :, :]
        nb = lab[q[..., 0], q[..., 1], q[..., 2]]
        ideal = (nb == lab_c[s0:s1, None]).to(torch.float32)
        idx = po[None, :] * V + vglob[s0:s1, None]
        u = _hash01_torch(idx, seed)
        v = ideal * 0.9 + 0.05
        v = v + (u * 2.0 - 1.0) * noise
        patches[s0:s1] = v.to(torch.float16)
    return gc.to(torch.int32), patches, ni_c.to(torch.uint8)


def discs_2d(shape_yx=(48, 48), centers=((16, 16), (30, 32)), radius=8):
    """the two-disc toy case of the survey probes."""
    Y, X = shape_yx
    labels = np.zeros((1, Y, X), np.int32)
    numinst = np.zeros((1, Y, X), np.uint8)
    yy, xx = np.mgrid[0:Y, 0:X]
    for i, (cy, cx) in enumerate(centers):
        m = (yy - cy) ** 2 + (xx - cx) ** 2 <= radius * radius
        numinst[0][m] += 1
        labels[0][m] = i + 1
    return labels, numinst


# ----------------------------------------------------------------------------
# labels -> patch predictions
# ----------------------------------------------------------------------------
def patches_from_labels(labels, patchshape, seed=0, noise=0.04, hard_frac=0.0,
                        device=None, out_dtype=None):
    """pred[po][c] = ideal*0.9 + 0.05 + U(-noise, noise), f16-rounded, f32.

    `hard_frac` > 0 replaces that fraction of entries by U(0,1) values so that
    thresholds, the "neither fg nor bg" band and ties get exercised.
    With `device` set the arithmetic runs in torch on that device and a torch
    tensor is returned; values are bit-identical to the numpy path.
    """
    ps = [int(p) for p in patchshape]
    Z, Y, X = labels.shape
    P = ps[0] * ps[1] * ps[2]
    r = [p // 2 for p in ps]
    V = Z * Y * X
    if device is None:
        lab = np.pad(labels, [(r[0], r[0]), (r[1], r[1]), (r[2], r[2])],
                     mode='constant', constant_values=-1)
        pred = np.empty((P, Z, Y, X), np.float32)
        base_idx = np.arange(V, dtype=np.int64).reshape(Z, Y, X)
        fgc = labels > 0
        po = 0
        for dz in range(ps[0]):
            for dy in range(ps[1]):
                for dx in range(ps[2]):
                    nb = lab[dz:dz + Z, dy:dy + Y, dx:dx + X]
                    ideal = ((nb == labels) & fgc).astype(np.float32)
                    idx = base_idx + np.int64(po) * np.int64(V)
                    u = _hash01_np(idx, seed)
                    v = ideal * np.float32(0.9) + np.float32(0.05)
                    v = v + (u * np.float32(2.0) - np.float32(1.0)) * np.float32(noise)
                    if hard_frac > 0:
                        u2 = _hash01_np(idx, seed + 7919)
                        u3 = _hash01_np(idx, seed + 104729)
                        v = np.where(u2 < np.float32(hard_frac), u3, v)
                    pred[po] = v.astype(np.float16).astype(np.float32)
                    po += 1
        return pred
    import torch
    dev = torch.device(device)
    lab_t = torch.as_tensor(labels, device=dev)
    lab = torch.nn.functional.pad(lab_t, (r[2], r[2], r[1], r[1], r[0], r[0]),
                                  value=-1)
    odt = out_dtype or torch.float32
    pred = torch.empty((P, Z, Y, X), dtype=odt, device=dev)
    base_idx = torch.arange(V, dtype=torch.int64, device=dev).reshape(Z, Y, X)
    fgc = lab_t > 0
    po = 0
    for dz in range(ps[0]):
        for dy in range(ps[1]):
            for dx in range(ps[2]):
                nb = lab[dz:dz + Z, dy:dy + Y, dx:dx + X]
                ideal = ((nb == lab_t) & fgc).to(torch.float32)
                idx = base_idx + po * V
                u = _hash01_torch(idx, seed)
                v = ideal * 0.9 + 0.05
                v = v + (u * 2.0 - 1.0) * noise
                if hard_frac > 0:
                    u2 = _hash01_torch(idx, seed + 7919)
                    u3 = _hash01_torch(idx, seed + 104729)
                    v = torch.where(u2 < hard_frac, u3, v)
                pred[po] = v.to(torch.float16).to(odt)
                po += 1
    return pred


def crop_case(labels, patchshape=(7, 7, 7), seed=31, hard_frac=0.1):
    """predictions around a given label volume (BASELINE configs[0]: the ground truth of
    the bundled flylight crop), nothing within patchshape//2 of the border."""
    ps = np.array(patchshape)
    pred = patches_from_labels(np.asarray(labels).astype(np.int32), ps, seed=seed,
                               hard_frac=hard_frac)
    r = ps // 2
    inner = np.zeros(labels.shape, bool)
    inner[r[0]:labels.shape[0] - r[0], r[1]:labels.shape[1] - r[1],
          r[2]:labels.shape[2] - r[2]] = True
    pred[:, ~inner] = 0
    return pred


def make_case(kind, patchshape, seed=0, hard_frac=0.0, noise=0.04, shape=None,
              device=None, **kw):
    """(pred f32 [P,Z,Y,X], numinst u8 [Z,Y,X], labels i32 [Z,Y,X])."""
    if kind == 'discs':
        labels, numinst = discs_2d(shape or (48, 48), **kw)
    elif kind == 'worms':
        labels, numinst = worms_2d(shape or (520, 696), seed=seed, **kw)
    elif kind == 'blobs':
        labels, numinst = blobs_3d(shape or (128, 512, 512), seed=seed, **kw)
    elif kind == 'neurites':
        labels, numinst = neurites_3d(shape or (256, 1024, 1024), seed=seed, **kw)
    else:
        raise ValueError(kind)
    pred = patches_from_labels(labels, patchshape, seed=seed, noise=noise,
                               hard_frac=hard_frac, device=device)
    return pred, numinst, labels
