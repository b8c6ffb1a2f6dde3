"""Index conventions shared by the CUDA path, the oracle and the tests.

Compact consensus layout: one row per *gated* foreground voxel
(`pred[mid] > TH` and not overlap, the gate of fillConsensusArray.cu:53-60),
rows in raster order, and inside a row one slot per lexicographically positive
offset `o` in raster order of the (2ps-1)^3 offset cube.  The reference
allocates the full `[NSZ][NSY][NSX]` cube per voxel (consensus_array.py:99-105)
but only ever writes the positive half.
"""
import numpy as np


def patch_geometry(patchshape):
    ps = np.array([int(p) for p in patchshape], dtype=np.int64)
    P = int(ps.prod())
    r = ps // 2
    n = 2 * ps - 1
    N = int(n.prod())
    K = (N - 1) // 2
    return ps, P, r, n, N, K


def neighshape(patchshape):
    """vote_instances.py:249-253."""
    ps = np.array(patchshape).copy()
    if ps[0] > 1:
        ps *= 2
    else:
        ps[1:] *= 2
    return ps


def offsets_of_k(patchshape):
    """int32 [K,3]: the offset (oz,oy,ox) stored in slot k."""
    ps, P, r, n, N, K = patch_geometry(patchshape)
    lin = np.arange((N - 1) // 2 + 1, N, dtype=np.int64)
    oz = lin // (n[1] * n[2]) - (ps[0] - 1)
    oy = (lin // n[2]) % n[1] - (ps[1] - 1)
    ox = lin % n[2] - (ps[2] - 1)
    return np.stack([oz, oy, ox], axis=1).astype(np.int32)


def k_of_offset(patchshape, oz, oy, ox):
    ps, P, r, n, N, K = patch_geometry(patchshape)
    lin = ((np.asarray(oz) + ps[0] - 1) * n[1] + (np.asarray(oy) + ps[1] - 1)) \
        * n[2] + (np.asarray(ox) + ps[2] - 1)
    return lin - (N - 1) // 2 - 1


def dense_to_compact(dense, gate, patchshape):
    """reference layout [NSZ][NSY][NSX][Z][Y][X] -> compact [F_b][K].

    The reference index of offset o is o_d + ps_d - 1 (fillConsensusArray.cu:88-90)."""
    ps, P, r, n, N, K = patch_geometry(patchshape)
    offs = offsets_of_k(patchshape)
    rows = np.flatnonzero(gate.reshape(-1))
    d = dense.reshape(dense.shape[0], dense.shape[1], dense.shape[2], -1)
    out = np.empty((len(rows), K), dense.dtype)
    for k, (oz, oy, ox) in enumerate(offs):
        out[:, k] = d[oz + ps[0] - 1, oy + ps[1] - 1, ox + ps[2] - 1][rows]
    return out


def compact_to_dense(comp, gate, patchshape):
    ps, P, r, n, N, K = patch_geometry(patchshape)
    ns = neighshape(patchshape)
    Z, Y, X = gate.shape
    dense = np.zeros((ns[0], ns[1], ns[2], Z * Y * X), comp.dtype)
    rows = np.flatnonzero(gate.reshape(-1))
    offs = offsets_of_k(patchshape)
    for k, (oz, oy, ox) in enumerate(offs):
        dense[oz + ps[0] - 1, oy + ps[1] - 1, ox + ps[2] - 1][rows] = comp[:, k]
    return dense.reshape(ns[0], ns[1], ns[2], Z, Y, X)
