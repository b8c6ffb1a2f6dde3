"""Output / post side of the blockwise stage on the device: the label-volume
operations stitch_patch_graph.main applies after stitching
(stitch_patch_graph.py:824-894) and the `only_bb` pre-crop (:745-764).

    remove_small_components, relabel    PatchPerPix/util/postprocess.py:24-52
    dilate_instances                    stitch_patch_graph.py:871-894
    clean_mask, foreground_bbox         stitch_patch_graph.py:46-57, 745-764

torch tensors in, torch tensors out (any device; the drivers pass CUDA tensors).
"""
import logging

import numpy as np

logger = logging.getLogger(__name__)


def remove_small_components(array, compsize=5):
    """labels that occur `compsize` times or fewer become 0 (postprocess.py:24-36)."""
    import torch
    labels, inv, counts = torch.unique(array, return_inverse=True, return_counts=True)
    lut = torch.where(counts <= compsize, torch.zeros_like(labels), labels)
    return lut[inv].reshape(array.shape)


def relabel(array, start=None):
    """consecutive labels in ascending order of the old ones, 0 stays
    (postprocess.py:39-52)."""
    import torch
    labels, inv = torch.unique(array, return_inverse=True)
    nz = labels != 0
    new = torch.cumsum(nz.to(torch.int64), 0) + ((start if start is not None else 1) - 1)
    lut = torch.where(nz, new, torch.zeros_like(new)).to(array.dtype)
    return lut[inv].reshape(array.shape)


def dilate_instances(instances):
    """stitch_patch_graph.py:873-881: for every label in ascending order, the voxels
    that CURRENTLY carry it are dilated by one step (6-neighbourhood,
    ndimage.binary_dilation's default) and everything under the dilated mask takes
    the label -- other labels included, so the order matters and is kept: one
    small box per label (a label never grows before its own turn, so its box is the
    box of its original voxels plus one)."""
    import torch
    inst = instances.clone()
    nzc = torch.nonzero(inst)
    if nzc.numel() == 0:
        return inst
    lab = inst[nzc[:, 0], nzc[:, 1], nzc[:, 2]].long()
    labels, inv = torch.unique(lab, return_inverse=True)
    n = int(labels.numel())
    big = int(max(inst.shape)) + 1
    lo = torch.full((n, 3), big, dtype=torch.int64, device=inst.device)
    hi = torch.full((n, 3), -1, dtype=torch.int64, device=inst.device)
    idx = inv[:, None].expand(-1, 3)
    lo.scatter_reduce_(0, idx, nzc, 'amin')
    hi.scatter_reduce_(0, idx, nzc, 'amax')
    lo = torch.clamp(lo - 1, min=0).cpu().numpy()
    hi = (hi + 2).cpu().numpy()
    shape = np.asarray(inst.shape)
    hi = np.minimum(hi, shape)
    for k, lbl in enumerate(labels.cpu().numpy().tolist()):
        sl = tuple(slice(int(a), int(b)) for a, b in zip(lo[k], hi[k]))
        box = inst[sl]
        m = box == lbl
        d = m.clone()
        d[1:] |= m[:-1]
        d[:-1] |= m[1:]
        d[:, 1:] |= m[:, :-1]
        d[:, :-1] |= m[:, 1:]
        d[:, :, 1:] |= m[:, :, :-1]
        d[:, :, :-1] |= m[:, :, 1:]
        box[d] = lbl
    return inst


def clean_mask(mask, structure, size):
    """connected components (scipy, `structure` connectivity) of `mask` with `size`
    voxels or fewer are dropped (stitch_patch_graph.py:46-57).  numpy in/out: runs
    once per volume on the foreground mask, before anything is on the device."""
    from scipy import ndimage
    labeled = ndimage.label(mask, structure)[0]
    counts = np.bincount(labeled.reshape(-1))
    small = counts <= size
    small[0] = True
    logger.info('removing %i of small components.', int(small.sum()))
    return ~small[labeled]


def foreground_bbox(mask, **kwargs):
    """bounding box of the (cleaned) foreground: (bb_offset, bb_shape) or None if
    the mask is empty (stitch_patch_graph.py:745-764).

    skeletonize_foreground: the reference shrinks the mask to its skeleton
    (skimage.morphology.skeletonize_3d) before taking the box.  skimage is an
    optional dependency here; without it the box of the un-skeletonised mask is
    used -- it CONTAINS the reference's box (a skeleton is a subset of its mask), the
    block grid may then start up to one neurite radius earlier (deviations;
    tests/test_boundary.py::test_bbox_without_skeleton_contains_reference_box)."""
    mask = np.squeeze(np.asarray(mask)) > 0
    if np.count_nonzero(mask) == 0:
        return None
    if kwargs.get('ignore_small_comps', 0) > 0:
        mask = clean_mask(mask, np.ones([3] * mask.ndim), kwargs.get('ignore_small_comps'))
    if kwargs.get('skeletonize_foreground'):
        try:
            from skimage.morphology import skeletonize_3d
            mask = skeletonize_3d(mask.astype(np.uint8)) > 0
        except ImportError:
            logger.warning("skeletonize_foreground: skimage is not installed, the bounding "
                           "box is taken from the un-skeletonised foreground")
    if np.count_nonzero(mask) == 0:
        return None
    nz = np.nonzero(mask)
    lo = np.array([int(a.min()) for a in nz])
    hi = np.array([int(a.max()) for a in nz])
    return lo, hi - lo + 1


def color(src, seed=0):
    """random colour per label, 0 stays black (util/postprocess.py:55-74; the reference
    draws unseeded random colours, here they are seeded)."""
    src = np.asarray(src)
    labels, inv = np.unique(src, return_inverse=True)
    lut = np.random.default_rng(seed).integers(0, 255, (len(labels), 3)).astype(np.uint8)
    lut[labels == 0] = 0
    return lut[inv.reshape(src.shape)]


def write_png(path, rgb):
    """8-bit RGB PNG with the standard library (the reference uses skimage.io.imsave,
    an optional package here)."""
    import struct
    import zlib
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    raw = b''.join(b'\x00' + rgb[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack('>I', len(data)) + tag + data
        return c + struct.pack('>I', zlib.crc32(tag + data) & 0xffffffff)
    with open(path, 'wb') as f:
        f.write(b'\x89PNG\r\n\x1a\n' + chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, 2, 0, 0, 0))
                + chunk(b'IDAT', zlib.compress(raw, 6)) + chunk(b'IEND', b''))
