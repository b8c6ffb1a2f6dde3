"""Operator seam of step 1, named like the reference's
PatchPerPix/vote_instances/consensus_array.py.

`create_consensus_array_cuda(pred_affs, overlap_mask, patchshape, neighshape,
**kwargs)` keeps the reference signature (consensus_array.py:71) but returns a
`ConsensusArray` (compact device layout) instead of the dense
`[NSZ][NSY][NSX][Z][Y][X]` managed array; `.to_dense()` gives the reference
layout for small blocks.
"""
import numpy as np

from .assembly import BlockAssembler
from . import layout


class ConsensusArray:
    def __init__(self, asm):
        self.asm = asm
        self.cons = asm.cons          # f32 [F][K] cuda
        self.cnt = asm.cnt            # i32 [F][K] cuda or None: (neg << 16) | pos
        self.fgidx = asm.fgidx
        self.flags = asm.flags
        self.shape = asm.shape
        self.patchshape = asm.ps

    def gate(self):
        """bool [Z,Y,X]: voxels that own a (possibly non-zero) row."""
        return ((self.flags & 2) != 0).reshape(self.shape).cpu().numpy()

    def rows_of_gate(self):
        """row index (into cons) of every gated voxel, raster order."""
        g = (self.flags & 2) != 0
        return self.fgidx[g].cpu().numpy()

    def compact(self, what='cons'):
        """numpy [F_gated][K] in the layout of patchperpix_b200/layout.py."""
        rows = self.rows_of_gate()
        if what == 'cons':
            return self.cons.cpu().numpy()[rows]
        c = self.cnt.cpu().numpy().view(np.uint32)[rows]
        if what == 'pos':
            return (c & 0xffff).astype(np.uint16)
        if what == 'neg':
            return (c >> 16).astype(np.uint16)
        return ((c & 0xffff) + (c >> 16)).astype(np.uint32)

    def to_dense(self):
        """reference layout [NSZ][NSY][NSX][Z][Y][X] (consensus_array.py:103-105)."""
        return layout.compact_to_dense(self.compact('cons'), self.gate(), self.patchshape)


def create_consensus_array_cuda(pred_affs, overlap_mask, patchshape, neighshape=None,
                                want_cnt=False, **kwargs):
    """consensus_array.py:71-206: fill (+ count) + normalise, fused."""
    import torch
    from .vote_instances import _to_device
    pred = _to_device(pred_affs, torch.float32)
    ov = _to_device(np.asarray(overlap_mask) != 0 if isinstance(overlap_mask, np.ndarray)
                    else overlap_mask, torch.uint8)
    mid = int(np.prod(patchshape)) // 2
    fg = (pred[mid] > float(np.float32(kwargs['patch_threshold']))).to(torch.uint8)
    asm = BlockAssembler(pred, fg, ov, patchshape, **kwargs)
    asm.prepare()
    asm.consensus(want_cnt=want_cnt)
    return ConsensusArray(asm)


def loadOrComputeConsensus(instances, patchshape, neighshape, all_patches, pred_affs,
                           rad, foreground, lookup, overlap_mask, **kwargs):
    """consensus_array.py:209-246 (CUDA branch only)."""
    if not kwargs.get('cuda', True):
        raise NotImplementedError("CPU consensus is not part of this build")
    return create_consensus_array_cuda(pred_affs, overlap_mask, patchshape, neighshape,
                                       **kwargs), None, None
