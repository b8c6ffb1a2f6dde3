"""Operator seam of step 2 (PatchPerPix/vote_instances/ranked_patches.py)."""
import numpy as np


def rank_patches_cuda(pred_affs, consensus_vote_array, patchshape, neighshape=None,
                      overlap_mask=None, **kwargs):
    """ranked_patches.py:33-74: score volume f32 [Z,Y,X] (numpy).

    `consensus_vote_array` is the ConsensusArray returned by
    create_consensus_array_cuda (it carries the device state of the block)."""
    asm = consensus_vote_array.asm
    return asm.rank().cpu().numpy()


def rank_patches_by_score(all_patches_idx, rank_scores):
    """ranked_patches.py:21-30 on the host: stable, descending."""
    s = rank_scores[tuple(np.asarray(all_patches_idx).T)]
    order = np.argsort(-s.astype(np.float64), kind='stable')
    return [(np.asarray(all_patches_idx)[i], s[i]) for i in order]
