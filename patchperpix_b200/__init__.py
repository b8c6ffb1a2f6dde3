"""patchperpix_b200 — B200-native instance assembly for PatchPerPix.

Drop-in for the reference's `PatchPerPix.vote_instances` package on the
`run_ppp.py --do decode label` path (PatchPerPix/vote_instances/__init__.py:1-3):
    patchperpix_b200.vote_instances.main(**kwargs)
    patchperpix_b200.stitch_patch_graph.main(pred_file, **kwargs)
"""
from . import vote_instances, stitch_patch_graph          # noqa: F401
from .stitch_patch_graph import main, get_offsets, get_offset_str  # noqa: F401  (as the reference)
from .postprocess import clean_mask                         # noqa: F401

__all__ = ['vote_instances', 'stitch_patch_graph', 'cuda_code', 'assembly', 'sharded', 'pipeline',
           'consensus_array', 'ranked_patches', 'aff_patch_graph', 'postprocess', 'io_util',
           'decoder', 'synth', 'layout', 'main', 'get_offsets', 'get_offset_str', 'clean_mask']
