"""Operator seam of step 5 (PatchPerPix/vote_instances/aff_patch_graph.py)."""
import numpy as np


def computePatchGraph_cuda(pred_affs, consensus_vote_array, selected_patch_pairsIDs,
                           patchshape, neighshape=None, **kwargs):
    """aff_patch_graph.py:113-187, matrix form: aff f32 [n] (numpy)."""
    import torch
    asm = consensus_vote_array.asm
    pairs = np.ascontiguousarray(selected_patch_pairsIDs, np.uint32)
    pd = torch.from_numpy(pairs.view(np.int32)).to(asm.dev)
    return asm.patch_graph(pd).cpu().numpy()


def setAffgraph(graphMat, computed_pairs):
    """aff_patch_graph.py:31-40 (host networkx graph, for callers that want it)."""
    import networkx as nx
    g = nx.Graph()
    for idx, p in enumerate(graphMat):
        if p != 0:
            g.add_edge(tuple(computed_pairs[idx, :3]), tuple(computed_pairs[idx, 3:6]), aff=p)
    return g
