"""input of the no_overlap_per_channel golden (tools/gen_golden.py nooverlap and
tests/test_mws.py regenerate it from here)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from patchperpix_b200 import synth  # noqa: E402


def no_overlap_case():
    """five discs; discs 1 and 2 overlap and BOTH claim the shared region (the patches of
    disc 1 come from a label volume where the overlap is theirs, those of disc 2 from one
    where it is disc 2's) -> their painted masks intersect; disc 5 is small."""
    ps = np.array([1, 9, 9])
    yy, xx = np.mgrid[0:120, 0:160]
    cen = ((40, 45), (60, 75), (85, 120), (30, 125), (100, 30))
    rad = (32, 32, 28, 28, 9)
    discs = [((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r) for (cy, cx), r in zip(cen, rad)]
    numinst = np.zeros((1, 120, 160), np.uint8)
    for d in discs:
        numinst[0][d] += 1
    la = np.zeros((1, 120, 160), np.int32)
    lb = np.zeros((1, 120, 160), np.int32)
    for i in (1, 0, 2, 3, 4):          # disc 1 painted last in la: owns the overlap
        la[0][discs[i]] = i + 1
    for i in (0, 1, 2, 3, 4):          # disc 2 painted after disc 1 in lb
        lb[0][discs[i]] = i + 1
    pa = synth.patches_from_labels(la, ps, seed=31)
    pb = synth.patches_from_labels(lb, ps, seed=31)
    own2 = discs[1] & ~discs[0]
    pred = pa.copy()
    pred[:, 0][:, own2] = pb[:, 0][:, own2]
    return pred, numinst, la


