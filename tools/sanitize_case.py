#!/usr/bin/env python
"""small 2-D (41x41 patches: tiled consensus) and 3-D (7^3: bit-guided consensus) cases
through the whole path, for compute-sanitizer runs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from patchperpix_b200 import synth, vote_instances as vi

for kind, ps, skw in (('worms', (1, 41, 41), dict(seed=7, shape=(96, 128), n_worms=4, width=(8, 11),
                                                  length=(60, 110))),
                      ('neurites', (7, 7, 7), dict(seed=11, shape=(20, 36, 36), n=4,
                                                   radius=(1.5, 2.5), seg_len=10.0, n_seg=6))):
    ps = np.array(ps)
    pred, numinst, _ = synth.make_case(kind=kind, patchshape=ps, **skw)
    fg = pred[int(np.prod(ps)) // 2] > np.float32(0.5)
    for mws in (False, True):
        inst, _ = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), numinst.copy(), ps.copy(),
                                     **dict(bench.KW, mws=mws))
        print(kind, tuple(ps), 'mws' if mws else 'cc', 'instances', len(np.unique(inst)) - 1, flush=True)
