"""helpers to load tests/golden/*.npz (written by tools/gen_golden.py from the
unmodified reference) and to regenerate the inputs they were computed on."""
import glob
import hashlib
import json
import os

import numpy as np

from patchperpix_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
# BASELINE-size cases (tools/gen_golden.py BIG_CASES): row samples + checksums, compared on
# the GPU only (tests/test_gpu_big_golden.py) -- the CPU oracle would need minutes on them
BIG = ('worms2d_c2_full', 'blobs3d_c3_block', 'blockwise3d_3x3x3_mws', 'c1_cpu_consensus',
       'flylight_default_kwargs', 'blockwise3d_bb_post')
NAMES = sorted(os.path.basename(f)[:-4]
               for f in glob.glob(os.path.join(GOLD, '*.npz'))
               if not os.path.basename(f).startswith(('blockwise', 'mws_', 'chan_'))
               and os.path.basename(f)[:-4] not in BIG)
_TUPLES = ('shape', 'width', 'length', 'radius', 'rad_xy', 'rad_z', 'centers')


def load(name):
    g = dict(np.load(os.path.join(GOLD, name + '.npz')))
    kw = json.loads(str(g['kwargs']))
    skw = json.loads(str(g['synth']))
    for k in _TUPLES:
        if k in skw:
            skw[k] = tuple(skw[k])
    ps = g['patchshape']
    if 'pred_f16' in g:
        pred = g['pred_f16'].astype(np.float32)
    else:
        pred = synth.make_case(patchshape=ps, **skw)[0]
    sha = hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest()
    assert sha == str(g['pred_sha1']), 'synthetic input drifted for ' + name
    return g, kw, ps, pred


def bb_case():
    """volume with an empty margin and two tiny spurious blobs (so that only_bb crops and
    ignore_small_comps matters); shared with tests/test_boundary.py through the golden."""
    ps = np.array([5, 5, 5])
    skw = dict(kind='neurites', seed=29, shape=(30, 60, 60), n=8, radius=(1.5, 2.5),
               seg_len=9.0, n_seg=8)
    pred, numinst, labels = synth.make_case(patchshape=ps, **skw)
    pred[:, :, :9, :] = 0
    numinst[:, :9, :] = 0
    pred[:, :, :, 52:] = 0
    numinst[:, :, 52:] = 0
    # a 2x2x2 blob of "foreground" far from everything, inside the cut margin
    pred[:, 3:5, 2:4, 2:4] = 0.9
    numinst[3:5, 2:4, 2:4] = 1
    return ps, skw, pred, numinst
