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
       'flylight_default_kwargs')
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
