#!/usr/bin/env python
"""BASELINE configs[0]: the reference's python CPU path (`cuda=false`) on the bundled
flylight crop (experiments/flylight/JRC_SS05008-20160318_24_B2_crop.zip).

Runs the UNMODIFIED reference functions fillLookup / computeFGBGsets /
create_consensus_array (utilVoteInstances.py:19-92, consensus_array.py:18-68) on seeded
patch predictions drawn around the crop's ground truth and records the int16 consensus
in the compact layout -> tests/golden/c1_cpu_consensus.npz.  SURVEY.md A.8: with plain
vote counting, the inverse-threshold background band and no overlap handling the CUDA
path's `cnt_pos - cnt_neg` must equal it (tests/test_gpu_big_golden.py).  Only runs
where /root/reference exists; the golden carries the labels so that the GPU box can
regenerate the predictions."""
import hashlib
import json
import os
import sys
import tempfile
import time
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_runner                       # noqa: E402
from patchperpix_b200 import io_util, layout, synth  # noqa: E402

ZIP = '/root/reference/experiments/flylight/JRC_SS05008-20160318_24_B2_crop.zip'
TH = 0.6
SEED = 31


def c1_inputs(labels):
    """predictions of the C1 case from its label volume (also used by the test)."""
    return synth.crop_case(labels, (7, 7, 7), seed=SEED, hard_frac=0.1)


def main():
    d = tempfile.mkdtemp()
    with zipfile.ZipFile(ZIP) as z:
        z.extractall(d)
    f = io_util.open_container(os.path.join(d, 'JRC_SS05008-20160318_24_B2_crop.zarr'))
    gt = np.array(f['volumes/gt_instances'])
    labels = np.zeros(gt.shape[1:], np.uint8)
    for c in range(gt.shape[0]):
        labels[gt[c] > 0] = c + 1
    ps = np.array([7, 7, 7])
    pred = c1_inputs(labels)
    mid = 343 // 2
    fg = pred[mid] > TH
    S = ref_runner.RefSession()
    uv, ca = S.mods['utilVoteInstances'], S.mods['consensus_array']
    neigh = 2 * ps
    rad = ps // 2
    allp = np.transpose(np.where(fg))
    t0 = time.time()
    lookup = uv.fillLookup(fg, ps, neigh, allp)
    t1 = time.time()
    allp = [p for p in allp if np.all(p >= rad) and np.all(p < np.array(fg.shape) - rad)]
    fgs, bgs = uv.computeFGBGsets(fg, allp, pred, ps, rad, isbiHack=False,
                                  patch_threshold=TH, sample=1.0)
    t2 = time.time()
    cons, _, _ = ca.create_consensus_array(fgs, bgs, fg.shape, ps, neigh, lookup)
    t3 = time.time()
    print('fg %d patches %d: lookup %.1fs sets %.1fs consensus %.1fs' % (
        int(fg.sum()), len(allp), t1 - t0, t2 - t1, t3 - t2), flush=True)
    # reference layout [code][Z][Y][X], code = oz*ns_y*ns_x + oy*ns_x + ox (may be negative
    # in y/x) -> compact [gated row][k]
    F = int(fg.sum())
    K = (13 ** 3 - 1) // 2
    comp = np.zeros((F, K), np.int16)
    k = 0
    for lin in range(K + 1, 13 ** 3):
        oz, oy, ox = lin // 169 - 6, (lin // 13) % 13 - 6, lin % 13 - 6
        code = oz * neigh[1] * neigh[2] + oy * neigh[2] + ox
        comp[:, k] = cons[code][fg]
        k += 1
    assert int(comp.astype(np.int64).sum()) == int(cons.astype(np.int64).sum()), \
        "votes outside the positive offsets"
    rows = np.sort(np.random.default_rng(0).choice(F, min(F, 600), replace=False))
    out = dict(labels=labels, rows=rows.astype(np.int32), cons=comp[rows],
               cons_sum=np.int64(comp.astype(np.int64).sum()),
               cons_abs_sum=np.int64(np.abs(comp.astype(np.int64)).sum()),
               pred_sha1=hashlib.sha1(pred.astype(np.float16).tobytes()).hexdigest(),
               th=np.float64(TH), seed=np.int64(SEED),
               timing=json.dumps(dict(fg=F, patches=len(allp), lookup_s=t1 - t0,
                                      sets_s=t2 - t1, consensus_s=t3 - t2, cores=1)))
    fn = os.path.join(ROOT, 'tests', 'golden', 'c1_cpu_consensus.npz')
    np.savez_compressed(fn, **out)
    print(fn, '%.2f MB' % (os.path.getsize(fn) / 1e6))


if __name__ == '__main__':
    main()
