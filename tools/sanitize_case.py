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

# the compact-rows path: 2x2x2 blocks + faces through the sharded driver (one rank), the
# block pipeline and the batched face jobs, then per-job host threads
import torch
from patchperpix_b200 import sharded

ps = np.array([7, 7, 7])
shape = (20, 36, 36)
pred, numinst, _ = synth.make_case(kind='neurites', patchshape=ps, seed=11, shape=shape, n=4,
                                   radius=(1.5, 2.5), seg_len=10.0, n_seg=6)
fg = pred[171] > np.float32(0.5)
for mws in (False, True):
    for extra in (dict(), dict(ppp_pipeline=False, ppp_batch_faces=False)):
        kw = dict(bench.KW, mws=mws, patchshape=[7, 7, 7], chunksize=[12, 20, 20], **extra)
        axis, slabs = sharded.slab_partition(shape, kw['chunksize'], 1)
        c, p, ni, f = sharded.rows_from_dense(pred, fg, numinst, axis, 0, shape[axis], 'cuda', 0.5)
        shard = sharded.RowShard(shape, axis, 0, shape[axis], c, p, ni, f)
        inst, info = sharded.stitch_shard(shard, slabs, workers=2, **kw)
        torch.cuda.synchronize()
        print('rows', 'mws' if mws else 'cc', sorted(extra), 'blocks', info['n_blocks'],
              'instances', len(torch.unique(inst)) - 1, flush=True)
