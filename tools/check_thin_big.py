#!/usr/bin/env python
"""big 3-D block: more selected patches than the per-round lists of the thinning kernel
hold -> exercises its overflow paths; rounds vs one-selection-per-step must agree, and the
parallel cover must agree with the serial walk."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from patchperpix_b200 import synth, cuda_code as cc
from patchperpix_b200.assembly import BlockAssembler
shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (96, 224, 224)
dev = torch.device('cuda', 0)
ps = np.array([7, 7, 7])
labels, numinst = synth.neurites_3d(shape, n=max(3, int(np.prod(shape)) // 25000), seed=9,
                                    radius=(2, 3), seg_len=12.0, n_seg=30)
pred = synth.patches_from_labels(labels, ps, seed=9, device=dev)
fg = (pred[171] > 0.5).to(torch.uint8)
overlap = torch.from_numpy((numinst > 1).astype(np.uint8)).to(dev)
mask = fg.clone(); mask[overlap > 0] = 0
asm = BlockAssembler(pred, fg, overlap, ps, **bench.KW)
asm.prepare(); asm.consensus(); asm.rank(); order = asm.ranked()
sel = asm.cover(mask, order)
asm.kwargs['ppp_cover_serial'] = True
sel_serial = asm.cover(mask, order)
del asm.kwargs['ppp_cover_serial']
print('fg', int(fg.sum()), 'cover', sel.numel(), 'cover parallel == serial:', bool(torch.equal(sel, sel_serial)))
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); thin = asm.thin(mask, sel); e1.record()
cfg0 = asm.cfg
asm.cfg = cc.make_cfg(asm.shape, asm.ps, **dict(asm.kwargs, ppp_tune=0x10000))
thin_serial = asm.thin(mask, sel); e2.record(); torch.cuda.synchronize()
asm.cfg = cfg0
print('thin', thin.numel(), 'rounds == serial:', bool(torch.equal(thin, thin_serial)),
      'rounds %.2f ms serial %.2f ms' % (e0.elapsed_time(e1), e1.elapsed_time(e2)))
