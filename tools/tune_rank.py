#!/usr/bin/env python
"""time ppp_rank on the bench workload for several values of the tuning knob."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from patchperpix_b200.assembly import BlockAssembler
dev = torch.device('cuda', 0)
ps = np.array([1, 41, 41])
pred, numinst, _ = bench.make_inputs(dev, 2)
P = int(np.prod(ps))
fg = (pred[P // 2] > 0.5).to(torch.uint8)
overlap = torch.from_numpy((numinst > 1).astype(np.uint8)).to(dev)
ref = None
for tune in [int(a) for a in sys.argv[1:]] or [0]:
    kw = dict(bench.KW, ppp_tune=tune)
    asm = BlockAssembler(pred, fg, overlap, ps, **kw)
    asm.prepare(); asm.consensus(); asm.rank()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        asm.rank()
    e1.record(); torch.cuda.synchronize()
    if ref is None:
        ref = asm.score.clone()
    print('tune %d: rank %.3f ms  identical=%s' % (tune, e0.elapsed_time(e1) / 3,
                                                 bool(torch.equal(ref, asm.score))))
