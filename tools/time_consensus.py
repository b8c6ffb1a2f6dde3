#!/usr/bin/env python
"""time ppp_consensus (count + sums) on the bench image, with debug switches."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from patchperpix_b200.assembly import BlockAssembler
dev = torch.device('cuda', 0)
ps = np.array(bench.WORKLOAD['patchshape'])
pred, numinst, _ = bench.make_inputs(dev, bench.WORKLOAD['seed'])
fg = (pred[int(np.prod(ps)) // 2] > 0.5).to(torch.uint8)
overlap = torch.from_numpy((numinst > 1).astype(np.uint8)).to(dev)
asm = BlockAssembler(pred, fg, overlap, ps, **bench.KW)
asm.prepare()
for dbg in [int(a) for a in sys.argv[1:]] or [0]:
    asm.cfg.reserved = dbg
    for _ in range(2):
        asm.consensus()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        asm.consensus()
    e1.record(); torch.cuda.synchronize()
    print('debug=%d  consensus %.2f ms' % (dbg, e0.elapsed_time(e1) / 3))
