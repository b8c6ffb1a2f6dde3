#!/usr/bin/env python
"""time ppp_consensus implementations (1 simple, 2 bit-guided gather, 3 tiled)
on a workload.  usage: python tools/tune_consensus.py c2|fly|nuclei [impl ...]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from patchperpix_b200 import synth
from patchperpix_b200.assembly import BlockAssembler
which = sys.argv[1]
dev = torch.device('cuda', 0)
if which == 'c2':
    ps = np.array([1, 41, 41]); pred, numinst, _ = bench.make_inputs(dev, 2)
elif which == 'fly':
    ps = np.array([7, 7, 7])
    labels, numinst = synth.neurites_3d((98, 98, 98), n=23, seed=4, radius=(2, 3), seg_len=12.0, n_seg=30)
    pred = synth.patches_from_labels(labels, ps, seed=4, device=dev)
else:
    ps = np.array([5, 21, 21])
    labels, numinst = synth.blobs_3d((32, 128, 128), n=40, seed=3)
    pred = synth.patches_from_labels(labels, ps, seed=3, device=dev)
P = int(np.prod(ps))
fg = (pred[P // 2] > 0.5).to(torch.uint8)
overlap = torch.from_numpy((numinst > 1).astype(np.uint8)).to(dev)
asm = BlockAssembler(pred, fg, overlap, ps, **bench.KW)
asm.prepare()
ref = None
for impl in [int(a) for a in sys.argv[2:]] or [1, 2, 3]:
    asm.consensus(impl=impl); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): asm.consensus(impl=impl)
    e1.record(); torch.cuda.synchronize()
    if ref is None: ref = (asm.cons.clone(), asm.cnt.clone())
    print('%s impl %d: %.3f ms identical=%s' % (which, impl, e0.elapsed_time(e1) / 3,
          bool(torch.equal(ref[0], asm.cons) and torch.equal(ref[1], asm.cnt))))
