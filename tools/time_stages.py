#!/usr/bin/env python
"""CUDA-event time of every stage of the assembly path on one block.
usage: python tools/time_stages.py c2 | fly [Z Y X] | nuclei [Z Y X]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from patchperpix_b200 import synth
from patchperpix_b200.assembly import BlockAssembler

which = sys.argv[1] if len(sys.argv) > 1 else 'c2'
dev = torch.device('cuda', 0)
if which == 'c2':
    ps = np.array([1, 41, 41])
    pred, numinst, _ = bench.make_inputs(dev, 2)
elif which == 'fly':
    shape = tuple(int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (98, 98, 98)
    ps = np.array([7, 7, 7])
    labels, numinst = synth.neurites_3d(shape, n=max(3, int(np.prod(shape)) // 40000), seed=4,
                                        radius=(2, 3), seg_len=12.0, n_seg=30)
    pred = synth.patches_from_labels(labels, ps, seed=4, device=dev)
else:
    shape = tuple(int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (32, 128, 128)
    ps = np.array([5, 21, 21])
    labels, numinst = synth.blobs_3d(shape, n=max(3, int(np.prod(shape)) // 13000), seed=3)
    pred = synth.patches_from_labels(labels, ps, seed=3, device=dev)
P = int(np.prod(ps))
fg = (pred[P // 2] > 0.5).to(torch.uint8)
overlap = torch.from_numpy((numinst > 1).astype(np.uint8)).to(dev)
mask = fg.clone(); mask[overlap > 0] = 0


def run(timed):
    ev = {}
    def tick(name):
        if timed:
            e = torch.cuda.Event(enable_timing=True); e.record(); ev[name] = e
    tick('start')
    asm = BlockAssembler(pred, fg, overlap, ps, **bench.KW)
    asm.prepare(); tick('prepare')
    asm.consensus(); tick('consensus')
    asm.rank(); tick('rank')
    order = asm.ranked(); tick('sort')
    sel = asm.cover(mask, order); tick('cover')
    sel = asm.thin(mask, sel); tick('thin')
    pairs = asm.patch_pairs(asm.coords(sel)); tick('pairs(host)')
    pd = torch.from_numpy(pairs.view(np.int32)).to(dev)
    aff = asm.patch_graph(pd); tick('patch_graph')
    inst, ncomp = asm.label(pd, aff, sel); tick('label')
    torch.cuda.synchronize()
    return asm, ev, int(sel.numel()), len(pairs), ncomp

run(False)
asm, ev, nsel, npairs, ncomp = run(True)
names = list(ev)
print('%s shape=%s ps=%s fg=%d rows=%d selected=%d pairs=%d instances=%d' % (
    which, tuple(pred.shape[1:]), tuple(ps), int(fg.sum()), asm.F, nsel, npairs, ncomp))
tot = ev[names[0]].elapsed_time(ev[names[-1]])
for a, b in zip(names[:-1], names[1:]):
    print('  %-12s %8.3f ms' % (b, ev[a].elapsed_time(ev[b])))
print('  %-12s %8.3f ms  -> %.3f M fg-voxel/s' % ('total', tot, int(fg.sum()) / tot / 1e3))
