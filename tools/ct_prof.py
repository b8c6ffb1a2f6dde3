import sys, ctypes, numpy as np, torch
sys.path.insert(0,'/root/repo')
import bench
from patchperpix_b200.assembly import BlockAssembler
from patchperpix_b200 import cuda_code as cc
dev = torch.device('cuda', 0)
ps = np.array([1, 41, 41]); pred, numinst, _ = bench.make_inputs(dev, 2)
P = int(np.prod(ps))
fg = (pred[P // 2] > 0.5).to(torch.uint8)
overlap = torch.from_numpy((numinst > 1).astype(np.uint8)).to(dev)
asm = BlockAssembler(pred, fg, overlap, ps, **bench.KW)
asm.prepare(); asm.consensus(impl=3); torch.cuda.synchronize()
lib = cc.load_library()
out = (ctypes.c_ulonglong * 8)()
lib.ppp_debug_ct_prof(out, 1)
asm.consensus(impl=3); torch.cuda.synchronize()
lib.ppp_debug_ct_prof(out, 0)
v = list(out)
print('cycles preamble %.3g main %.3g epilogue %.3g | CTAs %d items %d batches %d' % (v[0], v[1], v[2], v[4], v[5], v[6]))
print('per CTA: preamble %.0f main %.0f epi %.0f cycles; items/CTA %.1f batches/CTA %.2f fill %.2f' % (v[0]/v[4], v[1]/v[4], v[2]/v[4], v[5]/v[4], v[6]/v[4], v[5]/(v[6]*256.0)))
