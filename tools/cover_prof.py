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
mask = fg.clone(); mask[overlap > 0] = 0
asm = BlockAssembler(pred, fg, overlap, ps, **bench.KW)
asm.prepare(); asm.consensus(); asm.rank(); order = asm.ranked()
lib = cc.load_library()
out = (ctypes.c_ulonglong * 8)()
sel = asm.cover(mask, order); torch.cuda.synchronize()
lib.ppp_debug_cover_prof(out, 1)
sel = asm.cover(mask, order); torch.cuda.synchronize()
lib.ppp_debug_cover_prof(out, 0)
v = list(out)
print('cycles: init %.3g  phaseA %.3g  compact %.3g  phaseB %.3g | chunks %d survivors %d subbatches %d selected %d' % (v[0], v[1], v[2], v[3], v[4], v[5], v[6], sel.numel()))
