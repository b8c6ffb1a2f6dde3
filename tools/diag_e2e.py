#!/usr/bin/env python
"""per-rank timing of the end-to-end legs (serial, streamed) and of the bare H2D copy."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from patchperpix_b200 import cuda_code as cc, vote_instances as vi
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local); dev = torch.device('cuda', local)
if world > 1: dist.init_process_group('nccl', device_id=dev)
cc.init_cuda()
ps = np.array(bench.WORKLOAD['patchshape'])
pred, numinst, _ = bench.make_inputs(dev, 2 + rank)
fg = (pred[int(np.prod(ps)) // 2] > 0.5).to(torch.uint8)
pred_h = torch.empty(pred.shape, dtype=torch.float16).pin_memory(); pred_h.copy_(pred)
fg_h = fg.cpu().pin_memory(); numinst_h = torch.from_numpy(numinst).pin_memory()
del pred
def bar():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
buf = torch.empty(pred_h.shape, dtype=torch.float16, device=dev)
for _ in range(2): buf.copy_(pred_h, non_blocking=True)
bar(); t = time.perf_counter()
for _ in range(5): buf.copy_(pred_h, non_blocking=True)
torch.cuda.synchronize(); h2d = (time.perf_counter() - t) / 5 * 1e3
for _ in range(2): vi.to_instance_seg(pred_h, fg_h, fg_h, numinst_h, ps, **bench.KW)
bar(); t = time.perf_counter()
for _ in range(5): vi.to_instance_seg(pred_h, fg_h, fg_h, numinst_h, ps, **bench.KW)
torch.cuda.synchronize(); ser = (time.perf_counter() - t) / 5 * 1e3
def feed(n):
    for _ in range(n): yield (pred_h, fg_h, fg_h, numinst_h)
for _ in vi.to_instance_seg_stream(feed(2), ps, **bench.KW): pass
bar(); t = time.perf_counter(); marks = []
for _ in vi.to_instance_seg_stream(feed(6), ps, **bench.KW): marks.append(time.perf_counter() - t)
torch.cuda.synchronize(); st = (time.perf_counter() - t) / 6 * 1e3
print('rank %d: h2d %.1f ms (%.1f GB/s)  serial %.1f ms  stream %.1f ms  marks %s' % (
    rank, h2d, pred_h.numel() * 2 / h2d / 1e6, ser, st, ' '.join('%.0f' % (m * 1e3) for m in marks)), flush=True)
if world > 1: dist.destroy_process_group()
