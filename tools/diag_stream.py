#!/usr/bin/env python
"""per-sample timeline of the upload/assemble pipeline under torchrun: host time of the
upload call, device time of the H2D copy (events on the copy stream), completion marks."""
import os, sys, time, threading
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
pred, numinst, _ = bench.make_inputs(dev, 2)
fg = (pred[int(np.prod(ps)) // 2] > 0.5).to(torch.uint8)
pred_h = torch.empty(pred.shape, dtype=torch.float16).pin_memory(); pred_h.copy_(pred)
fg_h = fg.cpu().pin_memory(); numinst_h = torch.from_numpy(numinst).pin_memory()
del pred
copy = torch.cuda.Stream(dev)
bufs = [torch.empty(pred_h.shape, dtype=torch.float16, device=dev) for _ in range(2)]
def bar():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for _ in range(2): vi.to_instance_seg(pred_h, fg_h, fg_h, numinst_h, ps, **bench.KW)
bar()
t0 = time.perf_counter(); log = []
ev = []
def upload(i):
    th = time.perf_counter()
    with torch.cuda.stream(copy):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(copy); bufs[i & 1].copy_(pred_h, non_blocking=True); e1.record(copy)
    ev.append((e0, e1))
    return e1, (time.perf_counter() - th) * 1e3
n = 6
ready, hu = upload(0); hus = [hu]
for i in range(n):
    if i + 1 < n:
        nxt, hu = upload(i + 1); hus.append(hu)
    torch.cuda.current_stream().wait_event(ready)
    p32 = bufs[i & 1].float()
    vi.to_instance_seg(p32, fg_h, fg_h, numinst_h, ps, **bench.KW)
    log.append((time.perf_counter() - t0) * 1e3)
    if i + 1 < n: ready = nxt
torch.cuda.synchronize()
h2d = [a.elapsed_time(b) for a, b in ev]
print('rank %d inline-pipeline marks %s | upload host ms %s | h2d device ms %s' % (
    rank, ' '.join('%.0f' % m for m in log), ' '.join('%.1f' % h for h in hus),
    ' '.join('%.0f' % h for h in h2d)), flush=True)
bar()
def feed(k):
    for _ in range(k): yield (pred_h, fg_h, fg_h, numinst_h)
t0 = time.perf_counter(); marks = []
for _ in vi.to_instance_seg_stream(feed(n), ps, **bench.KW): marks.append((time.perf_counter() - t0) * 1e3)
print('rank %d library-pipeline marks %s' % (rank, ' '.join('%.0f' % m for m in marks)), flush=True)
if world > 1: dist.destroy_process_group()
