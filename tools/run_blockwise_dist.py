#!/usr/bin/env python
"""blockwise assembly of a synthetic 3-D volume under torchrun (NCCL): every
rank takes blocks / face jobs round-robin; rank 0 prints a digest of the labels
so that runs with different world sizes can be compared.
usage: [PPP_CHUNK=z,y,x] [PPP_MWS=1] torchrun --nproc-per-node N tools/run_blockwise_dist.py [Z Y X]"""
import hashlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from patchperpix_b200 import synth, stitch_patch_graph as spg  # noqa: E402
import bench  # noqa: E402


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (48, 160, 160)
    ps = np.array([7, 7, 7])
    labels, numinst = synth.neurites_3d(shape, n=max(4, int(np.prod(shape)) // 60000), seed=4,
                                        radius=(2, 3), seg_len=12.0, n_seg=30)
    pred = synth.patches_from_labels(labels, ps, seed=4).astype(np.float16)
    chunk = [int(v) for v in os.environ.get('PPP_CHUNK', '24,80,80').split(',')]
    kw = dict(bench.KW, patchshape=[7, 7, 7], chunksize=chunk, blockwise=True,
              numinst_key=None, fg_key=None, mws=os.environ.get('PPP_MWS', '0') == '1')
    inputs = spg.VolumeInputs(pred)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    inst, fg, info = spg.stitch_arrays(inputs, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print('world=%d shape=%s fg=%d blocks=%d faces=%d edges=%d instances=%d time=%.2fs sha1=%s'
              % (world, shape, int(fg.sum()), info['n_blocks'], info['n_faces'], info['n_edges'],
                 len(np.unique(inst)) - 1, dt, hashlib.sha1(inst.tobytes()).hexdigest()[:16]))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
