#!/usr/bin/env python
"""one block of the bench volume (compact rows) through to_instance_seg, for ncu:
  ncu --set full --import-source on -k regex:'consensus_small|thin_rounds|patch_graph_ref|rank_ref' \
      -c 8 -o gpurun_out/r2_block python tools/profile_block.py
usage: profile_block.py [Z Y X] [reps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from patchperpix_b200 import synth, sharded, vote_instances as vi  # noqa: E402


def main():
    shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (134, 70, 262)
    reps = int(sys.argv[4]) if len(sys.argv) >= 5 else 2
    w = dict(bench.WORKLOAD)
    dev = torch.device('cuda', 0)
    # a region in the middle of the bench volume, as its own little volume
    full = w['shape']
    start = [(f - s) // 2 for f, s in zip(full, shape)]
    if os.environ.get('PPP_START'):
        start = [int(v) for v in os.environ['PPP_START'].split(',')]
    c, p, ni = synth.neurite_rows(full, w['patchshape'], axis=0, lo=start[0], hi=start[0] + shape[0],
                                  device=dev, box=(np.array(start), np.array(start) + np.array(shape)),
                                  **bench.synth_kw(w))
    c = c - torch.tensor(start, dtype=torch.int32, device=dev)
    shard = sharded.RowShard(shape, 0, 0, shape[0], c, p, ni)
    shard.exchange_halo([(0, shape[0])], 0)
    kw = dict(bench.KW, patchshape=list(w['patchshape']))
    if os.environ.get('PPP_TUNE'):
        kw['ppp_tune'] = int(os.environ['PPP_TUNE'], 0)
    src, fg, mask, numinst, _ = shard.region(np.zeros(3, int), np.array(shape), **kw)
    print('rows', int(c.shape[0]))
    if os.environ.get('PPP_STAGES'):
        from patchperpix_b200 import cuda_code as cc
        kw['ppp_latency_stream'] = False
        vi.do_block(src, fg, mask, numinst, return_intermediates=True, **kw)
        with bench.CallTimer(cc, torch) as ct:
            for i in range(3):
                vi.do_block(src, fg, mask, numinst, return_intermediates=True, **kw)
        print('tune', kw.get('ppp_tune', 0), {k: round(v['ms'] / 3, 3) for k, v in ct.calls.items()})
        return
    if os.environ.get('PPP_CPROFILE'):
        import cProfile
        import pstats
        vi.do_block(src, fg, mask, numinst, return_intermediates=True, **kw)
        pr = cProfile.Profile()
        pr.enable()
        for i in range(5):
            vi.do_block(src, fg, mask, numinst, return_intermediates=True, **kw)
        pr.disable()
        pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
        return
    for i in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        pairs, aff = vi.do_block(src, fg, mask, numinst, return_intermediates=True, **kw)
        torch.cuda.synchronize()
        print('block %.2f ms, %d pairs' % ((time.perf_counter() - t) * 1e3, len(pairs)))


if __name__ == '__main__':
    main()
