#!/usr/bin/env python
"""one warm-up + one profiled step of the bench workload, bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees exactly one
step.  usage: ncu ... python tools/profile_step.py [--config c2|fly]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    ps = np.array(bench.WORKLOAD['patchshape'])
    pred, numinst, _ = bench.make_inputs(dev, bench.WORKLOAD['seed'])
    fg = (pred[int(np.prod(ps)) // 2] > 0.5).to(torch.uint8)
    overlap = torch.from_numpy((numinst > 1).astype(np.uint8)).to(dev)
    mask = fg.clone()
    mask[overlap > 0] = 0
    bench.device_step(pred, fg, overlap, mask, ps, bench.KW)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    bench.device_step(pred, fg, overlap, mask, ps, bench.KW)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == '__main__':
    main()
