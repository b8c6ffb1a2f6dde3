#!/usr/bin/env python
"""throughput of the ppp+dec decoder (ppp_decode): codes/s and dense fp16
TFLOP/s of the three tensor-core convolutions."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from patchperpix_b200.decoder import PatchDecoder


from patchperpix_b200.decoder import seeded_weights as random_weights  # noqa: E402


B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dec = PatchDecoder(random_weights())
codes = torch.rand((B, 176), device='cuda')
for _ in range(3):
    dec.decode(codes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for _ in range(n):
    dec.decode(codes)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flop = 2 * B * 64 * 27 * (128 * 64 + 64 * 64 + 64 * 64)
print('B=%d  %.3f ms  %.2f M codes/s  tensor-core convs %.1f TFLOP/s (of %.1f total MFLOP/code)'
      % (B, ms, B / ms / 1e3, flop / ms / 1e9, 58.6))
