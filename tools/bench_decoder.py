#!/usr/bin/env python
"""throughput of the ppp+dec decoder (ppp_decode): codes/s and dense fp16
TFLOP/s of the three tensor-core convolutions."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from patchperpix_b200.decoder import PatchDecoder


def random_weights(seed=0, gain=2.5):
    """seeded weights with the decoder's layer shapes (values do not matter for timing)."""
    rng = np.random.default_rng(seed)

    def conv(cout, cin, k):
        b = 1.0 / np.sqrt(cin * k ** 3)
        return (rng.uniform(-b, b, (cout, cin, k, k, k)).astype(np.float32) * gain,
                rng.uniform(-b, b, (cout,)).astype(np.float32))
    W = {}
    for name, (co, ci, k) in dict(from_code=(128, 22, 1), up0=(64, 128, 3), conv0a=(64, 64, 3),
                                  conv0b=(64, 64, 3), up1=(1, 64, 3), conv1a=(1, 1, 3),
                                  conv1b=(1, 1, 3)).items():
        W[name + '.w'], W[name + '.b'] = conv(co, ci, k)
    return W


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
