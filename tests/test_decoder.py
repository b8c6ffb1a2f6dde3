"""ppp+dec decoder (tcgen05 implicit-GEMM convolutions) against the pure-torch
restatement of `Autoencoder.forward` (oracle/decoder_torch.py; parity unpinned:
the layer library of the reference is not vendored, see its header).

Tolerance: operands are fp16 (the dtype the reference stores codes and patches
in), accumulation fp32: max abs error of the logits <= 2e-3 * max|logit|, and
<= 1e-3 on the sigmoid patches (the bar BASELINE.json states for 16-bit)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B', [1, 2, 7, 300])
def test_decoder_matches_torch(B):
    import torch
    from oracle import decoder_torch as dt
    from patchperpix_b200.decoder import PatchDecoder
    W = dt.make_weights(seed=3, gain=2.5)
    g = torch.Generator().manual_seed(B)
    codes = torch.rand((B, 176), generator=g).to(torch.float16).float().cuda()
    ref = dt.decode_ref(codes, W)
    dec = PatchDecoder(W)
    out = dec.decode(codes)
    assert out.shape == (B, 343)
    scale = float(ref.abs().max())
    err = float((out - ref).abs().max())
    assert err <= 2e-3 * max(scale, 1.0), (err, scale)
    perr = float((torch.sigmoid(out) - torch.sigmoid(ref)).abs().max())
    assert perr <= 1e-3, perr
    assert float(ref.std()) > 1e-2            # the test sees structure, not a constant


def test_decode_volume_scatters_foreground_only():
    import torch
    from oracle import decoder_torch as dt
    from patchperpix_b200.decoder import PatchDecoder
    W = dt.make_weights(seed=4, gain=2.5)
    dec = PatchDecoder(W)
    g = torch.Generator().manual_seed(0)
    code = torch.rand((176, 3, 10, 12), generator=g).cuda()
    fg = (torch.rand((3, 10, 12), generator=g) > 0.7).cuda()
    vol = dec.decode_volume(code, fg, sigmoid=True)
    assert vol.shape == (343, 3, 10, 12)
    assert float(vol[:, ~fg].abs().max()) == 0.0
    ref = torch.sigmoid(dt.decode_ref(code.reshape(176, -1).T[fg.reshape(-1)], W))
    assert float((vol[:, fg].T - ref).abs().max()) <= 1e-3
