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


@pytest.mark.gpu
def test_decode_rows_feed_the_rows_path():
    """configs[3] in small: codes of the foreground voxels -> f16 patch rows -> blockwise
    assembly, no dense [P,Z,Y,X] array anywhere; equals the dense route
    (decode_volume -> float16 -> to_instance_seg)."""
    import torch
    from patchperpix_b200.decoder import PatchDecoder
    from patchperpix_b200.assembly import RowSource
    from patchperpix_b200 import vote_instances as vi
    from oracle import decoder_torch
    W = decoder_torch.make_weights(seed=3)
    dec = PatchDecoder(W)
    rng = np.random.default_rng(0)
    shape = (14, 24, 24)
    fg = rng.random(shape) < 0.12
    fg[:3] = fg[-3:] = False
    fg[:, :3] = fg[:, -3:] = False
    fg[:, :, :3] = fg[:, :, -3:] = False
    c = np.argwhere(fg)
    codes = torch.from_numpy(rng.random((len(c), 176)).astype(np.float32)).cuda()
    rows = dec.decode_rows(codes)
    assert rows.dtype == torch.float16 and rows.shape == (len(c), 343)
    ref = dec.decode(codes, sigmoid=True)
    assert float((rows.float() - ref).abs().max()) <= 2.5e-4
    v2r = torch.full(shape, -1, dtype=torch.int32)
    v2r[c[:, 0], c[:, 1], c[:, 2]] = torch.arange(len(c), dtype=torch.int32)
    src = RowSource(rows, v2r.cuda())
    kw = dict(patch_threshold=0.5, fc_threshold=0.5, cuda=True, blockwise=False,
              select_patches_for_sparse_data=True, includeSinglePatchCCS=True, mws=False,
              consensus_norm_prob_product=True, consensus_prob_product=True,
              consensus_norm_aff=True, consensus_interleaved_cnt=False,
              vi_bg_use_inv_th=False, vi_bg_use_half_th=False, vi_bg_use_less_than_th=True,
              rank_norm_patch_score=True, rank_int_counter=False, patch_graph_norm_aff=True,
              overlapping_inst=False)
    fgt = torch.from_numpy(fg.astype(np.uint8)).cuda()
    a, _ = vi.to_instance_seg(src, fgt, fgt.clone(), fgt.clone(), np.array([7, 7, 7]), **kw)
    b, _ = vi.to_instance_seg(src.dense(), fgt, fgt.clone(), fgt.clone(), np.array([7, 7, 7]), **kw)
    assert np.array_equal(a, b)
