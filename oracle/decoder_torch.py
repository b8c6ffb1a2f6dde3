"""TEST INFRASTRUCTURE ONLY — pure-torch restatement of the ppp+dec decoder.

Reference: `Autoencoder.forward` (experiments/flylight/setups/setup01/
torch_model.py:523-544) with the flylight values of default_train_code.toml:61-85
(code_units 176 = 22 fmaps x 2^3, num_fmaps [64,128], kernel 3, 2 repetitions,
upsampling "resize_conv", padding "same", activation relu).

**Parity unpinned.**  `ConvPass` / `Upsample` come from the Kainmueller-Lab fork
of funlib.learn.torch (branch `ppp`, un-pinned in setup.py:31) which is neither
vendored nor installed, and no reference test pins decoder outputs.  Assumed
definitions:
  ConvPass(in, out, kernel_sizes, activation, padding='same')
      = one Conv3d per kernel size (padding k//2), each followed by the
        activation if it is not None;
  Upsample(factor, mode='resize_conv', in, out, activation, padding='same')
      = nearest-neighbour up-sampling by `factor`, then Conv3d(in, out, 3,
        padding 1), then the activation.
"""
import numpy as np
import torch


def make_weights(seed=0, code_fmaps=22, fmaps=(64, 128), gain=1.0):
    """seeded decoder weights (torch's default conv init x gain), float32 numpy.
    gain > 1 widens the logit range so that a test sees more than a constant."""
    g = torch.Generator().manual_seed(seed)

    def conv(cout, cin, k):
        fan_in = cin * k ** 3
        bound = 1.0 / np.sqrt(fan_in)
        w = (torch.rand((cout, cin, k, k, k), generator=g) * 2 - 1) * bound * gain
        b = (torch.rand((cout,), generator=g) * 2 - 1) * bound
        return w.numpy(), b.numpy()
    f1, f0 = fmaps[1], fmaps[0]
    W = {}
    W['from_code.w'], W['from_code.b'] = conv(f1, code_fmaps, 1)
    W['up0.w'], W['up0.b'] = conv(f0, f1, 3)
    W['conv0a.w'], W['conv0a.b'] = conv(f0, f0, 3)
    W['conv0b.w'], W['conv0b.b'] = conv(f0, f0, 3)
    W['up1.w'], W['up1.b'] = conv(1, f0, 3)
    W['conv1a.w'], W['conv1a.b'] = conv(1, 1, 3)
    W['conv1b.w'], W['conv1b.b'] = conv(1, 1, 3)
    return W


def decode_ref(codes, W, patchshape=(7, 7, 7), dtype=torch.float32):
    """codes [B, 176] -> patch logits [B, 343] (torch, CPU or CUDA)."""
    F = torch.nn.functional
    dev = codes.device
    t = {k: torch.as_tensor(v, dtype=dtype, device=dev) for k, v in W.items()}
    B = codes.shape[0]
    cf = W['from_code.w'].shape[1]
    s = round((codes.shape[1] / cf) ** (1 / 3))
    x = codes.to(dtype).reshape(B, cf, s, s, s)                     # torch_model.py:537
    x = F.relu(F.conv3d(x, t['from_code.w'], t['from_code.b']))     # from_code
    x = F.interpolate(x, scale_factor=2, mode='nearest')            # up[0]
    x = F.relu(F.conv3d(x, t['up0.w'], t['up0.b'], padding=1))
    x = F.relu(F.conv3d(x, t['conv0a.w'], t['conv0a.b'], padding=1))  # up_conv[0]
    x = F.relu(F.conv3d(x, t['conv0b.w'], t['conv0b.b'], padding=1))
    x = F.interpolate(x, scale_factor=2, mode='nearest')            # up[1]
    x = F.relu(F.conv3d(x, t['up1.w'], t['up1.b'], padding=1))
    x = F.conv3d(x, t['conv1a.w'], t['conv1a.b'], padding=1)        # up_conv[1], no activation
    x = F.conv3d(x, t['conv1b.w'], t['conv1b.b'], padding=1)
    pz, py, px = patchshape
    return x[:, 0, :pz, :py, :px].reshape(B, -1)                    # crop, torch_model.py:543
