"""ppp+dec decoder: per-voxel codes -> shape patches on the B200 tensor cores.

Counterpart of `decode_sample` / `Autoencoder.forward`
(experiments/flylight/setups/setup01/decode.py:16-65, torch_model.py:523-544):
the reference gathers the codes of the foreground voxels with a python list
comprehension, runs the decoder in batches of 1024 through cuDNN and scatters
the patches back voxel by voxel.  Here the gather is a torch index (plumbing),
the decoder is `ppp_decode` (tcgen05 implicit-GEMM convolutions, csrc/
ppp_decoder.cu) and the result stays compact [F][P] on the device.

The layer definitions of the un-vendored funlib.learn.torch fork are assumed
as documented in DESIGN.md §8 / oracle/decoder_torch.py (parity unpinned).
"""
import numpy as np

from . import cuda_code as cc


def _pack_conv(w):
    """torch conv weight [64][Cin][3][3][3] -> fp16 [27][Cin/64][64 cout][64 cin]."""
    cout, cin = w.shape[0], w.shape[1]
    assert cout == 64 and cin % 64 == 0
    t = w.reshape(cout, cin // 64, 64, 27)            # [co][cc][ci][tap]
    return np.ascontiguousarray(t.transpose(3, 1, 0, 2)).astype(np.float16)


def _fold_upsample_conv(w):
    """nearest x2 up-sampling followed by a 3^3 'same' convolution (weights
    [1][64][3][3][3]) = for each output parity (a,b,c) a convolution over the
    low-resolution input whose taps are sums of the original ones: along one
    axis, parity 0 reads {i-1: w0, i: w1+w2}, parity 1 reads {i: w0+w1, i+1: w2}.
    Returns fp16 [27][1][16][64] = [tap][cin chunk][parity (8 real, 8 zero)][cin]."""
    w = np.asarray(w, np.float64)[0]                 # [64][3][3][3]
    m = np.zeros((2, 3, 3))                          # [parity][low-res tap][original tap]
    m[0, 0, 0] = 1; m[0, 1, 1] = 1; m[0, 1, 2] = 1
    m[1, 1, 0] = 1; m[1, 1, 1] = 1; m[1, 2, 2] = 1
    out = np.zeros((27, 1, 16, 64), np.float64)
    for p in range(8):
        a, b, c = p >> 2, (p >> 1) & 1, p & 1
        f = np.einsum('cxyz,ix,jy,kz->ijkc', w, m[a], m[b], m[c])   # [3][3][3][64]
        out[:, 0, p, :] = f.reshape(27, 64)
    return out.astype(np.float16)


def seeded_weights(seed=0, gain=2.5):
    """seeded weights with the flylight decoder's layer shapes (no checkpoint ships with
    the reference; timing and plumbing tests only)."""
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


class PatchDecoder:
    """weights: dict with the keys of oracle.decoder_torch.make_weights /
    the decoder part of the reference checkpoint (`model.decoder`)."""

    def __init__(self, weights, device=None):
        import torch
        self.dev = torch.device(device or 'cuda')
        W = {k: np.asarray(v, np.float32) for k, v in weights.items()}
        assert W['from_code.w'].shape[:2] == (128, 22) and W['up0.w'].shape[:2] == (64, 128), \
            "only the flylight decoder geometry (22x2^3 -> 128 -> 64 -> 1, 7^3) is built"
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.w_fc = f(W['from_code.w'].reshape(128, 22))
        self.b_fc = f(W['from_code.b'])
        self.w_up0 = f(_pack_conv(W['up0.w']))
        self.b_up0 = f(W['up0.b'])
        self.w_c0a = f(_pack_conv(W['conv0a.w']))
        self.b_c0a = f(W['conv0a.b'])
        self.w_c0b = f(_pack_conv(W['conv0b.w']))
        self.b_c0b = f(W['conv0b.b'])
        self.w_up1 = f(_fold_upsample_conv(W['up1.w']))
        self.b_up1 = f(W['up1.b'].reshape(1))
        self.w_c1a = f(W['conv1a.w'].reshape(27))
        self.b_c1a = f(W['conv1a.b'].reshape(1))
        self.w_c1b = f(W['conv1b.w'].reshape(27))
        self.b_c1b = f(W['conv1b.b'].reshape(1))

    def decode(self, codes, sigmoid=False):
        """codes [B,176] (cuda, any float dtype) -> patches f32 [B,343] (cuda)."""
        import torch
        codes = codes.to(self.dev, torch.float32).contiguous()
        B = int(codes.shape[0])
        out = torch.empty((max(B, 1), 343), dtype=torch.float32, device=self.dev)
        scratch = torch.empty(cc.call('ppp_decode_scratch_bytes', B), dtype=torch.uint8,
                              device=self.dev)
        cc.call('ppp_decode', cc.ptr(codes), B, cc.ptr(self.w_fc), cc.ptr(self.b_fc),
                cc.ptr(self.w_up0), cc.ptr(self.b_up0), cc.ptr(self.w_c0a), cc.ptr(self.b_c0a),
                cc.ptr(self.w_c0b), cc.ptr(self.b_c0b), cc.ptr(self.w_up1), cc.ptr(self.b_up1),
                cc.ptr(self.w_c1a), cc.ptr(self.b_c1a), cc.ptr(self.w_c1b), cc.ptr(self.b_c1b),
                1 if sigmoid else 0, cc.ptr(out), cc.ptr(scratch), cc.current_stream_ptr())
        return out[:B]

    def decode_rows(self, codes, batch=1 << 16):
        """the codes of the stored voxels [G,176] -> compact patch ROWS f16 [G,343]
        (sigmoid applied), the input form of the rows path (assembly.RowSource,
        sharded.RowShard): decode.py:39-65 scatters the decoded patches into a dense
        [P,Z,Y,X] array and stores it as float16; here the rows stay rows and the
        dense array never exists.  The reference stores float16 LOGITS and applies
        expit after loading (utilVoteInstances.py:249-250); storing the float16
        PROBABILITY differs from that by at most 2.5e-4 (half an ulp below 1), inside
        the 1e-3 tolerance stated for decoded patches."""
        import torch
        G = int(codes.shape[0])
        out = torch.empty((G, 343), dtype=torch.float16, device=self.dev)
        for s0 in range(0, G, batch):
            out[s0:s0 + batch] = self.decode(codes[s0:s0 + batch], sigmoid=True)
        return out

    def decode_volume(self, pred_code, fg, sigmoid=True):
        """decode_sample (decode.py:16-65): codes [176,Z,Y,X] + fg mask ->
        dense patches f32 [343,Z,Y,X] (zeros outside the foreground)."""
        import torch
        pred_code = pred_code.to(self.dev)
        fg = fg.to(self.dev).bool()
        idx = torch.nonzero(fg.reshape(-1)).flatten()
        codes = pred_code.reshape(pred_code.shape[0], -1)[:, idx].T
        patches = self.decode(codes, sigmoid=sigmoid)
        out = torch.zeros((343, fg.numel()), dtype=torch.float32, device=self.dev)
        out[:, idx] = patches.T
        return out.reshape((343,) + tuple(fg.shape))
