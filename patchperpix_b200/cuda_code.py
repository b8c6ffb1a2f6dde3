"""Device boundary of the B200 build — the counterpart of the reference's
pycuda wrapper (PatchPerPix/vote_instances/cuda_code.py:5-59).

Same function names where the concept survives (`alloc_zero_array`, `sync`,
`init_cuda`, `delete_cuda`, `get_cuda_stream`); `make_kernel` is gone because
nothing is JIT-compiled: the kernels live in libppp_b200.so (C ABI declared in
include/ppp_b200.h, built for sm_100a by patchperpix_b200/build.py) and the
reference's -D variant flags (utilVoteInstances.py:389-449) become fields of
`ppp_cfg` (see `make_cfg`).  torch is used for device memory and streams only.

There is NO CPU fallback: importing is harmless, but any call raises if the
library or a CUDA device is missing.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'libppp_b200.so')


class PppCfg(ctypes.Structure):
    """mirror of `struct ppp_cfg` (include/ppp_b200.h)."""
    _fields_ = [
        ('Z', ctypes.c_int32), ('Y', ctypes.c_int32), ('X', ctypes.c_int32),
        ('psz', ctypes.c_int32), ('psy', ctypes.c_int32), ('psx', ctypes.c_int32),
        ('th_gt', ctypes.c_float), ('bg_lt', ctypes.c_float),
        ('fc_gt', ctypes.c_float), ('pt_gt', ctypes.c_float),
        ('th2', ctypes.c_double), ('one_m_th2', ctypes.c_double),
        ('prod_mode', ctypes.c_int32), ('norm_aff', ctypes.c_int32),
        ('use_overlap', ctypes.c_int32), ('rank_flags', ctypes.c_int32),
        ('graph_flags', ctypes.c_int32), ('reserved', ctypes.c_int32),
    ]


def _f32_floor(x):
    """largest float32 <= x:  (float)v > x  <=>  v > _f32_floor(x)."""
    f = np.float32(x)
    if float(f) > x:
        f = np.nextafter(f, np.float32(-np.inf))
    return float(f)


def _f32_ceil(x):
    """smallest float32 >= x:  (float)v < x  <=>  v < _f32_ceil(x)."""
    f = np.float32(x)
    if float(f) < x:
        f = np.nextafter(f, np.float32(np.inf))
    return float(f)


def make_cfg(shape, patchshape, **kwargs):
    """kwargs of the reference's `[vote_instances]` surface -> ppp_cfg.

    Follows setKernelBuildOptions (utilVoteInstances.py:389-449), the literal
    substitution of TH / THI (utilVoteInstances.py:361-365: the kernels compare
    floats against DOUBLE literals) and consensus_array.py:83-88,171."""
    th = float(kwargs['patch_threshold'])
    if kwargs.get('vi_bg_use_inv_th', True):
        if th < 0.5:
            bg = th                       # falls back to USE_LESS_THAN_TH (:396-398)
        else:
            bg = 1.0 - th                 # USE_INV_TH with THI = 1.0 - th
    elif kwargs.get('vi_bg_use_half_th', False):
        bg = th / 2
    elif kwargs.get('vi_bg_use_less_than_th', False):
        bg = th
    else:
        raise RuntimeError("how is bg defined for vote instances?")
    if kwargs.get('consensus_norm_prob_product', True):
        prod = 2
    elif kwargs.get('consensus_prob_product', True):
        prod = 1
    else:
        assert not kwargs.get('consensus_norm_aff', True) and \
            not kwargs.get('consensus_interleaved_cnt', True), \
            "no normalizing for accumulate consensus counter available"
        prod = 0
    c = PppCfg()
    c.Z, c.Y, c.X = (int(s) for s in shape)
    c.psz, c.psy, c.psx = (int(p) for p in patchshape)
    assert c.psz % 2 == 1 and c.psy % 2 == 1 and c.psx % 2 == 1, \
        "patchshape must be odd (the centre channel is P // 2)"
    c.th_gt = _f32_floor(th)
    c.bg_lt = _f32_ceil(bg)
    c.fc_gt = float(np.float32(kwargs.get('fc_threshold', 0.5)))
    c.pt_gt = float(np.float32(th))
    c.th2 = th * th
    c.one_m_th2 = 1.0 - th * th
    c.prod_mode = prod
    c.norm_aff = 1 if kwargs.get('consensus_norm_aff', True) else 0
    c.use_overlap = 1 if kwargs.get('overlapping_inst', False) else 0
    c.rank_flags = (1 if kwargs.get('rank_norm_patch_score', True) else 0) | \
                   (2 if kwargs.get('rank_int_counter', False) else 0) | \
                   (4 if kwargs.get('ppp_rank_fast', False) else 0)
    c.graph_flags = (1 if kwargs.get('patch_graph_norm_aff', True) else 0) | \
                    (4 if kwargs.get('ppp_graph_fast', False) else 0)
    c.reserved = int(kwargs.get('ppp_tune', 0))
    return c


_lib = None
_SIGS = {
    'ppp_last_error': (ctypes.c_char_p, []),
    'ppp_version': (ctypes.c_int, []),
    'ppp_launch_count': (ctypes.c_int64, []),
    'ppp_gate': (ctypes.c_int, ['p', 'p', 'p', 'cfg', 'p', 'p']),
    'ppp_compact_scratch_bytes': (ctypes.c_int64, ['i64']),
    'ppp_compact': (ctypes.c_int, ['p', 'i64', 'p', 'p', 'p', 'p', 'p']),
    'ppp_prepare_patches': (ctypes.c_int, ['p', 'p', 'p', 'i64', 'cfg', 'p', 'p', 'p', 'p', 'p']),
    'ppp_consensus_scratch_bytes': (ctypes.c_int64, ['cfg']),
    'ppp_consensus': (ctypes.c_int, ['p', 'p', 'p', 'p', 'p', 'i64', 'cfg', 'p', 'p', 'i32', 'p', 'p']),
    'ppp_rank_scratch_bytes': (ctypes.c_int64, ['cfg', 'i64']),
    'ppp_rank': (ctypes.c_int, ['p', 'p', 'p', 'p', 'i64', 'p', 'cfg', 'p', 'p', 'p']),
    'ppp_rank_sort_scratch_bytes': (ctypes.c_int64, ['i64']),
    'ppp_rank_sort': (ctypes.c_int, ['p', 'p', 'i64', 'p', 'p', 'p']),
    'ppp_cover_scratch_bytes': (ctypes.c_int64, ['cfg']),
    'ppp_cover': (ctypes.c_int, ['p', 'p', 'p', 'i64', 'p', 'p', 'cfg', 'p', 'i32', 'p', 'p', 'p']),
    'ppp_thin_scratch_bytes': (ctypes.c_int64, ['cfg', 'i64']),
    'ppp_thin': (ctypes.c_int, ['p', 'p', 'i64', 'p', 'p', 'cfg', 'p', 'p', 'p']),
    'ppp_patch_graph_scratch_bytes': (ctypes.c_int64, ['cfg', 'i64']),
    'ppp_patch_graph': (ctypes.c_int, ['p', 'p', 'p', 'p', 'p', 'i64', 'cfg', 'p', 'p', 'p']),
    'ppp_label_scratch_bytes': (ctypes.c_int64, ['i64', 'i64']),
    'ppp_label_cc': (ctypes.c_int, ['p', 'p', 'i64', 'cfg', 'p', 'p', 'p', 'p']),
    'ppp_mws_host': (ctypes.c_int, ['p', 'p', 'i64', 'cfg', 'p', 'p', 'p', 'p']),
    'ppp_pyset_order': (ctypes.c_int, ['p', 'i64', 'p']),
    'ppp_pyset_pairs': (ctypes.c_int, ['p', 'i64', 'p', 'p', 'p', 'p']),
    'ppp_paint': (ctypes.c_int, ['p', 'p', 'i64', 'p', 'cfg', 'p', 'p']),
    'ppp_paint_channels': (ctypes.c_int, ['p', 'p', 'i64', 'p', 'cfg', 'p', 'p']),
    'ppp_paint_patches': (ctypes.c_int, ['p', 'p', 'i64', 'p', 'cfg', 'p', 'p']),
    'ppp_received_row_words': (ctypes.c_int64, ['cfg']),
    'ppp_received': (ctypes.c_int, ['p', 'p', 'p', 'p', 'i64', 'cfg', 'p', 'p', 'p']),
    'ppp_received_rows': (ctypes.c_int, ['p', 'p', 'p', 'p', 'p', 'i64', 'cfg', 'p', 'p', 'p']),
    'ppp_consensus_small': (ctypes.c_int, ['p', 'p', 'p', 'p', 'p', 'p', 'i64', 'cfg', 'p', 'p', 'p']),
    'ppp_mark_windows': (ctypes.c_int, ['p', 'i64', 'cfg', 'i32', 'i32', 'i32', 'p', 'p', 'p']),
    'ppp_gate_rows': (ctypes.c_int, ['p', 'p', 'p', 'p', 'cfg', 'p', 'p']),
    'ppp_prepare_rows': (ctypes.c_int, ['p', 'p', 'p', 'p', 'i64', 'cfg', 'p', 'p', 'p', 'p', 'p']),
    'ppp_patch_graph_rows': (ctypes.c_int, ['p', 'p', 'p', 'p', 'p', 'p', 'p', 'i64', 'cfg', 'p', 'p', 'p']),
    'ppp_paint_rows': (ctypes.c_int, ['p', 'p', 'p', 'p', 'i64', 'cfg', 'p', 'p']),
    'ppp_decode_scratch_bytes': (ctypes.c_int64, ['i64']),
    'ppp_decode': (ctypes.c_int, ['p', 'i64'] + ['p'] * 14 + ['i32', 'p', 'p', 'p']),
}
_CT = {'p': ctypes.c_void_p, 'i64': ctypes.c_int64, 'i32': ctypes.c_int32,
       'cfg': ctypes.POINTER(PppCfg)}


def exported_symbols():
    """every entry point include/ppp_b200.h declares."""
    return sorted(_SIGS)


def load_library():
    """dlopen libppp_b200.so and type its entry points (no GPU needed)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise RuntimeError(
                "%s is missing: build it with `python -m patchperpix_b200.build` "
                "(there is no CPU fallback)" % SO)
        lib = ctypes.CDLL(SO)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = [_CT[a] for a in args]
        _lib = lib
    return _lib


class PppError(RuntimeError):
    pass


# set by a profiler (bench.py) to a callable(name, value): lets a stage report how many
# units a launch really processes when that is only known on the device
profile_hook = None


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return ctypes.c_void_p(t.data_ptr())


def call(name, *args):
    """invoke a C-ABI entry point; raise PppError on a non-zero return code."""
    lib = load_library()
    conv = []
    for a in args:
        if isinstance(a, PppCfg):
            conv.append(ctypes.byref(a))
        else:
            conv.append(a)
    rc = getattr(lib, name)(*conv)
    if _SIGS[name][0] is ctypes.c_int and rc != 0:
        raise PppError('%s failed (%d): %s' % (
            name, rc, lib.ppp_last_error().decode()))
    return rc


# ---------------------------------------------------------------------------
# the reference's cuda_code.py surface
# ---------------------------------------------------------------------------
def init_cuda():
    """cuda_code.py:17-48.  Returns the torch device used as "context"."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("patchperpix_b200 needs a CUDA device (no CPU fallback)")
    load_library()
    dev = torch.device('cuda', torch.cuda.current_device())
    torch.cuda.init()
    return dev


def delete_cuda(context):
    """cuda_code.py:50-55 (nothing to pop: torch owns the primary context)."""
    return None


def sync(context=None):
    """cuda_code.py:14-15."""
    import torch
    torch.cuda.synchronize(context)


def get_cuda_stream():
    """cuda_code.py:57-59."""
    import torch
    return torch.cuda.Stream()


def current_stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_TORCH_DT = None


def alloc_zero_array(shape, dtype, device=None):
    """cuda_code.py:9-11 allocates zeroed MANAGED memory; here plain device
    memory owned by torch (freed when the tensor dies, not leaked)."""
    import torch
    global _TORCH_DT
    if _TORCH_DT is None:
        _TORCH_DT = {np.dtype(np.float32): torch.float32, np.dtype(np.uint8): torch.uint8,
                     np.dtype(np.bool_): torch.uint8, np.dtype(np.int32): torch.int32,
                     np.dtype(np.uint32): torch.int32, np.dtype(np.int64): torch.int64,
                     np.dtype(np.uint16): torch.int16}
    if np.isscalar(shape):
        shape = (int(shape),)
    return torch.zeros(tuple(int(s) for s in shape), dtype=_TORCH_DT[np.dtype(dtype)],
                       device=device or 'cuda')
