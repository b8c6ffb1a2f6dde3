"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference on the CPU.

SURVEY.md §8c "Route 2b": the reference's python host code
(/root/reference/PatchPerPix/vote_instances/*.py) is imported where it lies
under a synthetic package name with stub modules for its absent dependencies,
and the six functions of its device boundary (cuda_code.py:5-59) are replaced
by a fake that compiles the ALREADY TEMPLATED kernel source the reference
hands to `make_kernel` with g++ through oracle/ref_shim/ and loops the launch
grid serially.  Nothing is copied from the reference; this module only works
where /root/reference exists (this container, not the GPU box) and is used by
tools/gen_golden.py to write tests/golden/*.npz, and by tests that pin the
oracle restatement against the live reference when it is present.

`build_ref_kernel` additionally builds a reference kernel straight from its
.cu file with -D macros (as the reference's cuda/Makefile does) into
oracle/_ref/, which DOES travel to the GPU box and serves as the
"reference" CPU baseline (bench.py --impl reference).
"""
import contextlib
import ctypes
import hashlib
import importlib
import os
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('PPP_REFERENCE_ROOT', '/root/reference')
REF_VI = os.path.join(REF_ROOT, 'PatchPerPix', 'vote_instances')
REF_OUT = os.path.join(HERE, '_ref')
SHIM = os.path.join(HERE, 'ref_shim')


def reference_available():
    return os.path.isdir(REF_VI)


# ----------------------------------------------------------------------------
# compiling reference kernels for the host
# ----------------------------------------------------------------------------
def _kind_of(code):
    if 'fillConsensusArray_allPatches' in code:
        return 'K_FILL'
    if 'normConsensusArray' in code:
        return 'K_NORM'
    if 'rankPatches' in code:
        return 'K_RANK'
    if 'computePatchGraph' in code:
        return 'K_GRAPH'
    raise ValueError('unknown kernel source')


def _compile(tu_text, so_path, defines, omp=False):
    os.makedirs(os.path.dirname(so_path), exist_ok=True)
    src = so_path[:-3] + '.cpp'
    with open(src, 'w') as f:
        f.write(tu_text)
    cmd = ['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-w',
           '-I', SHIM] + list(defines)
    if omp:
        cmd += ['-fopenmp', '-DPPP_SHIM_ATOMIC']
    cmd += [src, '-o', so_path]
    subprocess.run(cmd, check=True)
    os.remove(src)


class _Kernel:
    """callable mimicking pycuda's Function: kernel(*args, block=, grid=)."""

    def __init__(self, so_path):
        self.lib = ctypes.CDLL(so_path)
        self.fn = self.lib.ppp_ref_launch
        self.fn.restype = None
        self.on_launch = None

    def __call__(self, *args, block=None, grid=None):
        keep = []
        ptrs = (ctypes.c_void_p * len(args))()
        for i, a in enumerate(args):
            if isinstance(a, np.ndarray):
                assert a.flags['C_CONTIGUOUS']
                ptrs[i] = a.ctypes.data
            else:   # numpy scalar
                arr = np.array([a])
                keep.append(arr)
                ptrs[i] = arr.ctypes.data
        grid = tuple(int(g) for g in grid) + (1,) * (3 - len(grid))
        block = tuple(int(b) for b in block) + (1,) * (3 - len(block))
        self.fn(ptrs, ctypes.c_uint(grid[0]), ctypes.c_uint(grid[1]),
                ctypes.c_uint(grid[2]), ctypes.c_uint(block[0]),
                ctypes.c_uint(block[1]), ctypes.c_uint(block[2]))
        if self.on_launch is not None:
            self.on_launch(args)


def build_ref_kernel(kind, dims, patchshape, th, flags=(), omp=False):
    """compile a reference .cu file (read in place) with -D shape macros.

    kind in {'fill','norm','rank','graph'}; dims = (Z,Y,X).  Returns the path
    of oracle/_ref/refk_*.so (built if /root/reference is present, else the
    prebuilt file must already exist)."""
    fn = {'fill': 'fillConsensusArray.cu', 'norm': 'normConsensusArray.cu',
          'rank': 'rankPatches.cu', 'graph': 'computePatchGraph.cu'}[kind]
    ps = [int(p) for p in patchshape]
    ns = [2 * p if (ps[0] > 1 or i > 0) else p for i, p in enumerate(ps)]
    thi = th if th < 0.5 else 1.0 - th
    tag = '%s_%dx%dx%d_%dx%dx%d_%s_%s%s' % (
        kind, dims[0], dims[1], dims[2], ps[0], ps[1], ps[2], repr(th),
        '-'.join(sorted(f.replace('-D', '') for f in flags)) or 'none',
        '_omp' if omp else '')
    so = os.path.join(REF_OUT, 'refk_' + tag + '.so')
    if os.path.exists(so):
        return so
    if not reference_available():
        raise FileNotFoundError(
            '%s not prebuilt and %s is absent' % (so, REF_ROOT))
    defs = ['-DDATAZSIZE=%d' % dims[0], '-DDATAYSIZE=%d' % dims[1],
            '-DDATAXSIZE=%d' % dims[2],
            '-DPSZ=%d' % ps[0], '-DPSY=%d' % ps[1], '-DPSX=%d' % ps[2],
            '-DNSZ=%d' % ns[0], '-DNSY=%d' % ns[1], '-DNSX=%d' % ns[2],
            '-DTH=%s' % repr(th), '-DTHI=%s' % repr(thi),
            '-DL_Z=%d' % dims[0], '-DL_Y=%d' % dims[1], '-DL_X=%d' % dims[2],
            '-DL_NSY=%d' % ns[1], '-DL_NSX=%d' % ns[2],
            '-D' + {'fill': 'K_FILL', 'norm': 'K_NORM', 'rank': 'K_RANK',
                    'graph': 'K_GRAPH'}[kind]] + list(flags)
    tu = '#include "prelude.h"\n#include "%s"\n#include "launchers.inc"\n' % \
        os.path.join(REF_VI, 'cuda', fn)
    _compile(tu, so, defs, omp=omp)
    return so


def build_ref_kernel_cuda(kind, dims, patchshape, th, flags=()):
    """the same reference .cu file compiled by nvcc for sm_100a (unmodified, read in
    place) with a launcher that uses the reference's launch geometry: the "existing GPU
    implementation" timed beside ours (bench.py --impl refgpu).  Returns the path of
    oracle/_ref/refk_cuda_*.so (prebuilt here, travels to the GPU box)."""
    fn = {'fill': 'fillConsensusArray.cu', 'norm': 'normConsensusArray.cu',
          'rank': 'rankPatches.cu', 'graph': 'computePatchGraph.cu'}[kind]
    ps = [int(p) for p in patchshape]
    ns = [2 * p if (ps[0] > 1 or i > 0) else p for i, p in enumerate(ps)]
    thi = th if th < 0.5 else 1.0 - th
    tag = '%s_%dx%dx%d_%dx%dx%d_%s_%s' % (
        kind, dims[0], dims[1], dims[2], ps[0], ps[1], ps[2], repr(th),
        '-'.join(sorted(f.replace('-D', '') for f in flags)) or 'none')
    so = os.path.join(REF_OUT, 'refk_cuda_' + tag + '.so')
    if os.path.exists(so):
        return so
    if not reference_available():
        raise FileNotFoundError('%s not prebuilt and %s is absent' % (so, REF_ROOT))
    defs = ['-DDATAZSIZE=%d' % dims[0], '-DDATAYSIZE=%d' % dims[1],
            '-DDATAXSIZE=%d' % dims[2],
            '-DPSZ=%d' % ps[0], '-DPSY=%d' % ps[1], '-DPSX=%d' % ps[2],
            '-DNSZ=%d' % ns[0], '-DNSY=%d' % ns[1], '-DNSX=%d' % ns[2],
            '-DTH=%s' % repr(th), '-DTHI=%s' % repr(thi),
            '-DL_Z=%d' % dims[0], '-DL_Y=%d' % dims[1], '-DL_X=%d' % dims[2],
            '-DL_NSY=%d' % ns[1], '-DL_NSX=%d' % ns[2],
            '-D' + {'fill': 'K_FILL', 'norm': 'K_NORM', 'rank': 'K_RANK',
                    'graph': 'K_GRAPH'}[kind]] + list(flags)
    os.makedirs(REF_OUT, exist_ok=True)
    src = so[:-3] + '.cu'
    with open(src, 'w') as f:
        f.write('#include <cstdint>\n#include "%s"\n#include "launchers_cuda.inc"\n' %
                os.path.join(REF_VI, 'cuda', fn))
    cmd = ['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3',
           '-shared', '-Xcompiler', '-fPIC', '-w', '-I', SHIM] + defs + [src, '-o', so]
    subprocess.run(cmd, check=True)
    os.remove(src)
    return so


def load_ref_kernel_cuda(so_path):
    """callable(*device pointers / numpy scalars, block=, grid=) -> cudaError_t"""
    lib = ctypes.CDLL(so_path)
    fn = lib.ppp_ref_launch
    fn.restype = ctypes.c_int

    def launch(*args, block=None, grid=None):
        keep = []
        ptrs = (ctypes.c_void_p * len(args))()
        for i, a in enumerate(args):
            if isinstance(a, int):
                ptrs[i] = a
            else:
                arr = np.array([a])
                keep.append(arr)
                ptrs[i] = arr.ctypes.data
        grid = tuple(int(g) for g in grid) + (1,) * (3 - len(grid))
        block = tuple(int(b) for b in block) + (1,) * (3 - len(block))
        rc = fn(ptrs, *[ctypes.c_uint(v) for v in grid + block])
        if rc != 0:
            raise RuntimeError('reference kernel failed: cudaError %d' % rc)
    return launch


def load_ref_kernel(so_path):
    return _Kernel(so_path)


# ----------------------------------------------------------------------------
# fake device boundary for the reference's python host code
# ----------------------------------------------------------------------------
class _FakeModule:
    def __init__(self, so_path, recorder, kind, options):
        self.k = _Kernel(so_path)
        if recorder is not None:
            self.k.on_launch = lambda args: recorder(kind, options, args)

    def get_function(self, name):
        return self.k


class _Base:
    def free(self):
        pass


class _Managed(np.ndarray):
    """np.zeros view whose `.base.free()` exists (consensus_array.py:193)."""
    @property
    def base(self):
        return _Base()


class RefSession:
    """imports the reference package with stubs; exposes its modules."""

    def __init__(self, cache_dir=None, recorder=None):
        assert reference_available(), REF_ROOT + ' missing'
        self.cache = cache_dir or os.path.join('/tmp', 'ppp_r2b_cache')
        self.recorder = recorder
        os.makedirs(self.cache, exist_ok=True)
        self._install_stubs()
        self._load()

    # -- stubs ---------------------------------------------------------------
    def _install_stubs(self):
        if not hasattr(np, 'bool'):
            np.bool = bool
        if not hasattr(np, 'product'):
            np.product = np.prod

        def stub(name, **attrs):
            if name in sys.modules and not getattr(
                    sys.modules[name], '_ppp_stub', False):
                return sys.modules[name]
            m = types.ModuleType(name)
            m._ppp_stub = True
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
            return m

        def _na(*a, **k):
            raise RuntimeError('stubbed dependency called')

        stub('h5py', File=_na)
        stub('zarr', open=_na)
        sk = stub('skimage')
        sk.io = stub('skimage.io', imsave=_na)
        sk.morphology = stub('skimage.morphology', skeletonize_3d=_na,
                             binary_dilation=_na, ball=_na)
        sk.draw = stub('skimage.draw', line=_na)
        pc = stub('pycuda')
        pc.compiler = stub('pycuda.compiler')
        pc.driver = stub('pycuda.driver')
        class _Blosc:
            BITSHUFFLE = 2

            def __init__(self, **k):
                pass
        stub('numcodecs', Blosc=_Blosc)
        stub('colorcet')
        stub('nrrd')

    # -- load reference modules under a synthetic package ---------------------
    def _load(self):
        root = types.ModuleType('ppp_ref')
        root.__path__ = []
        sys.modules['ppp_ref'] = root
        util = types.ModuleType('ppp_ref.util')
        util.remove_small_components = lambda *a, **k: a[0]
        util.relabel = lambda *a, **k: a[0]
        util.color = lambda *a, **k: a[0]
        sys.modules['ppp_ref.util'] = util
        root.util = util
        pkg = types.ModuleType('ppp_ref.vote_instances')
        pkg.__path__ = [REF_VI]
        pkg.__package__ = 'ppp_ref.vote_instances'
        sys.modules['ppp_ref.vote_instances'] = pkg
        root.vote_instances = pkg
        names = ['cuda_code', 'get_patch_sets', 'utilVoteInstances',
                 'consensus_array', 'ranked_patches', 'foreground_cover',
                 'graph_mws', 'aff_patch_graph', 'graph_to_labeling',
                 'isbi_hacks', 'vote_instances']
        self.mods = {}
        for n in names:
            self.mods[n] = importlib.import_module(
                'ppp_ref.vote_instances.' + n)
        fakes = dict(make_kernel=self.make_kernel,
                     alloc_zero_array=self.alloc_zero_array,
                     sync=lambda ctx: None, init_cuda=lambda: None,
                     delete_cuda=lambda ctx: None)
        for m in self.mods.values():
            for k, v in fakes.items():
                if hasattr(m, k):
                    setattr(m, k, v)
        self.vi = self.mods['vote_instances']

    # -- the fake boundary -----------------------------------------------------
    def make_kernel(self, code, options=None):
        options = list(options or [])
        kind = _kind_of(code)
        # recover the literal sizes from the templated parameter list
        import re
        m = re.search(r'inPred\[\]\[(\d+)\]\[(\d+)\]\[(\d+)\]', code)
        Z, Y, X = (int(g) for g in m.groups())
        m = re.search(r'\[\]\[(\d+)\]\[(\d+)\]\[%d\]\[%d\]\[%d\]' % (Z, Y, X),
                      code)
        nsy, nsx = (int(g) for g in m.groups()) if m else (1, 1)
        h = hashlib.sha1((code + '|' + ' '.join(options)).encode()).hexdigest()
        so = os.path.join(self.cache, 'r2b_%s.so' % h[:20])
        if not os.path.exists(so):
            defs = ['-DL_Z=%d' % Z, '-DL_Y=%d' % Y, '-DL_X=%d' % X,
                    '-DL_NSY=%d' % nsy, '-DL_NSX=%d' % nsx, '-D' + kind]
            tu = '#include "prelude.h"\n' + code + '\n#include "launchers.inc"\n'
            _compile(tu, so, defs + options)
        return _FakeModule(so, self.recorder, kind, options)

    @staticmethod
    def alloc_zero_array(shape, dtype):
        if np.isscalar(shape):
            shape = (int(shape),)
        return np.zeros(shape, dtype=dtype).view(_Managed)

    # -- convenience -------------------------------------------------------------
    def default_kwargs(self, **over):
        """flylight [vote_instances] defaults (default.toml:114-169) for the
        array-level entry point do_block (vote_instances.py:455)."""
        kw = dict(
            patch_threshold=0.5, fc_threshold=0.5, cuda=True, blockwise=False,
            debug=False, select_patches_for_sparse_data=True,
            save_no_intermediates=True, includeSinglePatchCCS=True,
            one_instance_per_channel=False, sample=1.0,
            removeIntersection=True, mws=False, isbiHack=False,
            mask_fg_border=False, graphToInst=False, skipLookup=False,
            skipConsensus=False, skipRanking=False, skipThinCover=False,
            termAfterThinCover=False, pad_with_ps=False,
            consensus_interleaved_cnt=False, consensus_norm_prob_product=True,
            consensus_prob_product=True, consensus_norm_aff=True,
            vi_bg_use_inv_th=False, vi_bg_use_half_th=False,
            vi_bg_use_less_than_th=True, rank_norm_patch_score=True,
            rank_int_counter=False, patch_graph_norm_aff=True,
            flip_cons_arr_axes=False, overlapping_inst=True,
            num_parallel_samples=1, num_parallel_blocks=1,
            mutex=contextlib.nullcontext(), context=None,
            affinities='synthetic.npy', result_folder='/tmp',
        )
        kw.update(over)
        return kw


# ----------------------------------------------------------------------------
# in-memory stand-ins for zarr / h5py so that the reference's blockwise driver
# (stitch_patch_graph.py) can run unmodified without the real packages
# ----------------------------------------------------------------------------
class _Arr(np.ndarray):
    """ndarray with an `.attrs` dict (zarr / h5py datasets have one)."""
    def __new__(cls, a):
        obj = np.asarray(a).view(cls)
        obj.attrs = {}
        return obj

    def __array_finalize__(self, obj):
        self.attrs = getattr(obj, 'attrs', {})


class FakeGroup:
    _stores = {}

    def __init__(self):
        self.d = {}

    @classmethod
    def open(cls, path, mode='r', **kw):
        path = os.path.abspath(path)
        if mode == 'w' or path not in cls._stores:
            cls._stores[path] = FakeGroup()
            if mode in ('w', 'a') and not os.path.exists(path):
                os.makedirs(path)          # the reference tests os.path.exists(res_file)
        return cls._stores[path]

    def _split(self, key):
        return [k for k in key.split('/') if k]

    def __contains__(self, key):
        return any(k == '/'.join(self._split(key)) or
                   k.startswith('/'.join(self._split(key)) + '/') for k in self.d)

    def __getitem__(self, key):
        key = '/'.join(self._split(key))
        if key in self.d:
            return self.d[key]
        sub = FakeGroup()
        sub.d = {k[len(key) + 1:]: v for k, v in self.d.items() if k.startswith(key + '/')}
        if not sub.d:
            raise KeyError(key)
        return sub

    def __setitem__(self, key, value):
        self.d['/'.join(self._split(key))] = _Arr(value)

    def keys(self):
        return sorted({k.split('/')[0] for k in self.d})

    def create_dataset(self, name, data=None, shape=None, dtype=None, **kw):
        arr = np.array(data, dtype=dtype) if data is not None else np.zeros(shape, dtype)
        self[name] = arr
        return self['/'.join(self._split(name))]

    create = create_dataset

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def install_fake_io():
    """replace the zarr / h5py stubs by the in-memory stores above."""
    z = sys.modules['zarr']
    z.open = FakeGroup.open
    h = sys.modules['h5py']
    h.File = lambda path, mode='r', **kw: FakeGroup.open(path, mode)
    pc = sys.modules['pycuda']
    auto = types.ModuleType('pycuda.autoinit')
    auto.context = None
    auto._ppp_stub = True
    sys.modules['pycuda.autoinit'] = auto
    pc.autoinit = auto


def load_stitch_module(session):
    """import the reference's stitch_patch_graph.py under the synthetic package
    (needs the fake IO) and patch its device boundary like the other modules."""
    install_fake_io()
    for n in ('io_hdflike', 'stitch_patch_graph'):
        session.mods[n] = importlib.import_module('ppp_ref.vote_instances.' + n)
    return session.mods['stitch_patch_graph']
