"""Builds libppp_b200.so (the C-ABI library, include/ppp_b200.h) in-tree with
nvcc for sm_100a.  No torch involved: the library only needs the CUDA runtime."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, '_build')
SO = os.path.join(HERE, 'libppp_b200.so')
SOURCES = ['ppp_api.cu', 'ppp_prep.cu', 'ppp_consensus.cu', 'ppp_rank.cu',
           'ppp_cover.cu', 'ppp_graph.cu', 'ppp_mws.cu', 'ppp_decoder.cu']
EXTRA_DEFS = os.environ.get('PPP_EXTRA_DEFS', '').split()   # tuning experiments only
NVCC_FLAGS = EXTRA_DEFS + ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
              '-std=c++17', '-Xcompiler', '-fPIC', '--use_fast_math=false',
              '-Xptxas', '-v']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    return 'nvcc'


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math')]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'ppp_b200.h'))
    objs = []
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(BUILD, src[:-3] + '.o')
        objs.append(op)
        if force or _newer(sp, op) or any(_newer(d, op) for d in deps):
            cmd = [_nvcc()] + flags + ['-c', sp, '-o', op]
            procs.append((src, subprocess.Popen(
                cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    relink = force or bool(procs) or not os.path.exists(SO)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed for ' + src)
        with open(os.path.join(BUILD, src[:-3] + '.ptxas.log'), 'w') as f:
            f.write(out)
    if relink:
        # shared runtime + -Bsymbolic: the launch counter in ppp_api.cu interposes
        # cudaLaunchKernel for this library only
        cmd = [_nvcc(), '-shared', '--cudart', 'shared', '-Xlinker', '-Bsymbolic', '-gencode',
               'arch=compute_100a,code=sm_100a', '-o', SO] + objs + ['-lcudart', '-ldl']
        subprocess.run(cmd, check=True)
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
