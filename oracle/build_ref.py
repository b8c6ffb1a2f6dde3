"""TEST INFRASTRUCTURE ONLY — prebuilds the reference kernels for the host
(oracle/_ref/refk_*.so) from the sources where they lie under /root/reference,
for the shapes bench.py's `--impl reference` / cpu_baseline legs use.
oracle/_ref/ is git-ignored but travels to the GPU box with gpurun."""
from . import ref_runner

# (kind, dims, patchshape, th, flags, omp)
BASE = ['-DUSE_LESS_THAN_TH', '-DOVERLAP']
CONFIGS = []
for dims, ps in (((1, 160, 160), (1, 41, 41)), ((1, 256, 256), (1, 41, 41)), ((24, 48, 48), (7, 7, 7)),
                 ((72, 72, 72), (7, 7, 7))):
    for omp in (False, True):
        CONFIGS += [
            ('fill', dims, ps, 0.5, BASE + ['-DNORM_PROB_PRODUCT'], omp),
            ('fill', dims, ps, 0.5, BASE + ['-DNORM_PROB_PRODUCT', '-DOUTPUT_CNT'], omp),
            ('norm', dims, ps, 0.5, [], omp),
            ('rank', dims, ps, 0.5, BASE + ['-DNORM_PATCH_RANK'], omp),
            ('graph', dims, ps, 0.5, ['-DNORM_PATCH_AFFINITY'], omp),
        ]


# the same kernels for sm_100a (bench.py --impl refgpu), on the sample both arms share
CUDA_CONFIGS = [c[:5] for c in CONFIGS if c[1] == (72, 72, 72) and not c[5]]


def build_all():
    if not ref_runner.reference_available():
        print('reference sources absent: keeping prebuilt oracle/_ref')
        return []
    out = [ref_runner.build_ref_kernel(*c) for c in CONFIGS]
    out += [ref_runner.build_ref_kernel_cuda(*c) for c in CUDA_CONFIGS]
    return out


if __name__ == '__main__':
    for p in build_all():
        print(p)
