#!/usr/bin/env python
"""bench.py — consensus+assembly foreground Mvoxels/s (BASELINE.json metric).

Workload = BASELINE.json configs[4] (the north-star volume): ONE FlyLight-sized
3-D volume, 256x1024x1024, patchshape 7x7x7, synthetic neurites (seeded,
patchperpix_b200/synth.py), predictions in the compact row form a ppp+dec run
produces (float16 [G][343] for the G foreground voxels), assembled BLOCKWISE
(chunks 128x64x256 + patchshape//2 halo, one job per shared face, one global
partition) through patchperpix_b200.sharded.stitch_shard.  flylight
[vote_instances] flags, thresholded connected components.

N GPUs = the SAME volume cut into N slabs of whole block rows ("scaling": "strong");
NCCL moves halo rows (send/recv), block and face edge lists (all-gather).  A "step"
= the whole path over the whole volume: halo exchange -> per block gate/compact/
prepare/consensus/rank/cover/thin/pairs/patch graph -> face jobs -> global CC ->
painting of the own slab.

  value : fg voxels of the volume / s, patch rows resident in HBM, CUDA-event timed
  e2e   : same from PINNED HOST rows (H2D of coords+patches+numinst inside the
          timed region) to uint16 labels of the own slab back on the host
  roofline : the C-ABI call with the largest summed CUDA-event time over a step
          (events around every call of one extra, single-stream pass), algorithmic
          bytes per fg voxel of SURVEY.md 8d vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the reference kernels compiled for the host
          (oracle/_ref, OpenMP, all cores) + the oracle host logic on a bounded
          sample region of the same volume; the GPU path assembles the same
          region and the labels are compared ("parity_checked")
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(shape=(256, 1024, 1024), patchshape=(7, 7, 7), chunksize=(128, 64, 256),
                seed=4, n=600, seg_len=24.0, n_seg=40)
CPU_SAMPLE = (72, 72, 72)       # region the CPU arm works on (prebuilt oracle/_ref shapes)
KW = dict(patch_threshold=0.5, fc_threshold=0.5, cuda=True, blockwise=True,
          select_patches_for_sparse_data=True, includeSinglePatchCCS=True, mws=False,
          consensus_norm_prob_product=True, consensus_prob_product=True,
          consensus_norm_aff=True, consensus_interleaved_cnt=False,
          vi_bg_use_inv_th=False, vi_bg_use_half_th=False, vi_bg_use_less_than_th=True,
          rank_norm_patch_score=True, rank_int_counter=False, patch_graph_norm_aff=True,
          overlapping_inst=True, skipThinCover=False)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.stop_ev = threading.Event()
        self.rows = []
        self.index = index

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self.stop_ev.is_set():
            try:
                out = subprocess.run(
                    ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                     '--format=csv,noheader,nounits'],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([s.strip() for s in out.split(',')])
            except Exception:
                pass
            self.stop_ev.wait(0.5)

    def summary(self):
        self.stop_ev.set()
        self.join(timeout=6)
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names)
                   if any(r[2 + i].lower().startswith('active') for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None,
                    sm_max_mhz=float(self.rows[0][1]), reasons=reasons,
                    samples=len(self.rows))


def workload_from_args(args):
    w = dict(WORKLOAD)
    if args.shape:
        w['shape'] = tuple(int(v) for v in args.shape.split(','))
        w['n'] = max(4, int(WORKLOAD['n'] * np.prod(w['shape']) / np.prod(WORKLOAD['shape'])))
    if args.chunk:
        w['chunksize'] = tuple(int(v) for v in args.chunk.split(','))
    return w


def synth_kw(w):
    return dict(seed=w['seed'], n=w['n'], seg_len=w['seg_len'], n_seg=w['n_seg'])


# ---------------------------------------------------------------------------
# the bounded sample both arms can afford: a region of the same volume
# ---------------------------------------------------------------------------
CPU_REGIONS = 12                # regions the CPU arm assembles per step (about 10 s of host work)


def sample_regions(w, n=CPU_REGIONS):
    """n CPU_SAMPLE-sized boxes of the volume, each around the start of a neurite
    (deterministic): [(pred f32 [P,z,y,x], numinst u8 [z,y,x], start)]."""
    from patchperpix_b200 import synth
    shape = np.asarray(w['shape'])
    size = np.minimum(np.asarray(CPU_SAMPLE), shape)
    ps = w['patchshape']
    P = int(np.prod(ps))
    vol = synth.neurites_3d(w['shape'], **synth_kw(w))
    # the generator's own random sequence gives the first point of every polyline first
    rng = np.random.default_rng(w['seed'])
    out = []
    starts = []
    while len(out) < n:
        p = np.array([rng.uniform(0, shape[0]), rng.uniform(0, shape[1]),
                      rng.uniform(0, shape[2])])
        rng.normal(size=3)
        rng.uniform(0, 1)
        start = np.clip(p.astype(int) - size // 2, 0, shape - size)
        if any(np.all(np.abs(start - s0) < size) for s0 in starts):
            continue                                        # overlapping an earlier region
        lo, hi = int(start[0]), int(start[0] + size[0])
        coords, patches, numinst = synth.neurite_rows(
            w['shape'], ps, axis=0, lo=lo, hi=hi, device='cpu', box=(start, start + size),
            volume=vol, **synth_kw(w))
        if coords.shape[0] < 500:
            continue
        starts.append(start)
        c = coords.numpy().astype(np.int64) - start
        pred = np.zeros((P,) + tuple(int(s) for s in size), np.float32)
        pred[:, c[:, 0], c[:, 1], c[:, 2]] = patches.numpy().astype(np.float32).T
        ni = np.zeros(tuple(int(s) for s in size), np.uint8)
        ni[c[:, 0], c[:, 1], c[:, 2]] = numinst.numpy()
        out.append((pred, ni, start))
    return out


def sample_region(w):
    return sample_regions(w, 1)[0]


def cpu_reference_step(pred_np, numinst_np, ps, kw):
    """one pass of the reference CPU arm over a region: reference kernels (host
    build, OpenMP) + oracle host logic.  Returns (#fg, seconds, labels u16)."""
    from oracle import ref_runner, host_logic
    dims = pred_np.shape[1:]
    base = ['-DUSE_LESS_THAN_TH', '-DOVERLAP']
    th = kw['patch_threshold']
    ks = {}
    for kind, flags in (('fill', base + ['-DNORM_PROB_PRODUCT']),
                        ('cnt', base + ['-DNORM_PROB_PRODUCT', '-DOUTPUT_CNT']),
                        ('norm', []), ('rank', base + ['-DNORM_PATCH_RANK']),
                        ('graph', ['-DNORM_PATCH_AFFINITY'])):
        so = ref_runner.build_ref_kernel('fill' if kind == 'cnt' else kind, dims, ps, th,
                                         flags, omp=True)
        ks[kind] = ref_runner.load_ref_kernel(so)
    ns = [2 * p if (ps[0] > 1 or i > 0) else p for i, p in enumerate(ps)]
    Z, Y, X = dims
    grid = ((X + 7) // 8, (Y + 7) // 8, (Z + 7) // 8)
    blk = (8, 8, 8)
    mid = int(np.prod(ps)) // 2
    fg = pred_np[mid] > np.float32(th)
    overlap = np.ascontiguousarray(numinst_np > 1)
    t0 = time.perf_counter()
    cons = np.zeros(tuple(ns) + dims, np.float32)
    cnt = np.zeros(tuple(ns) + dims, np.float32)
    ks['fill'](pred_np, overlap, cons, block=blk, grid=grid)
    ks['cnt'](pred_np, overlap, cnt, block=blk, grid=grid)
    ks['norm'](pred_np, cons, cnt, block=blk, grid=grid)
    score = np.zeros(dims, np.float32)
    ks['rank'](pred_np, cons, overlap, score, block=blk, grid=grid)
    rad = np.array(ps) // 2
    allp = host_logic.interior_patches(fg, rad)
    ranked = host_logic.rank_by_score(allp, score)
    mask = fg.copy()
    mask[overlap] = 0
    fc = np.float32(kw['fc_threshold'])
    sel = host_logic.foreground_cover(1 * overlap, mask, np.array(ps), ranked, rad, pred_np, fc)
    sel = host_logic.thin_cover(mask, sel, np.array(ps), rad, pred_np, fc)
    pairs = host_logic.patch_pairs(sel, np.array(ps), True, 2)
    inst = np.zeros(dims, np.uint16)
    if pairs is not None:
        aff = np.zeros(len(pairs), np.float32)
        n = len(pairs)
        for i in range(0, n, 512):          # aff_patch_graph.py:141-159
            nb = min(512, n - i)
            ks['graph'](pred_np, cons, aff, pairs, np.uint64(nb), np.int32(i),
                        block=(min(512, n), 1, 1), grid=((nb + min(512, n) - 1) // min(512, n), 1, 1))
        inst, _ = host_logic.label_instances(pairs, aff, pred_np, np.array(ps), rad, dims,
                                             np.float32(th))
    dt = time.perf_counter() - t0
    return int(fg.sum()), dt, inst


def all_host_threads():
    """use every host core, also under torchrun (which exports OMP_NUM_THREADS=1)."""
    cores = os.cpu_count()
    os.environ['OMP_NUM_THREADS'] = str(cores)
    try:
        import ctypes
        ctypes.CDLL('libgomp.so.1').omp_set_num_threads(int(cores))
    except OSError:
        pass
    return cores


def workload_name(w):
    return ('configs[4]: 3-D FlyLight-sized volume %dx%dx%d, patchshape 7x7x7, %d synthetic '
            'neurites, compact f16 patch rows, blockwise chunks %s + halo, sharded in slabs'
            % (w['shape'] + (w['n'],) + ('x'.join(str(c) for c in w['chunksize']),)))


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    w = workload_from_args(args)
    cores = all_host_threads()
    regions = sample_regions(w)
    times = []
    nfg = 0
    for i in range(args.warmup + args.steps):
        nfg, dt = 0, 0.0
        for pred, ni, _ in regions:
            n1, d1, _ = cpu_reference_step(pred, ni, w['patchshape'], KW)
            nfg += n1
            dt += d1
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = nfg / (ms * 1e-3) / 1e6
    sample = '%d regions of %s voxels of the volume (%d fg voxels), all stages, one block each' % (
        len(regions), 'x'.join(str(s) for s in regions[0][0].shape[1:]), nfg)
    line = dict(metric='consensus+assembly fg Mvoxels/s', value=val, unit='Mvoxels/s',
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference',
                config=dict(workload=workload_name(w), sample=sample),
                cpu_baseline=dict(value=val, unit='Mvoxels/s', cores=cores, kind='reference',
                                  sample=sample),
                e2e=dict(value=val, unit='Mvoxels/s', h2d_bytes_per_step=0,
                         d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


def run_refgpu_arm(args):
    """--impl refgpu: the UNMODIFIED reference kernels compiled for sm_100a
    (oracle/_ref/refk_cuda_*.so), launched with the reference's geometry -- (8,8,8)
    blocks, one thread per voxel, 512 patch pairs per launch with a sync in between
    (utilVoteInstances.py:452-462, aff_patch_graph.py:137-159) -- on the sample region
    of the bench volume, next to this build's kernels on the same region.  The host
    steps between the kernels (cover, thinning, pairs, labels) are the oracle's, as in
    the CPU arm; kernel times are reported separately."""
    import torch
    from oracle import ref_runner, host_logic
    from patchperpix_b200 import cuda_code as cc, vote_instances as vi
    from patchperpix_b200.assembly import RowSource
    if int(os.environ.get('RANK', '0')) != 0:
        return 0
    w = workload_from_args(args)
    ps = tuple(int(p) for p in w['patchshape'])
    pred_np, ni_np, start = sample_region(w)
    dims = pred_np.shape[1:]
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    base = ['-DUSE_LESS_THAN_TH', '-DOVERLAP']
    th = KW['patch_threshold']
    ks = {}
    for kind, flags in (('fill', base + ['-DNORM_PROB_PRODUCT']),
                        ('cnt', base + ['-DNORM_PROB_PRODUCT', '-DOUTPUT_CNT']),
                        ('norm', []), ('rank', base + ['-DNORM_PATCH_RANK']),
                        ('graph', ['-DNORM_PATCH_AFFINITY'])):
        so = ref_runner.build_ref_kernel_cuda('fill' if kind == 'cnt' else kind, dims, ps, th,
                                              flags)
        ks[kind] = ref_runner.load_ref_kernel_cuda(so)
    ns = [2 * p for p in ps]
    Z, Y, X = dims
    grid = ((X + 7) // 8, (Y + 7) // 8, (Z + 7) // 8)
    blk = (8, 8, 8)
    P = int(np.prod(ps))
    fg = pred_np[P // 2] > np.float32(th)
    overlap_np = np.ascontiguousarray(ni_np > 1)
    pred = torch.from_numpy(pred_np).to(dev)
    overlap = torch.from_numpy(overlap_np).to(dev)
    rad = np.array(ps) // 2
    times = []
    kms = {}
    inst = None
    for it in range(args.warmup + args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        km = {}

        def timed(name, fn, *a, **k):
            t = time.perf_counter()
            fn(*a, **k)
            km[name] = km.get(name, 0.0) + (time.perf_counter() - t) * 1e3
        cons = torch.zeros(tuple(ns) + dims, dtype=torch.float32, device=dev)
        cnt = torch.zeros(tuple(ns) + dims, dtype=torch.float32, device=dev)
        timed('fill', ks['fill'], pred.data_ptr(), overlap.data_ptr(), cons.data_ptr(),
              block=blk, grid=grid)
        timed('fill_cnt', ks['cnt'], pred.data_ptr(), overlap.data_ptr(), cnt.data_ptr(),
              block=blk, grid=grid)
        timed('norm', ks['norm'], pred.data_ptr(), cons.data_ptr(), cnt.data_ptr(),
              block=blk, grid=grid)
        del cnt
        score = torch.zeros(dims, dtype=torch.float32, device=dev)
        timed('rank', ks['rank'], pred.data_ptr(), cons.data_ptr(), overlap.data_ptr(),
              score.data_ptr(), block=blk, grid=grid)
        score_np = score.cpu().numpy()
        allp = host_logic.interior_patches(fg, rad)
        ranked = host_logic.rank_by_score(allp, score_np)
        mask = fg.copy()
        mask[overlap_np] = 0
        fc = np.float32(KW['fc_threshold'])
        sel = host_logic.foreground_cover(1 * overlap_np, mask, np.array(ps), ranked, rad,
                                          pred_np, fc)
        sel = host_logic.thin_cover(mask, sel, np.array(ps), rad, pred_np, fc)
        pairs = host_logic.patch_pairs(sel, np.array(ps), True, 2)
        inst = np.zeros(dims, np.uint16)
        if pairs is not None:
            n = len(pairs)
            pd = torch.from_numpy(np.ascontiguousarray(pairs).view(np.int32)).to(dev)
            aff = torch.zeros(n, dtype=torch.float32, device=dev)
            for i in range(0, n, 512):
                nb = min(512, n - i)
                timed('graph', ks['graph'], pred.data_ptr(), cons.data_ptr(), aff.data_ptr(),
                      pd.data_ptr(), np.uint64(nb), np.int32(i),
                      block=(min(512, n), 1, 1), grid=((nb + min(512, n) - 1) // min(512, n), 1, 1))
            inst, _ = host_logic.label_instances(pairs, aff.cpu().numpy(), pred_np, np.array(ps),
                                                 rad, dims, np.float32(th))
        del cons
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            for k, v in km.items():
                kms[k] = kms.get(k, 0.0) + v / args.steps
    nfg = int(fg.sum())
    # this build's kernels on the same region (events around every C-ABI call)
    c = np.argwhere(fg)
    v2r = torch.full(fg.shape, -1, dtype=torch.int32)
    v2r[c[:, 0], c[:, 1], c[:, 2]] = torch.arange(len(c), dtype=torch.int32)
    rows = torch.from_numpy(np.ascontiguousarray(
        pred_np[:, c[:, 0], c[:, 1], c[:, 2]].T.astype(np.float16)))
    src = RowSource(rows.to(dev), v2r.to(dev))
    fg_t = torch.from_numpy(fg.astype(np.uint8)).to(dev)
    ni_t = torch.from_numpy(ni_np).to(dev)
    kw1 = dict(KW, blockwise=False, ppp_latency_stream=False)
    for _ in range(2):
        inst_gpu, _ = vi.to_instance_seg(src, fg_t, fg_t.clone(), ni_t, np.array(ps), **kw1)
    with CallTimer(cc, torch) as ct:
        inst_gpu, _ = vi.to_instance_seg(src, fg_t, fg_t.clone(), ni_t, np.array(ps), **kw1)
    ours = {k: round(v['ms'], 4) for k, v in ct.calls.items()}
    ours_kernels = sum(v for k, v in ours.items())
    ref_kernels = sum(kms.values())
    ms = 1e3 * float(np.mean(times))
    sample = 'region %s at %s of the volume (%d fg voxels), single block' % (
        'x'.join(str(s) for s in dims), tuple(int(v) for v in start), nfg)
    line = dict(metric='consensus+assembly fg Mvoxels/s', value=nfg / (ms * 1e-3) / 1e6,
                unit='Mvoxels/s', n_gpus=1, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling='strong', vs_baseline=None,
                dtype='f32', data='synthetic', impl='refgpu',
                config=dict(workload=workload_name(w), sample=sample,
                            note='unmodified reference .cu kernels, nvcc sm_100a, reference launch '
                                 'geometry; host steps = oracle python (cover / thinning dominate '
                                 'ms_per_step)'),
                reference_kernel_ms={k: round(v, 4) for k, v in kms.items()},
                reference_kernels_total_ms=ref_kernels,
                ours_call_ms=ours, ours_total_ms=ours_kernels,
                kernel_time_ratio=ref_kernels / ours_kernels if ours_kernels else None,
                labels_identical=bool(np.array_equal(inst, inst_gpu)))
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
class CallTimer:
    """CUDA events around every C-ABI call (single-stream profiling pass) and a
    count of the rows each consensus / rank call processed."""

    def __init__(self, cc, torch):
        self.cc, self.torch = cc, torch
        self.orig = cc.call
        self.ev = []
        self.calls = {}
        self.needed = []        # rows a masked consensus call really computes

    def __enter__(self):
        def timed(name, *a):
            if name.endswith('_bytes') or name in ('ppp_pyset_order', 'ppp_pyset_pairs',
                                                   'ppp_mws_host'):
                return self.orig(name, *a)
            e0 = self.torch.cuda.Event(enable_timing=True)
            e1 = self.torch.cuda.Event(enable_timing=True)
            e0.record()
            r = self.orig(name, *a)
            e1.record()
            units = 0
            if name == 'ppp_consensus_small':
                need, name = a[5], 'ppp_consensus'
                units = int(self.needed.pop())
            elif name in ('ppp_consensus', 'ppp_rank'):
                units = int(a[5] if name == 'ppp_consensus' else a[4])
            elif name in ('ppp_patch_graph', 'ppp_patch_graph_rows'):
                units = int(a[5] if name == 'ppp_patch_graph' else a[7])
            self.ev.append((name, e0, e1, units))
            return r
        self.cc.call = timed
        self.cc.profile_hook = lambda name, v: self.needed.append(v)
        return self

    def __exit__(self, *exc):
        self.cc.call = self.orig
        self.cc.profile_hook = None
        self.torch.cuda.synchronize()
        for name, e0, e1, units in self.ev:
            d = self.calls.setdefault(name, dict(ms=0.0, calls=0, units=0))
            d['ms'] += e0.elapsed_time(e1)
            d['calls'] += 1
            d['units'] += units
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'refgpu'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--shape', default=None, help='z,y,x (default: the FlyLight-sized volume)')
    ap.add_argument('--chunk', default=None, help='z,y,x chunksize')
    ap.add_argument('--workers', type=int, default=6)
    ap.add_argument('--mws', action='store_true')
    ap.add_argument('--no-decoder', action='store_true')
    ap.add_argument('--no-c2', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    if args.impl == 'refgpu':
        return run_refgpu_arm(args)

    import torch
    import torch.distributed as dist
    from patchperpix_b200 import cuda_code as cc, sharded, synth, vote_instances as vi
    from patchperpix_b200.assembly import RowSource
    from patchperpix_b200.layout import patch_geometry
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cc.init_cuda()
    w = workload_from_args(args)
    ps = np.array(w['patchshape'])
    _, P, _, _, _, K = patch_geometry(ps)
    kw = dict(KW, patchshape=list(w['patchshape']), chunksize=list(w['chunksize']), mws=args.mws)
    for k in ('PPP_INFLIGHT', 'PPP_STREAMS'):            # tuning experiments
        if os.environ.get(k):
            kw[k.lower()] = int(os.environ[k])
    shape = w['shape']

    # ---- inputs: the patch rows of my slab, resident in HBM ---------------------
    # slabs = contiguous runs of block rows balanced by their foreground count (every
    # rank draws the same label volume; a production run would histogram its fg mask)
    t_gen = time.perf_counter()
    vol = synth.neurites_3d(shape, **synth_kw(w))
    axis, _ = sharded.slab_partition(shape, w['chunksize'], world)
    csz = min(w['chunksize'][axis], shape[axis])
    per_plane = (vol[0] > 0).sum(axis=tuple(a for a in range(3) if a != axis))
    weights = [int(per_plane[i:i + csz].sum()) for i in range(0, shape[axis], csz)]
    axis, slabs = sharded.slab_partition(shape, w['chunksize'], world, axis=axis,
                                         weights=weights)
    lo, hi = slabs[rank]
    coords, patches, numinst = synth.neurite_rows(shape, ps, axis=axis, lo=lo, hi=hi,
                                                  device=dev, volume=vol, **synth_kw(w))
    del vol
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    n_own = int(coords.shape[0])
    steps, warm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        shard = sharded.RowShard(shape, axis, lo, hi, coords, patches, numinst)
        return sharded.stitch_shard(shard, slabs, workers=args.workers, **kw)

    # ---- device-resident timing -------------------------------------------------
    for _ in range(warm):
        inst, info = device_step()
    sampler = ClockSampler(local) if rank == 0 else None     # one nvidia-smi poller per box
    if sampler:
        sampler.start()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    l0 = int(cc.call('ppp_launch_count'))
    t0.record()
    tw = time.perf_counter()
    for _ in range(steps):
        inst, info = device_step()
    t1.record()
    barrier()
    launches_per_step = (int(cc.call('ppp_launch_count')) - l0) / steps
    wall_ms = (time.perf_counter() - tw) * 1e3 / steps
    ms = t0.elapsed_time(t1) / steps
    clocks = sampler.summary() if sampler else None

    # ---- end to end: pinned host rows -> uint16 labels of my slab on the host ------
    coords_h = coords.cpu().pin_memory()
    patches_h = patches.cpu().pin_memory()
    numinst_h = numinst.cpu().pin_memory()
    own_box = list(shape)
    own_box[axis] = hi - lo
    out_h = torch.empty(own_box, dtype=torch.int16).pin_memory()
    h2d = coords_h.numel() * 4 + patches_h.numel() * 2 + numinst_h.numel()
    d2h = out_h.numel() * 2

    def e2e_step():
        c = coords_h.to(dev, non_blocking=True)
        p = patches_h.to(dev, non_blocking=True)
        n = numinst_h.to(dev, non_blocking=True)
        shard = sharded.RowShard(shape, axis, lo, hi, c, p, n)
        inst_e, _ = sharded.stitch_shard(shard, slabs, workers=args.workers, **kw)
        out_h.copy_(inst_e.to(torch.int16), non_blocking=True)
        torch.cuda.synchronize()
    e2e_step()
    barrier()
    te = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - te) * 1e3 / steps
    assert torch.equal(out_h.to(torch.int32), inst.cpu().to(torch.int16).to(torch.int32)), \
        "end-to-end labels differ from the device-resident run"

    # ---- per-call CUDA-event times: one extra single-stream pass --------------------
    with CallTimer(cc, torch) as ct:
        shard = sharded.RowShard(shape, axis, lo, hi, coords, patches, numinst)
        sharded.stitch_shard(shard, slabs, workers=1, **dict(kw, ppp_pipeline=False,
                                                            ppp_latency_stream=False))
    calls = ct.calls

    # ---- digest of the whole label volume (equal for every N) ------------------------
    full = sharded.gather_slabs(inst, slabs, axis, shape)
    sha = None
    n_inst = None
    if rank == 0:
        fh = full.to(torch.int16).cpu().numpy()
        sha = hashlib.sha1(fh.tobytes()).hexdigest()[:16]
        n_inst = int(info.get('n_labels', 0))
    del full

    sys.stderr.write('[bench] rank %d: rows %d  step %.1f ms (wall %.1f)  e2e %.1f ms  blocks %d '
                     'faces %d  gen %.1f s  phases %s\n' % (
                         rank, n_own, ms, wall_ms, e2e_ms, info['my_blocks'], info['my_faces'],
                         t_gen, {k: round(v, 1) for k, v in info['phase_ms'].items()}))
    # ---- max over ranks ------------------------------------------------------------
    tot_fg = n_own
    halo = info['halo_bytes']
    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = (float(x) for x in t.tolist())
        c = torch.tensor([n_own, halo, h2d, d2h, int(launches_per_step)], device=dev,
                         dtype=torch.int64)
        dist.all_reduce(c)
        tot_fg, halo, h2d, d2h, launches_per_step = (int(x) for x in c.tolist())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, how = peaks()
    # algorithmic bytes per fg voxel, f16 input rows as consumed here:
    # consensus = own patch (P x 2 B) + 1 gate byte + K x (f32 affinity + integer counter);
    # rank = patch + K consensus values + 1 score
    bytes_unit = dict(ppp_consensus=P * 2 + 1 + K * 8, ppp_rank=P * 2 + K * 4 + 4)
    step_ms_profiled = sum(d['ms'] for d in calls.values())
    # DRAM bytes per row of the same kernels from the committed ncu --set full capture
    traffic = {}
    tp = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    roofs = []
    for name in ('ppp_consensus', 'ppp_rank'):
        d = calls.get(name)
        if not d or d['ms'] <= 0:
            continue
        ach = bytes_unit[name] * d['units'] / (d['ms'] * 1e-3) / 1e9
        tr = traffic.get(name, {}).get('dram_bytes_per_row')
        roofs.append(dict(bound='hbm', achieved=ach, peak=peak, unit='GB/s', frac=ach / peak,
                          traffic=(tr * d['units'] / d['calls']) if tr else None, kernel=name, kernel_ms=d['ms'], launches=d['calls'],
                          units=d['units'], bytes_per_fg_voxel=bytes_unit[name],
                          share_of_gpu_time=d['ms'] / step_ms_profiled, peak_source=how,
                          note='summed over the %d calls of one step (blocks + face regions); '
                               'units = rows those calls processed' % d['calls']))
    roofs.sort(key=lambda r: -r['kernel_ms'])
    line = dict(
        metric='consensus+assembly fg Mvoxels/s', value=tot_fg / (ms * 1e-3) / 1e6,
        unit='Mvoxels/s', n_gpus=world, steps=steps, warmup=warm, ms_per_step=ms,
        higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f32',
        data='synthetic',
        config=dict(workload=workload_name(w), fg_voxels=tot_fg, slab_axis=axis, slabs=slabs,
                    blocks=info['n_blocks'], faces=info['n_faces'], edges=info['n_edges'],
                    instances=n_inst, labels_sha1=sha, block_workers=args.workers,
                    halo_bytes_per_step=halo,
                    l2='inputs (%.1f GB of patch rows) larger than L2' % (
                        tot_fg * P * 2 / 1e9),
                    flags='flylight [vote_instances] defaults, mws=%s' % args.mws),
        e2e=dict(value=tot_fg / (e2e_ms * 1e-3) / 1e6, unit='Mvoxels/s',
                 h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                 ms_per_step=e2e_ms, api='sharded.stitch_shard',
                 input='pinned host rows: coords i32 [G,3], patches f16 [G,343], numinst u8 [G]',
                 output='uint16 labels of the own slab, pinned host'),
        gpu_launches=launches_per_step, clocks=clocks,
        phase_ms={k: round(v, 2) for k, v in info['phase_ms'].items()},
        stage_ms={k: round(v['ms'], 3) for k, v in sorted(calls.items(), key=lambda kv: -kv[1]['ms'])},
    )
    if roofs:
        line['roofline'] = roofs[0]
        line['roofline_other'] = roofs[1:]
    if world == 1 and not args.no_decoder:
        # configs[3] (ppp+dec): the code -> patch decoder on the tensor cores for as many
        # codes as the volume has foreground voxels, rows out (seeded weights and codes:
        # no checkpoint ships with the reference)
        try:
            from patchperpix_b200.decoder import PatchDecoder, seeded_weights
            dec = PatchDecoder(seeded_weights(), device=dev)
            nb = 1 << 16
            codes_d = torch.rand((nb, 176), device=dev)
            for _ in range(2):
                dec.decode_rows(codes_d)
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(1, min(40, tot_fg // nb))
            d0.record()
            for _ in range(reps):
                dec.decode_rows(codes_d)
            d1.record()
            torch.cuda.synchronize()
            dms = d0.elapsed_time(d1) / reps
            flop = 2.0 * nb * 64 * 27 * (128 * 64 + 64 * 64 + 64 * 64)
            tf_peak = 1607.7
            pp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
            if os.path.exists(pp):
                tf_peak = float(json.load(open(pp)).get('bf16_tflops_sustained', tf_peak))
            tf = flop / (dms * 1e-3) / 1e12
            line.setdefault('roofline_other', []).append(dict(
                bound='tensor', achieved=tf, peak=tf_peak, unit='TFLOP/s', frac=tf / tf_peak,
                traffic=None, kernel='ppp_decode (tcgen05 implicit-GEMM decoder, configs[3])',
                kernel_ms=dms, units=nb, mcodes_per_s=nb / dms / 1e3,
                whole_volume_ms=dms * tot_fg / nb,
                note='65 536 seeded codes per call, f16 rows out; flops = the three 3^3 '
                     'tensor-core convolutions (128->64, 64->64, 64->64) per code'))
        except Exception as e:
            line.setdefault('roofline_other', []).append(dict(kernel='ppp_decode',
                                                             failed=repr(e)))
    if world == 1 and not args.no_c2:
        # BASELINE configs[1] (round 1's workload), kept as a side leg: one 2-D image
        # 696x520, patchshape 1x41x41, dense float32 input, single block
        try:
            from patchperpix_b200.assembly import BlockAssembler
            ps2 = np.array([1, 41, 41])
            _, P2, _, _, _, K2 = patch_geometry(ps2)
            lab2, ni2 = synth.worms_2d((520, 696), n_worms=40, seed=2)
            pred2 = synth.patches_from_labels(lab2, ps2, seed=2, device=dev)
            fg2 = (pred2[P2 // 2] > 0.5).to(torch.uint8)
            ov2 = torch.from_numpy((ni2 > 1).astype(np.uint8)).to(dev)
            mask2 = fg2.clone()
            mask2[ov2 > 0] = 0
            kw2 = dict(KW, blockwise=False)
            n2 = int(fg2.sum().item())

            def c2_step(ev=None):
                asm = BlockAssembler(pred2, fg2, ov2, ps2, **kw2)
                asm.prepare()
                if ev:
                    ev[0].record()
                asm.consensus(want_cnt=True)
                if ev:
                    ev[1].record()
                asm.rank()
                if ev:
                    ev[2].record()
                order = asm.ranked()
                sel = asm.thin(mask2, asm.cover(mask2, order))
                pairs = asm.patch_pairs(asm.coords(sel))
                pd = torch.from_numpy(pairs.view(np.int32)).to(dev)
                inst2, _ = asm.label(pd, asm.patch_graph(pd), sel)
                return inst2
            for _ in range(3):
                c2_step()
            reps = 8
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            c0.record()
            for i in range(reps):
                c2_step(evs[i])
            c1.record()
            torch.cuda.synchronize()
            c2ms = c0.elapsed_time(c1) / reps
            cms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
            rms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
            bc, br = P2 * 4 + 1 + K2 * 8, P2 * 4 + K2 * 4 + 4
            line['configs1_leg'] = dict(
                workload='configs[1]: 2D worms 696x520, patchshape 1x41x41, dense f32 input',
                fg_voxels=n2, ms_per_image=c2ms, value=n2 / (c2ms * 1e-3) / 1e6, unit='Mvoxels/s',
                consensus=dict(kernel_ms=cms, bytes_per_fg_voxel=bc,
                               achieved=bc * n2 / (cms * 1e-3) / 1e9,
                               frac=bc * n2 / (cms * 1e-3) / 1e9 / peak),
                rank=dict(kernel_ms=rms, bytes_per_fg_voxel=br,
                          achieved=br * n2 / (rms * 1e-3) / 1e9,
                          frac=br * n2 / (rms * 1e-3) / 1e9 / peak))
            del pred2
        except Exception as e:
            line['configs1_leg'] = dict(failed=repr(e))
    if not args.no_cpu_baseline and world == 1:
        try:
            cores = all_host_threads()
            n_c, dt, same, n_inst = 0, 0.0, True, 0
            regions = sample_regions(w)
            for pred_s, ni_s, start in regions:
                n1, d1, inst_cpu = cpu_reference_step(pred_s, ni_s, tuple(int(p) for p in ps), KW)
                n_c += n1
                dt += d1
                n_inst += int(inst_cpu.max())
                # the same region through the CUDA path (rows form): identical labels
                m = pred_s[P // 2] > np.float32(0.5)
                c = np.argwhere(m)
                v2r = torch.full(m.shape, -1, dtype=torch.int32)
                v2r[c[:, 0], c[:, 1], c[:, 2]] = torch.arange(len(c), dtype=torch.int32)
                rows = torch.from_numpy(np.ascontiguousarray(
                    pred_s[:, c[:, 0], c[:, 1], c[:, 2]].T.astype(np.float16)))
                src = RowSource(rows.to(dev), v2r.to(dev))
                fg_s = torch.from_numpy(m.astype(np.uint8)).to(dev)
                inst_gpu, _ = vi.to_instance_seg(src, fg_s, fg_s.clone(),
                                                 torch.from_numpy(ni_s).to(dev), ps,
                                                 **dict(KW, blockwise=False))
                same = same and bool(np.array_equal(inst_gpu, inst_cpu))
            line['parity_checked'] = same
            line['cpu_baseline'] = dict(
                value=n_c / dt / 1e6, unit='Mvoxels/s', cores=cores, kind='reference',
                sample='%d regions of %s voxels of the volume (%d fg voxels, %d instances), all '
                       'stages, one block each, %.1f s; GPU labels on the same regions '
                       'identical: %s' % (len(regions), 'x'.join(str(v) for v in CPU_SAMPLE), n_c,
                                          n_inst, dt, same))
            if not same:
                raise AssertionError('GPU labels differ from the CPU reference arm on the sample')
        except AssertionError:
            raise
        except Exception as e:          # the baseline must not take the bench down
            line['cpu_baseline'] = dict(value=None, unit='Mvoxels/s', cores=os.cpu_count(),
                                        kind='reference', sample='failed: %r' % (e,))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
