#!/usr/bin/env python
"""bench.py — consensus+assembly foreground Mvoxels/s (BASELINE.json metric).

Workload at N=1 = BASELINE.json configs[1]: 2-D BBBC010-style worm bodies,
696x520, patchshape 1x41x41, synthetic seeded patch predictions
(patchperpix_b200/synth.py), flylight [vote_instances] flags, thresholded CC.
A "step" = the whole assembly path (gate -> consensus -> rank -> cover ->
thin -> patch graph -> CC -> paint) over one image.

  value : fg voxels / s, inputs resident in HBM, CUDA-event timed
  e2e   : same through patchperpix_b200.vote_instances.to_instance_seg_stream
          with PINNED HOST inputs (float16, the stored form): every step copies
          its own prediction H2D and reads its labels back inside the timed
          region; the copy of sample i+1 overlaps the assembly of sample i.
          serial_ms_per_step = one to_instance_seg call after the other.
  roofline : the stage with the largest CUDA-event time (consensus or rank),
             algorithmic bytes per fg voxel (consensus
             P*4 + 1 + K*8, rank P*4 + K*4 + 4) vs MEASURED_PEAKS.json;
             roofline_other = the other one
  cpu_baseline : the reference kernels compiled for the host (oracle/_ref,
             all cores) + the oracle host logic on a bounded crop

--impl reference runs only that CPU arm.  N>1 (torchrun): every rank
assembles its own image (independent objects, no data-path collective; weak).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(kind='worms', shape=(520, 696), patchshape=(1, 41, 41), seed=2,
                n_worms=40)
CPU_SAMPLE = (1, 256, 256)      # crop the CPU arm works on (prebuilt oracle/_ref shapes)
KW = dict(patch_threshold=0.5, fc_threshold=0.5, cuda=True, blockwise=False,
          select_patches_for_sparse_data=True, includeSinglePatchCCS=True, mws=False,
          consensus_norm_prob_product=True, consensus_prob_product=True,
          consensus_norm_aff=True, consensus_interleaved_cnt=False,
          vi_bg_use_inv_th=False, vi_bg_use_half_th=False, vi_bg_use_less_than_th=True,
          rank_norm_patch_score=True, rank_int_counter=False, patch_graph_norm_aff=True,
          overlapping_inst=True, skipThinCover=False)
# kernels launched per C-ABI call (counted to report gpu_launches)
# (own kernels + the CUB radix-sort passes the library launches; checked against
# the ncu launch list profiles/r1_v10_launches.csv: 51 per step)
LAUNCHES = dict(ppp_gate=1, ppp_compact=3, ppp_prepare_patches=2, ppp_consensus=3,
                ppp_rank=10, ppp_rank_sort=12, ppp_cover=4, ppp_thin=6, ppp_patch_graph=1,
                ppp_label_cc=8, ppp_paint=1)


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
            self.stop_ev.wait(0.2)

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


def make_inputs(device, seed):
    from patchperpix_b200 import synth
    w = WORKLOAD
    labels, numinst = synth.worms_2d(w['shape'], n_worms=w['n_worms'], seed=seed)
    pred = synth.patches_from_labels(labels, w['patchshape'], seed=seed, device=device)
    return pred, numinst, labels


# ---------------------------------------------------------------------------
# reference arm: reference kernels on the host cores + oracle host logic
# ---------------------------------------------------------------------------
def cpu_reference_step(pred_np, numinst_np, ps, kw):
    """one pass of the reference CPU arm over a crop; returns (#fg, seconds)."""
    from oracle import ref_runner, host_logic, cpu_oracle
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
    if pairs is not None:
        aff = np.zeros(len(pairs), np.float32)
        n = len(pairs)
        for i in range(0, n, 512):          # aff_patch_graph.py:141-159
            nb = min(512, n - i)
            ks['graph'](pred_np, cons, aff, pairs, np.uint64(nb), np.int32(i),
                        block=(min(512, n), 1, 1), grid=((nb + min(512, n) - 1) // min(512, n), 1, 1))
        host_logic.label_instances(pairs, aff, pred_np, np.array(ps), rad, dims,
                                   np.float32(th))
    dt = time.perf_counter() - t0
    return int(fg.sum()), dt


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    ps = WORKLOAD['patchshape']
    pred, numinst, _ = make_inputs(None, WORKLOAD['seed'])
    Z, Y, X = CPU_SAMPLE
    y0, x0 = (pred.shape[2] - Y) // 2, (pred.shape[3] - X) // 2
    crop = np.ascontiguousarray(pred[:, :, y0:y0 + Y, x0:x0 + X])
    ncrop = np.ascontiguousarray(numinst[:, y0:y0 + Y, x0:x0 + X])
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    cores = os.cpu_count()
    os.environ['OMP_NUM_THREADS'] = str(cores)
    try:
        import ctypes
        ctypes.CDLL('libgomp.so.1').omp_set_num_threads(int(cores))
    except OSError:
        pass
    times = []
    nfg = 0
    for i in range(args.warmup + args.steps):
        nfg, dt = cpu_reference_step(crop, ncrop, ps, KW)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = nfg / (ms * 1e-3) / 1e6
    sample = 'centre crop %dx%d of the %dx%d image (%d fg voxels), all stages' % (
        Y, X, pred.shape[2], pred.shape[3], nfg)
    line = dict(metric='consensus+assembly fg Mvoxels/s', value=val, unit='Mvoxels/s',
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference',
                config=dict(workload='configs[1]: 2D worms 696x520, patchshape 1x41x41',
                            sample=sample),
                cpu_baseline=dict(value=val, unit='Mvoxels/s', cores=cores, kind='reference',
                                  sample=sample),
                e2e=dict(value=val, unit='Mvoxels/s', h2d_bytes_per_step=0,
                         d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def device_step(pred, fg, overlap, mask, ps, kw, timers=None):
    """the hot path on device-resident inputs; returns (labels tensor, F)."""
    import torch
    from patchperpix_b200.assembly import BlockAssembler
    asm = BlockAssembler(pred, fg, overlap, ps, **kw)
    asm.prepare()
    if timers is not None:
        timers[0].record()
    asm.consensus(want_cnt=True)
    if timers is not None:
        timers[1].record()
    asm.rank()
    if timers is not None:
        timers[2].record()
    order = asm.ranked()
    sel = asm.cover(mask, order)
    sel = asm.thin(mask, sel)
    pairs = asm.patch_pairs(asm.coords(sel))
    if pairs is None:
        return torch.zeros(asm.shape, dtype=torch.int32, device=pred.device), asm.F
    pd = torch.from_numpy(pairs.view(np.int32)).to(pred.device)
    aff = asm.patch_graph(pd)
    inst, _ = asm.label(pd, aff, sel)
    return inst, asm.F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from patchperpix_b200 import cuda_code as cc, vote_instances as vi
    from patchperpix_b200.layout import patch_geometry
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cc.init_cuda()

    # count launches through the C ABI
    launches = [0]
    orig_call = cc.call

    def counting_call(name, *a):
        launches[0] += LAUNCHES.get(name, 0)
        return orig_call(name, *a)
    cc.call = counting_call
    import patchperpix_b200.assembly as asm_mod
    asm_mod.cc.call = counting_call

    ps = np.array(WORKLOAD['patchshape'])
    _, P, _, _, _, K = patch_geometry(ps)
    # every rank assembles its own copy of the SAME image: equal work per GPU (weak scaling)
    pred, numinst, _ = make_inputs(dev, WORKLOAD['seed'])
    mid = P // 2
    fg = (pred[mid] > 0.5).to(torch.uint8)
    overlap = torch.from_numpy((numinst > 1).astype(np.uint8)).to(dev)
    mask = fg.clone()
    mask[overlap > 0] = 0
    nfg = int(fg.sum().item())
    steps, warm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------
    for _ in range(warm):
        device_step(pred, fg, overlap, mask, ps, KW)
    sampler = ClockSampler(local)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    barrier()
    launches[0] = 0
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    F_rows = 0
    for i in range(steps):
        _, F_rows = device_step(pred, fg, overlap, mask, ps, KW, ev[i])
    t1.record()
    barrier()
    n_launch = launches[0] // steps
    ms = t0.elapsed_time(t1) / steps
    cons_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    rank_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    clocks = sampler.summary()

    # ---- end to end: pinned host inputs -> labels on the host -----------------
    # the prediction as the predict step stores it: float16 [P,Z,Y,X]
    # (predict_no_gp.py:243-257); to_instance_seg widens it on the device, which is
    # the exact conversion loadAffinities does on the host (utilVoteInstances.py:136-250)
    pred_h = torch.empty(pred.shape, dtype=torch.float16).pin_memory()
    pred_h.copy_(pred)
    assert torch.equal(pred_h.to(dev).float(), pred), "synthetic input is not f16-exact"
    fg_h = fg.cpu().pin_memory()
    numinst_h = torch.from_numpy(numinst).pin_memory()
    h2d = pred_h.numel() * 2 + fg_h.numel() * 2 + numinst_h.numel()
    d2h = fg_h.numel() * 2 + fg_h.numel()          # u16 labels + u8 foreground
    # one sample after the other (what the reference's file loop does) ...
    for _ in range(2):
        vi.to_instance_seg(pred_h, fg_h, fg_h, numinst_h, ps, **KW)
    barrier()
    te = time.perf_counter()
    for _ in range(steps):
        inst_e2e, _ = vi.to_instance_seg(pred_h, fg_h, fg_h, numinst_h, ps, **KW)
    torch.cuda.synchronize()
    e2e_serial_ms = (time.perf_counter() - te) * 1e3 / steps

    # ... and through the multi-sample entry point, which uploads sample i+1 on a
    # copy stream while sample i is assembled (every step still copies its own
    # input from pinned memory inside the timed region and reads its labels back)
    def feed(n):
        for _ in range(n):
            yield (pred_h, fg_h, fg_h, numinst_h)
    for _ in vi.to_instance_seg_stream(feed(2), ps, **KW):
        pass
    barrier()
    te = time.perf_counter()
    for inst_s, _ in vi.to_instance_seg_stream(feed(steps), ps, **KW):
        pass
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - te) * 1e3 / steps
    assert np.array_equal(inst_s, inst_e2e)
    sys.stderr.write('[bench] rank %d: step %.2f ms  consensus %.2f  rank %.2f  e2e serial %.2f  '
                     'e2e streamed %.2f ms\n' % (rank, ms, cons_ms, rank_ms, e2e_serial_ms, e2e_ms))

    # ---- max over ranks ------------------------------------------------------
    tot_fg = nfg
    if world > 1:
        t = torch.tensor([ms, e2e_ms, cons_ms, rank_ms, e2e_serial_ms], device=dev,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, cons_ms, rank_ms, e2e_serial_ms = (float(x) for x in t.tolist())
        c = torch.tensor([nfg], device=dev, dtype=torch.int64)
        dist.all_reduce(c)
        tot_fg = int(c.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, how = peaks()
    # algorithmic bytes per fg voxel: consensus = own patch (f32) +
    # 1 gate byte + K x (f32 affinity + integer counter); rank = patch + K consensus
    # values + 1 score
    b_unit = P * 4 + 1 + K * 8
    achieved = b_unit * nfg / (cons_ms * 1e-3) / 1e9
    b_rank = P * 4 + K * 4 + 4
    # DRAM bytes per launch of the same kernels from the committed ncu capture
    traffic = {}
    tp = os.path.join(ROOT, 'profiles', 'r1_v9_traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    line = dict(
        metric='consensus+assembly fg Mvoxels/s', value=tot_fg / (ms * 1e-3) / 1e6,
        unit='Mvoxels/s', n_gpus=world, steps=steps, warmup=warm, ms_per_step=ms,
        higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
        data='synthetic',
        config=dict(workload='configs[1]: 2D worms 696x520, patchshape 1x41x41, '
                             '%d worms, one image per step per GPU' % WORKLOAD['n_worms'],
                    fg_voxels_per_image=nfg, rows=F_rows,
                    l2='inputs (2.4 GB prediction) larger than L2',
                    flags='flylight [vote_instances] defaults, mws=False'),
        e2e=dict(value=tot_fg / (e2e_ms * 1e-3) / 1e6, unit='Mvoxels/s',
                 h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                 ms_per_step=e2e_ms, api='vote_instances.to_instance_seg_stream',
                 serial_ms_per_step=e2e_serial_ms,
                 input='float16 [P,Z,Y,X] pinned host buffer (the stored form), widened on the device'),
        gpu_launches=n_launch, clocks=clocks,
    )
    r_cons = dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s',
                  frac=achieved / peak, traffic=traffic.get('ppp_consensus'),
                  kernel='ppp_consensus (consensus_count_kernel + consensus_rows_kernel)',
                  kernel_ms=cons_ms, share_of_step=cons_ms / ms,
                  bytes_per_fg_voxel=b_unit, peak_source=how,
                  note='HBM is the bound SURVEY 8d assigns; the kernel itself is '
                       'issue/latency-bound on ~1.8e10 pair visits (DESIGN.md section 6)')
    a_rank = b_rank * nfg / (rank_ms * 1e-3) / 1e9
    r_rank = dict(bound='hbm', achieved=a_rank, peak=peak, unit='GB/s', frac=a_rank / peak,
                  traffic=traffic.get('ppp_rank'),
                  kernel='ppp_rank (rank_lists_kernel + sort + rank_ref_kernel, reference '
                         'summation order)',
                  kernel_ms=rank_ms, share_of_step=rank_ms / ms, bytes_per_fg_voxel=b_rank,
                  peak_source=how,
                  note='serial float order fixed by the reference; 4-byte gathers from a '
                       'consensus array larger than L2 (DESIGN.md section 6)')
    first, second = (r_cons, r_rank) if cons_ms >= rank_ms else (r_rank, r_cons)
    line['roofline'] = first
    line['roofline_other'] = [second]
    if not args.no_cpu_baseline:
        try:
            pn = pred.cpu().numpy()
            Z, Y, X = CPU_SAMPLE
            y0, x0 = (pn.shape[2] - Y) // 2, (pn.shape[3] - X) // 2
            crop = np.ascontiguousarray(pn[:, :, y0:y0 + Y, x0:x0 + X])
            ncrop = np.ascontiguousarray(numinst[:, y0:y0 + Y, x0:x0 + X])
            n_c, dt = cpu_reference_step(crop, ncrop, tuple(int(p) for p in ps), KW)
            line['cpu_baseline'] = dict(
                value=n_c / dt / 1e6, unit='Mvoxels/s', cores=os.cpu_count(),
                kind='reference',
                sample='centre crop %dx%d (%d fg voxels), all stages, %.1f s' % (Y, X, n_c, dt))
        except Exception as e:          # the baseline must not take the bench down
            line['cpu_baseline'] = dict(value=None, unit='Mvoxels/s', cores=os.cpu_count(),
                                        kind='reference', sample='failed: %r' % (e,))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
