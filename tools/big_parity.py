import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from patchperpix_b200 import synth, vote_instances as vi
from oracle import cpu_oracle, host_logic
ps = np.array([7, 7, 7])
pred, numinst, _ = synth.make_case(kind='neurites', patchshape=ps, seed=77, shape=(32, 72, 72), n=8,
                                   radius=(1.8, 2.8), seg_len=11.0, n_seg=10, hard_frac=0.05)
fg = pred[171] > np.float32(0.5)
print('fg', int(fg.sum()))
for mws in (False, True):
    kw = dict(bench.KW, mws=mws, return_intermediates=False, pad_with_ps=False)
    t0 = time.time()
    O = cpu_oracle.Oracle(pred, numinst > 1, ps, cpu_oracle.variant_from_kwargs(kw))
    want = host_logic.assemble(pred, fg, numinst, ps, kw, O)
    t1 = time.time()
    inst, _ = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), numinst.copy(), ps.copy(), **kw)
    pairs, aff = vi.to_instance_seg(pred.copy(), fg.copy(), fg.copy(), numinst.copy(), ps.copy(),
                                    **dict(kw, return_intermediates=True))
    print('mws' if mws else 'cc', 'oracle %.1fs' % (t1 - t0), 'instances', len(np.unique(inst)) - 1,
          'labels identical:', bool(np.array_equal(inst, want['instances'])),
          'pairs identical:', bool(np.array_equal(pairs, want['pairs'])),
          'max |aff diff| %.2e' % float(np.max(np.abs(aff - want['aff']))), flush=True)
