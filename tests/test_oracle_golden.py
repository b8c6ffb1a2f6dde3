"""The oracle (oracle/ppp_oracle.c + oracle/host_logic.py) against the golden
vectors recorded from the unmodified reference (tools/gen_golden.py): every
stage must be BIT-identical (same serial visiting order, same float/double
arithmetic)."""
import numpy as np
import pytest

from oracle import cpu_oracle, host_logic
from tests import golden_util as gu


@pytest.mark.parametrize('name', gu.NAMES)
def test_oracle_matches_reference_golden(name):
    g, kw, ps, pred = gu.load(name)
    var = cpu_oracle.variant_from_kwargs(kw)
    O = cpu_oracle.Oracle(pred, g['numinst'] > 1, ps, var)
    assert np.array_equal(O.fgidx >= 0, g['gate'])
    fg = pred[int(np.prod(ps)) // 2] > np.float32(kw['patch_threshold'])
    out = host_logic.assemble(pred, fg, g['numinst'], ps, kw, O)
    rows = g['rows']
    assert np.array_equal(O.cons_raw[rows], g['cons_raw'])
    assert np.isclose(O.cons_raw.astype(np.float64).sum(), g['cons_raw_sum'],
                      rtol=1e-9, atol=1e-6)
    if 'cnt' in g:
        cnt = O.cnt_pos.astype(np.int64) + O.cnt_neg
        assert np.array_equal(cnt[rows], g['cnt'])
        assert cnt.sum() == int(g['cnt_sum'])
    if 'cons_norm' in g:
        assert np.array_equal(O.cons[rows], g['cons_norm'])
    for k in ('score', 'ranked', 'cover', 'thin', 'pairs', 'aff', 'instances'):
        assert np.array_equal(out[k], g[k]), k


def test_counter_mode_is_pos_minus_neg():
    """with v3 = 1 the un-normalised consensus is exactly
    cnt_pos - cnt_neg (the CPU path's int16 consensus)."""
    g, kw, ps, pred = gu.load('worms2d_ps7_counter')
    var = cpu_oracle.variant_from_kwargs(kw)
    O = cpu_oracle.Oracle(pred, g['numinst'] > 1, ps, var)
    cr, cp, cn = O.consensus()
    assert np.array_equal(cr, cp.astype(np.float32) - cn.astype(np.float32))
