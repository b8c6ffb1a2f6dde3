"""TEST INFRASTRUCTURE ONLY — ctypes front end of oracle/ppp_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  See the header of ppp_oracle.c for
the reference file:line each function restates.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, '_build')
SO = os.path.join(BUILD, 'libppp_oracle.so')
SRC = os.path.join(HERE, 'ppp_oracle.c')


def build(force=False):
    if not force and os.path.exists(SO) and \
            os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    os.makedirs(BUILD, exist_ok=True)
    subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-shared', '-fPIC',
                    '-o', SO, SRC, '-lm'], check=True)
    return SO


class _Cfg(ctypes.Structure):
    _fields_ = [('Z', ctypes.c_int), ('Y', ctypes.c_int), ('X', ctypes.c_int),
                ('psz', ctypes.c_int), ('psy', ctypes.c_int),
                ('psx', ctypes.c_int), ('th', ctypes.c_double),
                ('thi', ctypes.c_double), ('bg_mode', ctypes.c_int),
                ('prod_mode', ctypes.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ppp_oracle_gate.restype = ctypes.c_int64
    return _lib


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def variant_from_kwargs(kw):
    """flag -> kernel variant mapping of setKernelBuildOptions
    (utilVoteInstances.py:389-449) and consensus_array.py:83-88."""
    th = float(kw['patch_threshold'])
    if kw.get('vi_bg_use_inv_th', True):
        bg = 2 if th < 0.5 else 0
    elif kw.get('vi_bg_use_half_th', False):
        bg = 1
    elif kw.get('vi_bg_use_less_than_th', False):
        bg = 2
    else:
        raise RuntimeError('how is bg defined for vote instances?')
    if kw.get('consensus_norm_prob_product', True):
        prod = 2
    elif kw.get('consensus_prob_product', True):
        prod = 1
    else:
        prod = 0
    return dict(
        th=th, thi=(th if th < 0.5 else 1.0 - th), bg_mode=bg, prod_mode=prod,
        overlap=bool(kw.get('overlapping_inst', False)),
        norm_aff=bool(kw.get('consensus_norm_aff', True)),
        rank_flags=(1 if kw.get('rank_norm_patch_score', True) else 0) |
                   (2 if kw.get('rank_int_counter', False) else 0),
        graph_flags=1 if kw.get('patch_graph_norm_aff', True) else 0,
    )


class Oracle:
    """consensus -> norm -> rank -> patch graph on one block, CPU, serial."""

    def __init__(self, pred, overlap, patchshape, var):
        self.pred = np.ascontiguousarray(pred, np.float32)
        P, Z, Y, X = self.pred.shape
        ps = [int(p) for p in patchshape]
        assert P == ps[0] * ps[1] * ps[2]
        self.ps = ps
        self.var = var
        self.cfg = _Cfg(Z, Y, X, ps[0], ps[1], ps[2], var['th'], var['thi'],
                        var['bg_mode'], var['prod_mode'])
        self.overlap = None
        if var['overlap'] and overlap is not None:
            self.overlap = np.ascontiguousarray(overlap != 0, np.uint8)
        self.fgidx = np.empty((Z, Y, X), np.int32)
        self.F = int(lib().ppp_oracle_gate(
            ctypes.byref(self.cfg), _p(self.pred), _p(self.overlap),
            _p(self.fgidx)))
        n = 2 * np.array(ps) - 1
        self.K = (int(n.prod()) - 1) // 2

    def consensus(self, want_cons=True, want_cnt=True):
        F, K = self.F, self.K
        self.cons_raw = np.zeros((F, K), np.float32) if want_cons else None
        self.cnt_pos = np.zeros((F, K), np.uint16) if want_cnt else None
        self.cnt_neg = np.zeros((F, K), np.uint16) if want_cnt else None
        lib().ppp_oracle_consensus(
            ctypes.byref(self.cfg), _p(self.pred), _p(self.overlap),
            _p(self.fgidx), _p(self.cons_raw), _p(self.cnt_pos),
            _p(self.cnt_neg))
        return self.cons_raw, self.cnt_pos, self.cnt_neg

    def norm(self):
        self.cons = self.cons_raw.copy()
        if self.var['norm_aff']:
            lib().ppp_oracle_norm(ctypes.c_int64(self.cons.size),
                                  _p(self.cons), _p(self.cnt_pos),
                                  _p(self.cnt_neg))
        return self.cons

    def rank(self, cons=None):
        cons = self.cons if cons is None else cons
        score = np.zeros(self.fgidx.shape, np.float32)
        lib().ppp_oracle_rank(
            ctypes.byref(self.cfg), _p(self.pred), _p(self.overlap),
            _p(self.fgidx), _p(cons), ctypes.c_int(self.var['rank_flags']),
            _p(score))
        return score

    def patch_graph(self, pairs, cons=None, exact_sum=False):
        """exact_sum: accumulate in double (the reference's float accumulator
        drifts by ~1e-3 relative at 41x41; diagnostic reference value)."""
        cons = self.cons if cons is None else cons
        pairs = np.ascontiguousarray(pairs, np.uint32)
        aff = np.zeros(len(pairs), np.float32)
        lib().ppp_oracle_patch_graph(
            ctypes.byref(self.cfg), _p(self.pred), _p(self.fgidx), _p(cons),
            _p(pairs), ctypes.c_int64(len(pairs)),
            ctypes.c_int(self.var['graph_flags'] | (4 if exact_sum else 0)), _p(aff))
        return aff
