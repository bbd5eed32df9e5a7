"""TEST INFRASTRUCTURE — ctypes binding of oracle/lsap.c (the CPU restatement of scipy.optimize.linear_sum_assignment and of
hungarian_assigner.py:229-270).  Only tests/, smoke() and bench.py's cpu_baseline may import this.

build(): gcc -O2 -shared -> oracle/_build/liblsap_oracle.so (git-ignored; travels to the GPU box with the snapshot).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'lsap.c')
OUT = os.path.join(HERE, '_build', 'liblsap_oracle.so')
_lib = None


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # -ffp-contract=off: the duals are sums/differences only, but keep the compiler from ever fusing anything
    subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', OUT, SRC, '-lm'], check=True)
    return OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i64, p = ctypes.c_int64, ctypes.c_void_p
        _lib.lsap_solve.argtypes = [i64, i64, p, p, p, ctypes.c_int]
        _lib.lsap_solve.restype = ctypes.c_int
        _lib.hungarian_v2.argtypes = [i64, i64, p, ctypes.c_int, p, ctypes.c_int]
        _lib.hungarian_v2.restype = ctypes.c_int
    return _lib


def _raise(rc):
    if rc == -1:
        raise ValueError('cost matrix is infeasible')
    if rc == -2:
        raise ValueError('matrix contains invalid numeric entries')


def linear_sum_assignment(cost, keyed=False):
    """(row_ind, col_ind) like scipy; cost = 2-D float32 array (the reference hands scipy an fp32 tensor)."""
    c = np.ascontiguousarray(cost, dtype=np.float32)
    nr, nc = c.shape
    m = min(nr, nc)
    r, k = np.zeros(m, np.int64), np.zeros(m, np.int64)
    _raise(lib().lsap_solve(nr, nc, c.ctypes.data, r.ctypes.data, k.ctypes.data, int(keyed)))
    return r, k


def hungarian_v2(cost, topk_k, keyed=False):
    """assigned_gt_inds (N,) int64 of HungarianAssignerV2.assign for an (N, n) fp32 cost."""
    c = np.ascontiguousarray(cost, dtype=np.float32)
    N, n = c.shape
    out = np.zeros(N, np.int64)
    _raise(lib().hungarian_v2(N, n, c.ctypes.data, int(topk_k), out.ctypes.data, int(keyed)))
    return out
