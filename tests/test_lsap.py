"""HungarianAssignerV2's matching (hungarian_assigner.py:229-270 over scipy.optimize.linear_sum_assignment).

CPU (`not gpu`): the C restatement oracle/lsap.c — sequential AND keyed (parallel-formulation) variants — and the ONE-thread host
emulation of the kernel body (csrc/lsap_core.cuh compiled with g++) are pinned against scipy itself: random fp32 costs, tie-heavy
small-integer and constant costs (where only scipy's exact tie rule gives the same answer), both orientations, the <= topk_k rounds.
GPU: ptb_hungarian_v2_batch against scipy — bit-exact assignments.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment as scipy_lsa

from oracle import lsap as olsap
from oracle import p2p as op2p

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cost(rng, N, n, kind):
    if kind == 'float':
        return rng.standard_normal((N, n)).astype(np.float32)
    if kind == 'int3':
        return rng.integers(0, 3, (N, n)).astype(np.float32)
    if kind == 'int2':
        return rng.integers(0, 2, (N, n)).astype(np.float32)
    if kind == 'const':
        return np.full((N, n), float(rng.integers(0, 3)), np.float32)
    if kind == 'p2p':      # focal-cost-like column pattern + L1 distance: many near-ties between neighbouring proposals
        side = max(int(np.ceil(np.sqrt(N))), 1)
        px = np.stack([np.arange(N) % side, np.arange(N) // side], 1).astype(np.float32) * 8
        g = rng.uniform(0, side * 8, (n, 2)).astype(np.float32)
        d = np.abs(px[:, None, :] - g[None]).sum(-1) / np.float32(side * 8)
        return (np.float32(0.1) * d + np.float32(2.0) * rng.uniform(0, 1e-3, (N, 1)).astype(np.float32)).astype(np.float32)
    raise ValueError(kind)


def _ref(cost, k):
    gi, _ = op2p.hungarian_v2_from_cost(torch.from_numpy(cost), torch.zeros(cost.shape[1], dtype=torch.long), k)
    return gi.numpy()


KINDS = ['float', 'int3', 'int2', 'const', 'p2p']


def test_oracle_lsap_matches_scipy():
    rng = np.random.default_rng(0)
    for trial in range(400):
        nr, nc = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        c = _cost(rng, nr, nc, KINDS[trial % len(KINDS)])
        r0, c0 = scipy_lsa(c)
        for keyed in (False, True):
            r1, c1 = olsap.linear_sum_assignment(c, keyed)
            assert np.array_equal(r0, r1) and np.array_equal(c0, c1), (trial, keyed, nr, nc)
    for N, n in [(2000, 60), (60, 2000), (300, 300)]:
        for kind in ('float', 'int3'):
            c = _cost(rng, N, n, kind)
            r0, c0 = scipy_lsa(c)
            r1, c1 = olsap.linear_sum_assignment(c, True)
            assert np.array_equal(r0, r1) and np.array_equal(c0, c1)


def test_oracle_lsap_errors_like_scipy():
    c = np.full((3, 4), np.inf, np.float32)
    with pytest.raises(ValueError, match='infeasible'):
        scipy_lsa(c)
    with pytest.raises(ValueError, match='infeasible'):
        olsap.linear_sum_assignment(c)
    c = np.zeros((3, 4), np.float32); c[1, 2] = np.nan
    with pytest.raises(ValueError, match='invalid numeric'):
        scipy_lsa(c)
    with pytest.raises(ValueError, match='invalid numeric'):
        olsap.linear_sum_assignment(c)


def test_oracle_hungarian_v2_rounds():
    rng = np.random.default_rng(1)
    for trial in range(200):
        N, n, k = int(rng.integers(1, 70)), int(rng.integers(1, 25)), int(rng.choice([1, 2, 5]))
        c = _cost(rng, N, n, KINDS[trial % len(KINDS)])
        assert np.array_equal(_ref(c, k), olsap.hungarian_v2(c, k, True)), (N, n, k)


@pytest.fixture(scope='module')
def emu():
    out = os.path.join(ROOT, 'tests', '_build', 'liblsap_emu.so')
    src = os.path.join(ROOT, 'tests', 'lsap_emu.cpp')
    hdr = os.path.join(ROOT, 'pointtinybenchmark_b200', 'csrc', 'lsap_core.cuh')
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', out, src], check=True)
    lib = ctypes.CDLL(out)
    lib.emu_hungarian_v2.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.emu_hungarian_v2.restype = ctypes.c_int
    return lib


def _emu(lib, c, k, row_idx=None, Q=None):
    N, n = c.shape
    out = np.zeros(Q if Q is not None else N, np.int64)
    rc = lib.emu_hungarian_v2(N, n, c.ctypes.data, k, row_idx.ctypes.data if row_idx is not None else None, out.ctypes.data)
    return rc, out


def test_kernel_body_host_emulation_matches_scipy(emu):
    """csrc/lsap_core.cuh with one emulated thread == scipy (bookkeeping of the GPU formulation, incl. tie-heavy costs)."""
    rng = np.random.default_rng(2)
    for trial in range(500):
        N, n, k = int(rng.integers(1, 60)), int(rng.integers(1, 30)), int(rng.choice([1, 2, 3, 5]))
        c = _cost(rng, N, n, KINDS[trial % len(KINDS)])
        rc, out = _emu(emu, c, k)
        assert rc == 0 and np.array_equal(_ref(c, k), out), (trial, N, n, k)
    for N, n, k in [(3000, 100, 5), (500, 500, 1), (300, 700, 1), (4096, 64, 5)]:
        c = _cost(rng, N, n, 'p2p' if N > n else 'float')
        rc, out = _emu(emu, c, k)
        assert rc == 0 and np.array_equal(_ref(c, k), out)
    # scatter through row_idx + status codes
    c = _cost(rng, 20, 4, 'float')
    ridx = np.sort(rng.choice(50, 20, replace=False)).astype(np.int32)
    rc, out = _emu(emu, c, 5, ridx, 50)
    full = np.zeros(50, np.int64); full[ridx] = _ref(c, 5)
    assert rc == 0 and np.array_equal(out, full)
    assert _emu(emu, np.full((3, 4), np.inf, np.float32), 1)[0] == 1
    bad = np.zeros((5, 2), np.float32); bad[0, 0] = np.nan
    assert _emu(emu, bad, 1)[0] == 2


# ---------------------------------------------------------------------------------------------------------------------
def _gpu_solve(costs, k, with_ridx=False, seed=0):
    from pointtinybenchmark_b200 import ops
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(seed)
    shapes = [c.shape for c in costs]
    flat = torch.from_numpy(np.concatenate([c.reshape(-1) for c in costs]) if sum(c.size for c in costs) else np.zeros(1, np.float32)).to(dev)
    if with_ridx:
        Q = max(s[0] for s in shapes) + 17
        ridx = [np.sort(rng.choice(Q, s[0], replace=False)).astype(np.int32) for s in shapes]
        roff = np.concatenate([[0], np.cumsum([s[0] for s in shapes])])
        rflat = torch.from_numpy(np.concatenate(ridx) if roff[-1] else np.zeros(1, np.int32)).to(dev)
        out = torch.zeros(len(costs) * Q, dtype=torch.int64, device=dev)
        st = ops.hungarian_v2_batch(flat, shapes, k, out, [b * Q for b in range(len(costs))], rflat, roff[:-1])
        out = out.cpu().numpy().reshape(len(costs), Q)
        res = [out[b][ridx[b]] for b in range(len(costs))]
        for b in range(len(costs)):                       # nothing written outside the listed slots
            mask = np.ones(Q, bool); mask[ridx[b]] = False
            assert not out[b][mask].any()
    else:
        ooff = np.concatenate([[0], np.cumsum([s[0] for s in shapes])])
        out = torch.zeros(max(int(ooff[-1]), 1), dtype=torch.int64, device=dev)
        st = ops.hungarian_v2_batch(flat, shapes, k, out, ooff[:-1])
        out = out.cpu().numpy()
        res = [out[ooff[b]:ooff[b + 1]] for b in range(len(costs))]
    return st.cpu().numpy(), res


@pytest.fixture(params=['one_cta', 'cluster', 'cluster6', 'cluster5'])
def lsap_kernel(request, monkeypatch):
    """both device formulations: one CTA per image (lsap_core.cuh) and a cluster of 8 / 6 / 5 CTAs per image (lsap_cluster.cuh: the library
    picks the size by occupancy, PTB_LSAP_NCTA forces it; problems beyond 17 600 columns or 1024 rows fall back to the former inside the
    library)"""
    monkeypatch.setenv('PTB_LSAP_CLUSTER', '0' if request.param == 'one_cta' else '1')
    if request.param in ('cluster6', 'cluster5'):
        monkeypatch.setenv('PTB_LSAP_NCTA', request.param[-1])
    else:
        monkeypatch.delenv('PTB_LSAP_NCTA', raising=False)
    return request.param


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_gpu_hungarian_matches_scipy_small_and_ties(lsap_kernel):
    rng = np.random.default_rng(3)
    for k in (1, 2, 5):
        costs = []
        for trial in range(60):
            N, n = int(rng.integers(1, 200)), int(rng.integers(1, 40))
            costs.append(_cost(rng, N, n, KINDS[trial % len(KINDS)]))
        for with_ridx in (False, True):
            st, res = _gpu_solve(costs, k, with_ridx, seed=k)
            assert not st.any()
            for c, r in zip(costs, res):
                assert np.array_equal(_ref(c, k), r), (c.shape, k, with_ridx)


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_gpu_hungarian_orientations_and_empty(lsap_kernel):
    rng = np.random.default_rng(4)
    costs = [_cost(rng, 300, 700, 'float'), _cost(rng, 256, 256, 'int3'), _cost(rng, 257, 256, 'int3'), np.zeros((0, 5), np.float32),
             np.zeros((7, 0), np.float32), _cost(rng, 1, 1, 'float'), _cost(rng, 1500, 1, 'float'), _cost(rng, 2, 1500, 'float'),
             _cost(rng, 1025, 33, 'p2p'), _cost(rng, 4099, 31, 'int2'),
             _cost(rng, 20000, 12, 'p2p')]          # > 19 200 columns: column state stays in the global workspace (no shared-memory copy)
    for k in (1, 5):
        st, res = _gpu_solve(costs, k)
        assert not st.any()
        for c, r in zip(costs, res):
            assert np.array_equal(_ref(c, k), r), (c.shape, k)


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_gpu_hungarian_status_codes(lsap_kernel):
    rng = np.random.default_rng(5)
    ok = _cost(rng, 50, 7, 'float')
    inf = np.full((6, 3), np.inf, np.float32)
    nan = _cost(rng, 40, 5, 'float'); nan[17, 2] = np.nan
    ninf = _cost(rng, 5, 40, 'float'); ninf[3, 30] = -np.inf
    st, res = _gpu_solve([ok, inf, nan, ninf, ok], 5)
    assert st.tolist() == [0, 1, 2, 2, 0]
    assert np.array_equal(_ref(ok, 5), res[0]) and np.array_equal(_ref(ok, 5), res[4])


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_gpu_hungarian_headline_shape(lsap_kernel):
    """P2P training shape: 16 800 valid proposals (100 x 168 map), 500 and 60 GTs, topk_k = 5 (configs2/COCO/p2p/...:117)."""
    import time
    rng = np.random.default_rng(6)
    costs = [_cost(rng, 16800, 500, 'p2p'), _cost(rng, 16800, 60, 'p2p'), _cost(rng, 16800, 200, 'float'), _cost(rng, 16800, 7, 'p2p')]
    t = time.time()
    refs = [_ref(c, 5) for c in costs]
    t_scipy = time.time() - t
    _gpu_solve(costs[3:], 5)                                   # warm-up
    torch.cuda.synchronize(); t = time.time()
    st, res = _gpu_solve(costs, 5)
    t_gpu = time.time() - t
    assert not st.any()
    for c, r, g in zip(costs, refs, res):
        assert np.array_equal(r, g), c.shape
        assert (g > 0).sum() == 5 * c.shape[1]
    print(f'hungarian 4 images: scipy {t_scipy * 1e3:.0f} ms, GPU incl. copies {t_gpu * 1e3:.0f} ms')
