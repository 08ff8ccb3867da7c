"""The device Hungarian routine (csrc/lsap.cuh) through its host twin against scipy.optimize.linear_sum_assignment
-- the un-vendored dependency the reference calls in match_mask_by_iou (losses/seg_loss_unsup.py:236).  Tie-heavy
inputs on purpose: empty slots give all-zero IoU rows, so identical assignments (not just equal totals) matter."""
import ctypes

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment


def solve(score):
    from ogc_b200 import _lib
    lib = _lib.load()
    n = score.shape[0]
    s = np.ascontiguousarray(score, dtype=np.float64)
    out = np.empty(n, dtype=np.int32)
    rc = lib.ogc_lsap_maximize_host(n, s.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 10, 16, 32])
def test_matches_scipy_on_random_and_tie_heavy_matrices(n):
    rng = np.random.default_rng(n)
    for trial in range(300):
        kind = trial % 5
        if kind == 0:
            a = rng.random((n, n))
        elif kind == 1:
            a = rng.integers(0, 3, (n, n)).astype(float)                 # many exact ties
        elif kind == 2:
            a = np.zeros((n, n))                                          # everything ties
        elif kind == 3:                                                   # IoU-like: a few matched slots, rest empty
            a = np.zeros((n, n))
            k = rng.integers(0, n + 1)
            rows, cols = rng.permutation(n)[:k], rng.permutation(n)[:k]
            a[rows, cols] = rng.random(k)
            a[rng.integers(0, n, 3), rng.integers(0, n, 3)] = rng.random(3) * 0.2
        else:
            a = np.round(rng.random((n, n)), 1)
        a = a.astype(np.float32).astype(np.float64)                       # the reference feeds fp32 IoUs
        _, col = linear_sum_assignment(a, maximize=True)
        np.testing.assert_array_equal(solve(a), col.astype(np.int32), err_msg=f"n={n} trial={trial} kind={kind}")


def test_iou_from_counts_matches_reference_formula():
    """Host-side restatement used by the composed path == the formula the device kernel evaluates."""
    from ogc_b200.losses import _hungarian_from_counts
    rng = np.random.default_rng(0)
    inter = rng.integers(0, 50, (4, 6, 6))
    inter[:, 2, :] = 0
    inter[:, :, 4] = 0
    got = _hungarian_from_counts(inter)
    for b in range(4):
        i = inter[b].astype(np.float32)
        union = i.sum(1, keepdims=True) + i.sum(0, keepdims=True) - i
        iou = i / np.maximum(union, np.float32(1e-10))
        np.testing.assert_array_equal(got[b], linear_sum_assignment(iou, maximize=True)[1])
        np.testing.assert_array_equal(got[b], solve(iou.astype(np.float64)))
