"""ogc_b200/metrics.py (device-wide evaluation, one D2H) against the unmodified reference's
metrics/seg_metric.accumulate_eval_results (golden: tests/golden/seg_metric.npz, tests/golden/make_golden_metric.py)."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"kitti_like": (4, 2048, 10, 50, 7), "tiny_objects": (3, 1024, 8, 120, 8), "no_ignore": (2, 512, 6, 0, 9),
         "sparse_labels": (2, 777, 12, 30, 10)}


def make_case(name):
    B, N, K, thresh, seed = CASES[name]
    rng = np.random.default_rng(seed)
    n_obj = 9 if name != "sparse_labels" else 5
    sizes = rng.dirichlet(np.ones(n_obj) * (0.4 if name == "tiny_objects" else 1.5), size=B)
    segm = np.stack([rng.choice(n_obj, size=N, p=sizes[b]) for b in range(B)])
    if name == "sparse_labels":
        segm = segm * 7 + 3                                   # arbitrary, non-contiguous label values
    logits = rng.normal(size=(B, N, K)) * 1.2
    for b in range(B):                                        # predictions correlated with the GT, slots permuted
        perm = rng.permutation(K)
        logits[b, np.arange(N), perm[(segm[b] // (7 if name == "sparse_labels" else 1)) % K]] += 2.5
    e = np.exp(logits - logits.max(-1, keepdims=True))
    mask = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    return torch.from_numpy(segm.astype(np.int32)), torch.from_numpy(mask), thresh


def _check(device):
    from ogc_b200.metrics import accumulate_eval_results
    g = dict(np.load(os.path.join(HERE, "golden", "seg_metric.npz")))
    for name in CASES:
        segm, mask, thresh = make_case(name)
        iou, matched, conf, n_gt = accumulate_eval_results(segm.to(device), mask.to(device), thresh)
        assert n_gt == int(g[name + ":n_gt"]), name
        assert iou.shape == g[name + ":iou"].shape, name
        np.testing.assert_allclose(iou, g[name + ":iou"], rtol=0, atol=1e-12, err_msg=name)
        np.testing.assert_array_equal(matched, g[name + ":matched"], err_msg=name)
        np.testing.assert_allclose(conf, g[name + ":conf"], rtol=0, atol=1e-6, err_msg=name)


def test_device_metrics_match_reference_cpu():
    _check("cpu")


@pytest.mark.gpu
def test_device_metrics_match_reference_gpu(b200):
    _check("cuda")
