"""Golden fixture for the multi-frame voting: the UNMODIFIED reference vote.py (/root/reference/vote.py:17-131) evaluated
in FLOAT64 on CPU (its fp32 cdist is ill-conditioned under the 1/0.01 temperature, like oa_icp: SURVEY.md 7).  Build
container only; the .npz is committed.

    python tests/golden/make_golden_vote.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.append("/root/reference")

from tests.golden.cases import CASES, make_inputs   # noqa: E402


def main():
    # vote.py imports its evaluation plumbing at module level; none of it is used by mask_voting
    sys.modules.setdefault("tqdm", types.SimpleNamespace(tqdm=lambda x, **k: x))
    sys.modules["metrics"] = types.ModuleType("metrics")
    sys.modules["metrics.seg_metric"] = types.SimpleNamespace(accumulate_eval_results=None, calculate_AP=None,
                                                              calculate_PQ_F1=None, ClusteringMetrics=None)
    sys.modules["utils.pytorch_util"] = types.SimpleNamespace(AverageMeter=None)
    import vote as ref_vote
    case = CASES["vote"]
    inp = make_inputs(case)
    d = {k: v.double() for k, v in inp.items()}
    # match_mask_by_cost builds its permutation matrix with an explicit float32 torch.eye (vote.py:88): widen it for
    # the float64 evaluation
    orig_eye = torch.eye
    torch.eye = lambda *a, **k: orig_eye(*a, **{**k, "dtype": torch.float64})
    try:
        voted64 = ref_vote.mask_voting(d["pc"], d["mask"], d["flows"], time_window_size=case["window"]).numpy()
    finally:
        torch.eye = orig_eye
    out = {"voted64": voted64,
           "voted32": ref_vote.mask_voting(inp["pc"], inp["mask"], inp["flows"], time_window_size=case["window"]).numpy()}
    np.savez_compressed(os.path.join(HERE, "vote.npz"), **out)
    print({k: v.shape for k, v in out.items()}, "fp32 reference vs fp64 reference:",
          float(np.abs(out["voted32"] - out["voted64"]).max()))


if __name__ == "__main__":
    main()
