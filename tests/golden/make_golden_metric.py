"""Golden vectors for ogc_b200/metrics.py from the UNMODIFIED reference `metrics/seg_metric.accumulate_eval_results`
(run in the build container only; matplotlib, which the reference imports at module level, is stubbed).

    python tests/golden/make_golden_metric.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.append("/root/reference")
mpl = types.ModuleType("matplotlib"); mpl.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot
from metrics.seg_metric import accumulate_eval_results          # noqa: E402
from tests.test_metrics import make_case, CASES                 # noqa: E402

out = {}
for name in CASES:
    segm, mask, thresh = make_case(name)
    iou, matched, conf, n_gt = accumulate_eval_results(segm, mask, thresh)
    out[name + ":iou"], out[name + ":matched"], out[name + ":conf"], out[name + ":n_gt"] = iou, matched, conf, np.int64(n_gt)
np.savez_compressed(os.path.join(HERE, "seg_metric.npz"), **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
