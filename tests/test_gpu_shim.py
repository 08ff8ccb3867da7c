"""GPU: INTEGRATION.md option B -- the reference's UNMODIFIED `pointnet2/pointnet2.py` (staged copy under oracle/_ref/py,
git-ignored, built by `python oracle/ref_arm.py stage`) on top of the committed `pointnet2_cuda` ctypes shim
(ogc_b200/shim/pointnet2_cuda.py) must reproduce this repository's operator layer bit for bit: same kernels underneath,
so any difference is a binding error (argument order, dtype, pre-fill conventions of pointnet2.py:32-33,61,99-100,251)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PY = os.path.join(ROOT, "oracle", "_ref", "py")

CHILD = r'''
import sys, json, hashlib
import torch
sys.path[:] = [p for p in sys.path if p not in ("", ".", ROOT)]
sys.path.insert(0, SHIM); sys.path.insert(0, REF_PY)
import pointnet2.pointnet2 as ops            # the reference's own file
assert ops.__file__.startswith(REF_PY), ops.__file__
torch.manual_seed(5)
xyz = (torch.rand(3, 1500, 3, device="cuda") - 0.5) * 20
feat = torch.randn(3, 7, 1500, device="cuda", requires_grad=True)
h = lambda t: hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()
out = {}
idx = ops.furthest_point_sample(xyz, 300); out["fps"] = h(idx)
new_xyz = ops.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous(); out["gather"] = h(new_xyz)
d, i = ops.knn(16, new_xyz, xyz); out["knn_d"], out["knn_i"] = h(d), h(i)
d3, i3 = ops.three_nn(xyz, new_xyz); out["nn3_d"], out["nn3_i"] = h(d3), h(i3)
w = torch.softmax(-d3, -1).contiguous()
g = ops.grouping_operation(feat, i); out["group"] = h(g)
itp = ops.three_interpolate(g.max(-1).values.contiguous(), i3, w); out["interp"] = h(itp)
(itp.sum() + g.sum()).backward(); out["dfeat_sum"] = float(feat.grad.sum())
out["ball"] = h(ops.ball_query(1.5, 24, xyz, new_xyz))
q = ops.QueryAndGroup(1.0, 16)(xyz, new_xyz, feat.detach())
out["qg"] = h(q[0]) if isinstance(q, tuple) else h(q)
print(json.dumps(out))
'''


def _run(paths):
    code = f"ROOT={ROOT!r}\nREF_PY={paths[0]!r}\nSHIM={paths[1]!r}\n" + CHILD
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.isdir(REF_PY), reason="staged reference python missing (python oracle/ref_arm.py stage)")
def test_reference_pointnet2_py_over_the_shim_equals_our_operator_layer(b200):
    via_shim = _run((REF_PY, os.path.join(ROOT, "ogc_b200", "shim")))
    # the same script over OUR pointnet2/pointnet2.py (repository root first on the path)
    code = (f"ROOT='/nonexistent'\nREF_PY={ROOT!r}\nSHIM={ROOT!r}\n" + CHILD).replace(
        "assert ops.__file__.startswith(REF_PY), ops.__file__", "")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    ours = json.loads(r.stdout.strip().splitlines()[-1])
    for k in via_shim:
        if k == "dfeat_sum":
            assert abs(via_shim[k] - ours[k]) <= 1e-3 * max(1.0, abs(ours[k])), (k, via_shim[k], ours[k])   # atomics order
        else:
            assert via_shim[k] == ours[k], k
