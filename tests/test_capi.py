"""The C-ABI library: it loads, and exports every symbol include/ogc_b200.h declares.
(CPU-only: no compute call is made without a GPU.)"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ogc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ogc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_ten_reference_ops():
    names = declared_symbols()
    for op in ["furthest_point_sampling", "gather_points", "gather_points_grad", "knn", "three_nn",
               "three_interpolate", "three_interpolate_grad", "group_points", "group_points_grad", "ball_query"]:
        assert f"ogc_{op}" in names


def test_library_builds_loads_and_exports_every_declared_symbol():
    from ogc_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/ogc_b200.h but not exported"
    loaded = _lib.load()
    assert loaded.ogc_version().decode().endswith("sm_100a")
    # every symbol bound by the Python side is declared in the header (no private ABI)
    assert set(_lib.SIGNATURES) <= set(declared_symbols())


def test_argument_validation_needs_no_gpu():
    """Rejected arguments return a negative ogc_status before anything is launched."""
    from ogc_b200 import _lib
    lib = _lib.load()
    assert lib.ogc_knn(1, 4, 4, 0, None, None, None, None, None) == -1        # k < 1
    assert lib.ogc_knn(1, 4, 4, 225, None, None, None, None, None) == -1      # k > 224
    assert lib.ogc_furthest_point_sampling(1, 0, 1, None, None, None, None) == -1
    assert lib.ogc_ball_query(1, 4, 4, 1.0, -1, None, None, None, None) == -1
    assert lib.ogc_group_points(-1, 1, 1, 1, 1, None, None, None, None) == -1
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(RuntimeError):
        _lib.check(700, "x")


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: the B200 back-end raises on non-CUDA tensors instead of computing elsewhere."""
    import torch
    from ogc_b200.backend import B200Backend
    be = B200Backend()
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        be.fps(torch.zeros(1, 8, 3), 2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        be.knn(2, torch.zeros(1, 8, 3), torch.zeros(1, 8, 3))


def test_pointnet2_cuda_shim_exports_the_reference_wrappers():
    """INTEGRATION.md option B: the shim module carries the ten names of pointnet2/src/pointnet2_api.cpp:10-25."""
    import importlib.util
    path = os.path.join(ROOT, "ogc_b200", "shim", "pointnet2_cuda.py")
    spec = importlib.util.spec_from_file_location("pointnet2_cuda_shim_check", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name in ["ball_query_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "gather_points_wrapper",
                 "gather_points_grad_wrapper", "furthest_point_sampling_wrapper", "knn_wrapper", "three_nn_wrapper",
                 "three_interpolate_wrapper", "three_interpolate_grad_wrapper"]:
        assert callable(getattr(mod, name)), name
