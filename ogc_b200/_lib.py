"""ctypes binding of libogc_b200.so -- the C ABI declared in include/ogc_b200.h.

The library is hand-written sm_100a CUDA; there is NO fallback.  If it is missing, or a call
is made with non-CUDA tensors, we raise: a silent CPU/eager path would void every parity and
performance claim of this repository.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libogc_b200.so")

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_LL = ctypes.c_longlong

# name -> argtypes; every function returns int (0 ok, <0 ogc_status, >0 cudaError_t)
SIGNATURES = {
    "ogc_furthest_point_sampling": [_I, _I, _I, _P, _P, _P, _P],
    "ogc_gather_points": [_I, _I, _I, _I, _P, _P, _P, _P],
    "ogc_gather_points_grad": [_I, _I, _I, _I, _P, _P, _P, _P],
    "ogc_knn": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "ogc_knn_sqrt": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "ogc_knn_bounded": [_I, _I, _I, _I, _F, _P, _P, _P, _P, _P],
    "ogc_three_nn": [_I, _I, _I, _P, _P, _P, _P, _P],
    "ogc_three_interpolate": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "ogc_three_interpolate_grad": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "ogc_group_points": [_I, _I, _I, _I, _I, _P, _P, _P, _P],
    "ogc_group_points_grad": [_I, _I, _I, _I, _I, _P, _P, _P, _P],
    "ogc_ball_query": [_I, _I, _I, _F, _I, _P, _P, _P, _P],
    "ogc_weighted_kabsch": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "ogc_dynamic_loss": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "ogc_apply_rigid_flow": [_I, _I, _I, _P, _P, _P, _P, _P],
    "ogc_neighbor_l1": [_I, _I, _I, _I, _P, _P, _P, _F, _F, _P, _P, _P],
    "ogc_mask_contingency": [_I, _I, _I, _P, _P, _P, _P],
    "ogc_invariance_loss": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "ogc_count_nan": [_LL, _P, _P, _P],
    "ogc_adam_step": [_LL, _P, _P, _P, _P, _F, _F, _F, _F, _F, _I, _F, _P, _P],
    "ogc_sa_mlp_layer_fwd": [_I] * 8 + [_P] * 14,
    "ogc_gn_finalize": [_I, _I, _LL, _P, _P, _P, _P, _P, _P],
    "ogc_sa_finish": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P],
    "ogc_sa_last_stats": [_I, _I, _I, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "ogc_gn_bwd_coef": [_I, _I, _LL, _P, _P, _P, _P, _P],
    "ogc_sa_mlp_layer_dx": [_I] * 8 + [_P, _P, _I, _I] + [_P] * 14 + [_I, _I, _P],
    "ogc_sa_mlp_layer_dw": [_I] * 7 + [_P, _P, _I, _I] + [_P] * 11,
    "ogc_grid_build": [_I, _I, _F, _P, _P, _P, _P, _P],
    "ogc_knn_grid": [_I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P],
    "ogc_ball_query_grid": [_I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P],
    "ogc_sa_chain_debug": [_P],
    "ogc_sa_chain_fits": [_I] * 5 + [_P],
    "ogc_sa_chain_fwd": [_I] * 7 + [_P] * 19,
    "ogc_sa_pool_finish": [_I] * 3 + [_P] * 7 + [_I, _I, _P, _P, _P],
    "ogc_sa_mlp_layer_fwd_tc": [_I] * 8 + [_P] * 14,
    "ogc_sa_mlp_layer_dx_tc": [_I] * 8 + [_P, _P, _I, _I] + [_P] * 14 + [_I, _I, _P],
    "ogc_sa_chain_dx_debug": [_P],
    "ogc_sa_chain_dx": [_I] * 8 + [_P, _P, _I, _I] + [_P] * 14 + [_I, _I, _P, _I, _I, _P],
    "ogc_sa_mlp_layer_dw_tc": [_I] * 7 + [_P, _P, _I, _I] + [_P] * 11,
    "ogc_icp_correspond": [_I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P],
    "ogc_mask_match": [_I, _I, _P, _P, _P, _P],
    "ogc_lsap_maximize_host": [_I, _P, _P],
    "ogc_mask_nuclear_norm": [_I, _I, _I, _P, _P, _P, _P],
    "ogc_fp_interp_concat": [_I] * 5 + [_P] * 7,
    "ogc_pw_mlp_layer_fwd": [_I] * 4 + [_P] * 6,
    "ogc_gn_relu_apply": [_I] * 3 + [_P] * 4,
    "ogc_gn_relu_bwd_stats": [_I] * 3 + [_P] * 10,
    "ogc_pw_mlp_input_grad": [_I] * 6 + [_P] * 5 + [_I, _I, _P],
    "ogc_sa_mlp_narrow_fwd": [_I] * 6 + [_P] * 11,
    "ogc_sa_mlp_narrow_dx": [_I] * 5 + [_P, _P, _I, _I] + [_P] * 13,
    "ogc_sa_dx_tma": [_I] * 7 + [_P, _P, _I, _I] + [_P] * 13,
    "ogc_sa_fwd_tma": [_I] * 6 + [_P] * 10,
    "ogc_sa_dw_tma": [_I] * 5 + [_P, _P, _I, _I] + [_P] * 7,
    "ogc_sa_mlp_narrow_dw": [_I] * 5 + [_P, _P, _I, _I] + [_P] * 7,
    "ogc_mask_head_fwd": [_I] * 4 + [_F] + [_P] * 4,
    "ogc_mask_head_bwd": [_I] * 4 + [_F] + [_P] * 7,
    "ogc_softmax_transfer": [_I] * 4 + [_F] + [_P] * 5,
    "ogc_adam_step_dev": [_LL, _P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _P, _P],
    "ogc_bn_stats": [_I] * 3 + [_P] * 3,
    "ogc_bn_finalize": [_I, _I, _LL] + [_P] * 7 + [_F, _P],
    "ogc_bn_pool": [_I] * 4 + [_P] * 5,
    "ogc_bn_pool_bwd": [_I] * 4 + [_P] * 7,
    "ogc_bn_bwd_stats": [_I] * 3 + [_P] * 5,
    "ogc_bn_bwd_coef": [_I, _I, _LL] + [_P] * 7,
}

_lib = None


class OgcLibraryError(RuntimeError):
    pass


def load():
    """Load libogc_b200.so once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OgcLibraryError(
            f"{LIB_PATH} not found. Build it with `python -m ogc_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / eager fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = header / library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = _I
    lib.ogc_version.restype = ctypes.c_char_p
    lib.ogc_version.argtypes = []
    for name, argtypes in (("ogc_grid_sorted_bytes", [_I, _I]), ("ogc_grid_table_bytes", [_I]), ("ogc_grid_params_bytes", [_I])):
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _LL
    _lib = lib
    return lib


_STATUS = {-1: "OGC_ERR_INVALID_ARG", -2: "OGC_ERR_UNSUPPORTED", -3: "OGC_ERR_WORKSPACE"}


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        raise ValueError(f"{what}: {_STATUS.get(rc, rc)}")
    raise RuntimeError(f"{what}: CUDA error {rc}")
