"""ctypes binding of ``csrc/libpcst.so`` (the C ABI declared in ``include/pcst.h``).

There is no fallback of any kind: if the shared library is missing or a call returns a non-zero
status this module raises.  Build the library with ``python -m pointcloud_style_transfer_b200.build``
(or ``__graft_entry__.build()``), which runs ``make`` in ``csrc/`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCST_LIB points at another build of the same C ABI (A/B timing of two builds on one box)
LIB_PATH = os.environ.get("PCST_LIB") or os.path.join(_HERE, "csrc", "libpcst.so")

PCST_OK = 0
STATUS_NAMES = {0: "PCST_OK", -1: "PCST_ERR_INVALID", -2: "PCST_ERR_UNSUPPORTED", -3: "PCST_ERR_CUDA",
                -4: "PCST_ERR_WORKSPACE"}


class PcstError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class Mlp3(ctypes.Structure):
    """``pcst_mlp3_t``."""
    _fields_ = [("w", c_void_p * 3), ("scale", c_void_p * 3), ("shift", c_void_p * 3), ("cout", c_int * 3)]


class Mlp3Train(ctypes.Structure):
    """``pcst_mlp3_train_t``."""
    _fields_ = [("w", c_void_p * 3), ("bias", c_void_p * 3), ("gamma", c_void_p * 3), ("beta", c_void_p * 3),
                ("running_mean", c_void_p * 3), ("running_var", c_void_p * 3), ("num_batches_tracked", c_void_p * 3),
                ("cout", c_int * 3), ("eps", c_float), ("momentum", c_float)]


class NoiseMlp(ctypes.Structure):
    """``pcst_noise_mlp_t``."""
    _fields_ = [("pe_w", c_void_p * 3), ("pe_b", c_void_p * 3), ("time_w", c_void_p), ("time_b", c_void_p),
                ("style_w", c_void_p), ("style_b", c_void_p), ("blk_w1", c_void_p * 8), ("blk_b1", c_void_p * 8),
                ("blk_w2", c_void_p * 8), ("blk_b2", c_void_p * 8), ("out_w", c_void_p * 3), ("out_b", c_void_p * 3),
                ("feature_dim", c_int), ("time_dim", c_int), ("nblocks", c_int)]


class Mlp3Grads(ctypes.Structure):
    """``pcst_mlp3_grads_t``."""
    _fields_ = [("w", c_void_p * 3), ("bias", c_void_p * 3), ("gamma", c_void_p * 3), ("beta", c_void_p * 3)]


# name -> (restype, argtypes); the single source of truth for the symbol table test
SIGNATURES = {
    "pcst_version": (c_char_p, []),
    "pcst_last_error": (c_char_p, []),
    "pcst_device_check": (c_int, []),
    "pcst_set_tuning": (c_int, [c_char_p, c_int]),
    "pcst_get_tuning": (c_int, [c_char_p, POINTER(c_int)]),
    "pcst_fp32_probe": (ctypes.c_longlong, [c_int, c_void_p, c_void_p]),
    "pcst_sa_mlp_set_probe": (None, [c_void_p, c_int]),
    "pcst_l2_prefetch": (c_int, [c_void_p, c_size_t, c_void_p]),
    "pcst_fps_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcst_fps_max_concurrent_clouds": (c_int, [c_int]),
    "pcst_fps_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pcst_ball_query_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcst_ball_query_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "pcst_square_distance_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pcst_index_points_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pcst_index_points_bwd_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pcst_group_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                               c_void_p]),
    "pcst_sa_mlp_pick_cluster": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_int), c_int]),
    "pcst_sa_mlp_packed_bytes": (c_size_t, [c_int, POINTER(c_int), c_int, c_int]),
    "pcst_sa_mlp_pack_f32": (c_int, [POINTER(Mlp3), c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "pcst_sa_mlp_max_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_int]),
    "pcst_sa_mlp_max_kernel_launches": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_int]),
    "pcst_sa_mlp_max_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    POINTER(c_int), c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pcst_sa_mlp_train_saved_bytes": (c_size_t, [c_int, c_int, c_int, c_int, POINTER(c_int)]),
    "pcst_sa_mlp_train_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, POINTER(c_int), c_int, c_int]),
    "pcst_sa_mlp_max_bnstats_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                             POINTER(Mlp3Train), c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t,
                                             c_void_p]),
    "pcst_sa_mlp_max_bwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                         POINTER(Mlp3Train), c_int, c_void_p, c_size_t, c_void_p, POINTER(Mlp3Grads), c_void_p,
                                         c_void_p, c_size_t, c_void_p]),
    "pcst_noise_predictor_packed_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcst_noise_predictor_pack_launches": (c_int, [c_int, c_int, c_int]),
    "pcst_noise_predictor_plan_selfcheck": (c_int, [c_int, c_int, c_int]),
    "pcst_noise_predictor_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcst_noise_predictor_pack_f32": (c_int, [POINTER(NoiseMlp), c_void_p, c_size_t, c_void_p]),
    "pcst_noise_predictor_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p, c_size_t, c_void_p]),
    "pcst_nn_min_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcst_nn_min_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_size_t, c_void_p]),
    "pcst_nn_min_pair_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcst_nn_min_pair_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    "pcst_nn_min_pair_arg_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcst_nn_min_pair_arg_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_size_t, c_void_p]),
    "pcst_chamfer_bwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                     c_void_p, c_void_p]),
    "pcst_chamfer_shard_pack_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pcst_chamfer_shard_payload_floats": (c_int, [c_int]),
    "pcst_chamfer_shard_finish_f32": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.c_longlong, c_int, c_void_p, c_void_p,
                                              c_size_t, c_void_p]),
    "pcst_knn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "pcst_knn_kernel_launches": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "pcst_knn_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                             c_void_p]),
    "pcst_knn_interpolate_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                         c_void_p]),
    "pcst_minmax_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pcst_voxel_representatives_workspace_bytes": (c_size_t, [c_int, c_int]),
    "pcst_voxel_representatives_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_size_t, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libpcst.so (once) and declare every prototype.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m "
            "pointcloud_style_transfer_b200.build` (nvcc, sm_100a). There is no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    # PCST_TUNE="key=value,key=value": tuning knobs for a whole process (A/B measurements without code changes)
    for kv in filter(None, os.environ.get("PCST_TUNE", "").split(",")):
        key, _, value = kv.partition("=")
        status = lib.pcst_set_tuning(key.strip().encode(), int(value))
        if status != PCST_OK:
            raise PcstError(status, lib.pcst_last_error().decode("utf-8", "replace"))
    return lib


def last_error() -> str:
    return load().pcst_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    if status != PCST_OK:
        raise PcstError(status, last_error())


def set_tuning(key: str, value: int) -> None:
    check(load().pcst_set_tuning(key.encode(), int(value)))


def get_tuning(key: str) -> int:
    v = c_int(0)
    check(load().pcst_get_tuning(key.encode(), ctypes.byref(v)))
    return v.value


def version() -> str:
    return load().pcst_version().decode()
