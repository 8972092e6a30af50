"""ctypes binding of libprobpose_b200.so (the C ABI in include/probpose_b200.h).

There is no fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PROBPOSE_B200_LIB") or os.path.join(HERE, "libprobpose_b200.so")  # override: kernel-variant experiments

PREC_FP16X3, PREC_BF16, PREC_FP16, PREC_FP32_SIMT = 0, 1, 2, 3
PRECISIONS = {"fp16x3": PREC_FP16X3, "bf16": PREC_BF16, "fp16": PREC_FP16, "fp32_simt": PREC_FP32_SIMT}
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
OUT_F32, OUT_OPERAND, OUT_PLANES = 0, 1, 2
RECORD_FLOATS = 7
HEAD_PROBMAP, HEAD_HEATMAP = 0, 1
MAX_KEYPOINTS = 17

# every symbol include/probpose_b200.h declares (tests check they are all exported)
EXPORTS = [
    "pp_last_error", "pp_version", "pp_decode", "pp_gemm", "pp_operand_bytes", "pp_operand_from_f32",
    "pp_engine_workspace_bytes", "pp_engine_create", "pp_engine_destroy", "pp_engine_load", "pp_engine_finalize",
    "pp_engine_backbone", "pp_engine_head", "pp_engine_infer", "pp_engine_last_launch_count",
    "pp_engine_profile_begin", "pp_engine_profile_end", "pp_engine_set_graph", "pp_engine_graph_replay_count", "pp_crop_warp", "pp_attention", "pp_decode_udp", "pp_revert_heatmaps",
    "pp_allgather", "pp_operand_overflow",
]
KERNEL_CLASSES = ("gemm", "attention", "decode", "other")


class DecodeCfg(C.Structure):
    _fields_ = [("num_keypoints", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("input_is_logits", C.c_int32), ("temperature", C.c_float), ("normalize", C.c_float),
                ("error_divisor", C.c_float)]


class UdpCfg(C.Structure):
    _fields_ = [("num_keypoints", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("blur_kernel_size", C.c_int32)]


class GemmArgs(C.Structure):
    _fields_ = [("precision", C.c_int32), ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
                ("a", C.c_void_p), ("w", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("residual", C.c_void_p), ("act", C.c_int32), ("out_kind", C.c_int32), ("d", C.c_void_p),
                ("ldd", C.c_int32), ("plane", C.c_int32), ("up_hin", C.c_int32), ("up_win", C.c_int32),
                ("up_py", C.c_int32), ("up_px", C.c_int32), ("tile_n", C.c_int32), ("res_mod", C.c_int32),
                ("a_taps", C.c_int32), ("a_tap_shift", C.c_int32 * 9), ("in_pad", C.c_int32), ("in_h", C.c_int32),
                ("in_w", C.c_int32), ("out_pad", C.c_int32), ("cta_pair", C.c_int32)]


class EngineCfg(C.Structure):
    _fields_ = [("precision", C.c_int32), ("max_batch", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
                ("patch", C.c_int32), ("patch_pad", C.c_int32), ("embed_dim", C.c_int32), ("depth", C.c_int32),
                ("heads", C.c_int32), ("ffn_dim", C.c_int32), ("num_keypoints", C.c_int32),
                ("deconv_channels", C.c_int32), ("ln_eps", C.c_float), ("bn_eps", C.c_float),
                ("temperature", C.c_float), ("normalize", C.c_float), ("mean", C.c_float * 3), ("std", C.c_float * 3),
                ("head_kind", C.c_int32), ("blur_kernel_size", C.c_int32)]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * 4), ("launches", C.c_int64 * 4), ("gemm_flops", C.c_double)]


class PPError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PPError(f"{LIB_PATH} is missing: build it with `python -m probpose_code_b200.build` "
                      "(there is no CPU or PyTorch fallback)")
    l = C.CDLL(LIB_PATH)
    l.pp_last_error.restype = C.c_char_p
    l.pp_version.restype = C.c_char_p
    l.pp_operand_bytes.restype = C.c_size_t
    l.pp_operand_bytes.argtypes = [C.c_int32, C.c_int64, C.c_int64]
    l.pp_operand_from_f32.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    l.pp_decode.argtypes = [C.POINTER(DecodeCfg), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p,
                            C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    l.pp_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    l.pp_crop_warp.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                               C.c_int32, C.c_void_p]
    l.pp_decode_udp.argtypes = [C.POINTER(UdpCfg), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p]
    l.pp_revert_heatmaps.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                     C.c_int32, C.c_void_p, C.c_void_p]
    l.pp_attention.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                               C.c_void_p]
    l.pp_operand_overflow.argtypes = [C.c_int32, C.POINTER(C.c_int32)]
    l.pp_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    if True:
        l.pp_engine_workspace_bytes.restype = C.c_size_t
        l.pp_engine_workspace_bytes.argtypes = [C.POINTER(EngineCfg)]
        l.pp_engine_create.argtypes = [C.POINTER(EngineCfg), C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        l.pp_engine_destroy.argtypes = [C.c_void_p]
        l.pp_engine_destroy.restype = None
        l.pp_engine_load.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.pp_engine_finalize.argtypes = [C.c_void_p, C.c_void_p]
        l.pp_engine_backbone.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        l.pp_engine_head.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        l.pp_engine_infer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        l.pp_engine_last_launch_count.restype = C.c_int64
        l.pp_engine_last_launch_count.argtypes = [C.c_void_p]
        l.pp_engine_set_graph.argtypes = [C.c_void_p, C.c_int32]
        l.pp_engine_graph_replay_count.restype = C.c_int64
        l.pp_engine_graph_replay_count.argtypes = [C.c_void_p]
        l.pp_engine_profile_begin.argtypes = [C.c_void_p]
        l.pp_engine_profile_end.argtypes = [C.c_void_p, C.POINTER(Profile), C.c_void_p]
    _lib = l
    return l


def check(status: int, what: str = "") -> None:
    """Map a pp_status to the exception the reference's Python layer would raise."""
    if status == 0:
        return
    msg = lib().pp_last_error().decode(errors="replace")
    if status == -1:
        raise ValueError(f"{what}: {msg}")
    raise PPError(f"{what}: status {status}: {msg}")
