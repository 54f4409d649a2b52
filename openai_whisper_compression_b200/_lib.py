"""ctypes loader for libwhisperq.so (the C ABI declared in include/whisperq.h).

The product path has NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  The library is never built implicitly at import time on a GPU box;
run ``python -m openai_whisper_compression_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# WQ_LIB_PATH: an alternative build of the same library (A/B experiments with compile-time variants)
LIB_PATH = os.environ.get("WQ_LIB_PATH") or os.path.join(_HERE, "libwhisperq.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPE); mirrors include/whisperq.h
_PROTOS = {
    "wq_version": [],
    "wq_device_info": [c_ptr, c_ptr, c_ptr],
    "wq_quant_4bit": [c_ptr, c_int, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr],
    "wq_dequant_4bit": [c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr, c_int, c_ptr],
    "wq_quant_absmax_double": [c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_dequant_absmax_double": [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr],
    "wq_quant_i8_rowwise_bnb": [c_ptr, c_i64, c_i64, c_f32, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_outlier_columns": [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_quant_i8_rowwise_quanto": [c_ptr, c_int, c_i64, c_i64, c_ptr, c_ptr, c_ptr],
    "wq_quant_u4_group_quanto": [c_ptr, c_int, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_quant_ubits_group_quanto": [c_ptr, c_int, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_quant_i8_tensor_torch": [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_quant_act_u8_tensor": [c_ptr, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_gemm_llmint8": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr],
    "wq_gemm_llmint8_shared": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_int,
                               c_ptr],
    "wq_gemm_llmint8_residual": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_int,
                                 c_ptr, c_f32, c_int, c_ptr],
    "wq_linear_llmint8_small": [c_ptr, c_i64, c_i64, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr],
    "wq_gemm_w8a16": [c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_i64, c_i64, c_i64, c_ptr],
    "wq_gemm_w4a16": [c_ptr, c_int, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_int, c_i64, c_i64, c_i64, c_ptr],
    "wq_gemm_u4a16": [c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_int, c_i64, c_i64, c_i64, c_ptr],
    "wq_gemm_dyn_i8": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr],
    "wq_add_layernorm_quant": [c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_f32, c_i64, c_i64, c_ptr, c_ptr, c_f32, c_ptr,
                               c_ptr, c_ptr, c_ptr],
    "wq_gelu_quant": [c_ptr, c_int, c_i64, c_i64, c_ptr, c_f32, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_self_attn_decode": [c_ptr, c_ptr, c_ptr, c_i64, c_int, c_f32, c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr,
                            c_ptr, c_f32, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_cross_attn_decode": [c_ptr, c_i64, c_int, c_f32, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_int, c_ptr, c_f32,
                             c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_decode_fused_llmint8": [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr],
    "wq_masked_argmax": [c_ptr, c_int, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr],
    "wq_quant_f8_rowwise_quanto": [c_ptr, c_int, c_i64, c_i64, c_ptr, c_ptr, c_ptr],
    "wq_gemm_wf8a16": [c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_i64, c_i64, c_i64, c_ptr],
    "wq_quant_act_static": [c_ptr, c_int, c_i64, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr],
    "wq_gemm_w8a8": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_i64, c_i64, c_i64, c_ptr],
    "wq_scatter_dense_f32": [c_ptr, c_ptr, c_int, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr],
    "wq_gemv_weightonly": [c_ptr, c_int, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr, c_i64, c_ptr],
    "wq_gemm_f16": [c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_int, c_i64, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_ptr],
    "wq_argmax_finalize": [c_ptr, c_i64, c_ptr, c_ptr],
    "wq_logmel": [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_int, c_ptr, c_int, c_ptr, c_ptr],
    "wq_edit_distance": [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr],
}
EXPORTS = tuple(["wq_last_error", *_PROTOS.keys()])


class DecodeLinear(ctypes.Structure):       # wq_decode_linear
    _fields_ = [("cb", c_ptr), ("scb", c_ptr), ("bias", c_ptr), ("N", c_int), ("K", c_int)]


class DecodeLayer(ctypes.Structure):        # wq_decode_layer
    _fields_ = [("qkv", DecodeLinear), ("o", DecodeLinear), ("cq", DecodeLinear), ("co", DecodeLinear),
                ("fc1", DecodeLinear), ("fc2", DecodeLinear),
                ("ln1_g", c_ptr), ("ln1_b", c_ptr), ("ln2_g", c_ptr), ("ln2_b", c_ptr), ("ln3_g", c_ptr), ("ln3_b", c_ptr),
                ("eps1", c_f32), ("eps2", c_f32), ("eps3", c_f32), ("kcache", c_ptr), ("vcache", c_ptr)]


class DecodeArgs(ctypes.Structure):         # wq_decode_args
    _fields_ = [("M", c_int), ("d", c_int), ("ffn", c_int), ("H", c_int), ("t_max", c_int),
                ("threshold", c_f32), ("scaling", c_f32), ("pos", c_ptr), ("x", c_ptr),
                ("h", c_ptr), ("att", c_ptr), ("qkv", c_ptr), ("f1", c_ptr), ("g", c_ptr),
                ("ca_d", c_ptr), ("ca_f", c_ptr), ("sca", c_ptr), ("flags", c_ptr), ("q_out", c_ptr),
                ("xa", c_ptr), ("xa_ca", c_ptr), ("xa_sca", c_ptr), ("xa_flags", c_ptr),
                ("lnf_g", c_ptr), ("lnf_b", c_ptr), ("epsf", c_f32), ("hfinal", c_ptr), ("bar", c_ptr)]

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
            "(run `python -m openai_whisper_compression_b200.build`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.wq_last_error.restype = ctypes.c_char_p
    lib.wq_last_error.argtypes = []
    for name, argtypes in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().wq_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libwhisperq {what} failed (status {rc}): {msg}")
