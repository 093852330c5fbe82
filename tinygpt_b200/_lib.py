"""ctypes binding of include/b200_decode.h.  Fails loudly when the CUDA library is missing — there is no fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

LIB_PATH = Path(os.environ.get("B200_LIB") or (Path(__file__).resolve().parent / "lib" / "libb200decode.so"))
IPC_HANDLE_BYTES = 64

_lib = None


class B200Error(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("layers", C.c_int32), ("q_heads", C.c_int32), ("kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("intermediate", C.c_int32), ("vocab", C.c_int32), ("max_ctx", C.c_int32),
        ("rms_eps", C.c_float), ("qkv_bias", C.c_int32), ("qk_norm", C.c_int32),
        ("tp_rank", C.c_int32), ("tp_world", C.c_int32), ("tp_shard_attn", C.c_int32),
    ]


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("input_norm", "qkv_w", "qkv_b", "q_norm", "k_norm", "o_w", "post_norm", "gate_up_w", "down_w")]


class WeightTable(C.Structure):
    _fields_ = [("embed", C.c_void_p), ("final_norm", C.c_void_p), ("lm_head", C.c_void_p),
                ("rope_table", C.c_void_p), ("layers_host", C.POINTER(LayerWeights))]


# name -> (restype, argtypes); every symbol include/b200_decode.h declares (tests/test_abi.py checks the two agree)
P, I64, I32, F = C.c_void_p, C.c_int64, C.c_int, C.c_float
PROTOTYPES = {
    "b200_abi_version": (I32, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_launch_count": (I64, []),
    "b200_device_check": (I32, []),
    "b200_gemv_bf16": (I32, [P, P, P, P, I64, I64, I64, P]),
    "b200_gemm_bf16": (I32, [P, P, P, I64, I64, I64, P]),
    "b200_rmsnorm_bf16": (I32, [P, P, P, I64, I64, F, P]),
    "b200_rope_bf16": (I32, [P, P, P, I64, I64, I64, I64, I64, I32, P]),
    "b200_rope_init_f32": (I32, [P, I64, I64, F, F, F, F, I64, P]),
    "b200_attn_bf16": (I32, [P, P, P, P, I64, I64, I64, I64, I64, I64, I32, P]),
    "b200_silu_mul_bf16": (I32, [P, P, I64, I64, P]),
    "b200_add_bf16": (I32, [P, P, P, I64, P]),
    "b200_embedding_bf16": (I32, [P, P, P, I64, I64, I64, P]),
    "b200_argmax_workspace_bytes": (I64, [I64, I64]),
    "b200_argmax_bf16": (I32, [P, P, I64, I64, P, P]),
    "b200_sample_workspace_bytes": (I64, []),
    "b200_sample_bf16": (I32, [P, P, I64, F, I64, F, F, F, P, P]),
    "b200_gemv_fused_bf16": (I32, [P, P, P, I64, I64, I32, P, F, P, P, I32, P]),
    "b200_attn_decode_workspace_bytes": (I64, [I64, I64, I64, I64]),
    "b200_attn_decode_bf16": (I32, [P, P, P, P, F, P, P, I64, P, P, I64, I64, I64, I64, P, P]),
    "b200_engine_create": (I32, [C.POINTER(ModelDesc), C.POINTER(WeightTable), C.POINTER(P)]),
    "b200_engine_destroy": (None, [P]),
    "b200_engine_reset": (I32, [P, P]),
    "b200_engine_seek": (I32, [P, I64, P]),
    "b200_engine_forward": (I32, [P, P, I64, I64, P, I32, P]),
    "b200_engine_decode": (I32, [P, I64, P, P]),
    "b200_engine_last_token": (I32, [P, P, P]),
    "b200_engine_debug_trace": (I64, [P, P, I64]),
    "b200_engine_position": (I64, [P]),
    "b200_engine_generated": (I64, [P]),
    "b200_engine_set_mailbox": (I32, [P, P, I64, P]),
    "b200_engine_set_sampler": (I32, [P, F, I64, F, F, C.c_uint64, P]),
    "b200_engine_launches_per_token": (I64, [P]),
    "b200_engine_options": (I64, [P]),
    "b200_engine_bytes_per_token": (I64, [P, I64]),
    "b200_tp_window_bytes": (I64, [C.POINTER(ModelDesc)]),
    "b200_tp_window_create": (I32, [I64, C.POINTER(P), C.c_char_p]),
    "b200_tp_window_open": (I32, [C.c_char_p, C.POINTER(P)]),
    "b200_tp_window_close": (I32, [P]),
    "b200_tp_window_destroy": (I32, [P]),
    "b200_engine_create_tp": (I32, [C.POINTER(ModelDesc), C.POINTER(WeightTable), C.POINTER(P), C.POINTER(P)]),
}


def lib() -> C.CDLL:
    """Load lib/libb200decode.so (built by `python -m tinygpt_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise B200Error(f"{LIB_PATH} is missing: build it with `python -m tinygpt_b200.build` "
                            "(nvcc, sm_100a). There is no CPU fallback for the decode path.")
        h = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)  # AttributeError if the library does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().b200_last_error().decode(errors="replace")
        raise B200Error(f"{what or 'b200 call'} failed with status {rc}: {msg}")


def require_device() -> None:
    """Raise unless an sm_100 device is current (the product path never silently degrades)."""
    check(lib().b200_device_check(), "b200_device_check")
