"""ctypes loader for libmingb200.so — the only place the C ABI (include/mingb200.h) is bound.

Every entry point is declared with explicit argtypes so a signature drift between the header and the Python side
fails at load time (tests/test_abi.py additionally checks every symbol the header declares is exported).
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmingb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mingb200.h")

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes  (restype is int for all but mb_last_error)
SIGNATURES = {
    "mb_abi_version": [],
    "mb_device_ok": [],
    "mb_num_sms": [],
    "mb_gemm_bf16": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _i64, _i, _i, _i, _vp],
    "mb_gemm_bf16_ex": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _i64, _i, _i, _i, _vp, _i, _vp, _vp,
                        _f, _vp, _vp],
    "mb_row_stats": [_vp, _i64, _vp, _i, _i, _vp],
    "mb_gemm_force_tile": [_i, _i],
    "mb_attn_set_backend": [_i],
    "mb_attn_set_debug": [_vp],
    "mb_gemm_set_debug": [_vp],
    "mb_gemv_bf16": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, _vp],
    "mb_gemv_bf16_norm": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, _i, _vp, _vp,
                          _vp, _i64, _vp, _i64, _f, _vp],
    "mb_adaln_modulate": [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _f, _vp],
    "mb_silu_add_rows": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "mb_rf_euler_step": [_vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _vp],
    "mb_rf_fused_supported": [_i, _i, _i, _i],
    "mb_rf_set_debug": [_vp],
    "mb_rf_pack_weights": [_vp, _i, _i, _i, _i, _vp, _vp],
    "mb_rf_sample_fused": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _f,
                           _i, _vp],
    "mb_rmsnorm": [_vp, _i64, _vp, _vp, _i64, _i, _i, _f, _vp],
    "mb_rope_kv_append": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _f, _vp],
    "mb_rope3d_kv_append": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _f, _i, _i, _i, _vp],
    "mb_attn_decode_gqa": [_vp, _vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _vp, _i, _f, _vp, _i, _vp],
    "mb_argmax_f32": [_vp, _vp, _i, _i, _vp, _vp, _i, _vp],
    "mb_router_topk": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "mb_moe_sort": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "mb_moe_gate_up": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "mb_moe_down": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "mb_moe_ffn_workspace_bytes": [_i, _i, _i, _i, _i, _vp],
    "mb_moe_ffn": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mb_moe_combine": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "mb_moe_plan": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mb_moe_gather_rows": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "mb_moe_grouped_gemm": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "mb_moe_finalize": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "mb_moe_peer_area_bytes": [_i, _i, _i, _vp],
    "mb_moe_combine_push": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "mb_moe_reduce_finalize": [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "mb_ep_area_layout": [_i, _i, _i, _i, _vp],
    "mb_ep_dispatch_push": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "mb_ep_dispatch_wait": [_vp, _i, _i, _i, _i, _i, _vp],
    "mb_ep_wait_sort": [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp],
    "mb_ep_combine_push": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mb_ep_reduce_finalize": [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "mb_image_preprocess_workspace_bytes": [_i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mb_image_preprocess_u8": [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f, _vp, _i, _vp, _i64,
                               _vp],
    "mb_image_postprocess_u8": [_vp, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f, _vp, _vp],
    "mb_unpatchify_to_u8": [_vp, _vp, _i, _i, _i, _f, _f, _f, _f, _f, _f, _vp],
    "mb_pack_swiglu_rows": [_vp, _vp, _i, _i, _i, _vp],
    "mb_layernorm": [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _f, _i, _i, _i64, _vp],
    "mb_attn_hd64": [_vp, _vp, _i, _i, _i, _f, _i, _vp],
    "mb_attn_fwd": [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i, _i,
                    _i, _i, _i, _i, _f, _i, _vp],
    "mb_attn_hd64_decode": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp],
    "mb_patchify": [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "mb_fill_cls_row": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "mb_group_mean": [_vp, _i64, _vp, _i, _i, _i, _vp],
    "mb_affine": [_vp, _i, _vp, _i64, _f, _f, _vp],
    "mb_inproj_repeat": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "mb_pixel_shuffle": [_vp, _vp, _i, _i, _i, _i, _vp],
    "mb_unpatchify_clamp": [_vp, _vp, _i, _i, _i, _i, _vp],
}

_lib = None
_launches = 0  # successful C-ABI calls == kernel launches issued by this process (each entry point launches one)


def launch_count() -> int:
    return _launches


def count_replay(n_kernels: int) -> None:
    """A CUDA-graph replay launches the kernels that were captured into it: the owners of the graphs record how many
    C-ABI calls the capture made and add them here on every replay, so launch_count() keeps meaning "kernels of this
    library launched by this process"."""
    global _launches
    _launches += int(n_kernels)


def header_symbols() -> list[str]:
    """Function names declared in include/mingb200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    """Loads libmingb200.so (built by ``__graft_entry__.build()`` / ``make -C ming_univision_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the ming_univision_b200 operators)")
    lib = C.CDLL(LIB_PATH)
    lib.mb_last_error.restype = C.c_char_p
    lib.mb_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    global _launches
    _launches += 1
    if rc != 0:
        msg = load().mb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with code {rc}: {msg}")


def require_device() -> None:
    lib = load()
    if not lib.mb_device_ok():
        raise RuntimeError("ming_univision_b200 needs an sm_100 (B200) CUDA device; there is no CPU fallback")
