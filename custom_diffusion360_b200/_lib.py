"""ctypes binding of libcd360.so (the C ABI in include/cd360.h).

No torch types cross the boundary: wrappers in `ops.py` pass `tensor.data_ptr()` integers, sizes
and the raw `cudaStream_t` of torch's current stream.  The product path never falls back to CPU or
to torch ops: if the library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcd360.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "cd360.h")

OK = 0
ACT_NONE, ACT_SILU, ACT_GELU, ACT_QUICK_GELU = 0, 1, 2, 3


class Cd360Error(RuntimeError):
    pass


class GemmArgs(C.Structure):
    """Mirror of `cd360_gemm_args` (include/cd360.h)."""

    _fields_ = [
        ("a0", C.c_void_p), ("lda0", C.c_int64), ("k0", C.c_int32),
        ("a1", C.c_void_p), ("lda1", C.c_int64), ("k1", C.c_int32),
        ("w", C.c_void_p),
        ("bias", C.c_void_p),
        ("row_bias", C.c_void_p), ("rows_per_group", C.c_int32), ("ld_row_bias", C.c_int64),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_fp32", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32),
        ("conv", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("act", C.c_int32), ("geglu", C.c_int32), ("block_n", C.c_int32), ("max_ctas", C.c_int32),
        ("ln_stats", C.c_void_p), ("ln_slabs", C.c_int32), ("ln_eps", C.c_float),
        ("ln_colsum", C.c_void_p), ("stats_out", C.c_void_p), ("k_splits", C.c_int32), ("split_stride", C.c_int64),
        ("tn", C.c_int32), ("ldw", C.c_int64),
    ]


_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol declared in include/cd360.h
SIGNATURES = {
    "cd360_abi_version": (C.c_int, []),
    "cd360_strerror": (C.c_char_p, [C.c_int]),
    "cd360_gemm_bf16": (C.c_int, [C.POINTER(GemmArgs), _P]),
    "cd360_geglu_pack_block": (C.c_int, [_I]),
    "cd360_attention_bf16": (C.c_int, [_P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _P]),
    "cd360_groupnorm_workspace_floats": (C.c_int64, [_I, _I]),
    "cd360_groupnorm_silu_bf16": (C.c_int, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _F, _I, _P]),
    "cd360_layernorm_bf16": (C.c_int, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "cd360_small_linear": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "cd360_timestep_embedding": (C.c_int, [_P, _P, _I, _I, _P]),
    "cd360_im2col3x3_nchw_f32": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "cd360_im2col3x3_s2_bf16": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "cd360_upsample_nearest2x_bf16": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "cd360_cfg_euler_step": (C.c_int, [_P, _P, _P, _I, _I, _I, _F, _F, _F, _F, _F, _P]),
    "cd360_cfg_euler_step_dev": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _F, _F, _P]),
    "cd360_nerf_points": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "cd360_nerf_combine": (C.c_int, [_P, _L, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "cd360_nerf_volrender": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "cd360_nerf_mask_ref": (C.c_int, [_P, _P, _P, _L, _I, _I, _I, _I, _P]),
    "cd360_embed_tokens": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "cd360_attention_causal_bf16": (C.c_int, [_P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _I, _P]),
    "cd360_gather_rows_bf16_f32": (C.c_int, [_P, _L, _P, _P, _I, _I, _L, _P]),
    "cd360_cast_f32_to_bf16": (C.c_int, [_P, _P, _L, _P]),
    "cd360_cast_bf16_to_f32": (C.c_int, [_P, _P, _L, _P]),
    "cd360_splitk_slices": (C.c_int, [_I, _I, _I]),
    "cd360_splitk_finish": (C.c_int, [_P, _L, _L, _I, _P, _P, _L, _P, _L, _I, _L, _I, _P]),
    "cd360_pointwise_conv_nchw_f32": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _L, C.c_float, _P]),
    "cd360_softmax_rows_f32_bf16": (C.c_int, [_P, _L, _P, _L, _L, _I, C.c_float, _P]),
    "cd360_nhwc_to_nchw_f32": (C.c_int, [_P, _I, _P, _I, _I, _I, _P]),
    "cd360_nchw_f32_to_nhwc_bf16": (C.c_int, [_P, _P, _I, _I, _I, _P]),
    # training step (backward / loss / optimiser)
    "cd360_attention_bwd_bf16": (C.c_int, [_P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L,
                                           _P, _P, _I, _I, _I, _I, _P]),
    "cd360_attention_bwd_kv_split_bf16": (C.c_int, [_P, _L, _P, _L, _P, _L, _P, _L, _P, _P, _P, _P, _L, _P, _L,
                                                    _I, _I, _I, _I, _I, _P]),
    "cd360_layernorm_bwd_bf16": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "cd360_groupnorm_bwd_workspace_floats": (C.c_int64, [_I]),
    "cd360_groupnorm_silu_bwd_bf16": (C.c_int, [_P, _I, _P, _I, _P, _P, _P, _P, _L, _P, _L, _P, _P, _P,
                                                _I, _I, _F, _I, _P]),
    "cd360_geglu_bwd_bf16": (C.c_int, [_P, _P, _P, _L, _I, _I, _P]),
    "cd360_geglu_fwd_bf16": (C.c_int, [_P, _P, _L, _I, _I, _P]),
    "cd360_add_bf16": (C.c_int, [_P, _P, _P, _L, _P]),
    "cd360_silu_bwd_f32": (C.c_int, [_P, _P, _P, _L, _P]),
    "cd360_transpose_to_bf16": (C.c_int, [_P, _I, _L, _P, _L, _I, _I, _P]),
    "cd360_colsum_bf16": (C.c_int, [_P, _L, _P, _L, _I, _P]),
    "cd360_col2im3x3_s2_bf16": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "cd360_upsample_nearest2x_bwd_bf16": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "cd360_nerf_volrender_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "cd360_nerf_combine_bwd": (C.c_int, [_P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "cd360_nerf_nviews_geo_bwd": (C.c_int, [_P, _P, _P, _I, _I, _L, _P]),
    "cd360_diffusion_loss": (C.c_int, [_P, _P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _P]),
    "cd360_nerf_aux_loss": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "cd360_resize_bilinear_aa": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _F, _F, _P]),
    "cd360_adamw_step": (C.c_int, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P]),
}

_lib = None


def header_symbols() -> list[str]:
    """Every function name declared in include/cd360.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cd360_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libcd360.so, building it in-tree first when absent or stale and nvcc is available."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import _build

        try:
            if _build.is_stale():
                _build.build()
        except Exception as e:  # stale-but-present library still loads; missing one is fatal
            if not os.path.exists(LIB_PATH):
                raise Cd360Error(f"libcd360.so missing and could not be built: {e}") from e
    if not os.path.exists(LIB_PATH):
        raise Cd360Error(
            f"{LIB_PATH} not found: the CUDA extension is required (there is no CPU fallback); "
            "run `python __graft_entry__.py build`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != OK:
        msg = load().cd360_strerror(code).decode()
        raise Cd360Error(f"{what} failed: {msg} ({code})")
