"""ctypes binding of libchadavit_b200.so (the C ABI declared in include/chadavit_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get("CB_VARIANT", "")   # debug builds only (chadavit_b200/build.py); the product library has no suffix
LIB_PATH = os.path.join(_HERE, "lib", "libchadavit_b200" + ("_" + _VARIANT if _VARIANT else "") + ".so")

_vp, _i, _f, _l = C.c_void_p, C.c_int, C.c_float, C.c_long

# name -> argtypes (all return int unless noted).  Must mirror include/chadavit_b200.h exactly;
# tests/test_abi.py checks that every symbol declared in the header is listed here and exported.
SIGNATURES = {
    "cb_version": [],
    "cb_num_sms": [],
    "cb_sync_check": [_vp],
    "cb_attn_schedule": [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "cb_gemm_bf16": [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _f, _i, _vp, _vp],
    "cb_gemm_ln_fwd": [_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "cb_ffn_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "cb_ffn_bwd": [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp],
    "cb_im2col_bf16": [_vp, _vp, _i, _i, _i, _i, _vp],
    "cb_tokenize_fwd": [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "cb_tokenize_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "cb_small_matmul_f32": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "cb_layernorm_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp],
    "cb_layernorm2_fwd": [_vp, _vp, _vp, _f, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "cb_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],  # dy x idx g mean rstd dres dx32 dx16 dg db dc
    "cb_colsum_bf16": [_vp, _i, _vp, _i, _i, _vp],
    "cb_cast_f32_bf16": [_vp, _vp, _l, _vp],
    "cb_gather_rows_f32": [_vp, _vp, _vp, _i, _i, _vp],
    "cb_attn_varlen_fwd": [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _f, _vp],
    "cb_attn_varlen_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _f, _vp],
    "cb_gelu_fwd": [_vp, _vp, _l, _vp],
    "cb_gelu_bwd": [_vp, _vp, _vp, _l, _vp],
    "cb_bn_gelu_fwd": [_vp, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "cb_bn_gelu_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp],
    "cb_l2norm_fwd": [_vp, _vp, _vp, _i, _i, _f, _vp],
    "cb_l2norm_bwd": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "cb_weightnorm_fwd": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "cb_weightnorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "cb_dino_loss_fwd_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _vp],
    "cb_colsum_f32": [_vp, _vp, _i, _i, _vp],
    "cb_dino_center_ema": [_vp, _vp, _f, _f, _i, _vp],
    "cb_ema_update": [_vp, _vp, _vp, _f, _l, _vp],
    "cb_adamw_step": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _l, _f, _f, _f, _f, _f, _i, _i, _f, _f, _vp, _vp],
    "cb_attn_probs": [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp],
    "cb_attn_cls_fwd": [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp],
    "cb_attn_cls_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp],
    "cb_split_bf16x3": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "cb_inv_euclid": [_vp, _vp, _vp, _i, _i, _i, _f, _vp],
    "cb_param_norms": [_vp, _vp, _vp, _vp, _vp, _vp, _l, _i, _f, _f, _vp],
    "cb_scale_grads": [_vp, _vp, _vp, _l, _vp],
    "cb_lars_step": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _l, _f, _f, _f, _i, _f, _f, _f, _i, _f, _f, _vp, _vp],
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (raises RuntimeError with build instructions if it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library must be built (python -m chadavit_b200.build). "
            "chadavit_b200 has no CPU / PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.cb_last_error.restype = C.c_char_p
    lib.cb_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = lib
    return lib


class CudaLibError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cb_last_error().decode(errors="replace")
        raise CudaLibError(f"{what} failed (rc={rc}): {msg}")


# launch counter: bench.py reports how many of OUR kernels ran in the timed region
launch_count = 0
