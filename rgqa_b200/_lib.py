"""ctypes binding of ``librgqa_b200.so`` (the C ABI declared in ``include/rgqa_b200.h``).

There is no CPU fallback anywhere in this package: if the library is missing or a call fails the
caller gets a ``RuntimeError`` carrying ``rgqa_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
ABI_VERSION = 3  # include/rgqa_b200.h: RGQA_ABI_VERSION
LIB_PATH = _PKG / "_C" / "librgqa_b200.so"

# constants mirrored from include/rgqa_b200.h
ARCH_LXMERT, ARCH_UNITER = 0, 1
PREC_FP32, PREC_BF16 = 0, 1
DT_F32, DT_BF16, DT_F16 = 0, 1, 2
EPI_BIAS, EPI_GELU, EPI_TANH, EPI_RESID = 0, 1, 2, 3
SCORE_MSP, SCORE_ENERGY_TOP2, SCORE_ENERGY_SUM, SCORE_LSE, SCORE_MAHA = 0, 1, 2, 3, 4
SCORE_MODES = {"msp": SCORE_MSP, "energy": SCORE_ENERGY_TOP2, "energy_top2": SCORE_ENERGY_TOP2,
               "energy_sum": SCORE_ENERGY_SUM, "lse": SCORE_LSE, "maha": SCORE_MAHA}


class RgqaConfig(C.Structure):
    _fields_ = [
        ("arch", C.c_int32), ("l_layers", C.c_int32), ("x_layers", C.c_int32), ("r_layers", C.c_int32),
        ("hidden", C.c_int32), ("heads", C.c_int32), ("intermediate", C.c_int32), ("vocab_size", C.c_int32),
        ("max_position", C.c_int32), ("type_vocab", C.c_int32), ("visual_feat_dim", C.c_int32),
        ("visual_pos_dim", C.c_int32), ("num_rois", C.c_int32), ("num_answers", C.c_int32),
        ("precision", C.c_int32), ("ln_eps", C.c_float),
    ]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
SIGNATURES = {
    "rgqa_abi_version": (_i, []),
    "rgqa_last_error": (C.c_char_p, []),
    "rgqa_create": (_i, [C.POINTER(RgqaConfig), C.POINTER(_vp)]),
    "rgqa_destroy": (None, [_vp]),
    "rgqa_num_weights": (_i, [_vp]),
    "rgqa_weight_name": (C.c_char_p, [_vp, _i]),
    "rgqa_load_weight": (_i, [_vp, C.c_char_p, _vp, _i, C.POINTER(_i64), _i]),
    "rgqa_finalize_weights": (_i, [_vp]),
    "rgqa_lxmert_forward": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "rgqa_uniter_forward": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "rgqa_forward_scored": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, C.c_uint, _vp, _vp, _vp, _vp]),
    "rgqa_set_maha": (_i, [_vp, _vp, _vp, _i]),
    "rgqa_score": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "rgqa_set_option": (_i, [C.c_char_p, _i]),
    "rgqa_launch_count": (C.c_longlong, []),
    "rgqa_graph_captures": (C.c_longlong, [_vp]),
    "rgqa_fold_residue": (_f, [_vp, C.POINTER(_i)]),
    "rgqa_gemm_bf16_reduce_logits": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp, _i64, _vp, _vp]),
    "rgqa_reduce_work_bytes": (_i64, [_i, _i]),
    "rgqa_qkv_attention_fused": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _f, _vp, _vp, _i64, _i, _i, _i]),
    "rgqa_visn_embed": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i]),
    "rgqa_prof_enable": (None, [_i]),
    "rgqa_prof_collect": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "rgqa_gemm_bf16": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i]),
    "rgqa_gemm_bf16_ln": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _f, _vp, _vp, _i, _i, _i]),
    "rgqa_fold_ln_weight": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "rgqa_gemm_bf16_folded": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _vp, _vp, _f]),
    "rgqa_gemm_bf16_resid_hl": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i]),
    "rgqa_fold_ln_weight_f16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "rgqa_gemm_bf16_folded_a16": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp, _vp, _f]),
    "rgqa_qkv_attention_fused_a16": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _f, _vp, _vp, _i64, _i, _i, _i]),
    "rgqa_gemm_bf16_resid_h16": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _i, _i]),
    "rgqa_gemm_f32": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i]),
    "rgqa_attention": (_i, [_vp, _i, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i]),
    "rgqa_embed_ln": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f]),
    "rgqa_layernorm": (_i, [_vp, _i, _vp, _i64, _vp, _vp, _vp, _i, _i, _f]),
    "rgqa_visn_combine": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f]),
    "rgqa_cast_f32_to_bf16": (_i, [_vp, _vp, _vp, _i64]),
    "rgqa_gelu_fwd_bf16": (_i, [_vp, _vp, _vp, _i64]),
    "rgqa_layernorm_dual": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i, _i, _f]),
    "rgqa_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f]),
    "rgqa_gelu_fwd": (_i, [_vp, _vp, _vp, _i64]),
    "rgqa_gelu_bwd": (_i, [_vp, _vp, _vp, _vp, _i64]),
    "rgqa_tanh_bwd": (_i, [_vp, _vp, _vp, _vp, _i64]),
    "rgqa_attention_bwd_bf16": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i]),
    "rgqa_attention_bwd": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _f, C.c_uint64]),
    "rgqa_attention_dropout": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _f, C.c_uint64]),
    "rgqa_transpose_cast_bf16": (_i, [_vp, _vp, _i64, _vp, _i, _i]),
    "rgqa_gemm_bf16_accum_tn": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i]),
    "rgqa_cast_bf16_colsum": (_i, [_vp, _vp, _i64, _vp, _i64, _i, _i, _vp]),
    "rgqa_gemm_bf16_accum": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i]),
    "rgqa_transpose_cast_bf16_colsum": (_i, [_vp, _vp, _i64, _vp, _i, _i, _vp]),
    "rgqa_layernorm_bwd_params": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f]),
    "rgqa_sign_step": (_i, [_vp, _vp, _vp, _vp, _i64, _f]),
    "rgqa_dropout_add": (_i, [_vp, _vp, _vp, _vp, _i64, _f, C.c_uint64]),
    "rgqa_gemm_f32_tn": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i]),
    "rgqa_colsum": (_i, [_vp, _vp, _i64, _vp, _i, _i]),
    "rgqa_ln_param_grad": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f]),
    "rgqa_embedding_scatter": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "rgqa_bce_loss_grad": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "rgqa_grad_sumsq": (_i, [_vp, _vp, _i64, _vp]),
    "rgqa_bertadam_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, C.c_double, C.c_double, _f, _f, _vp, _f]),
    "rgqa_estimator_accumulate": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "rgqa_score_ensemble": (_i, [_vp, C.POINTER(_vp), _i, _i, _i64, _i, _i, _vp, _vp]),
    "rgqa_score_logits": (_i, [_vp, _i, _vp, _i64, _i, _i, _vp, _vp]),
    "rgqa_mahalanobis": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
}

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load the shared library (building it with nvcc first if it is absent and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and os.environ.get("RGQA_B200_NO_BUILD") != "1":
        # a library older than its sources would be called with a stale ABI: rebuild (a no-op when it is up to date;
        # skipped when there is no nvcc and a library exists, e.g. a deployment box that received the prebuilt .so)
        from . import build as _build
        if not LIB_PATH.exists() or (_build.have_nvcc() and _build.stale()):
            _build.build()
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"rgqa_b200: CUDA extension {LIB_PATH} is missing; run `python -m rgqa_b200.build` "
            "(this package has no CPU or PyTorch fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    got = lib.rgqa_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"rgqa_b200: {LIB_PATH} has ABI version {got}, the Python binding expects {ABI_VERSION}; "
                           "rebuild with `python -m rgqa_b200.build --force`")
    # RGQA_OPTS=name=value,... : kernel-selection switches for A/B measurements (rgqa_set_option), applied once at load
    for kv in filter(None, os.environ.get("RGQA_OPTS", "").split(",")):
        k, _, v = kv.partition("=")
        if lib.rgqa_set_option(k.strip().encode(), int(v)) != 0:
            raise RuntimeError(f"rgqa_b200: RGQA_OPTS: {lib.rgqa_last_error().decode('utf-8', 'replace')}")
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().rgqa_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"rgqa_b200: {what} failed: {last_error()}")


def ptr(t) -> int:
    """Raw data pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
