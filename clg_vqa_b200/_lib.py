"""ctypes binding of libclgvqa.so (include/clgvqa.h).  Loads loudly: a missing library or a
non-sm_100 device raises - there is no CPU / eager fallback in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libclgvqa.so")

CLG_F32, CLG_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH, ACT_DGELU, ACT_DRELU, ACT_DTANH, ACT_GELU_GRAD, ACT_MUL_AUX = range(9)
GEMM_AUTO, GEMM_TC, GEMM_SIMT = 0, 1, 2
CHUNK = 2048

c_i32, c_i64, c_f32, c_vp, c_u64, c_u32 = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_uint64, C.c_uint32


class Dropout(C.Structure):
    _fields_ = [("p", c_f32), ("_pad", c_u32), ("seed", c_u64), ("offset", c_u64)]


class RowMap(C.Structure):
    _fields_ = [("group_rows", c_i32), ("row_offset", c_i32), ("group_stride", c_i64)]


class GemmParams(C.Structure):
    _fields_ = [
        ("A", c_vp), ("B", c_vp), ("C", c_vp),
        ("a_dtype", c_i32), ("b_dtype", c_i32), ("c_dtype", c_i32),
        ("M", c_i32), ("N", c_i32), ("K", c_i32),
        ("lda", c_i64), ("ldb", c_i64), ("ldc", c_i64),
        ("a_kmajor", c_i32), ("b_kmajor", c_i32),
        ("alpha", c_f32), ("act", c_i32),
        ("bias", c_vp), ("residual", c_vp), ("ld_res", c_i64),
        ("aux_in", c_vp), ("ld_aux_in", c_i64),
        ("aux_in_dtype", c_i32), ("aux_out_dtype", c_i32),
        ("aux_out", c_vp), ("ld_aux_out", c_i64),
        ("mask_bits", c_vp),
        ("accumulate", c_i32), ("drop_before_act", c_i32),
        ("dropout", Dropout),
        ("backend", c_i32), ("split_k", c_i32),
        ("colsum", c_vp),
        ("A2", c_vp), ("B2", c_vp), ("C2", c_vp),
        ("lda2", c_i64), ("ldb2", c_i64), ("ldc2", c_i64),
        ("mask_bits2", c_vp),
        ("M2", c_i32), ("_pad2", c_i32),
    ]


class PruneTensor(C.Structure):
    _fields_ = [("w", c_vp), ("mask", c_vp), ("mask_out", c_vp), ("n", c_i64)]


class LayerNormFwdParams(C.Structure):
    _fields_ = [
        ("x", c_vp), ("ldx", c_i64), ("gamma", c_vp), ("beta", c_vp),
        ("y", c_vp), ("ldy", c_i64),
        ("y_lowp", c_vp), ("ldy_lowp", c_i64), ("lowp_dtype", c_i32), ("_pad0", c_i32),
        ("mean", c_vp), ("rstd", c_vp),
        ("rows", c_i32), ("N", c_i32), ("eps", c_f32), ("_pad1", c_i32),
        ("out_map", RowMap),
        ("add_mat", c_vp), ("ld_add", c_i64), ("add_map", RowMap),
        ("add_vec", c_vp), ("row_valid", c_vp),
        ("valid_pre", c_i32), ("valid_post", c_i32),
        ("dropout", Dropout),
    ]


class LayerNormBwdParams(C.Structure):
    _fields_ = [
        ("dy", c_vp), ("lddy", c_i64), ("dy_map", RowMap),
        ("x", c_vp), ("ldx", c_i64), ("gamma", c_vp), ("mean", c_vp), ("rstd", c_vp),
        ("dx", c_vp), ("lddx", c_i64),
        ("dx_lowp", c_vp), ("lddx_lowp", c_i64), ("lowp_dtype", c_i32), ("_pad0", c_i32),
        ("dgamma", c_vp), ("dbeta", c_vp), ("dbias", c_vp),
        ("rows", c_i32), ("N", c_i32),
        ("row_valid", c_vp), ("valid_pre", c_i32), ("valid_post", c_i32),
        ("dropout_out", Dropout), ("dropout_in", Dropout),
    ]


class AttentionFwdParams(C.Structure):
    _fields_ = [
        ("qkv", c_vp), ("ld_qkv", c_i64), ("key_bias", c_vp),
        ("ctx", c_vp), ("ld_ctx", c_i64), ("lse", c_vp),
        ("dtype", c_i32), ("B", c_i32), ("S", c_i32), ("H", c_i32),
        ("scale", c_f32), ("_pad", c_i32), ("dropout", Dropout),
        ("cu_rows", c_vp), ("total_rows", c_i32), ("_pad2", c_i32),
    ]


class AttentionBwdParams(C.Structure):
    _fields_ = [
        ("qkv", c_vp), ("ld_qkv", c_i64), ("key_bias", c_vp),
        ("ctx", c_vp), ("ld_ctx", c_i64), ("dctx", c_vp), ("ld_dctx", c_i64),
        ("lse", c_vp), ("dqkv", c_vp), ("ld_dqkv", c_i64),
        ("dtype", c_i32), ("B", c_i32), ("S", c_i32), ("H", c_i32),
        ("scale", c_f32), ("_pad", c_i32), ("dropout", Dropout),
        ("cu_rows", c_vp), ("total_rows", c_i32), ("_pad2", c_i32),
    ]


class EmbedTextParams(C.Structure):
    _fields_ = [
        ("ids", c_vp), ("seg", c_vp), ("word", c_vp), ("pos", c_vp), ("type", c_vp),
        ("out", c_vp), ("ld_out", c_i64), ("out_map", RowMap),
        ("B", c_i32), ("T", c_i32), ("N", c_i32), ("pad_id", c_i32),
        ("pos_mode", c_i32), ("pos_base", c_i32), ("vocab", c_i64),
    ]


class EmbedTextBwdParams(C.Structure):
    _fields_ = [
        ("ids", c_vp), ("seg", c_vp), ("dout", c_vp), ("ld_dout", c_i64), ("dout_map", RowMap),
        ("dword", c_vp), ("dpos", c_vp), ("dtype_tab", c_vp),
        ("B", c_i32), ("T", c_i32), ("N", c_i32), ("pad_id", c_i32),
        ("pos_mode", c_i32), ("pos_base", c_i32), ("vocab", c_i64),
        ("type_row0_has_total", c_i32), ("_pad", c_i32),
    ]


class GqaLossParams(C.Structure):
    _fields_ = [
        ("logits", c_vp), ("ld_logits", c_i64), ("target", c_vp), ("ld_target", c_i64),
        ("dist", c_vp), ("ld_dist", c_i64), ("dlogits", c_vp), ("ld_dlogits", c_i64),
        ("row_buf", c_vp), ("out", c_vp), ("pred", c_vp),
        ("B", c_i32), ("A", c_i32), ("topk", c_i32), ("lambda_", c_f32),
        ("gscale", c_f32), ("score_mean", c_i32),
    ]


class ColsumParams(C.Structure):          # clg_colsum's arguments as a struct (launch plans)
    _fields_ = [("x", c_vp), ("out", c_vp), ("ld", c_i64), ("dtype", c_i32), ("M", c_i32), ("N", c_i32), ("_pad", c_i32)]


class PlanEntry(C.Structure):             # one launch of a recorded plan (clg_run_plan)
    _fields_ = [("op", c_i32), ("off", c_i32), ("rows_off", c_i32 * 2), ("seed_off", c_i32 * 2)]


class TensorMeta(C.Structure):
    _fields_ = [("lr", c_f32), ("wd", c_f32), ("shadow_off", c_i64), ("mask_off", c_i64),
                ("start", c_i64), ("numel", c_i64)]


class AdamWParams(C.Structure):
    _fields_ = [
        ("p", c_vp), ("g", c_vp), ("m", c_vp), ("v", c_vp), ("n", c_i64),
        ("chunk_tensor", c_vp), ("tensors", c_vp), ("sumsq", c_vp),
        ("max_norm", c_f32), ("gscale", c_f32),
        ("beta1", c_f32), ("beta2", c_f32), ("eps", c_f32),
        ("bias_c1", c_f32), ("bias_c2", c_f32), ("zero_grad", c_i32),
        ("shadow", c_vp), ("mask_bits", c_vp),
        ("hot", c_vp),
    ]


_SIZEOF_CHECKS = {
    "clg_sizeof_gemm_params": GemmParams,
    "clg_sizeof_layernorm_fwd_params": LayerNormFwdParams,
    "clg_sizeof_layernorm_bwd_params": LayerNormBwdParams,
    "clg_sizeof_attention_fwd_params": AttentionFwdParams,
    "clg_sizeof_attention_bwd_params": AttentionBwdParams,
    "clg_sizeof_embed_text_params": EmbedTextParams,
    "clg_sizeof_embed_text_bwd_params": EmbedTextBwdParams,
    "clg_sizeof_gqa_loss_params": GqaLossParams,
    "clg_sizeof_adamw_params": AdamWParams,
    "clg_sizeof_tensor_meta": TensorMeta,
    "clg_sizeof_prune_tensor": PruneTensor,
    "clg_sizeof_plan_entry": PlanEntry,
    "clg_sizeof_colsum_params": ColsumParams,
}

# every symbol include/clgvqa.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "clg_strerror": (C.c_char_p, [C.c_int]),
    "clg_last_cuda_error": (C.c_char_p, []),
    "clg_version": (C.c_int, []),
    "clg_device_check": (C.c_int, [C.c_int]),
    "clg_num_sms": (C.c_int, [C.c_int]),
    "clg_gemm": (C.c_int, [C.POINTER(GemmParams), c_vp]),
    "clg_set_gemm_sm_limit": (C.c_int, [C.c_int]),
    "clg_layernorm_fwd": (C.c_int, [C.POINTER(LayerNormFwdParams), c_vp]),
    "clg_layernorm_bwd": (C.c_int, [C.POINTER(LayerNormBwdParams), c_vp]),
    "clg_set_attention_backend": (C.c_int, [C.c_int]),
    "clg_attention_fwd": (C.c_int, [C.POINTER(AttentionFwdParams), c_vp]),
    "clg_attention_bwd": (C.c_int, [C.POINTER(AttentionBwdParams), c_vp]),
    "clg_segment_sum_rows": (C.c_int, [c_vp, c_i64, c_vp, c_vp, C.c_int32, c_vp, C.c_int32, c_i64, C.c_int32, c_vp, c_vp]),
    "clg_embed_text_fwd": (C.c_int, [C.POINTER(EmbedTextParams), c_vp]),
    "clg_embed_text_bwd": (C.c_int, [C.POINTER(EmbedTextBwdParams), c_vp]),
    "clg_gqa_prior_loss": (C.c_int, [C.POINTER(GqaLossParams), c_vp]),
    "clg_sumsq": (C.c_int, [c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "clg_run_plan": (C.c_int, [c_vp, c_i32, c_vp, c_i32, C.c_uint64, c_vp]),
    "clg_sizeof_plan_entry": (C.c_size_t, []),
    "clg_sizeof_colsum_params": (C.c_size_t, []),
    "clg_sumsq_sparse": (C.c_int, [c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "clg_mark_rows": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "clg_adamw_step": (C.c_int, [C.POINTER(AdamWParams), c_vp]),
    "clg_cast_shadow": (C.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "clg_cast_f32_to_bf16": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "clg_prune_hist": (C.c_int, [c_vp, C.c_int32, C.c_uint32, C.c_uint32, C.c_int32, c_vp, c_vp]),
    "clg_prune_apply": (C.c_int, [c_vp, C.c_int32, C.c_uint32, C.c_uint64, c_vp, c_vp]),
    "clg_pack_index": (C.c_int, [c_vp, c_vp, C.c_int32, C.c_int32, C.c_int32, c_vp, c_vp, c_vp]),
    "clg_rows_gather": (C.c_int, [c_vp, c_i64, c_vp, C.c_int32, c_vp, c_i64, C.c_int32, C.c_int32, c_vp]),
    "clg_rows_scatter": (C.c_int, [c_vp, c_i64, c_vp, C.c_int32, c_vp, c_i64, C.c_int32, C.c_int32, c_vp]),
    "clg_cast_f32_to_bf16_2d": (C.c_int, [c_vp, c_i64, c_vp, c_i64, C.c_int32, C.c_int32, c_vp]),
    "clg_pack_mask": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "clg_apply_mask": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "clg_colsum": (C.c_int, [c_vp, c_i32, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "clg_build_key_bias": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "clg_dropout_inplace": (C.c_int, [c_vp, c_i64, C.POINTER(Dropout), c_vp]),
    "clg_act_bwd": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
}
for _n in _SIZEOF_CHECKS:
    SYMBOLS[_n] = (C.c_size_t, [])


class ClgError(RuntimeError):
    pass


_lib = None


def load():
    """Load libclgvqa.so, bind every declared symbol, verify struct layouts.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ClgError(
            "libclgvqa.so is missing (%s). Build it with `python -m clg_vqa_b200.build`; this package has no "
            "CPU or eager fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    for name, struct in _SIZEOF_CHECKS.items():
        want = getattr(lib, name)()
        if want != C.sizeof(struct):
            raise ClgError("ABI mismatch for %s: library %d bytes, binding %d bytes" % (struct.__name__, want, C.sizeof(struct)))
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        lib = load()
        msg = lib.clg_strerror(status).decode()
        detail = lib.clg_last_cuda_error().decode()
        raise ClgError("%s failed: %s%s" % (what or "clgvqa call", msg, (" [" + detail + "]") if detail and status == -3 else ""))
