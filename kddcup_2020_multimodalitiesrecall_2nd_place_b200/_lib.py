"""ctypes binding of include/mmrecall.h (libmmrecall.so).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libmmrecall.so"

MMR_OK = 0
DT_FP16, DT_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU_TANH, ACT_GELU_ERF, ACT_TANH = range(5)
MODEL_ZK, MODEL_LDS, MODEL_LXMERT = range(3)
TUNE_GEMM_PAIR, TUNE_GEMM_P16, TUNE_GEMM_TAIL, TUNE_GEMM_CLUSTER, TUNE_GEMM_LN, TUNE_PDL, TUNE_ATTN_TMA, TUNE_ATTN_TC, TUNE_LN_ROW_CFG, TUNE_LABEL_DEDUP, TUNE_LX_MERGE, TUNE_PRUNE_LAST, TUNE_LX_QUERY_DEDUP = range(13)
PRECISION_FAST, PRECISION_STRICT = 0, 1

# every symbol include/mmrecall.h declares (tests check the .so exports all of them)
EXPORTS = [
    "mmr_last_error", "mmr_abi_version", "mmr_experimental_build", "mmr_device_check", "mmr_set_tuning", "mmr_get_tuning", "mmr_tuning_generation",
    "mmr_gemm", "mmr_gemm_layernorm", "mmr_gemm_layernorm_supported", "mmr_layernorm", "mmr_attention", "mmr_cast16", "mmr_cls_attention", "mmr_split3", "mmr_attention_f32", "mmr_am_softmax_head", "mmr_linear_head", "mmr_decode_tsv", "mmr_decode_tsv_reuse", "mmr_crc32c", "mmr_boxes_normalize", "mmr_ensemble_topk", "mmr_ndcg_at_k",
    "mmr_create", "mmr_workspace_bytes", "mmr_destroy", "mmr_forward", "mmr_set_debug_taps", "mmr_get_activation",
    "mmr_launches_per_forward", "mmr_set_profiling", "mmr_get_profile",
]


class MmrConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "model_kind", "dtype", "hidden", "heads", "intermediate", "vocab", "max_pos", "type_vocab", "feat_dim",
        "label_len", "n_layers", "n_r_layers", "n_x_layers", "lq", "nbox", "max_batch", "precision")]


class MmrTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("dims", C.c_int64 * 4)]


class MmrInputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "query_ids", "segment_ids", "label_ids", "feats", "boxes", "len_query", "num_boxes", "query_mask",
        "visn_mask", "labels", "region_sum", "lang_unique", "lang_slot")] + [("n_lang_unique", C.c_int32)]


class MmrError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = False) -> C.CDLL:
    """Loads libmmrecall.so.  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if build_if_missing:
            from .csrc.build import build
            build()
        else:
            raise MmrError(
                f"{LIB_PATH} not found: build it with `python -m {__package__}.csrc.build` "
                "(there is no CPU / PyTorch fallback path)")
    lib = C.CDLL(str(LIB_PATH))
    lib.mmr_last_error.restype = C.c_char_p
    lib.mmr_abi_version.restype = C.c_int
    lib.mmr_experimental_build.restype = C.c_int
    lib.mmr_device_check.argtypes = [C.c_int]
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
    lib.mmr_gemm.argtypes = [vp, i64, vp, i64, i32, i32, i32, vp, vp, i64, vp, i64, vp, i64, i32, i32, vp]
    lib.mmr_layernorm.argtypes = [vp, i64, vp, vp, f32, i32, i32, vp, i64, vp, i64, f32, i32, i32, vp]
    lib.mmr_attention.argtypes = [vp, i64, vp, i64, vp, i64, vp, vp, i64, i32, i32, i32, i32, i32, vp]
    lib.mmr_cast16.argtypes = [vp, vp, i64, i32, vp]
    lib.mmr_cls_attention.argtypes = [vp, i64, vp, vp, i64, vp, vp, i64, i32, i32, i32, i32, vp]
    lib.mmr_split3.argtypes = [vp, i64, i32, i32, vp, i64, i32, i32, i32, vp]
    lib.mmr_attention_f32.argtypes = [vp, i64, i64, vp, vp, i64, i64, vp, vp, i64, i64, i32, i32, i32, i32, vp]
    lib.mmr_gemm_layernorm.argtypes = [vp, i64, vp, i64, i32, i32, vp, vp, i64, vp, vp, f32, vp, i64, vp, i64, i32, vp]
    lib.mmr_gemm_layernorm_supported.argtypes = [i32, i32, i32]
    lib.mmr_set_tuning.argtypes = [i32, i32]
    lib.mmr_get_tuning.argtypes = [i32]
    lib.mmr_get_tuning.restype = i32
    lib.mmr_tuning_generation.restype = C.c_uint
    lib.mmr_crc32c.argtypes = [vp, C.c_size_t, C.c_uint32]
    lib.mmr_crc32c.restype = C.c_uint32
    lib.mmr_am_softmax_head.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    lib.mmr_linear_head.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, vp, vp]
    if True:
        lib.mmr_create.argtypes = [C.POINTER(MmrConfig), C.POINTER(MmrTensor), i32, i32, C.POINTER(vp)]
        lib.mmr_workspace_bytes.argtypes = [C.POINTER(MmrConfig), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        lib.mmr_destroy.argtypes = [vp]
        lib.mmr_destroy.restype = None
        lib.mmr_forward.argtypes = [vp, C.POINTER(MmrInputs), i32, vp, vp, vp, vp]
        lib.mmr_set_debug_taps.argtypes = [vp, i32]
        lib.mmr_get_activation.argtypes = [vp, i32, vp, i64, vp]
        lib.mmr_launches_per_forward.argtypes = [vp]
        lib.mmr_set_profiling.argtypes = [vp, i32]
        lib.mmr_get_profile.argtypes = [vp, i32, vp, vp, vp]
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != MMR_OK:
        raise MmrError(f"mmrecall error {status}: {load().mmr_last_error().decode()}")
