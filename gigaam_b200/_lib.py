"""ctypes binding of libgigaam_b200.so (include/gigaam_b200.h).  There is no CPU fallback: if the library is
missing it is built with nvcc, and if that is impossible the import of any compute class fails loudly."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

_PKG = Path(__file__).resolve().parent
_LIB: Optional[C.CDLL] = None

c_f32p = C.c_void_p
c_vp = C.c_void_p


class GamConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "sample_rate", "n_mels", "n_fft", "win_length", "hop_length", "center",
        "feat_in", "n_layers", "d_model", "n_heads", "d_ff",
        "subsampling", "subs_kernel_size", "conv_kernel_size", "conv_norm", "self_attention", "pos_emb_max_len",
        "head", "num_classes", "pred_hidden", "joint_hidden", "max_symbols")]


LAYER_FIELDS = (
    "ln_ff1_g", "ln_ff1_b", "ff1_w1", "ff1_b1", "ff1_w2", "ff1_b2",
    "ln_att_g", "ln_att_b", "w_qk", "b_qk", "w_v", "b_v", "w_o", "b_o",
    "ln_conv_g", "ln_conv_b", "pw1_w", "pw1_b", "dw_w", "dw_b", "cn_g", "cn_b", "pw2_w", "pw2_b",
    "ln_ff2_g", "ln_ff2_b", "ff2_w1", "ff2_b1", "ff2_w2", "ff2_b2", "ln_out_g", "ln_out_b",
    "w_qkv_rel", "b_qkv_rel", "pos_proj")

REL_POS_MAX_T = 768   # GAM_REL_POS_MAX_T in include/gigaam_b200.h


class GamLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in LAYER_FIELDS]


WEIGHT_FIELDS_HEAD = ("window", "dft_cos", "dft_sin", "mel_fb", "sub1_w", "sub1_b", "sub2_w", "sub2_b",
                      "sub_out_w", "sub_out_b", "rope_cos", "rope_sin")
WEIGHT_FIELDS_TAIL = ("ctc_w", "ctc_b", "rnnt_enc_w", "rnnt_enc_b", "rnnt_emb_gates", "rnnt_whh_t", "rnnt_wp_t",
                      "rnnt_bp", "rnnt_wo", "rnnt_bo", "c1d_w1", "c1d_b1", "c1d_w2", "c1d_b2", "dft_w", "mel_lo", "mel_hi")


class GamWeights(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in WEIGHT_FIELDS_HEAD] + [("layers", C.POINTER(GamLayerWeights))]
                + [(n, C.c_void_p) for n in WEIGHT_FIELDS_TAIL])


EXPORTS = ("gam_create", "gam_destroy", "gam_last_error", "gam_version", "gam_logmel_frames", "gam_encoded_frames",
           "gam_workspace_bytes", "gam_logmel", "gam_encode", "gam_ctc_greedy", "gam_rnnt_greedy", "gam_test_gemm",
           "gam_test_attention", "gam_launch_count", "gam_profile_begin", "gam_profile_end", "gam_profile_class_count",
           "gam_profile_class_name", "gam_logmel_workspace_bytes", "gam_logmel_tc", "gam_test_attention_relpos",
           "gam_decode_workspace_bytes", "gam_group_words", "gam_comm_unique_id", "gam_comm_init",
           "gam_comm_nccl_version", "gam_gather_hyps", "gam_test_attention_varlen")


def lib_path() -> Path:
    return Path(os.environ.get("GIGAAM_B200_LIB", _PKG / "libgigaam_b200.so"))


def load() -> C.CDLL:
    """Load (building first if needed) the shared library and declare the prototypes."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not path.exists():
        from ._build import build_library
        path = build_library()
    lib = C.CDLL(str(path))
    H = C.c_void_p
    i32, i64 = C.c_int32, C.c_int64
    lib.gam_create.argtypes = [C.POINTER(GamConfig), C.POINTER(GamWeights), C.c_int, C.POINTER(H)]
    lib.gam_create.restype = C.c_int
    lib.gam_destroy.argtypes = [H]
    lib.gam_destroy.restype = None
    lib.gam_last_error.argtypes = [H]
    lib.gam_last_error.restype = C.c_char_p
    lib.gam_version.restype = C.c_int
    lib.gam_launch_count.argtypes = [H]
    lib.gam_launch_count.restype = i64
    lib.gam_logmel_frames.argtypes = [H, i64]
    lib.gam_logmel_frames.restype = i64
    lib.gam_encoded_frames.argtypes = [H, i64]
    lib.gam_encoded_frames.restype = i64
    lib.gam_workspace_bytes.argtypes = [H, i32, i64]
    lib.gam_workspace_bytes.restype = i64
    lib.gam_decode_workspace_bytes.argtypes = [H, i32, i32]
    lib.gam_decode_workspace_bytes.restype = i64
    lib.gam_comm_unique_id.argtypes = [c_vp]
    lib.gam_comm_unique_id.restype = C.c_int
    lib.gam_comm_init.argtypes = [H, c_vp, i32, i32]
    lib.gam_comm_init.restype = C.c_int
    lib.gam_comm_nccl_version.restype = i32
    lib.gam_gather_hyps.argtypes = [H, c_vp, i64, c_vp, c_vp]
    lib.gam_gather_hyps.restype = C.c_int
    lib.gam_group_words.argtypes = [H, c_vp, c_vp, c_vp, i32, i32, c_vp, i32, i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.gam_group_words.restype = C.c_int
    lib.gam_logmel.argtypes = [H, c_vp, i32, i64, c_vp, c_vp]
    lib.gam_logmel.restype = C.c_int
    lib.gam_encode.argtypes = [H, c_vp, c_vp, i32, i64, c_vp, i64, c_vp, c_vp, i32, c_vp]
    lib.gam_encode.restype = C.c_int
    for fn in (lib.gam_ctc_greedy, lib.gam_rnnt_greedy):
        fn.argtypes = [H, c_vp, c_vp, i32, i32, c_vp, i64, c_vp, c_vp, c_vp, i32, c_vp]
        fn.restype = C.c_int
    lib.gam_test_gemm.argtypes = [H, i32, c_vp, c_vp, c_vp, c_vp, c_vp, i32, i32, i32, i32, C.c_float, c_vp]
    lib.gam_test_gemm.restype = C.c_int
    lib.gam_test_attention.argtypes = [H, c_vp, c_vp, c_vp, i32, i32, c_vp]
    lib.gam_test_attention.restype = C.c_int
    lib.gam_test_attention_relpos.argtypes = [H, c_vp, c_vp, c_vp, c_vp, i32, i32, c_vp]
    lib.gam_test_attention_relpos.restype = C.c_int
    lib.gam_test_attention_varlen.argtypes = [H, c_vp, c_vp, c_vp, c_vp, c_vp, i32, i32, i32, c_vp]
    lib.gam_test_attention_varlen.restype = C.c_int
    lib.gam_logmel_workspace_bytes.argtypes = [H, i32, i64]
    lib.gam_logmel_workspace_bytes.restype = i64
    lib.gam_logmel_tc.argtypes = [H, c_vp, i32, i64, c_vp, c_vp, i64, c_vp]
    lib.gam_logmel_tc.restype = C.c_int
    lib.gam_profile_begin.argtypes = [H]
    lib.gam_profile_begin.restype = C.c_int
    lib.gam_profile_end.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_int64), i32]
    lib.gam_profile_end.restype = C.c_int
    lib.gam_profile_class_count.restype = C.c_int
    lib.gam_profile_class_name.argtypes = [i32]
    lib.gam_profile_class_name.restype = C.c_char_p
    _LIB = lib
    return lib


class GamError(RuntimeError):
    pass


def check(lib: C.CDLL, handle, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.gam_last_error(handle)
        raise GamError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
