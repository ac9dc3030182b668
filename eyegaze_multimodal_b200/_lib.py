"""ctypes binding of libeyegaze_b200.so (the C ABI declared in include/eyegaze_b200.h).

The library is loaded from the package directory (built in-tree by ``csrc/build.py``).  There is no
fallback of any kind: if the shared object is missing or a kernel entry point fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeyegaze_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_GELU_DGRAD = 0, 1, 2, 3
ACTBWD_NONE, ACTBWD_RELU_MASK, ACTBWD_GELU, ACTBWD_MUL = 0, 1, 2, 3

vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float


class Operand(C.Structure):
    _fields_ = [("ptr", vp), ("major", i32), ("rows_per_group", i32), ("row_stride", i64), ("group_stride", i64),
                ("seg_len", i32), ("seg_row_shift", i32)]


class Matrix(C.Structure):
    _fields_ = [("ptr", vp), ("dtype", i32), ("rows_per_group", i32), ("row_stride", i64), ("group_stride", i64)]


class GemmDesc(C.Structure):
    _fields_ = [("M", i32), ("N", i32), ("K", i32), ("in_dtype", i32), ("a", Operand), ("b", Operand),
                ("c", Matrix), ("c_pre", Matrix), ("residual", Matrix), ("aux", Matrix), ("bias", vp),
                ("alpha", f32), ("act", i32), ("act_bwd", i32), ("aux_scale", f32), ("dropout_p", f32),
                ("dropout_seed", u64), ("accumulate", i32), ("split_k", i32), ("c_colsum", vp)]


class AttentionDesc(C.Structure):
    _fields_ = [("q", vp), ("k", vp), ("v", vp), ("o", vp),
                ("q_bs", i64), ("q_rs", i64), ("k_bs", i64), ("k_rs", i64), ("v_bs", i64), ("v_rs", i64),
                ("o_bs", i64), ("o_rs", i64),
                ("d_o", vp), ("do_bs", i64), ("do_rs", i64),
                ("dq", vp), ("dk", vp), ("dv", vp),
                ("dq_bs", i64), ("dq_rs", i64), ("dk_bs", i64), ("dk_rs", i64), ("dv_bs", i64), ("dv_rs", i64),
                ("lse", vp), ("delta", vp), ("probs", vp),
                ("dtype", i32), ("S", i32), ("H", i32), ("Lq", i32), ("Lk", i32), ("head_dim", i32), ("kv_shift", i32),
                ("scale", f32), ("dropout_p", f32), ("seed", u64),
                ("dq_colsum", vp), ("dk_colsum", vp), ("dv_colsum", vp)]


class FuzzyDesc(C.Structure):
    _fields_ = [("tau_img", vp), ("tau_eeg", vp), ("c_reliable", vp), ("c_unreliable_img", vp),
                ("c_unreliable_eeg", vp), ("log_sigma_reliable_img", vp), ("log_sigma_reliable_eeg", vp),
                ("log_sigma_unreliable_img", vp), ("log_sigma_unreliable_eeg", vp), ("beta", vp),
                ("mode", i32), ("B", i32), ("num_classes", i32), ("eps_temp", f32), ("eps_log", f32), ("eps_div", f32)]


class AdamwState(C.Structure):
    _fields_ = [("lr", vp), ("ctrl", vp), ("grad_scale", vp), ("use_device_step", i32)]


# name -> argtypes (all return int except where noted); mirrors include/eyegaze_b200.h one to one
_SIGNATURES = {
    "egb_gemm": [C.POINTER(GemmDesc), vp],
    "egb_prof_enable": [i32],
    "egb_prof_read": [i32, C.POINTER(C.c_double), i32],
    "egb_prof_dump": [i32, C.POINTER(C.c_double), i32, C.POINTER(i32)],
    "egb_cast_from_f32": [vp, vp, i32, i64, vp],
    "egb_cast_to_f32": [vp, i32, vp, i64, vp],
    "egb_copy_strided4": [vp, i32, vp, i32, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64), vp],
    "egb_zero": [vp, i64, vp],
    "egb_eeg_pack": [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "egb_seq_assemble_fwd": [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "egb_seq_assemble_bwd": [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "egb_add_rows_broadcast": [vp, vp, vp, i32, i64, i32, i32, vp],
    "egb_tail_pool_fwd": [vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "egb_tail_pool_bwd": [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "egb_colsum": [C.POINTER(Matrix), i32, i32, vp, i32, vp],
    "egb_act_bwd": [C.POINTER(Matrix), vp, C.POINTER(Matrix), i32, i32, i32, f32, vp],
    "egb_dropout_bwd": [vp, vp, i32, i64, f32, u64, vp],
    "egb_cross_entropy": [vp, vp, vp, vp, i32, i32, vp],
    "egb_scale_by_device_scalar": [vp, vp, vp, i64, vp],
    "egb_layernorm_fwd": [vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, vp],
    "egb_layernorm_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "egb_eeg_window_normalize": [vp, vp, i32, i32, i32, i32, vp],
    "egb_image_u8_normalize": [vp, vp, i32, i32, i32, C.POINTER(f32), C.POINTER(f32), vp],
    "egb_multi_tensor_cast_bf16": [vp, vp, i32, vp],
    "egb_multi_tensor_sqnorm": [vp, vp, i32, vp, vp],
    "egb_multi_tensor_adamw": [vp, vp, i32, f32, f32, f32, f32, f32, f32, vp, vp],
    "egb_layernorm_bwd_res": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "egb_layernorm_bwd_ex": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "egb_dropout_bwd_colsum": [vp, vp, i32, i32, i32, f32, u64, vp, vp],
    "egb_attention_fwd": [C.POINTER(AttentionDesc), vp],
    "egb_attention_bwd": [C.POINTER(AttentionDesc), vp],
    "egb_debug_attention_timing": [vp],
    "egb_debug_gemm_timing": [vp],
    "egb_stft_logmag": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "egb_conv3x3_c64_c32": [vp, i64, vp, vp, i64, i64, i32, vp],
    "egb_conv3x3_c32_c64": [vp, i64, vp, vp, vp, i64, i64, i32, vp],
    "egb_conv3x3_dw_c32_c64": [vp, i64, vp, i64, vp, i64, i64, i32, vp],
    "egb_spec_conv1_pool_fwd": [vp, vp, vp, vp, i32, i32, i32, i32, i64, vp, vp],
    "egb_spec_conv1_pool_bwd": [vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, vp, vp],
    "egb_relu_avgpool_fwd": [vp, vp, i32, i32, i32, i32, vp],
    "egb_relu_avgpool_bwd": [vp, vp, vp, i32, i32, i32, i32, vp, vp],
    "egb_ibs_connectivity": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32),
                             C.POINTER(i32), i32, vp],
    "egb_ibs_scalar_features": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), vp],
    "egb_instnorm_tokens_fwd": [vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, vp],
    "egb_instnorm_tokens_bwd": [vp, vp, i32, vp, vp, i32, i32, i32, f32, vp],
    "egb_fuzzy_fwd": [C.POINTER(FuzzyDesc), vp, vp, vp, vp, vp, vp],
    "egb_fuzzy_bwd": [C.POINTER(FuzzyDesc), vp, vp, vp, vp, vp, vp, vp, vp],
    "egb_vit_patchify": [vp, vp, i64, i64, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "egb_fill_row0": [vp, vp, vp, i32, i32, i32, i32, vp],
    "egb_seed_epoch_enable": [C.POINTER(vp)],
    "egb_seed_epoch_advance": [vp],
    "egb_multi_tensor_adamw_ex": [vp, vp, i32, f32, f32, f32, f32, f32, f32, vp, C.POINTER(AdamwState), vp],
    "egb_adamw_prepare": [vp, vp, vp, vp, i32, i32, vp],
    "egb_lr_schedule_step": [vp, vp, vp, vp, i32, i32, f32, f32, i32, i32, vp],
    "egb_accum_scalars": [C.POINTER(vp), i32, vp, vp],
    "egb_argmax_count": [vp, vp, vp, vp, i32, i32, vp],
    "egb_l2norm_rows_fwd": [vp, vp, vp, i32, i32, f32, vp],
    "egb_l2norm_rows_bwd": [vp, vp, vp, vp, i32, i32, vp],
    "egb_infonce_rows": [vp, vp, i32, i32, vp],
    "egb_supcon_rows": [vp, vp, vp, vp, vp, i32, vp],
    "egb_mse_loss": [vp, vp, vp, vp, i64, vp],
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["egb_last_error", "egb_version", "egb_launch_count"])

_lib = None


def load():
    """Loads the shared library (once) and declares every prototype.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "libeyegaze_b200.so not found at %s: build it with `python -m eyegaze_multimodal_b200.csrc.build` "
            "(nvcc, sm_100a).  This package has no CPU or eager fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.egb_last_error.restype = C.c_char_p
    lib.egb_version.restype = C.c_int
    lib.egb_launch_count.restype = C.c_int64
    _lib = lib
    return lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (name, lib.egb_last_error().decode("utf-8", "replace")))


def launch_count() -> int:
    return int(load().egb_launch_count())


def prof_enable(on: bool) -> None:
    call("egb_prof_enable", 1 if on else 0)


def prof_dump(kind: int = 0, max_records: int = 8192) -> list:
    """Per-launch records (launch order): dicts with ms, flops, bytes and the launcher's tag (GEMM: M, N, K, variant)."""
    out = (C.c_double * (8 * max_records))()
    n = i32(0)
    call("egb_prof_dump", kind, out, max_records, C.byref(n))
    return [{"ms": out[8 * i], "flops": out[8 * i + 1], "bytes": out[8 * i + 2], "tag": tuple(out[8 * i + 3:8 * i + 7])}
            for i in range(n.value)]


def prof_read(kind: int = 0, reset: bool = True) -> dict:
    out = (C.c_double * 4)()
    call("egb_prof_read", kind, out, 1 if reset else 0)
    return {"launches": int(out[0]), "ms": out[1], "flops": out[2], "bytes": out[3]}
