"""ctypes binding of libnlc_b200.so (the C ABI declared in include/nlc_b200.h).

There is deliberately no fallback: if the shared object is missing or a call fails, an exception is
raised.  PyTorch is used by the callers only to own device memory and streams.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnlc_b200.so")

NLC_F32 = 0
NLC_BF16 = 1
NLC_F16 = 3  # fp16 operands: kind::f16 at the bf16 rate, 3 more mantissa bits
NLC_F32X3 = 2  # plain fp32 operands, split into tf32 hi+lo inside nlc_conv_tc (accuracy mode)
EDM_PARTS = 16
MAX_SRC = 3
MAX_SEG = 24


class NlcError(RuntimeError):
    pass


class Operand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
                ("ld", C.c_int), ("sh", C.c_int64), ("sn", C.c_int64)]


class KSeg(C.Structure):
    _fields_ = [("src", C.c_int), ("dh", C.c_int), ("dw", C.c_int), ("c0", C.c_int), ("nch", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("nsrc", C.c_int), ("src", Operand * MAX_SRC),
        ("nseg", C.c_int), ("seg", KSeg * MAX_SEG),
        ("weight", C.c_void_p), ("wbatched", Operand), ("Cout", C.c_int), ("stride", C.c_int),
        ("B", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int),
        ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("ld_rowvec", C.c_int),
        ("resid", C.c_void_p), ("ld_resid", C.c_int), ("out_scale", C.c_float),
        ("out_f32", C.c_void_p), ("ld_out_f32", C.c_int), ("out_op", C.c_void_p), ("ld_out_op", C.c_int),
        ("out_head_split", C.c_int), ("stats", C.c_void_p), ("stats_nblk", C.c_int), ("resid_mode", C.c_int),
        ("out_up", C.c_int), ("resid_is_op", C.c_int), ("act", C.c_int),
    ]


class OpDesc(C.Structure):
    _fields_ = [
        ("task", C.c_int), ("channels", C.c_int), ("R", C.c_int), ("ratio", C.c_int),
        ("idx_host", C.c_void_p), ("n_idx", C.c_int64),
        ("U_small_host", C.c_void_p), ("V_small_host", C.c_void_p), ("sing_small_host", C.c_void_p),
        ("m_small", C.c_int), ("mult_host", C.c_void_p), ("pinv_mult_host", C.c_void_p),
        ("U_small2_host", C.c_void_p), ("V_small2_host", C.c_void_p), ("lambda_sing_host", C.c_void_p),
    ]


class DdnmCoef(C.Structure):
    """nlc_ddnm_coef (include/nlc_b200.h)."""
    _fields_ = [("a", C.c_float), ("sigma_t", C.c_float), ("sigma_y", C.c_double), ("eta", C.c_double)]


_lib = None
_ctx = {}

_I, _F, _P, _SZ, _I64 = C.c_int, C.c_float, C.c_void_p, C.c_size_t, C.c_int64

_SIGNATURES = {
    "nlc_last_error": (C.c_char_p, []),
    "nlc_abi_version": (_I, []),
    "nlc_create": (_P, [_I]),
    "nlc_destroy": (None, [_P]),
    "nlc_sm_count": (_I, [_P]),
    "nlc_ctx_set": (_I, [_P, C.c_char_p, _I]),
    "nlc_conv_tc": (_I, [_P, C.POINTER(ConvDesc), _P]),
    "nlc_conv_in_nchw": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P, _I, _I, _P]),
    "nlc_im2col_in": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "nlc_conv_out_nchw": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    "nlc_nhwc_head_to_nchw": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nlc_image_metrics": (_I, [_P, _P, _P, _I, _I64, _P, _P, _P, _P]),
    "nlc_train_prepare": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I64, _P, _P, _P, _P]),
    "nlc_adamw_ema_step": (_I, [_P, _P, _P, _P, _P, _P, _I64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                _I64, C.c_double, C.c_double, _P]),
    "nlc_ssim3d_ws": (_SZ, [_I, _I, _I]),
    "nlc_ssim3d": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "nlc_groupnorm": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _P]),
    "nlc_groupnorm_ws": (_SZ, [_I, _I, _I, _I]),
    "nlc_resample": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P]),
    "nlc_resample_op": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    "nlc_sgemm": (_I, [_P, _I, _I, _I, _I, _P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _P, _P, _P]),
    "nlc_unfold3x3": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nlc_fold3x3": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _F, _P]),
    "nlc_gn_train_fwd": (_I, [_P, _P, _I, _I, _I, _I, _F, _P, _P, _I, _P, _P, _P]),
    "nlc_gn_train_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P, _P]),
    "nlc_softmax_rows": (_I, [_P, _P, _P, _I, _I, _F, _P, _P]),
    "nlc_bias_add": (_I, [_P, _P, _P, _I64, _I, _P]),
    "nlc_colsum": (_I, [_P, _P, _I64, _I, _P, _P]),
    "nlc_axpby": (_I, [_P, _F, _P, _F, _P, _P, _I64, _P]),
    "nlc_permute_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "nlc_bn1d_gelu_train": (_I, [_P, _P, _P, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nlc_bn1d_act_train": (_I, [_P, _P, _P, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nlc_head_loss": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "nlc_head_loss_weighted": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "nlc_fid_preprocess": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    "nlc_im2col_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I64, _P]),
    "nlc_pool2d": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    "nlc_global_avgpool": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nlc_cov_accumulate": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "nlc_attention": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P, _I, _P, _P]),
    "nlc_attention_ws": (_SZ, [_I, _I, _I, _I, _I]),
    "nlc_linear": (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _P, _I, _P]),
    "nlc_timestep_embedding": (_I, [_P, _P, _I, _P, _I, _I, _P, _I, _P]),
    "nlc_row_norm": (_I, [_P, _P, _I, _I, _P, _P]),
    "nlc_refine_sigma": (_I, [_P, _P, _I, _I, _P, _I, _F, _F, _I, _F, _P, _P, _I, _I, _P, _P, _P, _P]),
    "nlc_sigma_correct": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P]),
    "nlc_dynamic_threshold": (_I, [_P, _P, _I, _I, C.c_double, _F, _P, _P]),
    "nlc_sigma_estimate": (_I, [_P, _P, _P, _I, _I, _F, _F, _P, _I, _P, _I, C.POINTER(C.c_float * 4), _P, _P, _I, _P,
                                _P, _P]),
    "nlc_normalize_rows": (_I, [_P, _P, _I, _I, _P]),
    "nlc_pred_xstart": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "nlc_pred_xprev": (_I, [_P, _I, C.c_double, _P, _P, _P, _P, _P, _I, _F, _P, _I, _P, _I, _I, _I, _P, _P, _P]),
    "nlc_best_update": (_I, [_P, _P, _F, _P, _P, _P, _P, _I64, _P]),
    "nlc_edm_prepare": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "nlc_edm_eps": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "nlc_edm_mix": (_I, [_P, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, _I, _I, _P, _P, _P]),
    "nlc_edm_axpy": (_I, [_P, _P, _P, _P, C.c_double, _P, _P, _I, _I, _P, _P]),
    "nlc_op_create": (_I, [_P, C.POINTER(OpDesc), C.POINTER(_P)]),
    "nlc_op_destroy": (None, [_P]),
    "nlc_op_ydim": (_I64, [_P]),
    "nlc_op_ws": (_SZ, [_P, _I]),
    "nlc_op_A": (_I, [_P, _P, _I, _P, _P, _P]),
    "nlc_op_At": (_I, [_P, _P, _I, _P, _P, _P]),
    "nlc_op_Apinv": (_I, [_P, _P, _I, _P, _P, _P]),
    "nlc_op_project": (_I, [_P, _P, _P, _I, _P, _P, _P]),
    "nlc_op_Apinv_eta": (_I, [_P, _P, _I, C.c_double, _P, _P, _P]),
    "nlc_op_lambda": (_I, [_P, _P, _I, C.POINTER(DdnmCoef), _P, _P, _P]),
    "nlc_op_lambda_noise": (_I, [_P, _P, _P, _I, C.POINTER(DdnmCoef), _P, _P, _P]),
    "nlc_ddnm_step": (_I, [_P, _P, _P, _I64, _P, _P, _I, _F, _F, C.c_double, C.c_double, _I, _P, _P, _P, _P]),
    "nlc_ddnm_renoise": (_I, [_P, _P, _P, _I64, _F, _P, _P]),
    "nlc_l1_diff_rows": (_I, [_P, _P, _P, _I, _I64, _P, _P]),
}


def exported_symbols():
    """Names include/nlc_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def lib():
    """Load libnlc_b200.so once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NlcError("%s is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                           "There is no CPU or PyTorch fallback for the hot path." % LIB_PATH)
        import torch  # noqa: F401  (loads libcudart.so.12 first so both sides share one runtime)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGNATURES.items():
            if os.environ.get("NLC_PARTIAL") == "1" and not hasattr(L, name):
                continue  # development only: a partially built library
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise NlcError("libnlc_b200 call failed (%d): %s" % (rc, lib().nlc_last_error().decode()))


def ctx(device_index):
    """One nlc_ctx per device of this process."""
    h = _ctx.get(device_index)
    if h is None:
        h = lib().nlc_create(int(device_index))
        if not h:
            raise NlcError("nlc_create(%d) failed: %s" % (device_index, lib().nlc_last_error().decode()))
        _ctx[device_index] = h
    return h
