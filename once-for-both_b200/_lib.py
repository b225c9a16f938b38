"""ctypes binding of libofb_b200.so (the C ABI declared in include/ofb_b200.h).

There is deliberately no fallback: if the shared library is missing, importing a symbol raises.  Build it with
``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OFB_B200_LIB: load another build of the same sources instead (debug builds only, e.g. the phase-trace build of tools/attn_trace.py)
LIB_PATH = os.environ.get("OFB_B200_LIB") or os.path.join(_HERE, "libofb_b200.so")

_lib = None


class OfbError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    """Mirror of ``ofb_gemm_args`` (include/ofb_b200.h)."""
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("k_splits", C.c_int32),
        ("out0", C.c_void_p), ("ld0", C.c_int32),
        ("out1", C.c_void_p), ("ld1", C.c_int32),
        ("out_fp32", C.c_int32), ("bias_rowscaled", C.c_int32),
        ("bias", C.c_void_p), ("colscale", C.c_void_p), ("colscale_period", C.c_int32),
        ("rowscale", C.c_void_p), ("rows_per_scale", C.c_int32),
        ("res", C.c_void_p), ("ldres", C.c_int32),
        ("aux", C.c_void_p), ("ldaux", C.c_int32),
        ("colpart0", C.c_void_p), ("colpart1", C.c_void_p),
        ("scale_ptr", C.c_void_p),
        ("pos", C.c_void_p), ("mask_token", C.c_void_p), ("rowmask", C.c_void_p), ("target", C.c_void_p),
        ("tokens", C.c_int32),
        ("splitk_ws", C.c_void_p),
    ]


class BimaskModule(C.Structure):
    """Mirror of ``ofb_bimask_module``."""
    _fields_ = [
        ("kind", C.c_int32), ("dim", C.c_int32), ("heads", C.c_int32), ("n_i", C.c_int32), ("n_j", C.c_int32),
        ("switch_off", C.c_int32), ("width_off", C.c_int32), ("gate_off", C.c_int32),
        ("stride", C.c_int32), ("pad_", C.c_int32),
        ("alpha_off", C.c_int64), ("score_off", C.c_int64),
        ("coef", C.c_float), ("loss_w", C.c_float),
    ]


class SplitkJob(C.Structure):
    """Mirror of ``ofb_splitk_job``."""
    _fields_ = [("ws", C.c_void_p), ("out", C.c_void_p), ("n4", C.c_int64), ("splits", C.c_int32), ("pad_", C.c_int32)]


class ReduceJob(C.Structure):
    """Mirror of ``ofb_reduce_job``."""
    _fields_ = [("part", C.c_void_p), ("out", C.c_void_p), ("div_by", C.c_void_p), ("R", C.c_int32), ("N", C.c_int32),
                ("scale", C.c_float), ("accumulate", C.c_int32)]


_P, _I, _F, _L, _D = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_double
# every symbol include/ofb_b200.h declares, with its argument types
SIGNATURES = {
    "ofb_version": [],
    "ofb_num_sms": [],
    "ofb_gemm_bf16": [_I, _I, _I, _I, _P, _I, _P, _I, _P, _P],
    "ofb_gemm_mlp_partial_rows": [_I, _I],
    "ofb_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _F, _P],
    "ofb_layernorm_fwd_ex": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "ofb_layernorm_bwd_parts": [_I],
    "ofb_layernorm_bwd_ex": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ofb_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "ofb_reduce_partials": [_P, _I, _I, _P, _F, _P, _I, _P],
    "ofb_reduce_partials_multi": [_P, _I, _P],
    "ofb_gemm_wgrad_splits": [_I, _I, _I, _I, _I],
    "ofb_splitk_reduce": [_P, _I, _P],
    "ofb_patchify": [_P, _P, _I, _I, _I, _P],
    "ofb_mixup_batch": [_P, _P, _I, _I, _D, _I, _I, _I, _I, _I, _P],
    "ofb_patchify_mixup": [_P, _P, _I, _I, _I, _D, _I, _I, _I, _I, _I, _P],
    "ofb_mixup_target": [_P, _P, _I, _I, _D, _D, _P],
    "ofb_pmim_mask": [_P, _P, _I, _I, _I, _P],
    "ofb_droppath_scale": [_P, _P, _P, _I, _I, _P],
    "ofb_cls_rows": [_P, _P, _P, _P, _I, _I, _I, _P],
    "ofb_embed_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "ofb_norm_targets": [_P, _P, _P, _I, _I, _P],
    "ofb_ls_cross_entropy": [_P, _P, _P, _P, _I, _I, _F, _F, _P],
    "ofb_soft_target_cross_entropy": [_P, _P, _P, _P, _I, _I, _F, _P],
    "ofb_eval_metrics": [_P, _P, _P, _I, _I, _P],
    "ofb_loss_finalize": [_P, _I, _P, _I, _P, _I, _P, _F, _P, _P],
    "ofb_adamw": [_P, _P, _P, _P, _P, _P, _I, _P, _L, _I, _P],
    "ofb_cast_bf16": [_P, _P, _L, _P],
    "ofb_copy_f32": [_P, _P, _I, _P],
    "ofb_colsum_bf16": [_P, _I, _I, _I, _P, _F, _P, _P],
    "ofb_bimask_fwd": [_P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "ofb_arch_finalize": [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P, _P, _P],
    "ofb_bimask_bwd": [_P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P],
    "ofb_attention_fwd": [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    "ofb_attention_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P],
}


def lib():
    """Load (once) and return the shared library; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OfbError(
                f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback. "
                "Run __graft_entry__.build().")
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)       # AttributeError here = header / library mismatch
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = l
    return _lib


def check(code, what):
    if code != 0:
        raise OfbError(f"{what} failed with code {code}")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else t.data_ptr()


def cur_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
