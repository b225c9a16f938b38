"""ctypes binding of libofb_b200.so (the C ABI declared in include/ofb_b200.h).

There is deliberately no fallback: if the shared library is missing, importing a symbol raises.  Build it with
``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libofb_b200.so")

_lib = None


class OfbError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    """Mirror of ``ofb_gemm_args`` (include/ofb_b200.h)."""
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("k_splits", C.c_int32),
        ("out0", C.c_void_p), ("ld0", C.c_int32),
        ("out1", C.c_void_p), ("ld1", C.c_int32),
        ("out_fp32", C.c_int32),
        ("bias", C.c_void_p), ("colscale", C.c_void_p),
        ("rowscale", C.c_void_p), ("rows_per_scale", C.c_int32),
        ("res", C.c_void_p), ("ldres", C.c_int32),
        ("aux", C.c_void_p), ("ldaux", C.c_int32),
        ("colpart0", C.c_void_p), ("colpart1", C.c_void_p),
        ("scale_ptr", C.c_void_p),
        ("pos", C.c_void_p), ("mask_token", C.c_void_p), ("rowmask", C.c_void_p), ("target", C.c_void_p),
        ("tokens", C.c_int32),
    ]


def lib():
    """Load (once) and return the shared library; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OfbError(
                f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback. "
                "Run __graft_entry__.build().")
        _lib = C.CDLL(LIB_PATH)
    return _lib


def check(code, what):
    if code != 0:
        raise OfbError(f"{what} failed with code {code}")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
