"""Thin Python wrappers over the C ABI (include/ofb_b200.h).  Tensors in, tensors out; no math happens here.

Every function launches on the current torch CUDA stream and never synchronises.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, check, cur_stream, lib, ptr

EPI_STORE, EPI_FC1, EPI_FC2_DGRAD, EPI_WGRAD, EPI_PATCH, EPI_DECODER = range(6)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.OfbError("ofb_b200 kernels need CUDA tensors (no CPU fallback)")


def gemm(epi, A, B, *, M, N, K, out0=None, ld0=0, out1=None, ld1=0, out_fp32=False, bias=None, colscale=None,
         rowscale=None, rows_per_scale=1, res=None, aux=None, colpart0=None, colpart1=None, scale_ptr=None, pos=None,
         mask_token=None, rowmask=None, target=None, tokens=1, a_mn=False, b_mn=False, bn=0, k_splits=0, lda=None,
         ldb=None):
    """D[M,N] = sum_k A[m,k] B[n,k] with a fused epilogue (see ofb_b200.h).  A/B are bf16, 2-D, last dim contiguous."""
    _need_cuda(A, B, out0)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    g = GemmArgs()
    g.M, g.N, g.K, g.k_splits = M, N, K, k_splits
    g.out0, g.ld0 = ptr(out0), ld0 or (out0.stride(0) if out0 is not None and out0.dim() == 2 else 0)
    g.out1, g.ld1 = ptr(out1), ld1 or (out1.stride(0) if out1 is not None and out1.dim() == 2 else 0)
    g.out_fp32 = 1 if out_fp32 else 0
    g.bias, g.colscale = ptr(bias), ptr(colscale)
    g.rowscale, g.rows_per_scale = ptr(rowscale), rows_per_scale
    g.res, g.ldres = ptr(res), (res.stride(0) if res is not None else 0)
    g.aux, g.ldaux = ptr(aux), (aux.stride(0) if aux is not None else 0)
    g.colpart0, g.colpart1, g.scale_ptr = ptr(colpart0), ptr(colpart1), ptr(scale_ptr)
    g.pos, g.mask_token, g.rowmask, g.target = ptr(pos), ptr(mask_token), ptr(rowmask), ptr(target)
    g.tokens = tokens
    lda = lda if lda is not None else A.stride(0)
    ldb = ldb if ldb is not None else B.stride(0)
    check(lib().ofb_gemm_bf16(epi, int(a_mn), int(b_mn), bn, ptr(A), lda, ptr(B), ldb, C.byref(g), cur_stream()),
          "ofb_gemm_bf16")
