// capi.cu — extern "C" boundary (include/ofb_b200.h) over the kernels in this directory.
#include "../../include/ofb_b200.h"
#include "gemm.cuh"
#include <cuda_runtime.h>

namespace ofb {
int launch_gemm(int epi, int a_mn, int b_mn, int bn_hint, const void* A, int lda, const void* B, int ldb, GemmArgs g,
                cudaStream_t stream);
int num_sms();
}  // namespace ofb

extern "C" {

int ofb_version(void) { return 1; }
int ofb_num_sms(void) { return ofb::num_sms(); }

int ofb_gemm_bf16(int epilogue, int a_mn, int b_mn, int bn_hint, const void* A, int lda, const void* B, int ldb,
                  const ofb_gemm_args* a, void* stream) {
    if (a == nullptr || A == nullptr || B == nullptr) return 1000;
    ofb::GemmArgs g;
    g.M = a->M; g.N = a->N; g.K = a->K; g.k_splits = a->k_splits;
    g.out0 = a->out0; g.ld0 = a->ld0; g.out1 = a->out1; g.ld1 = a->ld1; g.out_fp32 = a->out_fp32;
    g.bias = a->bias; g.colscale = a->colscale; g.rowscale = a->rowscale;
    g.rows_per_scale = a->rows_per_scale > 0 ? a->rows_per_scale : 1;
    g.res = reinterpret_cast<const __nv_bfloat16*>(a->res); g.ldres = a->ldres;
    g.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux); g.ldaux = a->ldaux;
    g.colpart0 = a->colpart0; g.colpart1 = a->colpart1; g.scale_ptr = a->scale_ptr;
    g.pos = a->pos; g.mask_token = a->mask_token; g.rowmask = a->rowmask; g.target = a->target;
    g.tokens = a->tokens > 0 ? a->tokens : 1;
    return ofb::launch_gemm(epilogue, a_mn, b_mn, bn_hint, A, lda, B, ldb, g, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
