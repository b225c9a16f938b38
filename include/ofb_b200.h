/* ofb_b200.h — C ABI of the B200-native bi-mask DeiT search-step kernels.
 *
 * The reference (HankYe/Once-for-Both) is pure Python/PyTorch and has no FFI of its own; the seam it offers is the
 * nn.Module surface + the ModuleInjection factory (models/layers.py:1052-1081). This header is the C boundary a
 * maintainer binds underneath that seam (ctypes stub in INTEGRATION.md). Every entry point
 *   - takes plain device pointers, sizes and a cudaStream_t passed as void*,
 *   - never allocates, never synchronises, launches on the given stream,
 *   - returns 0 on success, a cudaError_t value (or a code >= 1000 for argument errors) otherwise.
 * All activations are bf16 row-major, all parameters / gradients / reductions are fp32 unless stated.
 * Each function cites the reference code it replaces (file:line relative to the reference repo).
 */
#ifndef OFB_B200_H
#define OFB_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* library / device info -------------------------------------------------------------------------------------- */
int ofb_version(void);                 /* ABI version */
int ofb_num_sms(void);                 /* SM count of the current device */

/* ------------------------------------------------------------------------------------------------------------
 * GEMM   D[M,N] = sum_k A[m,k] * B[n,k]   (bf16 x bf16 -> fp32 in TMEM, tcgen05 + TMA)
 * epilogue ids: */
enum {
    OFB_EPI_STORE = 0,     /* out0 = rowscale*(acc+bias)*colscale + res
                              nn.Linear fwd / dgrad: layers.py:491 (qkv + gate 507-509), 515 (proj), 863 (fc2),
                              vision_transformer.py:744 (head); residual + DropPath vision_transformer.py:197,201 */
    OFB_EPI_FC1 = 1,       /* out0 = u = acc+bias ; out1 = gelu(u*gate)            layers.py:845-861 */
    OFB_EPI_FC2_DGRAD = 2, /* backward of layers.py:858-863: du, column partials of d gate and d bias */
    OFB_EPI_WGRAD = 3,     /* out0(fp32) += scale * A^T B, split-K (weight gradients of every Linear / conv) */
    OFB_EPI_PATCH = 4,     /* layers.py:177-191 + vision_transformer.py:628-637 fused */
    OFB_EPI_DECODER = 5    /* vision_transformer.py:720-729 fused (1x1 conv + pixel-shuffle + masked L1) */
};

typedef struct ofb_gemm_args {
    int32_t M, N, K;
    int32_t k_splits;          /* 0 = auto (WGRAD only) */
    void* out0; int32_t ld0;
    void* out1; int32_t ld1;
    int32_t out_fp32;
    const float* bias;
    const float* colscale;
    const float* rowscale; int32_t rows_per_scale;
    const void* res; int32_t ldres;       /* bf16 */
    const void* aux; int32_t ldaux;       /* bf16 */
    float* colpart0;
    float* colpart1;
    const float* scale_ptr;
    const float* pos;
    const float* mask_token;
    const float* rowmask;
    const float* target;
    int32_t tokens;
} ofb_gemm_args;

/* a_mn / b_mn: 0 = operand is [rows, K] row-major (K-major), 1 = operand is [K, rows] row-major (MN-major).
 * bn_hint: 0 = auto, else 64/128/192/256. */
int ofb_gemm_bf16(int epilogue, int a_mn, int b_mn, int bn_hint, const void* A, int lda, const void* B, int ldb,
                  const ofb_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OFB_B200_H */
