// gemm.cuh — argument block shared by the tcgen05 GEMM kernel and its launchers.
#pragma once
#include <stdint.h>
#include <cuda_bf16.h>

namespace ofb {

enum GemmEpilogue : int {
    EPI_STORE = 0,      // out0 = rowscale*(acc + bias)*colscale + res           (bf16 or fp32 out)
                        //   (bias_rowscaled: out0 = (acc + rowscale*bias)*colscale + res)
    EPI_FC1 = 1,        // TRANSPOSED hidden layout (rows = hidden units, columns = tokens):
                        //   out0 = u^T = acc + bias[row] ; out1 = h^T = rowscale[col] * gelu(u^T * gate[row])   (layers.py:845-861)
    EPI_FC2_DGRAD = 2,  // transposed: dh = rowscale[col]*acc ; du^T = dh*gelu'(u g)*g -> out0 ; per-row token sums of d gate, d bias
    EPI_WGRAD = 3,      // out0(fp32) += scale * acc   (split-K, red.global.add)
    EPI_PATCH = 4,      // patch-embed: gate, pos-embed, PMIM mask-token select, row remap (skip cls row)
    EPI_DECODER = 5,    // PMIM decoder: x_rec = acc + bias ; masked L1 vs normalised target ; sign*mask -> out0
};

struct GemmArgs {
    int M, N, K;            // D[M,N] = sum_k A[m,k] * B[n,k]
    int k_splits;           // >1 only for EPI_WGRAD
    void* out0; int ld0;
    void* out1; int ld1;
    int out_fp32;           // EPI_STORE: 1 -> out0 is float
    int bias_rowscaled;     // EPI_STORE: 1 -> out0 = acc + rowscale*bias + res (input rows already carry rowscale)
    const float* bias;      // [N] or null
    const float* colscale;  // [N] or null (bi-mask gate)
    int colscale_period;    // >0: gate index = col % period (q,k,v share one [H*d] gate, layers.py:507-509)
    const float* rowscale;  // [ceil(M/rows_per_scale)] or null (drop-path keep/scale per sample)
    int rows_per_scale;
    const __nv_bfloat16* res; int ldres;   // residual or null
    const __nv_bfloat16* aux; int ldaux;   // EPI_FC2_DGRAD: saved pre-gate fc1 output u
    float* colpart0;        // EPI_FC2_DGRAD: [2 * n_tiles][M] token-sum partials of d gate[row]  / EPI_DECODER: [tiles*8] loss partials (one per epilogue warp)
    float* colpart1;        // EPI_FC2_DGRAD: [2 * n_tiles][M] token-sum partials of d bias[row]
    const float* scale_ptr; // EPI_WGRAD / EPI_STORE: optional device scalar multiplied into acc
    // EPI_PATCH / EPI_DECODER
    const float* pos;        // [tokens+1, N] fp32 positional embedding (row 0 = cls)
    const float* mask_token; // [N]
    const float* rowmask;    // [B*tokens] PMIM mask (1 = masked/removed)
    const float* target;     // EPI_DECODER: [B*tokens, N] normalised pixel targets, patch-major
    int tokens;              // patches per image (196)
    // EPI_WGRAD, deterministic split-K: when non-null every (split, tile) stores its fp32 partial tile to
    // splitk_ws[split][M][N] (dense, plain stores) instead of red.global.add into out0; launch_splitk_reduce then adds the
    // splits in fixed order. k_splits must then be the value wgrad_plan() returns (no empty split).
    float* splitk_ws;
};

}  // namespace ofb
