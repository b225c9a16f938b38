// capi.cu — extern "C" boundary (include/ofb_b200.h) over the kernels in this directory.
#include "../../include/ofb_b200.h"
#include "gemm.cuh"
#include <cuda_runtime.h>

namespace ofb {
int launch_gemm(int epi, int a_mn, int b_mn, int bn_hint, const void* A, int lda, const void* B, int ldb, GemmArgs g,
                cudaStream_t stream);
int num_sms();
int mlp_partial_rows(int N, int bn);
int launch_ln_fwd(const void*, const float*, const float*, void*, float*, float*, int, int, float, int, cudaStream_t);
int ln_bwd_grid(int M);
int launch_ln_bwd(const void*, const void*, const float*, const float*, const float*, void*, float*, float*, float*, const float*, int,
                  int, int, int, const void*, cudaStream_t);
int launch_soft_ce(const float*, const float*, float*, void*, int, int, float, cudaStream_t);
int launch_eval_metrics(const float*, const int64_t*, float*, int, int, cudaStream_t);
int launch_reduce_partials(const float*, int, int, float*, float, const float*, int, cudaStream_t);
int launch_reduce_partials_multi(const void*, int, cudaStream_t);
int launch_splitk_reduce(const void*, int, cudaStream_t);
int wgrad_plan(int M, int N, int K, int b_mn, int bn_hint);
int launch_patchify(const float*, void*, int, int, int, cudaStream_t);
int launch_copy_f32(const float*, float*, int, cudaStream_t);
int launch_mixup_batch(const float*, float*, int, int, double, int, int, int, int, int, cudaStream_t);
int launch_patchify_mixup(const float*, void*, int, int, int, double, int, int, int, int, int, cudaStream_t);
int launch_mixup_target(const long long*, float*, int, int, double, double, cudaStream_t);
int launch_pmim_mask(const float*, float*, int, int, int, cudaStream_t);
int launch_droppath_scale(const float*, const float*, float*, int, int, cudaStream_t);
int launch_cls_rows(const float*, const float*, const float*, void*, int, int, int, cudaStream_t);
int launch_embed_bwd(const void*, const void*, const float*, const float*, void*, float*, float*, float*, int, int, int, cudaStream_t);
int launch_norm_targets(const float*, const float*, float*, int, int, cudaStream_t);
int launch_ce(const float*, const int64_t*, float*, void*, int, int, float, float, cudaStream_t);
int launch_loss_finalize(const float*, int, const float*, int, const float*, int, const float*, float, float*, cudaStream_t);
int launch_adamw(float*, float*, float*, float*, void*, const float*, int, const long long*, long long, int, cudaStream_t);
int launch_cast_bf16(const float*, void*, long long, cudaStream_t);
int launch_colsum_bf16(const void*, int, int, int, float*, float, const float*, cudaStream_t);
int launch_bimask_fwd(const void*, int, int, const float*, const uint8_t*, const int*, const float*, float*, int*, float*, float*, float*,
                      cudaStream_t);
int launch_arch_finalize(const void*, int, const float*, const float*, int, int, int, int, int, int, int, int, float, float, float*, float*,
                         cudaStream_t);
int launch_bimask_bwd(const void*, int, int, const float*, const uint8_t*, const int*, const float*, const float*, const int*,
                      const float*, const float*, float, float*, cudaStream_t);
int launch_attn_fwd(const void*, void*, float*, const float*, int, int, int, float, cudaStream_t);
int launch_attn_bwd(const void*, const void*, const void*, const float*, const float*, const float*, void*, float*, float*, int, int, int,
                    float, cudaStream_t);
}  // namespace ofb

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int ofb_version(void) { return 3; }
int ofb_num_sms(void) { return ofb::num_sms(); }
int ofb_gemm_mlp_partial_rows(int n_tokens, int bn) { return ofb::mlp_partial_rows(n_tokens, bn); }

int ofb_gemm_bf16(int epilogue, int a_mn, int b_mn, int bn_hint, const void* A, int lda, const void* B, int ldb,
                  const ofb_gemm_args* a, void* stream) {
    if (a == nullptr || A == nullptr || B == nullptr) return 1000;
    ofb::GemmArgs g;
    g.M = a->M; g.N = a->N; g.K = a->K; g.k_splits = a->k_splits;
    g.out0 = a->out0; g.ld0 = a->ld0; g.out1 = a->out1; g.ld1 = a->ld1; g.out_fp32 = a->out_fp32; g.bias_rowscaled = a->bias_rowscaled;
    g.bias = a->bias; g.colscale = a->colscale; g.colscale_period = a->colscale_period; g.rowscale = a->rowscale;
    g.rows_per_scale = a->rows_per_scale > 0 ? a->rows_per_scale : 1;
    g.res = reinterpret_cast<const __nv_bfloat16*>(a->res); g.ldres = a->ldres;
    g.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux); g.ldaux = a->ldaux;
    g.colpart0 = a->colpart0; g.colpart1 = a->colpart1; g.scale_ptr = a->scale_ptr;
    g.pos = a->pos; g.mask_token = a->mask_token; g.rowmask = a->rowmask; g.target = a->target;
    g.tokens = a->tokens > 0 ? a->tokens : 1;
    g.splitk_ws = a->splitk_ws;
    return ofb::launch_gemm(epilogue, a_mn, b_mn, bn_hint, A, lda, B, ldb, g, reinterpret_cast<cudaStream_t>(stream));
}


int ofb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, int D, float eps,
                      void* stream) {
    return ofb::launch_ln_fwd(x, gamma, beta, y, mean, rstd, M, D, eps, D, ST(stream));
}
int ofb_layernorm_fwd_ex(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, int D,
                         int D_valid, float eps, void* stream) {
    return ofb::launch_ln_fwd(x, gamma, beta, y, mean, rstd, M, D, eps, D_valid, ST(stream));
}
int ofb_layernorm_bwd_parts(int M) { return ofb::ln_bwd_grid(M); }
int ofb_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                      float* part_dgamma, float* part_dbeta, float* part_dbias, const float* rowscale, int rows_per_scale, int M, int D,
                      void* stream) {
    return ofb::launch_ln_bwd(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale, rows_per_scale, M, D, D,
                              nullptr, ST(stream));
}
int ofb_layernorm_bwd_ex(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, const void* dres,
                         void* dx, float* part_dgamma, float* part_dbeta, float* part_dbias, const float* rowscale,
                         int rows_per_scale, int M, int D, int D_valid, void* stream) {
    return ofb::launch_ln_bwd(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale, rows_per_scale, M, D,
                              D_valid, dres, ST(stream));
}
int ofb_reduce_partials(const float* part, int R, int N, float* out, float scale, const float* div_by, int accumulate, void* stream) {
    return ofb::launch_reduce_partials(part, R, N, out, scale, div_by, accumulate, ST(stream));
}
int ofb_gemm_wgrad_splits(int M, int N, int K, int b_mn, int bn_hint) { return ofb::wgrad_plan(M, N, K, b_mn, bn_hint); }
int ofb_splitk_reduce(const ofb_splitk_job* jobs, int njobs, void* stream) {
    static_assert(sizeof(ofb_splitk_job) == 32, "ofb_splitk_job layout");
    if (jobs == nullptr) return 1000;
    return ofb::launch_splitk_reduce(jobs, njobs, ST(stream));
}
int ofb_reduce_partials_multi(const ofb_reduce_job* jobs, int njobs, void* stream) {
    static_assert(sizeof(ofb_reduce_job) == 40, "ofb_reduce_job layout");
    if (jobs == nullptr) return 1000;
    return ofb::launch_reduce_partials_multi(jobs, njobs, ST(stream));
}
int ofb_patchify(const float* images, void* patches, int B, int img, int patch, void* stream) {
    return ofb::launch_patchify(images, patches, B, img, patch, ST(stream));
}
int ofb_mixup_batch(const float* images, float* out, int B, int img, double lam, int cutmix, int yl, int yh, int xl, int xh,
                    void* stream) {
    return ofb::launch_mixup_batch(images, out, B, img, lam, cutmix, yl, yh, xl, xh, ST(stream));
}
int ofb_patchify_mixup(const float* images, void* patches, int B, int img, int patch, double lam, int cutmix, int yl, int yh, int xl,
                       int xh, void* stream) {
    return ofb::launch_patchify_mixup(images, patches, B, img, patch, lam, cutmix, yl, yh, xl, xh, ST(stream));
}
int ofb_mixup_target(const int64_t* labels, float* target, int B, int C, double lam, double smoothing, void* stream) {
    return ofb::launch_mixup_target(reinterpret_cast<const long long*>(labels), target, B, C, lam, smoothing, ST(stream));
}
int ofb_pmim_mask(const float* noise, float* mask, int B, int L, int keep, void* stream) {
    return ofb::launch_pmim_mask(noise, mask, B, L, keep, ST(stream));
}
int ofb_droppath_scale(const float* u, const float* drop_prob, float* scale, int n_rows, int B, void* stream) {
    return ofb::launch_droppath_scale(u, drop_prob, scale, n_rows, B, ST(stream));
}
int ofb_cls_rows(const float* cls, const float* pos, const float* gate, void* x, int B, int T, int D, void* stream) {
    return ofb::launch_cls_rows(cls, pos, gate, x, B, T, D, ST(stream));
}
int ofb_embed_bwd(const void* g0, const void* x0, const float* gate, const float* mask, void* dconv, float* part_gx, float* part_pos,
                  float* part_mt, int B, int T, int D, void* stream) {
    return ofb::launch_embed_bwd(g0, x0, gate, mask, dconv, part_gx, part_pos, part_mt, B, T, D, ST(stream));
}
int ofb_norm_targets(const float* images, const float* mask, float* target, int B, int img, void* stream) {
    return ofb::launch_norm_targets(images, mask, target, B, img, ST(stream));
}
int ofb_ls_cross_entropy(const float* logits, const int64_t* labels, float* loss_rows, void* dlogits, int B, int C, float smoothing,
                         float grad_scale, void* stream) {
    return ofb::launch_ce(logits, labels, loss_rows, dlogits, B, C, smoothing, grad_scale, ST(stream));
}
int ofb_soft_target_cross_entropy(const float* logits, const float* target, float* loss_rows, void* dlogits, int B, int C,
                                  float grad_scale, void* stream) {
    return ofb::launch_soft_ce(logits, target, loss_rows, dlogits, B, C, grad_scale, ST(stream));
}
int ofb_eval_metrics(const float* logits, const int64_t* labels, float* out_rows, int B, int C, void* stream) {
    return ofb::launch_eval_metrics(logits, labels, out_rows, B, C, ST(stream));
}
int ofb_loss_finalize(const float* loss_rows, int B, const float* dec_part, int n_dec_part, const float* mask, int n_mask,
                      const float* arch_loss, float grad_scale, float* scal, void* stream) {
    return ofb::launch_loss_finalize(loss_rows, B, dec_part, n_dec_part, mask, n_mask, arch_loss, grad_scale, scal, ST(stream));
}
int ofb_adamw(float* p, float* g, float* m, float* v, void* shadow, const float* hyper, int nseg, const int64_t* seg_end, int64_t n,
              int zero_grad, void* stream) {
    long long ends[OFB_ADAMW_MAX_SEGMENTS];
    if (nseg < 1 || nseg > OFB_ADAMW_MAX_SEGMENTS) return 1013;
    for (int i = 0; i < nseg; ++i) ends[i] = seg_end[i];
    return ofb::launch_adamw(p, g, m, v, shadow, hyper, nseg, ends, n, zero_grad, ST(stream));
}
int ofb_copy_f32(const float* src, float* dst, int n, void* stream) { return ofb::launch_copy_f32(src, dst, n, ST(stream)); }
int ofb_cast_bf16(const float* src, void* dst, int64_t n, void* stream) { return ofb::launch_cast_bf16(src, dst, n, ST(stream)); }
int ofb_colsum_bf16(const void* x, int ld, int R, int N, float* out, float scale, const float* scale_dev, void* stream) {
    return ofb::launch_colsum_bf16(x, ld, R, N, out, scale, scale_dev, ST(stream));
}

int ofb_bimask_fwd(const ofb_bimask_module* mods, int nmod, int max_n, const float* params, const uint8_t* switches,
                   const int32_t* widths, const float* w_p, float* gate, int32_t* rank, float* aprob, float* wsum, float* sp_loss,
                   void* stream) {
    return ofb::launch_bimask_fwd(mods, nmod, max_n, params, switches, widths, w_p, gate, rank, aprob, wsum, sp_loss, ST(stream));
}
int ofb_arch_finalize(const ofb_bimask_module* mods, int nmod, const float* wsum, const float* sp_loss, int depth, int D, int H, int d,
                      int hidden, int D_active, int L, int C, float target_flops, float w_flops, float* arch, float* dwsum,
                      void* stream) {
    static_assert(sizeof(ofb_bimask_module) == 64, "ofb_bimask_module layout");
    return ofb::launch_arch_finalize(mods, nmod, wsum, sp_loss, depth, D, H, d, hidden, D_active, L, C, target_flops, w_flops, arch,
                                     dwsum, ST(stream));
}
int ofb_bimask_bwd(const ofb_bimask_module* mods, int nmod, int max_n, const float* params, const uint8_t* switches,
                   const int32_t* widths, const float* w_p, const float* dgate, const int32_t* rank, const float* aprob,
                   const float* dwsum, float grad_scale, float* grads, void* stream) {
    return ofb::launch_bimask_bwd(mods, nmod, max_n, params, switches, widths, w_p, dgate, rank, aprob, dwsum, grad_scale, grads,
                                  ST(stream));
}
int ofb_attention_fwd(const void* qkv, void* o, float* lse, const float* drop_scale, int B, int T, int H, float scale, void* stream) {
    return ofb::launch_attn_fwd(qkv, o, lse, drop_scale, B, T, H, scale, ST(stream));
}
int ofb_attention_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, const float* gate, const float* drop_scale,
                      void* dqkv, float* part_gate, float* part_bias, int B, int T, int H, float scale, void* stream) {
    return ofb::launch_attn_bwd(qkv, o, d_o, lse, gate, drop_scale, dqkv, part_gate, part_bias, B, T, H, scale, ST(stream));
}

}  // extern "C"
