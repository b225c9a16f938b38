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
    OFB_EPI_FC1 = 1,       /* hidden activations TRANSPOSED [hidden, tokens]: rows (M) = hidden units = the fc1 weight,
                              columns (N) = tokens: out0 = u^T = acc+bias[row] ; out1 = h^T = rowscale[col/rows_per_scale] *
                              gelu(u^T*gate[row])                                   layers.py:845-861 */
    OFB_EPI_FC2_DGRAD = 2, /* backward of layers.py:858-863 in the same transposed layout (A = fc2 weight, MN-major):
                              out0 = du^T ; colpart0/1[ofb_gemm_mlp_partial_rows()][M] = per-tile token sums of d gate[row], d bias[row] */
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
    int32_t bias_rowscaled;    /* STORE: out0 = (acc + rowscale*bias)*colscale + res */
    const float* bias;
    const float* colscale; int32_t colscale_period;   /* >0: gate index = col % period (multiple of 32) */
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
    float* splitk_ws;          /* WGRAD, deterministic split-K: [k_splits][M][N] fp32 workspace or null (see ofb_splitk_reduce) */
} ofb_gemm_args;

/* a_mn / b_mn: 0 = operand is [rows, K] row-major (K-major), 1 = operand is [K, rows] row-major (MN-major).
 * bn_hint: 0 = auto, else 64/128/192/256. */
int ofb_gemm_bf16(int epilogue, int a_mn, int b_mn, int bn_hint, const void* A, int lda, const void* B, int ldb,
                  const ofb_gemm_args* args, void* stream);
/* rows R of the colpart0 / colpart1 buffers ([R][M] floats) the OFB_EPI_FC2_DGRAD epilogue writes for n_tokens columns at
 * column tile bn (64/128/192/256, must be passed as bn_hint): one row per (column tile, epilogue column group); returns -1
 * for an invalid bn. (Value, not an error code.) */
int ofb_gemm_mlp_partial_rows(int n_tokens, int bn);

/* ------------------------------------------------------------------------------------------------------------
 * LayerNorm (reference LayerNorm.forward, models/layers.py:96-98, eps 1e-6) — bf16 in/out, fp32 statistics.
 * D % 8 == 0, D <= 1024. */
int ofb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      int M, int D, float eps, void* stream);
/* Physically pruned embeddings (finetune of a searched subnet, models/vision_transformer.py:157-160 Block; pruned widths are
 * multiples of 12): the row holds D_valid real channels followed by D - D_valid (< 8) zero channels; statistics
 * run over D_valid. */
int ofb_layernorm_fwd_ex(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                         int M, int D, int D_valid, float eps, void* stream);
/* number of partial rows ofb_layernorm_bwd writes (size the part_* buffers as [ofb_layernorm_bwd_parts(M), D]) */
int ofb_layernorm_bwd_parts(int M);
/* backward + fused column partials: part_dgamma/part_dbeta always; part_dbias (may be NULL) = sum_rows
 * rowscale[row / rows_per_scale] * dx — the bias gradient of the Linear whose output (through the DropPath-scaled
 * residual add, vision_transformer.py:197,201) is this LayerNorm's input. */
int ofb_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                      void* dx, float* part_dgamma, float* part_dbeta, float* part_dbias, const float* rowscale,
                      int rows_per_scale, int M, int D, void* stream);
/* same with D_valid (padding columns of dx are zero) and dres (may be NULL): the gradient of the residual branch that
 * bypasses this LayerNorm in a pre-norm Block (vision_transformer.py:157-160), added to dx before it is stored and
 * column-summed into part_dbias. */
int ofb_layernorm_bwd_ex(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                         const void* dres, void* dx, float* part_dgamma, float* part_dbeta, float* part_dbias,
                         const float* rowscale, int rows_per_scale, int M, int D, int D_valid, void* stream);
/* Deterministic split-K of the weight-gradient GEMMs (the reference's autograd sums weight gradients in a fixed order;
 * red.global.add does not). With args->splitk_ws set, OFB_EPI_WGRAD stores each split's partial tile into
 * splitk_ws[split][M][N] instead of accumulating atomically; args->k_splits must be ofb_gemm_wgrad_splits(M, N, K, b_mn,
 * bn_hint) (the factor the library would pick itself; value, not an error code). ofb_splitk_reduce then does
 * out[i] += sum_s ws[s][i] (s ascending) for up to 8 GEMMs in one launch; n4 = M*N/4 (N % 4 == 0, out dense). */
int ofb_gemm_wgrad_splits(int M, int N, int K, int b_mn, int bn_hint);
typedef struct ofb_splitk_job {
    const float* ws;
    float* out;
    int64_t n4;
    int32_t splits;
    int32_t pad_;
} ofb_splitk_job;
int ofb_splitk_reduce(const ofb_splitk_job* jobs, int njobs, void* stream);

/* out[col] (+)= scale * sum_r part[r,col] / (div_by ? div_by[col] : 1) */
int ofb_reduce_partials(const float* part, int R, int N, float* out, float scale, const float* div_by, int accumulate,
                        void* stream);
/* up to 12 such reductions in one launch (the gradient pieces that fall out of the backward kernels of one block) */
typedef struct ofb_reduce_job {
    const float* part;
    float* out;
    const float* div_by;
    int32_t R, N;
    float scale;
    int32_t accumulate;
} ofb_reduce_job;
int ofb_reduce_partials_multi(const ofb_reduce_job* jobs, int njobs, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * token assembly (models/layers.py:177 im2col; vision_transformer.py:586-612 PMIM mask, 646-651 cls row; timm
 * DropPath used at vision_transformer.py:183) */
int ofb_patchify(const float* images, void* patches_bf16, int B, int img, int patch, void* stream);
/* timm.data.Mixup mode='batch' as applied by the post-search phase and the finetune loop (search.py:651-655, engine.py:98-99,
 * finetune.py:360-366): images <- lam x + (1 - lam) x.flip(0), or (cutmix != 0) the box [yl,yh) x [xl,xh) pasted from
 * x.flip(0). lam and the box are drawn on the host (numpy RNG, as timm does). out may alias images (in place, like timm).
 * ofb_patchify_mixup is ofb_patchify of the mixed batch without materialising it; ofb_mixup_target builds timm's
 * mixup_target: lam * smoothed one-hot(y) + (1 - lam) * smoothed one-hot(y.flip(0)), fp32 [B, C]. */
int ofb_mixup_batch(const float* images, float* out, int B, int img, double lam, int cutmix, int yl, int yh, int xl, int xh,
                    void* stream);
int ofb_patchify_mixup(const float* images, void* patches_bf16, int B, int img, int patch, double lam, int cutmix, int yl,
                       int yh, int xl, int xh, void* stream);
int ofb_mixup_target(const int64_t* labels, float* target, int B, int C, double lam, double smoothing, void* stream);
int ofb_pmim_mask(const float* noise, float* mask, int B, int L, int keep, void* stream);
int ofb_droppath_scale(const float* u, const float* drop_prob, float* scale, int n_rows, int B, void* stream);
int ofb_cls_rows(const float* cls, const float* pos, const float* gate, void* x_bf16, int B, int T, int D, void* stream);
/* backward of the embed stage: d(conv out), and [T, D] partials of d gate (sum g*x), d pos_embed, d mask_token */
int ofb_embed_bwd(const void* g0, const void* x0, const float* gate, const float* mask, void* dconv, float* part_gx,
                  float* part_pos, float* part_mt, int B, int T, int D, void* stream);
/* norm_targets (vision_transformer.py:121-141, window 47) at masked patches only, patch-major [B*L, 768] fp32;
   img % 16 == 0 and a 16-byte aligned image base (the halo is read as float4), else 1012 */
int ofb_norm_targets(const float* images, const float* mask, float* target, int B, int img, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * losses: timm LabelSmoothingCrossEntropy fwd+bwd (search.py:584); loss assembly of engine.py:134-144.
 * scal[0..6] = {base, arch, decoder, total, w_dec, decoder_grad_scale, n_masked} */
int ofb_ls_cross_entropy(const float* logits, const int64_t* labels, float* loss_rows, void* dlogits_bf16, int B, int C,
                         float smoothing, float grad_scale, void* stream);
/* timm SoftTargetCrossEntropy fwd+bwd on Mixup targets (finetune.py:388-389, search.py:655): target fp32 [B, C] */
int ofb_soft_target_cross_entropy(const float* logits, const float* target, float* loss_rows, void* dlogits_bf16, int B,
                                  int C, float grad_scale, void* stream);
/* evaluate() metrics (engine.py:222-257): out_rows[b] = {cross entropy, top-1 hit, top-5 hit} */
int ofb_eval_metrics(const float* logits, const int64_t* labels, float* out_rows, int B, int C, void* stream);
int ofb_loss_finalize(const float* loss_rows, int B, const float* dec_part, int n_dec_part, const float* mask,
                      int n_mask, const float* arch_loss, float grad_scale, float* scal, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * fused multi-segment AdamW (optim.py:56-120; groups search.py:486-559). hyper[seg*8 + {0..6}] =
 * {lr, weight_decay, beta1, beta2, eps, 1-beta1^t, 1-beta2^t}; seg_end[] exclusive ends (multiples of 4).
 * Writes the bf16 shadow copy used by the GEMMs and optionally zeroes the gradient. Up to OFB_ADAMW_MAX_SEGMENTS
 * segments: the finetune optimizer (torch.optim.AdamW over lr_decay.param_groups_lrd, finetune.py:378-383) has
 * 2 x (depth + 2) of them; same update rule. */
#define OFB_ADAMW_MAX_SEGMENTS 64
int ofb_adamw(float* p, float* g, float* m, float* v, void* shadow_bf16, const float* hyper, int nseg,
              const int64_t* seg_end, int64_t n, int zero_grad, void* stream);
int ofb_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
/* dst[0..n) = src[0..n) by a kernel; src may be pinned HOST memory (UVA). Used for the per-step hyper vector so that the upload
 * never waits on a copy engine behind the bulk H2D copy of the next batch (engine.py:118-120 lr / schedule values). */
int ofb_copy_f32(const float* src, float* dst, int n, void* stream);
/* out[col] += scale * (scale_dev ? *scale_dev : 1) * sum_rows x[row, col]  (bias gradients of head / decoder) */
int ofb_colsum_bf16(const void* x, int ld, int R, int N, float* out, float scale, const float* scale_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * bi-mask gates of all searchable modules + architecture losses, one launch each
 * (models/layers.py:179-191, 494-509, 847-858; base_model.py:31-86; vision_transformer.py:759-783). */
typedef struct ofb_bimask_module {
    int32_t kind;        /* 0 embed, 1 mlp, 2 attention */
    int32_t dim, heads, n_i, n_j;
    int32_t switch_off, width_off, gate_off;
    int32_t stride;      /* distance between heads in the score tensor and in gate / rank / dgate (>= dim; 64 for an attention
                            module whose head dim has been pruned below 64: heads stay 64 wide physically) */
    int32_t pad_;
    int64_t alpha_off, score_off;     /* offsets (floats) into the parameter / gradient arenas */
    float coef, loss_w;
} ofb_bimask_module;
#define OFB_BIMASK_MAX_CELLS 64
int ofb_bimask_fwd(const ofb_bimask_module* mods_dev, int nmod, int max_n, const float* params, const uint8_t* switches,
                   const int32_t* widths, const float* w_p_dev, float* gate, int32_t* rank, float* aprob, float* wsum,
                   float* sp_loss, void* stream);
/* arch[0..6] = {loss_arch, l_attn, l_mlp, l_embed, l_flops, searched GFLOPs, original GFLOPs}; dwsum[nmod].
 * D, H, d, hidden: ORIGINAL dims (total-FLOPs side, layers.py:747-753); D_active: current LayerNorm width after truncating
 * prune events (vision_transformer.py:210; 0 = D); active head counts are read from the attention module records. */
int ofb_arch_finalize(const ofb_bimask_module* mods_dev, int nmod, const float* wsum, const float* sp_loss, int depth,
                      int D, int H, int d, int hidden, int D_active, int L, int C, float target_flops, float w_flops,
                      float* arch, float* dwsum, void* stream);
int ofb_bimask_bwd(const ofb_bimask_module* mods_dev, int nmod, int max_n, const float* params, const uint8_t* switches,
                   const int32_t* widths, const float* w_p_dev, const float* dgate, const int32_t* rank,
                   const float* aprob, const float* dwsum, float grad_scale, float* grads, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * fused attention on gated q,k,v (models/layers.py:507-514), head_dim 64, T <= 208.
 * qkv: bf16 [B, T, 3, H, 64] (the qkv GEMM output as is); o: bf16 [B, T, H*64] multiplied by drop_scale[b];
 * lse: fp32 [B, H, T]. Backward writes d(pre-gate qkv) and per-image partials of d gate / d bias. */
int ofb_attention_fwd(const void* qkv, void* o, float* lse, const float* drop_scale, int B, int T, int H, float scale,
                      void* stream);
int ofb_attention_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, const float* gate,
                      const float* drop_scale, void* dqkv, float* part_gate, float* part_bias, int B, int T, int H,
                      float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OFB_B200_H */
