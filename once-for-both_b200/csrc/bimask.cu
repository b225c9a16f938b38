// bimask.cu — the bi-mask gate of every searchable module in ONE launch, its backward, and the architecture losses.
//
// Reference (per module, every forward): layers.py:179-191 (embed), 494-509 (attention), 847-858 (MLP) build
//   gate = w_p*sigmoid(score) + (1-w_p)*rank_gather(sum_alive softmax(alpha)_ij * mask_ij)
// with ~100 micro-kernels and H2D copies; base_model.py:37-86 (sparsity loss) and vision_transformer.py:759-783
// (FLOPs loss) add another ~25x20.  Here: closed forms
//   * sum_alive a_ij*mask_ij over prefix masks  ==  table[hr][cr] = sum_ij a_ij [n_i > hr][w_j > cr]
//   * double argsort + gather                   ==  rank by counting (ties: lower index first)
// one CTA per module, fp32 throughout, FLOPs polynomial in double.
#include "ptx.cuh"
#include "launch.cuh"
#include <math.h>

namespace ofb {

struct BimaskModule {
    int kind;             // 0 embed, 1 mlp, 2 attention
    int dim;              // 1-D modules: width; attention: head_dim
    int heads;            // attention: H, else 1
    int n_i, n_j;         // alpha shape
    int switch_off;       // into uint8 switch array (n_i*n_j entries, row-major)
    int width_off;        // into int widths array: n_j channel widths, then n_i head counts (attention)
    int gate_off;         // into gate / rank / dgate buffers (heads*stride entries)
    int stride;           // distance between heads in the score tensor and the gate / rank / dgate buffers (>= dim): a pruned
                          // attention module keeps its heads 64 wide physically (zero-padded layout, finetune_engine.py)
    long long alpha_off;  // into the fp32 parameter arena
    long long score_off;
    float coef;           // score-norm coefficient: 4e-4 attention, 1e-4 otherwise (base_model.py:72-75)
    float loss_w;         // w_head / w_mlp / w_embedding (search.py:173-175)
};

struct ArchDims {
    int depth, D, H, d, hidden, L, C;   // ORIGINAL dims: the reference keeps them for the total-FLOPs side (layers.py:747-753)
    int D_active;                       // current LayerNorm width (vt:210 active_dim); active heads come from the module records
    float target_flops, w_flops;
};

static constexpr int BM_THREADS = 256;
static constexpr int MAX_CELLS = 64;

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < BM_THREADS / 32; ++i) t += red[i];
    return t;
}

// shared: a[MAX_CELLS], sig[heads*dim], (attention) headsum[heads], rank_h[heads]
__global__ void __launch_bounds__(BM_THREADS) bimask_fwd_kernel(const BimaskModule* __restrict__ mods, const float* __restrict__ params,
                                                                const uint8_t* __restrict__ switches, const int* __restrict__ widths,
                                                                const float* __restrict__ w_p_ptr, float* __restrict__ gate,
                                                                int* __restrict__ rank, float* __restrict__ aprob, float* __restrict__ wsum,
                                                                float* __restrict__ sp_loss) {
    pdl_wait();
    extern __shared__ float bsm[];
    __shared__ float a[MAX_CELLS];
    __shared__ float red[BM_THREADS / 32];
    __shared__ float headsum[64];
    __shared__ int rank_h[64];
    __shared__ int s_alive;
    const BimaskModule md = mods[blockIdx.x];
    const int tid = threadIdx.x;
    const int n = md.heads * md.dim;
    // blockIdx.y: a module's units are split over gridDim.y CTAs (the rank of a unit is a count over the whole row: 1536^2
    // comparisons for an MLP module, which one CTA per module turned into a 120 us kernel at the head of every step). Every part
    // loads all scores (cheap); part 0 alone writes the per-module outputs.
    const int part = blockIdx.y, per_part = (n + int(gridDim.y) - 1) / int(gridDim.y);
    const int i_begin = part * per_part, i_end = min(n, i_begin + per_part);
    const int ncell = md.n_i * md.n_j;
    const float* alpha = params + md.alpha_off;
    const float* score = params + md.score_off;
    const uint8_t* sw = switches + md.switch_off;
    const int* wj = widths + md.width_off;
    const int* ni = wj + md.n_j;
    const float w_p = *w_p_ptr;
    float* sig = bsm;            // [n]
    float* sc = bsm + n;         // [n] raw scores

    // ---- alive softmax + sparsity loss (thread 0; <= 64 cells) ----
    if (tid == 0) {
        float mx = -INFINITY;
        int alive = 0;
        for (int k = 0; k < ncell; ++k) if (sw[k]) { mx = fmaxf(mx, alpha[k]); ++alive; }
        float se = 0.f;
        for (int k = 0; k < ncell; ++k) { a[k] = sw[k] ? expf(alpha[k] - mx) : 0.f; se += a[k]; }
        float ent = 0.f, var = 0.f;
        const float mean = 1.f / alive;
        for (int k = 0; k < ncell; ++k) {
            a[k] /= se;
            if (sw[k]) { ent -= a[k] * logf(a[k]); var += (a[k] - mean) * (a[k] - mean); }
        }
        float l = 0.f;
        if (alive > 1) {
            const float sigma = var / (1.f - 1.f / alive);
            l = ent + tanf(1.5707963267948966f - 3.14159265358979323846f * sigma) / alive;
        }
        if (part == 0) {
            sp_loss[blockIdx.x] = l;     // score-norm term added below
            for (int k = 0; k < ncell; ++k) aprob[blockIdx.x * MAX_CELLS + k] = a[k];
        }
        s_alive = alive;
    }
    float ssum = 0.f;
    for (int i = tid; i < n; i += BM_THREADS) {
        const float s = score[(i / md.dim) * md.stride + (i % md.dim)];
        sc[i] = s;
        const float sg = 1.f / (1.f + expf(-s));
        sig[i] = sg;
        ssum += sg;
    }
    ssum = block_sum(ssum, red);   // includes __syncthreads -> a[], sig[], sc[] visible
    const int alive = s_alive;
    if (tid == 0 && part == 0 && alive > 1) sp_loss[blockIdx.x] += md.coef * ssum;

    // ---- head ranks (attention): descending by sum_c sigmoid(score[h,c]) ----
    if (md.kind == 2) {
        if (tid < md.heads) {
            float hs = 0.f;
            for (int c = 0; c < md.dim; ++c) hs += sig[tid * md.dim + c];
            headsum[tid] = hs;
        }
        __syncthreads();
        if (tid < md.heads) {
            int r = 0;
            for (int h = 0; h < md.heads; ++h) r += (headsum[h] > headsum[tid]) || (headsum[h] == headsum[tid] && h < tid);
            rank_h[tid] = r;
        }
        __syncthreads();
    }

    // ---- per-channel rank (within head), table lookup, gate ----
    for (int i = i_begin + tid; i < i_end; i += BM_THREADS) {
        const int h = i / md.dim, c = i % md.dim;
        const float s = sc[i];
        const float* row = sc + h * md.dim;
        int r = 0;
        for (int k = 0; k < md.dim; ++k) r += (row[k] > s) || (row[k] == s && k < c);
        const int hr = (md.kind == 2) ? rank_h[h] : 0;
        float t = 0.f;   // table[hr][r]
        for (int ii = 0; ii < md.n_i; ++ii) {
            const bool hok = (md.kind == 2) ? (ni[ii] > hr) : true;
            if (!hok) continue;
            for (int jj = 0; jj < md.n_j; ++jj)
                if (wj[jj] > r) t += a[ii * md.n_j + jj];
        }
        // a finished module (one cell left: its score has been finalised, layers.py:629 / 939 / 275) gates with the frozen score
        // itself (layers.py:196-197, 518-521, 859-860)
        gate[md.gate_off + h * md.stride + c] = alive > 1 ? w_p * sig[i] + (1.f - w_p) * t : s;
        rank[md.gate_off + h * md.stride + c] = hr * md.dim + r;
    }
    // weighted_mask.sum(): the ranks of a row are a permutation of 0..dim-1, so the sum over units of table[head rank][rank] does
    // not depend on the scores: sum_h sum_ii [n_i > head rank] sum_jj a[ii][jj] * min(w_jj, dim)
    if (tid == 0 && part == 0) {
        float tsum = 0.f;
        for (int h = 0; h < md.heads; ++h) {
            const int hr = (md.kind == 2) ? rank_h[h] : 0;
            for (int ii = 0; ii < md.n_i; ++ii) {
                if (md.kind == 2 && !(ni[ii] > hr)) continue;
                for (int jj = 0; jj < md.n_j; ++jj) tsum += a[ii * md.n_j + jj] * float(min(wj[jj], md.dim));
            }
        }
        wsum[blockIdx.x] = tsum;
    }
}

// ---- FLOPs loss + total architecture loss + d loss / d wsum (vision_transformer.py:759-783, losses.py:93-102) ----
// module order: [0] = embed, then per block: attention (1+2l), mlp (2+2l).
// out: arch[0] = loss_arch, arch[1] = l_attn, arch[2] = l_mlp, arch[3] = l_embed, arch[4] = l_flops,
//      arch[5] = searched GFLOPs, arch[6] = original GFLOPs
__global__ void arch_finalize_kernel(const BimaskModule* __restrict__ mods, int nmod, const float* __restrict__ wsum,
                                     const float* __restrict__ sp_loss, ArchDims ad, float* __restrict__ arch, float* __restrict__ dwsum) {
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double n = ad.L, D = ad.D, H = ad.H, d = ad.d, hid = ad.hidden, C = ad.C, Da = ad.D_active;
    const double ae = wsum[0];
    double f_ori = ad.L * D * 768.0, f_s = ad.L * ae * 768.0;
    double dae = ad.L * 768.0;
    double l_attn = 0, l_mlp = 0, l_embed = sp_loss[0];
    for (int l = 0; l < ad.depth; ++l) {
        const double sd = wsum[1 + 2 * l], sm = wsum[2 + 2 * l];
        const double Ha = mods[1 + 2 * l].heads;        // active_H (layers.py:749)
        f_ori += 2 * D * n;
        f_s += 2 * Da * n;
        f_ori += n * (D * 3 * D) + 3 * n * D + H * n * d * n + H * n * n + 5 * H * n * n + H * n * n * d + n * D * D + n * D;
        f_s += n * (ae * 3 * sd) + 3 * n * sd + n * n * sd + Ha * n * n + 5 * Ha * n * n + n * n * sd + n * (sd * ae) + n * ae;
        f_ori += (2 * D * hid + D + hid) * n;
        f_s += (ae * sm * 2 + ae + sm) * n;
        dae += n * 4 * sd + n + (2 * sm + 1) * n;
        l_attn += sp_loss[1 + 2 * l];
        l_mlp += sp_loss[2 + 2 * l];
    }
    f_ori += D * C;
    f_s += ae * C;
    dae += C;
    const double fo = f_ori / 1e9, fs = f_s / 1e9;
    const double r = (fs - ad.target_flops) / fo;
    const double l_flops = r * r;
    const double dl_dfs = ad.w_flops * 2.0 * r / fo / 1e9;   // d (w_flops*l_flops) / d f_s (raw flops)
    dwsum[0] = float(dl_dfs * dae);
    for (int l = 0; l < ad.depth; ++l) {
        dwsum[1 + 2 * l] = float(dl_dfs * (n * (4 * ae + 3 + 2 * n)));
        dwsum[2 + 2 * l] = float(dl_dfs * ((2 * ae + 1) * n));
    }
    double la = 0, lm = 0, le = 0;
    // loss weights live in the module records (all attention modules share w_attn, ...)
    la = mods[1].loss_w * l_attn;
    lm = mods[2].loss_w * l_mlp;
    le = mods[0].loss_w * l_embed;
    arch[0] = float(la + lm + le + ad.w_flops * l_flops);
    arch[1] = float(l_attn); arch[2] = float(l_mlp); arch[3] = float(l_embed); arch[4] = float(l_flops);
    arch[5] = float(fs); arch[6] = float(fo);
    (void)nmod;
}

// ---- backward: dgate (from the network) + d wsum (FLOPs loss) + sparsity loss -> d score, d alpha (accumulated) ----
__global__ void __launch_bounds__(BM_THREADS) bimask_bwd_kernel(const BimaskModule* __restrict__ mods, const float* __restrict__ params,
                                                                const uint8_t* __restrict__ switches, const int* __restrict__ widths,
                                                                const float* __restrict__ w_p_ptr, const float* __restrict__ dgate,
                                                                const int* __restrict__ rank, const float* __restrict__ aprob,
                                                                const float* __restrict__ dwsum, float grad_scale, float* __restrict__ grads) {
    pdl_wait();
    extern __shared__ float bsm[];
    __shared__ float red[BM_THREADS / 32];
    __shared__ float da[MAX_CELLS];
    const BimaskModule md = mods[blockIdx.x];
    const int tid = threadIdx.x;
    const int n = md.heads * md.dim;
    const int ncell = md.n_i * md.n_j;
    const float* score = params + md.score_off;
    const uint8_t* sw = switches + md.switch_off;
    const int* wj = widths + md.width_off;
    const int* ni = wj + md.n_j;
    const float* a = aprob + blockIdx.x * MAX_CELLS;
    const float w_p = *w_p_ptr;
    float* dtable = bsm;   // [n], indexed by combined rank hr*dim + cr
    int alive = 0;
    for (int k = 0; k < ncell; ++k) alive += sw[k] ? 1 : 0;
    const float dws = dwsum[blockIdx.x] * grad_scale;

    for (int i = tid; i < n; i += BM_THREADS) {
        const int ph = (i / md.dim) * md.stride + (i % md.dim);       // physical position (head stride)
        const float dg = dgate[md.gate_off + ph];
        const float s = score[ph];
        const float sg = 1.f / (1.f + expf(-s));
        if (alive > 1) {
            const float dsig = dg * w_p + md.loss_w * md.coef * grad_scale;
            grads[md.score_off + ph] += dsig * sg * (1.f - sg);
        } else {
            grads[md.score_off + ph] += dg;                            // finished module: gate == score
        }
        dtable[rank[md.gate_off + ph]] = dg * (1.f - w_p) + dws;
    }
    __syncthreads();
    if (alive <= 1) return;   // alpha frozen (finish_search): no alpha gradient
    // da_ij = sum over the prefix rectangle of dtable
    for (int cell = 0; cell < ncell; ++cell) {
        float s = 0.f;
        if (sw[cell]) {
            const int ii = cell / md.n_j, jj = cell % md.n_j;
            const int hlim = (md.kind == 2) ? ni[ii] : 1;
            const int clim = wj[jj];
            for (int i = tid; i < n; i += BM_THREADS) {
                const int hr = i / md.dim, cr = i % md.dim;
                if (hr < hlim && cr < clim) s += dtable[i];
            }
        }
        s = block_sum(s, red);
        if (tid == 0) da[cell] = s;
    }
    __syncthreads();
    if (tid == 0) {
        // sparsity-loss gradient w.r.t. p (base_model.py:58-70) joins before the softmax backward
        const float mean = 1.f / alive;
        const float ts = 1.f - 1.f / alive;
        float var = 0.f;
        for (int k = 0; k < ncell; ++k) if (sw[k]) var += (a[k] - mean) * (a[k] - mean);
        const float sigma = var / ts;
        const float cs = cosf(1.5707963267948966f - 3.14159265358979323846f * sigma);
        const float dvar_dsigma = -3.14159265358979323846f / (cs * cs) / alive;
        float dot = 0.f;
        for (int k = 0; k < ncell; ++k) {
            if (!sw[k]) { da[k] = 0.f; continue; }
            da[k] += md.loss_w * grad_scale * (-(logf(a[k]) + 1.f) + dvar_dsigma * 2.f * (a[k] - mean) / ts);
            dot += a[k] * da[k];
        }
        for (int k = 0; k < ncell; ++k)
            if (sw[k]) grads[md.alpha_off + k] += a[k] * (da[k] - dot);
    }
}

// =============================================================================================
int launch_bimask_fwd(const void* mods, int nmod, int max_n, const float* params, const uint8_t* switches, const int* widths,
                      const float* w_p_ptr, float* gate, int* rank, float* aprob, float* wsum, float* sp_loss, cudaStream_t s) {
    const size_t smem = size_t(2) * max_n * sizeof(float);
    if (smem > 48 * 1024) return 1020;
    const int parts = max_n > 512 ? 8 : (max_n > 256 ? 2 : 1);
    OFB_LAUNCH(bimask_fwd_kernel, dim3(nmod, parts), BM_THREADS, smem, s, reinterpret_cast<const BimaskModule*>(mods), params, switches, widths,
                                                                 w_p_ptr, gate, rank, aprob, wsum, sp_loss);
    return int(cudaGetLastError());
}

int launch_arch_finalize(const void* mods, int nmod, const float* wsum, const float* sp_loss, int depth, int D, int H, int d, int hidden,
                         int D_active, int L, int C, float target_flops, float w_flops, float* arch, float* dwsum, cudaStream_t s) {
    ArchDims ad{depth, D, H, d, hidden, L, C, D_active > 0 ? D_active : D, target_flops, w_flops};
    OFB_LAUNCH(arch_finalize_kernel, 1, 32, 0, s, reinterpret_cast<const BimaskModule*>(mods), nmod, wsum, sp_loss, ad, arch, dwsum);
    return int(cudaGetLastError());
}

int launch_bimask_bwd(const void* mods, int nmod, int max_n, const float* params, const uint8_t* switches, const int* widths,
                      const float* w_p_ptr, const float* dgate, const int* rank, const float* aprob, const float* dwsum,
                      float grad_scale, float* grads, cudaStream_t s) {
    const size_t smem = size_t(max_n) * sizeof(float);
    if (smem > 48 * 1024) return 1020;
    OFB_LAUNCH(bimask_bwd_kernel, nmod, BM_THREADS, smem, s, reinterpret_cast<const BimaskModule*>(mods), params, switches, widths, w_p_ptr, dgate,
                                                     rank, aprob, dwsum, grad_scale, grads);
    return int(cudaGetLastError());
}

}  // namespace ofb
