// elementwise.cu — bandwidth-bound kernels of the bi-mask search step (sm_100a):
// LayerNorm fwd/bwd (+ fused column reductions), patchify/cast, PMIM mask + DropPath scales, cls-row assembly,
// embed backward, local 47x47 target normalisation (masked patches only), label-smoothing CE fwd+bwd,
// loss finalisation, column-partial reduction, fused multi-segment AdamW (+ bf16 shadow weights + grad zeroing).
// All are vectorised (16-byte accesses), warp-shuffle reduced, grid-strided over 148 SMs.
#include "ptx.cuh"
#include "launch.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace ofb {

int num_sms();
static inline int err() { return int(cudaGetLastError()); }

// =============================================================================================
// LayerNorm forward   (reference: LayerNorm.forward layers.py:96-98, eps 1e-6; 25 per step)
//   x, y: bf16 [M, D]; gamma/beta fp32 [D]; mean/rstd fp32 [M].  One warp per row, D % 8 == 0, D <= 1024.
// =============================================================================================
// LN_MAXC = 16-byte chunks per lane, a template parameter (1..4 <-> D <= 256, 512, 768, 1024) so that narrow models
// do not pay registers (and occupancy) for columns they do not have.

__device__ __forceinline__ void unpack8(const uint4& p, float* f) {
    float2 t;
    t = unpack_bf16x2(p.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16x2(p.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16x2(p.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16x2(p.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 p;
    p.x = pack_bf16x2(f[0], f[1]); p.y = pack_bf16x2(f[2], f[3]);
    p.z = pack_bf16x2(f[4], f[5]); p.w = pack_bf16x2(f[6], f[7]);
    return p;
}

template <int LN_MAXC>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, int M, int D, float eps,
                                                     int Dv) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    // Dv <= D: number of real channels; columns [Dv, D) are zero padding of a physically pruned embedding (they hold zeros,
    // carry gamma = beta = 0 and must not enter the statistics: the reference normalises over Dv channels)
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int nchunk = D >> 3;
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + size_t(row) * D);
        float v[LN_MAXC][8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunk) {
                unpack8(__ldg(xr + c), v[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += v[i][j];
            }
        }
        const float mu = warp_sum(s) / Dv;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunk) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mu; q += (c * 8 + j < Dv) ? d * d : 0.f; }
            }
        }
        const float rs = rsqrtf(warp_sum(q) / Dv + eps);
        uint4* yr = reinterpret_cast<uint4*>(y + size_t(row) * D);
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunk) {
                float o[8];
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c);
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c + 1);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c);
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c + 1);
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mu) * rs * gg[j] + bb[j];
                yr[c] = pack8(o);
            }
        }
        if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    }
}

// =============================================================================================
// LayerNorm backward (autograd of layers.py:96-98) with fused column reductions:
//   dx = rstd * (dy*gamma - mean_D(dy*gamma) - xhat * mean_D(dy*gamma*xhat))
//   part_dgamma[blk] = sum_rows dy*xhat, part_dbeta[blk] = sum_rows dy,
//   part_dbias[blk]  = sum_rows rowscale[row/rows_per_scale] * dx   (bias gradient of the Linear that produced x:
//                      x = res + droppath*(.. + bias), vision_transformer.py:197,201)   [optional]
//   partial buffers are [gridDim.x, D]; a second kernel reduces them (deterministic).
// =============================================================================================
// Rows are staged by 1-D bulk async copies (cp.async.bulk + mbarrier) into a per-warp ring of LN_STAGES rows, so every warp
// keeps LN_STAGES x 2 row loads in flight without holding them in registers.
static constexpr int LN_STAGES = 4;
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar)
                 : "memory");
}
template <int LN_MAXC>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
                                                     float* __restrict__ part_dgamma, float* __restrict__ part_dbeta,
                                                     float* __restrict__ part_dbias, const float* __restrict__ rowscale,
                                                     int rows_per_scale, int M, int D, int Dv,
                                                     const __nv_bfloat16* __restrict__ dres) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    // Dv: real channels (see ln_fwd_kernel); dres (optional): gradient arriving over the residual connection that bypasses
    // this LayerNorm (pre-norm blocks, vision_transformer.py:157-160), added to dx before it is stored / column-summed
    extern __shared__ __align__(128) uint8_t ln_smem_raw[];
    // layout: ring [8 warps][LN_STAGES][2][D bf16] | barriers [8][LN_STAGES]; the ring is reused as float [3][8][D] at the end
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int WPB = 8;
    const int nchunk = D >> 3;
    const uint32_t row_bytes = uint32_t(D) * 2u;
    const uint32_t ring_bytes = WPB * LN_STAGES * 2 * row_bytes;
    const uint32_t red_bytes = 3u * WPB * uint32_t(D) * 4u;
    const uint32_t bar_off = (ring_bytes > red_bytes ? ring_bytes : red_bytes);
    uint8_t* my_ring = ln_smem_raw + size_t(warp) * LN_STAGES * 2 * row_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem_raw + bar_off) + warp * LN_STAGES;

    if (lane == 0) {
        for (int sidx = 0; sidx < LN_STAGES; ++sidx) mbar_init(smem_u32(&bars[sidx]), 1);
        mbar_fence_init();
    }
    __syncwarp();

    float ag[LN_MAXC][8], ab[LN_MAXC][8], ad[LN_MAXC][8];
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = ad[i][j] = 0.f;
    float gam[LN_MAXC][8];
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunk) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c);
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c + 1);
            gam[i][0] = g0.x; gam[i][1] = g0.y; gam[i][2] = g0.z; gam[i][3] = g0.w;
            gam[i][4] = g1.x; gam[i][5] = g1.y; gam[i][6] = g1.z; gam[i][7] = g1.w;
        }
    }

    const int row_stride = gridDim.x * WPB;
    const int first = blockIdx.x * WPB + warp;
    auto issue = [&](int row, int stage) {
        const uint32_t b = smem_u32(&bars[stage]);
        const uint32_t dst = smem_u32(my_ring + size_t(stage) * 2 * row_bytes);
        mbar_arrive_expect_tx(b, 2 * row_bytes);
        bulk_load_1d(dst, dy + size_t(row) * D, row_bytes, b);
        bulk_load_1d(dst + row_bytes, x + size_t(row) * D, row_bytes, b);
    };
    if (lane == 0) {
        for (int sidx = 0; sidx < LN_STAGES; ++sidx) {
            const int row = first + sidx * row_stride;
            if (row < M) issue(row, sidx);
        }
    }
    int k = 0;
    for (int row = first; row < M; row += row_stride, ++k) {
        const int stage = k % LN_STAGES;
        const uint32_t parity = (k / LN_STAGES) & 1;
        const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
        const float rsc = (part_dbias != nullptr) ? (rowscale != nullptr ? __ldg(rowscale + row / rows_per_scale) : 1.f) : 0.f;
        mbar_wait(smem_u32(&bars[stage]), parity);
        const uint4* sdy = reinterpret_cast<const uint4*>(my_ring + size_t(stage) * 2 * row_bytes);
        const uint4* sx = reinterpret_cast<const uint4*>(my_ring + size_t(stage) * 2 * row_bytes + row_bytes);
        float dyv[LN_MAXC][8], xh[LN_MAXC][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunk) {
                unpack8(sdy[c], dyv[i]);
                unpack8(sx[c], xh[i]);
            }
        }
        // the stage is consumed (values are in registers): refill it with the row LN_STAGES iterations ahead
        __syncwarp();
        if (lane == 0) {
            const int nxt = row + LN_STAGES * row_stride;
            if (nxt < M) { fence_proxy_async_smem(); issue(nxt, stage); }
        }
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunk) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xh[i][j] = (xh[i][j] - mu) * rs;
                    const float dg = dyv[i][j] * gam[i][j];
                    s1 += dg;
                    s2 += dg * xh[i][j];
                    ag[i][j] += dyv[i][j] * xh[i][j];
                    ab[i][j] += dyv[i][j];
                }
            }
        }
        s1 = warp_sum(s1) / Dv;
        s2 = warp_sum(s2) / Dv;
        uint4* dxr = reinterpret_cast<uint4*>(dx + size_t(row) * D);
#pragma unroll
        for (int i = 0; i < LN_MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunk) {
                float o[8], rr[8];
                if (dres != nullptr) {
                    unpack8(__ldg(reinterpret_cast<const uint4*>(dres + size_t(row) * D) + c), rr);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) rr[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j] = (c * 8 + j < Dv ? rs * (dyv[i][j] * gam[i][j] - s1 - xh[i][j] * s2) : 0.f) + rr[j];
                    ad[i][j] += rsc * o[j];
                }
                dxr[c] = pack8(o);
            }
        }
    }
    // cross-warp reduction of the column accumulators (the ring is idle now: every issued copy has been consumed)
    __syncthreads();
    float* sg = reinterpret_cast<float*>(ln_smem_raw);
    float* sb = sg + WPB * D;
    float* sd = sg + 2 * WPB * D;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunk) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sg[warp * D + c * 8 + j] = ag[i][j];
                sb[warp * D + c * 8 + j] = ab[i][j];
                sd[warp * D + c * 8 + j] = ad[i][j];
            }
        }
    }
    __syncthreads();
    for (int col = threadIdx.x; col < D; col += blockDim.x) {
        float a = 0.f, b = 0.f, d = 0.f;
        for (int w = 0; w < WPB; ++w) {
            a += sg[w * D + col]; b += sb[w * D + col]; d += sd[w * D + col];
        }
        part_dgamma[size_t(blockIdx.x) * D + col] = a;
        part_dbeta[size_t(blockIdx.x) * D + col] = b;
        if (part_dbias != nullptr) part_dbias[size_t(blockIdx.x) * D + col] = d;
    }
}

// =============================================================================================
// Fast paths for D = 192 W (W = 1, 2, 4 <-> DeiT-T/S/B): every lane owns exactly three chunks of W packed bf16x2 words, so
// there is no ragged last chunk, and the arithmetic runs on packed fp32x2 (the generic kernels above are instruction-issue
// bound: 47 thread-instructions per element measured with ncu, r01). Same math, same partial-buffer layout.
// =============================================================================================
template <int W> struct WordVec;
template <> struct WordVec<1> { using T = uint32_t; };
template <> struct WordVec<2> { using T = uint2; };
template <> struct WordVec<4> { using T = uint4; };
template <int W>
__device__ __forceinline__ void words_to_pairs(const typename WordVec<W>::T& v, float2* out) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
    for (int k = 0; k < W; ++k) out[k] = unpack_bf16x2(w[k]);
}
template <int W>
__device__ __forceinline__ typename WordVec<W>::T pairs_to_words(const float2* in) {
    typename WordVec<W>::T v;
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int k = 0; k < W; ++k) w[k] = pack_bf16x2(in[k].x, in[k].y);
    return v;
}
__device__ __forceinline__ float2 warp_sum2(float2 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

template <int W>
__global__ void __launch_bounds__(256) ln_fwd3_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                      float* __restrict__ mean, float* __restrict__ rstd, int M, float eps) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    using V = typename WordVec<W>::T;
    constexpr int D = 192 * W, NP = 3 * W;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    float2 gam[NP], bet[NP];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < W; ++k) {
            gam[i * W + k] = __ldg(reinterpret_cast<const float2*>(gamma) + (lane + 32 * i) * W + k);
            bet[i * W + k] = __ldg(reinterpret_cast<const float2*>(beta) + (lane + 32 * i) * W + k);
        }
    const int stride = gridDim.x * wpb;
    int row = blockIdx.x * wpb + (threadIdx.x >> 5);
    V cur[3], nxt[3];
    if (row < M) {
#pragma unroll
        for (int i = 0; i < 3; ++i) cur[i] = __ldg(reinterpret_cast<const V*>(x + size_t(row) * D) + lane + 32 * i);
    }
    for (; row < M; row += stride) {
        const int rn = row + stride;
        if (rn < M) {
#pragma unroll
            for (int i = 0; i < 3; ++i) nxt[i] = __ldg(reinterpret_cast<const V*>(x + size_t(rn) * D) + lane + 32 * i);
        }
        float2 v[NP];
#pragma unroll
        for (int i = 0; i < 3; ++i) words_to_pairs<W>(cur[i], v + i * W);
        float2 s = v[0];
#pragma unroll
        for (int k = 1; k < NP; ++k) s = add2(s, v[k]);
        const float mu = warp_sum(s.x + s.y) * (1.f / D);
        const float2 nmu = splat2(-mu);
        float2 q = splat2(0.f);
#pragma unroll
        for (int k = 0; k < NP; ++k) { v[k] = add2(v[k], nmu); q = fma2(v[k], v[k], q); }
        const float rs = rsqrtf(warp_sum(q.x + q.y) * (1.f / D) + eps);
        const float2 rs2 = splat2(rs);
#pragma unroll
        for (int k = 0; k < NP; ++k) v[k] = fma2(mul2(v[k], rs2), gam[k], bet[k]);
#pragma unroll
        for (int i = 0; i < 3; ++i) reinterpret_cast<V*>(y + size_t(row) * D)[lane + 32 * i] = pairs_to_words<W>(v + i * W);
        if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
        for (int i = 0; i < 3; ++i) cur[i] = nxt[i];
    }
}

// Two rows per warp iteration with the next two rows in flight: at D = 192 a row is only 12 bytes per lane, so the one-row
// kernel is bound by shuffle / load latency (in situ, DeiT-Tiny batch 1024: 40.3 -> 33.5 us); at D >= 384 the extra registers
// cost an occupancy step and the one-row kernel stays faster (16.8 vs 18.9 us), so only W = 1 is dispatched here.
template <int W>
__global__ void __launch_bounds__(256) ln_fwd3x2_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                        float* __restrict__ mean, float* __restrict__ rstd, int M, float eps) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    using V = typename WordVec<W>::T;
    constexpr int D = 192 * W, NP = 3 * W, RB = 2;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    float2 gam[NP], bet[NP];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < W; ++k) {
            gam[i * W + k] = __ldg(reinterpret_cast<const float2*>(gamma) + (lane + 32 * i) * W + k);
            bet[i * W + k] = __ldg(reinterpret_cast<const float2*>(beta) + (lane + 32 * i) * W + k);
        }
    const int stride = gridDim.x * wpb;            // rows r, r + stride of one iteration; the next iteration is 2 * stride on
    int row = blockIdx.x * wpb + (threadIdx.x >> 5);
    V cur[RB][3], nxt[RB][3];
#pragma unroll
    for (int b = 0; b < RB; ++b) {
        const int r = row + b * stride;
        if (r < M) {
#pragma unroll
            for (int i = 0; i < 3; ++i) cur[b][i] = __ldg(reinterpret_cast<const V*>(x + size_t(r) * D) + lane + 32 * i);
        }
    }
    for (; row < M; row += RB * stride) {
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int rn = row + (RB + b) * stride;
            if (rn < M) {
#pragma unroll
                for (int i = 0; i < 3; ++i) nxt[b][i] = __ldg(reinterpret_cast<const V*>(x + size_t(rn) * D) + lane + 32 * i);
            }
        }
        const bool ok1 = row + stride < M;
        float2 v[RB][NP];
        float sum[RB];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
#pragma unroll
            for (int i = 0; i < 3; ++i) words_to_pairs<W>(cur[b][i], v[b] + i * W);
            float2 s = v[b][0];
#pragma unroll
            for (int k = 1; k < NP; ++k) s = add2(s, v[b][k]);
            sum[b] = s.x + s.y;
        }
        if (!ok1) sum[1] = 0.f;
        const float2 mu2 = warp_sum2(make_float2(sum[0], sum[1]));
        float qs[RB];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const float2 nmu = splat2(-(b == 0 ? mu2.x : mu2.y) * (1.f / D));
            float2 q = splat2(0.f);
#pragma unroll
            for (int k = 0; k < NP; ++k) { v[b][k] = add2(v[b][k], nmu); q = fma2(v[b][k], v[b][k], q); }
            qs[b] = q.x + q.y;
        }
        const float2 q2 = warp_sum2(make_float2(qs[0], qs[1]));
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int r = row + b * stride;
            if (r < M) {
                const float mu = (b == 0 ? mu2.x : mu2.y) * (1.f / D);
                const float rs = rsqrtf((b == 0 ? q2.x : q2.y) * (1.f / D) + eps);
                const float2 rs2 = splat2(rs);
#pragma unroll
                for (int k = 0; k < NP; ++k) v[b][k] = fma2(mul2(v[b][k], rs2), gam[k], bet[k]);
#pragma unroll
                for (int i = 0; i < 3; ++i) reinterpret_cast<V*>(y + size_t(r) * D)[lane + 32 * i] = pairs_to_words<W>(v[b] + i * W);
                if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
            }
        }
#pragma unroll
        for (int b = 0; b < RB; ++b)
#pragma unroll
            for (int i = 0; i < 3; ++i) cur[b][i] = nxt[b][i];
    }
}

template <int W>
__global__ void __launch_bounds__(256) ln_bwd3_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                      const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
                                                      float* __restrict__ part_dgamma, float* __restrict__ part_dbeta,
                                                      float* __restrict__ part_dbias, const float* __restrict__ rowscale,
                                                      int rows_per_scale, int M) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    using V = typename WordVec<W>::T;
    constexpr int D = 192 * W, NP = 3 * W, WPB = 8;
    extern __shared__ __align__(128) uint8_t ln_smem_raw[];
    // layout: ring [8 warps][LN_STAGES][2][D bf16] | barriers [8][LN_STAGES]; the ring is reused as float [3][8][D] at the end
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t row_bytes = D * 2u;
    constexpr uint32_t ring_bytes = WPB * LN_STAGES * 2 * row_bytes;
    constexpr uint32_t red_bytes = 3u * WPB * D * 4u;
    constexpr uint32_t bar_off = (ring_bytes > red_bytes ? ring_bytes : red_bytes);
    uint8_t* my_ring = ln_smem_raw + size_t(warp) * LN_STAGES * 2 * row_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem_raw + bar_off) + warp * LN_STAGES;
    if (lane == 0) {
        for (int sidx = 0; sidx < LN_STAGES; ++sidx) mbar_init(smem_u32(&bars[sidx]), 1);
        mbar_fence_init();
    }
    __syncwarp();

    float2 ag[NP], ab[NP], ad[NP], gam[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) ag[k] = ab[k] = ad[k] = splat2(0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < W; ++k) gam[i * W + k] = __ldg(reinterpret_cast<const float2*>(gamma) + (lane + 32 * i) * W + k);

    const int row_stride = gridDim.x * WPB;
    const int first = blockIdx.x * WPB + warp;
    auto issue = [&](int row, int stage) {
        const uint32_t b = smem_u32(&bars[stage]);
        const uint32_t dst = smem_u32(my_ring + size_t(stage) * 2 * row_bytes);
        mbar_arrive_expect_tx(b, 2 * row_bytes);
        bulk_load_1d(dst, dy + size_t(row) * D, row_bytes, b);
        bulk_load_1d(dst + row_bytes, x + size_t(row) * D, row_bytes, b);
    };
    if (lane == 0) {
        for (int sidx = 0; sidx < LN_STAGES; ++sidx) {
            const int row = first + sidx * row_stride;
            if (row < M) issue(row, sidx);
        }
    }
    const bool want_bias = part_dbias != nullptr;
    auto row_stats = [&](int row, float& mu, float& rs, float& rsc) {
        mu = __ldg(mean + row); rs = __ldg(rstd + row);
        rsc = want_bias ? (rowscale != nullptr ? __ldg(rowscale + row / rows_per_scale) : 1.f) : 0.f;
    };
    float mu = 0.f, rs = 0.f, rsc = 0.f;
    if (first < M) row_stats(first, mu, rs, rsc);
    int k = 0;
    for (int row = first; row < M; row += row_stride, ++k) {
        const int stage = k % LN_STAGES;
        const uint32_t parity = (k / LN_STAGES) & 1;
        float mu_n = 0.f, rs_n = 0.f, rsc_n = 0.f;
        if (row + row_stride < M) row_stats(row + row_stride, mu_n, rs_n, rsc_n);     // one row ahead
        mbar_wait(smem_u32(&bars[stage]), parity);
        const V* sdy = reinterpret_cast<const V*>(my_ring + size_t(stage) * 2 * row_bytes);
        const V* sx = reinterpret_cast<const V*>(my_ring + size_t(stage) * 2 * row_bytes + row_bytes);
        float2 dyv[NP], xh[NP];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            words_to_pairs<W>(sdy[lane + 32 * i], dyv + i * W);
            words_to_pairs<W>(sx[lane + 32 * i], xh + i * W);
        }
        // the stage is consumed (values are in registers): refill it with the row LN_STAGES iterations ahead
        __syncwarp();
        if (lane == 0) {
            const int nxt = row + LN_STAGES * row_stride;
            if (nxt < M) { fence_proxy_async_smem(); issue(nxt, stage); }
        }
        const float2 rs2 = splat2(rs), nmr2 = splat2(-mu * rs);
        float2 s1 = splat2(0.f), s2 = splat2(0.f);
        float2 dg[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            xh[j] = fma2(xh[j], rs2, nmr2);
            dg[j] = mul2(dyv[j], gam[j]);
            s1 = add2(s1, dg[j]);
            s2 = fma2(dg[j], xh[j], s2);
            ag[j] = fma2(dyv[j], xh[j], ag[j]);
            ab[j] = add2(ab[j], dyv[j]);
        }
        const float2 ss = warp_sum2(make_float2(s1.x + s1.y, s2.x + s2.y));
        const float2 c1 = splat2(-ss.x * (1.f / D) * rs), c2 = splat2(-ss.y * (1.f / D) * rs);
        const float2 rsc2 = splat2(rsc);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            dg[j] = fma2(xh[j], c2, fma2(dg[j], rs2, c1));          // rs * (dy*gamma - s1 - xhat*s2)
            ad[j] = fma2(rsc2, dg[j], ad[j]);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) reinterpret_cast<V*>(dx + size_t(row) * D)[lane + 32 * i] = pairs_to_words<W>(dg + i * W);
        mu = mu_n; rs = rs_n; rsc = rsc_n;
    }
    // cross-warp reduction of the column accumulators (the ring is idle now: every issued copy has been consumed)
    __syncthreads();
    float2* sg = reinterpret_cast<float2*>(ln_smem_raw);
    float2* sb = sg + WPB * D / 2;
    float2* sd = sg + 2 * WPB * D / 2;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const int idx = warp * (D / 2) + (lane + 32 * i) * W + j;
            sg[idx] = ag[i * W + j]; sb[idx] = ab[i * W + j]; sd[idx] = ad[i * W + j];
        }
    __syncthreads();
    const float* fg = reinterpret_cast<const float*>(sg);
    const float* fb = reinterpret_cast<const float*>(sb);
    const float* fd = reinterpret_cast<const float*>(sd);
    for (int col = threadIdx.x; col < D; col += blockDim.x) {
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int w = 0; w < WPB; ++w) { a += fg[w * D + col]; b += fb[w * D + col]; d += fd[w * D + col]; }
        part_dgamma[size_t(blockIdx.x) * D + col] = a;
        part_dbeta[size_t(blockIdx.x) * D + col] = b;
        if (want_bias) part_dbias[size_t(blockIdx.x) * D + col] = d;
    }
}

// =============================================================================================
// Packed paths for ANY even width (the pruned embeddings of the search: multiples of 12, or 6 at DeiT-T, zero-padded to a
// multiple of 8; configs[4] finetune, search steps after truncating prune events, post-search phase): a lane owns NW bf16x2
// words at word index lane + 32 i. Only the last word slot is ragged, so lanes stay 81-100 % busy where the 16-byte-chunk
// kernels above drop to 56 % at D = 288 (36 chunks over 64 slots) - and the arithmetic is packed fp32x2 like the 192 W paths.
// D: physical row width (multiple of 8), Dv: real channels (even, D - 8 < Dv <= D); words >= Dv / 2 are zero padding that
// stays out of the statistics and gets y = 0 / dx = dres.
// =============================================================================================
template <int NW>
__global__ void __launch_bounds__(256) ln_fwdw_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                      float* __restrict__ mean, float* __restrict__ rstd, int M, int D, int Dv,
                                                      float eps) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int nw = D >> 1, nv = Dv >> 1;
    float2 gam[NW], bet[NW];
    bool ok[NW], sv[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const int w = lane + 32 * i;
        ok[i] = w < nw; sv[i] = w < nv;
        gam[i] = ok[i] ? __ldg(reinterpret_cast<const float2*>(gamma) + w) : splat2(0.f);
        bet[i] = ok[i] ? __ldg(reinterpret_cast<const float2*>(beta) + w) : splat2(0.f);
    }
    const float inv = 1.f / Dv;
    const int stride = gridDim.x * wpb;
    int row = blockIdx.x * wpb + (threadIdx.x >> 5);
    uint32_t cur[NW], nxt[NW];
    if (row < M) {
#pragma unroll
        for (int i = 0; i < NW; ++i) cur[i] = ok[i] ? __ldg(reinterpret_cast<const uint32_t*>(x + size_t(row) * D) + lane + 32 * i) : 0u;
    }
    for (; row < M; row += stride) {
        const int rn = row + stride;
        if (rn < M) {
#pragma unroll
            for (int i = 0; i < NW; ++i) nxt[i] = ok[i] ? __ldg(reinterpret_cast<const uint32_t*>(x + size_t(rn) * D) + lane + 32 * i) : 0u;
        }
        float2 v[NW];
        float2 s = splat2(0.f);
#pragma unroll
        for (int i = 0; i < NW; ++i) { v[i] = unpack_bf16x2(cur[i]); s = add2(s, v[i]); }
        const float mu = warp_sum(s.x + s.y) * inv;
        const float2 nmu = splat2(-mu);
        float2 q = splat2(0.f);
#pragma unroll
        for (int i = 0; i < NW; ++i) { v[i] = sv[i] ? add2(v[i], nmu) : splat2(0.f); q = fma2(v[i], v[i], q); }
        const float rs = rsqrtf(warp_sum(q.x + q.y) * inv + eps);
        const float2 rs2 = splat2(rs);
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            v[i] = fma2(mul2(v[i], rs2), gam[i], bet[i]);
            if (ok[i]) reinterpret_cast<uint32_t*>(y + size_t(row) * D)[lane + 32 * i] = pack_bf16x2(v[i].x, v[i].y);
        }
        if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
        for (int i = 0; i < NW; ++i) cur[i] = nxt[i];
    }
}

template <int NW>
__global__ void __launch_bounds__(256) ln_bwdw_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                      const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
                                                      float* __restrict__ part_dgamma, float* __restrict__ part_dbeta,
                                                      float* __restrict__ part_dbias, const float* __restrict__ rowscale,
                                                      int rows_per_scale, int M, int D, int Dv,
                                                      const __nv_bfloat16* __restrict__ dres) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    constexpr int WPB = 8;
    extern __shared__ __align__(128) uint8_t ln_smem_raw[];
    // layout as in ln_bwd_kernel: ring [8 warps][LN_STAGES][2][D bf16] | barriers; the ring is reused as float [3][8][D] at the end
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = D >> 1, nv = Dv >> 1;
    const uint32_t row_bytes = uint32_t(D) * 2u;
    const uint32_t ring_bytes = WPB * LN_STAGES * 2 * row_bytes;
    const uint32_t red_bytes = 3u * WPB * uint32_t(D) * 4u;
    const uint32_t bar_off = (ring_bytes > red_bytes ? ring_bytes : red_bytes);
    uint8_t* my_ring = ln_smem_raw + size_t(warp) * LN_STAGES * 2 * row_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem_raw + bar_off) + warp * LN_STAGES;
    if (lane == 0) {
        for (int sidx = 0; sidx < LN_STAGES; ++sidx) mbar_init(smem_u32(&bars[sidx]), 1);
        mbar_fence_init();
    }
    __syncwarp();

    float2 ag[NW], ab[NW], ad[NW], gam[NW];
    bool ok[NW], sv[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const int w = lane + 32 * i;
        ok[i] = w < nw; sv[i] = w < nv;
        ag[i] = ab[i] = ad[i] = splat2(0.f);
        gam[i] = ok[i] ? __ldg(reinterpret_cast<const float2*>(gamma) + w) : splat2(0.f);
    }
    const int row_stride = gridDim.x * WPB;
    const int first = blockIdx.x * WPB + warp;
    auto issue = [&](int row, int stage) {
        const uint32_t b = smem_u32(&bars[stage]);
        const uint32_t dst = smem_u32(my_ring + size_t(stage) * 2 * row_bytes);
        mbar_arrive_expect_tx(b, 2 * row_bytes);
        bulk_load_1d(dst, dy + size_t(row) * D, row_bytes, b);
        bulk_load_1d(dst + row_bytes, x + size_t(row) * D, row_bytes, b);
    };
    if (lane == 0) {
        for (int sidx = 0; sidx < LN_STAGES; ++sidx) {
            const int row = first + sidx * row_stride;
            if (row < M) issue(row, sidx);
        }
    }
    const bool want_bias = part_dbias != nullptr, has_res = dres != nullptr;
    auto row_stats = [&](int row, float& mu, float& rs, float& rsc) {
        mu = __ldg(mean + row); rs = __ldg(rstd + row);
        rsc = want_bias ? (rowscale != nullptr ? __ldg(rowscale + row / rows_per_scale) : 1.f) : 0.f;
    };
    auto load_res = [&](int row, uint32_t* r) {
#pragma unroll
        for (int i = 0; i < NW; ++i)
            r[i] = (has_res && ok[i]) ? __ldg(reinterpret_cast<const uint32_t*>(dres + size_t(row) * D) + lane + 32 * i) : 0u;
    };
    const float inv = 1.f / Dv;
    float mu = 0.f, rs = 0.f, rsc = 0.f;
    uint32_t res[NW], res_n[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) res[i] = res_n[i] = 0u;
    if (first < M) { row_stats(first, mu, rs, rsc); load_res(first, res); }
    int k = 0;
    for (int row = first; row < M; row += row_stride, ++k) {
        const int stage = k % LN_STAGES;
        const uint32_t parity = (k / LN_STAGES) & 1;
        float mu_n = 0.f, rs_n = 0.f, rsc_n = 0.f;
        if (row + row_stride < M) { row_stats(row + row_stride, mu_n, rs_n, rsc_n); load_res(row + row_stride, res_n); }   // one row ahead
        mbar_wait(smem_u32(&bars[stage]), parity);
        const uint32_t* sdy = reinterpret_cast<const uint32_t*>(my_ring + size_t(stage) * 2 * row_bytes);
        const uint32_t* sx = reinterpret_cast<const uint32_t*>(my_ring + size_t(stage) * 2 * row_bytes + row_bytes);
        float2 dyv[NW], xh[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            dyv[i] = ok[i] ? unpack_bf16x2(sdy[lane + 32 * i]) : splat2(0.f);
            xh[i] = ok[i] ? unpack_bf16x2(sx[lane + 32 * i]) : splat2(0.f);
        }
        // the stage is consumed (values are in registers): refill it with the row LN_STAGES iterations ahead
        __syncwarp();
        if (lane == 0) {
            const int nxt = row + LN_STAGES * row_stride;
            if (nxt < M) { fence_proxy_async_smem(); issue(nxt, stage); }
        }
        const float2 rs2 = splat2(rs), nmr2 = splat2(-mu * rs);
        float2 s1 = splat2(0.f), s2 = splat2(0.f);
        float2 dg[NW];
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            xh[j] = sv[j] ? fma2(xh[j], rs2, nmr2) : splat2(0.f);
            dg[j] = mul2(dyv[j], gam[j]);
            s1 = add2(s1, dg[j]);
            s2 = fma2(dg[j], xh[j], s2);
            ag[j] = fma2(dyv[j], xh[j], ag[j]);
            ab[j] = add2(ab[j], dyv[j]);
        }
        const float2 ss = warp_sum2(make_float2(s1.x + s1.y, s2.x + s2.y));
        const float2 c1 = splat2(-ss.x * inv * rs), c2 = splat2(-ss.y * inv * rs);
        const float2 rsc2 = splat2(rsc);
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            dg[j] = sv[j] ? fma2(xh[j], c2, fma2(dg[j], rs2, c1)) : splat2(0.f);          // rs * (dy*gamma - s1 - xhat*s2)
            if (has_res) dg[j] = add2(dg[j], unpack_bf16x2(res[j]));
            ad[j] = fma2(rsc2, dg[j], ad[j]);
            if (ok[j]) reinterpret_cast<uint32_t*>(dx + size_t(row) * D)[lane + 32 * j] = pack_bf16x2(dg[j].x, dg[j].y);
        }
        mu = mu_n; rs = rs_n; rsc = rsc_n;
#pragma unroll
        for (int i = 0; i < NW; ++i) res[i] = res_n[i];
    }
    // cross-warp reduction of the column accumulators (the ring is idle now: every issued copy has been consumed)
    __syncthreads();
    float2* sg = reinterpret_cast<float2*>(ln_smem_raw);
    float2* sb = sg + WPB * nw;
    float2* sd = sg + 2 * WPB * nw;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        if (ok[i]) {
            const int idx = warp * nw + lane + 32 * i;
            sg[idx] = ag[i]; sb[idx] = ab[i]; sd[idx] = ad[i];
        }
    }
    __syncthreads();
    const float* fg = reinterpret_cast<const float*>(sg);
    const float* fb = reinterpret_cast<const float*>(sb);
    const float* fd = reinterpret_cast<const float*>(sd);
    for (int col = threadIdx.x; col < D; col += blockDim.x) {
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int w = 0; w < WPB; ++w) { a += fg[w * D + col]; b += fb[w * D + col]; d += fd[w * D + col]; }
        part_dgamma[size_t(blockIdx.x) * D + col] = a;
        part_dbeta[size_t(blockIdx.x) * D + col] = b;
        if (want_bias) part_dbias[size_t(blockIdx.x) * D + col] = d;
    }
}

// =============================================================================================
// out[col] (+)= scale * (colscale ? 1/colscale[col] : 1) * sum_r part[r, col]      (deterministic tree per column)
// =============================================================================================
__global__ void reduce_partials_kernel(const float* __restrict__ part, int R, int N, float* __restrict__ out, float scale,
                                       const float* __restrict__ inv_colscale, int accumulate) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = threadIdx.x >> 5, ny = blockDim.x >> 5;
    __shared__ float sm[32][33];
    float s = 0.f;
    if (col < N)
        for (int r = ty; r < R; r += ny) s += part[size_t(r) * N + col];
    sm[ty][threadIdx.x & 31] = s;
    __syncthreads();
    if (ty == 0 && col < N) {
        float t = 0.f;
        for (int i = 0; i < ny; ++i) t += sm[i][threadIdx.x];
        t *= scale;
        if (inv_colscale != nullptr) { const float dv = inv_colscale[col]; t = dv != 0.f ? t / dv : 0.f; }   // zero gate = padding channel
        out[col] = accumulate ? out[col] + t : t;
    }
}

// several independent reductions in one launch (blockIdx.y = job): the gradient pieces that fall out of one backward
// kernel (LayerNorm: d gamma, d beta, d bias; fc2 dgrad: d gate, d bias; ...) are finished together.
struct ReduceJob { const float* part; float* out; const float* div_by; int R, N; float scale; int accumulate; };
struct ReduceJobs { ReduceJob j[12]; };
__global__ void reduce_partials_multi_kernel(const ReduceJobs jobs) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const ReduceJob jb = jobs.j[blockIdx.y];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    if (blockIdx.x * 32 >= jb.N) return;
    const int ty = threadIdx.x >> 5, ny = blockDim.x >> 5;
    __shared__ float sm[32][33];
    float s = 0.f;
    if (col < jb.N) {
        // eight independent row loads in flight per thread (the longest job - 788 partial rows of the fc2 data-gradient GEMM - is
        // 25 rows per thread, each a DRAM round trip); fixed addition order
        const float* p = jb.part + col;
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = 0.f;
        int r = ty;
        for (; r + 7 * ny < jb.R; r += 8 * ny) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(p + size_t(r + i * ny) * jb.N);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] += v[i];
        }
        for (; r < jb.R; r += ny) a[0] += __ldg(p + size_t(r) * jb.N);
        s = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    }
    sm[ty][threadIdx.x & 31] = s;
    __syncthreads();
    if (ty == 0 && col < jb.N) {
        float t = 0.f;
        for (int i = 0; i < ny; ++i) t += sm[i][threadIdx.x];
        t *= jb.scale;
        if (jb.div_by != nullptr) { const float dv = jb.div_by[col]; t = dv != 0.f ? t / dv : 0.f; }       // zero gate = padding channel
        jb.out[col] = jb.accumulate ? jb.out[col] + t : t;
    }
}

// =============================================================================================
// deterministic split-K finish of the weight-gradient GEMMs: out[i] += sum_s ws[s][i], s ascending - the run-to-run order of
// red.global.add is gone, two runs of a step are bit-identical. Up to 8 GEMMs (one block's four weight gradients, or the
// embed / head / decoder ones) per launch; float4 columns, the splits of a column stay in registers.
// =============================================================================================
struct SplitkJob { const float* ws; float* out; long long n4; int splits; int pad_; };
struct SplitkJobs { SplitkJob j[8]; };
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const SplitkJobs jobs) {
    pdl_wait();
    const SplitkJob jb = jobs.j[blockIdx.y];
    const float4* ws = reinterpret_cast<const float4*>(jb.ws);
    float4* out = reinterpret_cast<float4*>(jb.out);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < jb.n4; i += (long long)gridDim.x * blockDim.x) {
        float4 a = ws[i];
        // ascending split order (bit-reproducible); the partial loads go out four at a time
        int s = 1;
        for (; s + 3 < jb.splits; s += 4) {
            const float4 b0 = ws[s * jb.n4 + i], b1 = ws[(s + 1) * jb.n4 + i], b2 = ws[(s + 2) * jb.n4 + i], b3 = ws[(s + 3) * jb.n4 + i];
            a.x += b0.x; a.y += b0.y; a.z += b0.z; a.w += b0.w;
            a.x += b1.x; a.y += b1.y; a.z += b1.z; a.w += b1.w;
            a.x += b2.x; a.y += b2.y; a.z += b2.z; a.w += b2.w;
            a.x += b3.x; a.y += b3.y; a.z += b3.z; a.w += b3.w;
        }
        for (; s < jb.splits; ++s) {
            const float4 b = ws[s * jb.n4 + i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        float4 o = out[i];
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        out[i] = o;
    }
}

// =============================================================================================
// patchify + cast: images fp32 [B,3,224,224] -> bf16 [B*196, 768] with column c*256 + i*16 + j
// (the im2col of the 16x16/16 conv, layers.py:177)
// =============================================================================================
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int HW, int P) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    // one thread = 8 consecutive pixels of one patch row
    const int G = HW / P;                       // 14
    const int per_img = 3 * HW * HW / 8;
    const size_t total = size_t(B) * per_img;
    for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
        const int b = idx / per_img;
        int r = idx % per_img;
        const int x8 = r % (HW / 8); r /= (HW / 8);
        const int yy = r % HW; const int c = r / HW;
        const float4* src = reinterpret_cast<const float4*>(img + ((size_t(b) * 3 + c) * HW + yy) * HW + x8 * 8);
        const float4 a = __ldg(src), d = __ldg(src + 1);
        const float f[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
        const int ph = yy / P, i = yy % P, pw = (x8 * 8) / P, j = (x8 * 8) % P;
        const size_t orow = size_t(b) * G * G + ph * G + pw;
        *reinterpret_cast<uint4*>(out + orow * (3 * P * P) + c * P * P + i * P + j) = pack8(f);
    }
}

// =============================================================================================
// Mixup / CutMix of a batch, timm.data.Mixup mode='batch' as the post-search phase applies it (search.py:651-655,
// engine.py:98-99; finetune.py:360-366):  x <- lam x + (1 - lam) x.flip(0)   or   x[:, :, yl:yh, xl:xh] <- x.flip(0)[same box].
// A thread owns 4 pixels of the PAIR (b, B-1-b): both images are read once and both results written (in place allowed; an
// odd batch's middle image pairs with itself). fp32 arithmetic in timm's order: fl(fl(x lam) + fl(x' (1 - lam))).
// =============================================================================================
struct MixParams {
    float lam, oml;            // lam and 1 - lam (both rounded from the host double, as torch does with a Python scalar)
    int cutmix, yl, yh, xl, xh;
};
__device__ __forceinline__ void mix_pair4(const float4& a, const float4& b, int x0, int yy, const MixParams& mp, float4& oa, float4& ob) {
    if (mp.cutmix) {
        const bool row_in = yy >= mp.yl && yy < mp.yh;
        const bool i0 = row_in && x0 >= mp.xl && x0 < mp.xh, i1 = row_in && x0 + 1 >= mp.xl && x0 + 1 < mp.xh;
        const bool i2 = row_in && x0 + 2 >= mp.xl && x0 + 2 < mp.xh, i3 = row_in && x0 + 3 >= mp.xl && x0 + 3 < mp.xh;
        oa = make_float4(i0 ? b.x : a.x, i1 ? b.y : a.y, i2 ? b.z : a.z, i3 ? b.w : a.w);
        ob = make_float4(i0 ? a.x : b.x, i1 ? a.y : b.y, i2 ? a.z : b.z, i3 ? a.w : b.w);
    } else {
        oa = make_float4(__fadd_rn(__fmul_rn(a.x, mp.lam), __fmul_rn(b.x, mp.oml)), __fadd_rn(__fmul_rn(a.y, mp.lam), __fmul_rn(b.y, mp.oml)),
                         __fadd_rn(__fmul_rn(a.z, mp.lam), __fmul_rn(b.z, mp.oml)), __fadd_rn(__fmul_rn(a.w, mp.lam), __fmul_rn(b.w, mp.oml)));
        ob = make_float4(__fadd_rn(__fmul_rn(b.x, mp.lam), __fmul_rn(a.x, mp.oml)), __fadd_rn(__fmul_rn(b.y, mp.lam), __fmul_rn(a.y, mp.oml)),
                         __fadd_rn(__fmul_rn(b.z, mp.lam), __fmul_rn(a.z, mp.oml)), __fadd_rn(__fmul_rn(b.w, mp.lam), __fmul_rn(a.w, mp.oml)));
    }
}
__global__ void __launch_bounds__(256) mixup_batch_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int HW, MixParams mp) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const int per_img = 3 * HW * HW / 4;
    const int pairs = (B + 1) / 2;
    const size_t total = size_t(pairs) * per_img;
    for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
        const int b = idx / per_img, r = idx % per_img;
        const int f = B - 1 - b;
        const int x0 = (r % (HW / 4)) * 4, yy = (r / (HW / 4)) % HW;
        const float4 a = reinterpret_cast<const float4*>(in)[size_t(b) * per_img + r];
        const float4 c = reinterpret_cast<const float4*>(in)[size_t(f) * per_img + r];
        float4 oa, ob;
        mix_pair4(a, c, x0, yy, mp, oa, ob);
        reinterpret_cast<float4*>(out)[size_t(b) * per_img + r] = oa;
        if (f != b) reinterpret_cast<float4*>(out)[size_t(f) * per_img + r] = ob;
    }
}
// the same fused into the im2col: the mixed batch is never materialised (the post-search phase has no other consumer of the
// images: PMIM is off, search.py:645)
__global__ void __launch_bounds__(256) patchify_mixup_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int HW, int P,
                                                             MixParams mp) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const int G = HW / P;
    const int per_img = 3 * HW * HW / 8;
    const int pairs = (B + 1) / 2;
    const size_t total = size_t(pairs) * per_img;
    for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
        const int b = idx / per_img;
        const int f = B - 1 - b;
        int r = idx % per_img;
        const int x8 = r % (HW / 8); r /= (HW / 8);
        const int yy = r % HW; const int c = r / HW;
        const size_t off = (size_t(c) * HW + yy) * HW + x8 * 8;
        const float4* sa = reinterpret_cast<const float4*>(img + size_t(b) * 3 * HW * HW + off);
        const float4* sb = reinterpret_cast<const float4*>(img + size_t(f) * 3 * HW * HW + off);
        const float4 a0 = __ldg(sa), a1 = __ldg(sa + 1), b0 = __ldg(sb), b1 = __ldg(sb + 1);
        float4 oa0, oa1, ob0, ob1;
        mix_pair4(a0, b0, x8 * 8, yy, mp, oa0, ob0);
        mix_pair4(a1, b1, x8 * 8 + 4, yy, mp, oa1, ob1);
        const int ph = yy / P, i = yy % P, pw = (x8 * 8) / P, j = (x8 * 8) % P;
        const size_t col = size_t(c) * P * P + i * P + j;
        const float fa[8] = {oa0.x, oa0.y, oa0.z, oa0.w, oa1.x, oa1.y, oa1.z, oa1.w};
        *reinterpret_cast<uint4*>(out + (size_t(b) * G * G + ph * G + pw) * (3 * P * P) + col) = pack8(fa);
        if (f != b) {
            const float fb[8] = {ob0.x, ob0.y, ob0.z, ob0.w, ob1.x, ob1.y, ob1.z, ob1.w};
            *reinterpret_cast<uint4*>(out + (size_t(f) * G * G + ph * G + pw) * (3 * P * P) + col) = pack8(fb);
        }
    }
}
// timm mixup_target: lam * smoothed one-hot(y) + (1 - lam) * smoothed one-hot(y.flip(0)), fp32 [B, C]
__global__ void __launch_bounds__(256) mixup_target_kernel(const long long* __restrict__ labels, float* __restrict__ target, int B, int C, float lam,
                                                           float oml, float on, float off) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const size_t total = size_t(B) * C;
    for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
        const int b = idx / C, c = idx % C;
        const float y1 = labels[b] == c ? on : off, y2 = labels[B - 1 - b] == c ? on : off;
        target[idx] = __fadd_rn(__fmul_rn(y1, lam), __fmul_rn(y2, oml));
    }
}

// =============================================================================================
// PMIM mask + DropPath scales from uniform randoms
//   mask[b,l] = 1 if noise[b,l] is NOT among the `keep` smallest of its row (vision_transformer.py:597-607)
//   drop_scale[i] = floor(keep_i + u_i) / keep_i   (timm DropPath)
// =============================================================================================
__global__ void pmim_mask_kernel(const float* __restrict__ noise, float* __restrict__ mask, int L, int keep) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    extern __shared__ float nz[];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < L; i += blockDim.x) nz[i] = noise[size_t(b) * L + i];
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const float v = nz[i];
        int rank = 0;
        for (int j = 0; j < L; ++j) rank += (nz[j] < v) || (nz[j] == v && j < i);
        mask[size_t(b) * L + i] = rank >= keep ? 1.f : 0.f;
    }
}
__global__ void droppath_scale_kernel(const float* __restrict__ u, const float* __restrict__ drop_prob, float* __restrict__ scale,
                                      int n_layers2, int B) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n_layers2 * B) {
        const float p = drop_prob[idx / B];
        const float keep = 1.f - p;
        scale[idx] = (p > 0.f) ? floorf(keep + u[idx]) / keep : 1.f;
    }
}

// =============================================================================================
// cls rows of the token matrix: x[b, 0, :] = (cls + pos[0]) * gate   (vision_transformer.py:646-651)
// =============================================================================================
__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, const float* __restrict__ gate,
                                __nv_bfloat16* __restrict__ x, int B, int T, int D) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < B * D) {
        const int b = idx / D, c = idx % D;
        x[size_t(b) * T * D + c] = __float2bfloat16((cls[c] + pos[c]) * gate[c]);
    }
}

// =============================================================================================
// embed backward (autograd of layers.py:179-191 + vision_transformer.py:628-651).  With
//   x0[b,t,:] = gate * s[b,t,:],  s = cls+pos0 (t=0) | mask_token (masked) | conv+bias+pos_t (kept):
//   dconv[b,l,:]   = g0[b,1+l,:] * gate * (1-mask)                          -> bf16 [B*L, D]
//   part_gx[t,:]   = sum_b g0*x0            (d gate = sum_t part_gx / gate)
//   part_pos[t,:]  = sum_b g0*gate*(t==0 ? 1 : 1-mask)   (d pos_embed; row 0 is also d cls_token)
//   part_mt[t,:]   = sum_b g0*gate*mask                    (d mask_token, rows t >= 1)
// grid = (T, ceil(D/128)); thread = one column, loops over the batch.
// =============================================================================================
// one block per token position t: thread (vc, gq) owns 8 columns and the images b = gq, gq + G, ...; 16-byte loads, several rows
// in flight per thread, then a shared-memory reduction over the G image groups
__global__ void __launch_bounds__(512) embed_bwd_kernel(const __nv_bfloat16* __restrict__ g0, const __nv_bfloat16* __restrict__ x0,
                                 const float* __restrict__ gate, const float* __restrict__ mask, __nv_bfloat16* __restrict__ dconv,
                                 float* __restrict__ part_gx, float* __restrict__ part_pos, float* __restrict__ part_mt, int B, int T,
                                 int D, int G, int VCc) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    extern __shared__ float eb_red[];                  // [G][Dc]
    const int t = blockIdx.x;
    // blockIdx.y: column slice of VCc 16-byte vectors (Dc = 8 VCc channels) - T CTAs alone are 1.33 per SM on 148 SMs
    const int Dc = VCc << 3, c0 = blockIdx.y * Dc;
    const int vl = threadIdx.x % VCc, vc = blockIdx.y * VCc + vl, gq = threadIdx.x / VCc;
    const int L = T - 1;
    float agx[8], apos[8], amt[8], gt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { agx[j] = apos[j] = amt[j] = 0.f; gt[j] = 0.f; }
    if (gq < G) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gt[j] = __ldg(gate + vc * 8 + j);
#pragma unroll 4
        for (int b = gq; b < B; b += G) {
            const size_t off = (size_t(b) * T + t) * D + vc * 8;
            float gr[8], xv[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(g0 + off)), gr);
            unpack8(__ldg(reinterpret_cast<const uint4*>(x0 + off)), xv);
            if (t == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { agx[j] = fmaf(gr[j], xv[j], agx[j]); apos[j] = fmaf(gr[j], gt[j], apos[j]); }
            } else {
                const float mk = __ldg(mask + size_t(b) * L + t - 1);
                float dc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    agx[j] = fmaf(gr[j], xv[j], agx[j]);
                    const float gg = gr[j] * gt[j];
                    dc[j] = gg * (1.f - mk);
                    apos[j] += dc[j];
                    amt[j] = fmaf(gg, mk, amt[j]);
                }
                *reinterpret_cast<uint4*>(dconv + (size_t(b) * L + t - 1) * D + vc * 8) = pack8(dc);
            }
        }
    }
    // three reductions over the image groups through the same buffer
    float* outs[3] = {part_gx, part_pos, part_mt};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* src = k == 0 ? agx : (k == 1 ? apos : amt);
        __syncthreads();
        if (gq < G) {
#pragma unroll
            for (int j = 0; j < 8; ++j) eb_red[gq * Dc + vl * 8 + j] = src[j];
        }
        __syncthreads();
        for (int col = threadIdx.x; col < Dc; col += blockDim.x) {
            float a = 0.f;
            for (int q2 = 0; q2 < G; ++q2) a += eb_red[q2 * Dc + col];
            outs[k][size_t(t) * D + c0 + col] = a;
        }
    }
}

// =============================================================================================
// norm_targets (vision_transformer.py:121-141), evaluated only at masked patches, written patch-major in the
// decoder's output order: tgt[(b*L + l), c*256 + i*16 + j].  Window = 47; a CTA takes NT_PPC consecutive patches and skips the
// unmasked ones (5-25 % are masked over the warm-up: one CTA per patch spent most of the launch on CTAs that exit at once).
// The 16 window sums a row (then a column) of the 62-wide halo needs share their 32 middle elements:
//     sum_j = [elements 15..30 + elements j..14] + [elements 31..46 + elements 47..j+46]
// Two threads walk one line, each over 31 elements (16 shared + a running sum over its 15 outer elements, walking outwards),
// on packed (x, x^2) pairs, and exchange their 8 + 8 partial sums by shuffle: 31 shared-memory reads and ~55 packed additions
// per thread instead of 16 x 47 taps per line, in a fixed order.
// =============================================================================================
static constexpr int NT_P = 16, NT_K = 47, NT_R = 23, NT_W = NT_P + 2 * NT_R;  // 62
static constexpr int NT_PPC = 4;              // patches per CTA
static constexpr int NT_WS = 68;              // halo row pitch in floats: 64 columns [px - 24, px + 40) as 16 float4 + 4 of padding
static_assert(NT_W == 2 * (NT_P - 1) + 32 && NT_K == NT_W - (NT_P - 1), "core / prefix / suffix split of the window sums");
// f(e): packed (x, x^2) of element e of the line, e = 0 .. 61.  half 0 walks elements 30 .. 0, half 1 elements 31 .. 61.
// Returns in t[n], n = 0 .. 7, the window sums of outputs j = n (half 0) or j = 15 - n (half 1); the two threads of a line are
// neighbouring lanes.
template <typename F>
__device__ __forceinline__ void window_sums_half(F f, int half, float2 (&t)[8]) {
    const int e0 = half ? 31 : 30, st = half ? 1 : -1;
    float2 core = splat2(0.f);
#pragma unroll
    for (int m = 0; m < 16; ++m) core = add2(core, f(e0 + st * m));
    float2 p[16];                             // p[n] = core + the n outer elements next to the core
    p[0] = core;
#pragma unroll
    for (int n = 0; n < 15; ++n) p[n + 1] = add2(p[n], f(e0 + st * (16 + n)));
    // output j = (elements 15..30 + j..14) + (elements 31..46 + 47..j+46) = p0[15 - j] + p1[j]; with mine / other for the two
    // halves, t[n] = mine[15 - n] + other[n] is output n for half 0 and output 15 - n for half 1
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        float2 o;
        o.x = __shfl_xor_sync(0xffffffffu, p[n].x, 1);
        o.y = __shfl_xor_sync(0xffffffffu, p[n].y, 1);
        t[n] = add2(p[15 - n], o);
    }
}
static constexpr int NT_THREADS = 128;        // the line sums keep 124 threads busy: 4-warp CTAs, four per SM (128 registers: spills cost 15 %)
__global__ void __launch_bounds__(NT_THREADS, 4) norm_targets_kernel(const float* __restrict__ img, const float* __restrict__ mask,
                                                                     float* __restrict__ tgt, int HW, int n_patches) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    __shared__ __align__(16) float win[NT_W][NT_WS];
    __shared__ float2 h[NT_W][NT_P + 1];      // horizontal window sums (x, x^2)
    __shared__ float2 v[NT_P][NT_P + 1];      // full window sums
    const int G = HW / NT_P, L = G * G;
    const int tid = threadIdx.x;
    const int oj = tid % NT_P, oi0 = tid / NT_P;                // this thread's two output pixels of every patch: rows oi0, oi0 + 8
    constexpr int NLD = (NT_W * 16 + NT_THREADS - 1) / NT_THREADS;     // float4 loads per thread and channel (8)
    float mk[NT_PPC];                                           // the CTA's mask values in one round trip
#pragma unroll
    for (int pp = 0; pp < NT_PPC; ++pp) {
        const int patch = blockIdx.x * NT_PPC + pp;
        mk[pp] = patch < n_patches ? __ldg(mask + patch) : 0.f;
    }
#pragma unroll 1
    for (int pp = 0; pp < NT_PPC; ++pp) {
        const int patch = blockIdx.x * NT_PPC + pp;             // b*L + l
        if (mk[pp] == 0.f) continue;
        const int b = patch / L, l = patch % L;
        const int py = (l / G) * NT_P, px = (l % G) * NT_P;
        // pixels of the (border-clipped) window around this thread's output pixels: the same for the three channels
        float inv_cnt[2], bessel[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int gy = py + oi0 + 8 * k, gx = px + oj;
            const int cy = min(gy + NT_R, HW - 1) - max(gy - NT_R, 0) + 1;
            const int cx = min(gx + NT_R, HW - 1) - max(gx - NT_R, 0) + 1;
            const float cnt = float(cy * cx);
            inv_cnt[k] = 1.f / cnt;
            bessel[k] = cnt / (cnt - 1.f);
        }
        // halo rows as 16 aligned float4 (image rows and px - 24 are multiples of 4 floats: a float4 lies inside or outside);
        // a thread's loads are issued back to back, those of the next channel before the sums of the current one
        float4 r[NLD];
        auto load = [&](int c) {
            const float* plane = img + (size_t(b) * 3 + c) * HW * HW;
#pragma unroll
            for (int i = 0; i < NLD; ++i) {
                const int sl = tid + NT_THREADS * i, row = sl >> 4, vec = sl & 15;
                const int gy = py - NT_R + row, gx = px - NT_R - 1 + 4 * vec;
                const bool ok = sl < NT_W * 16 && gy >= 0 && gy < HW && gx >= 0 && gx < HW;
                r[i] = ok ? __ldg(reinterpret_cast<const float4*>(plane + size_t(gy) * HW + gx)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        load(0);
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int i = 0; i < NLD; ++i) {
                const int sl = tid + NT_THREADS * i;
                if (sl < NT_W * 16) *reinterpret_cast<float4*>(&win[sl >> 4][4 * (sl & 15)]) = r[i];
            }
            __syncthreads();
            if (c + 1 < 3) load(c + 1);
            // horizontal: threads (row, half); column e of the line is win[row][e + 1]; the shuffles name every lane, so
            // threads 124 .. 127 repeat line 61 without storing
            {
                const int row = min(tid >> 1, NT_W - 1), half = tid & 1;
                float2 t[8];
                window_sums_half([&](int e) { const float x = win[row][e + 1]; return make_float2(x, x * x); }, half, t);
                if (tid < 2 * NT_W) {
#pragma unroll
                    for (int n = 0; n < 8; ++n) h[row][half ? 15 - n : n] = t[n];
                }
            }
            __syncthreads();
            // vertical: threads (column, half) - one warp
            if (tid < 2 * NT_P) {
                const int j = tid >> 1, half = tid & 1;
                float2 t[8];
                window_sums_half([&](int e) { return h[e][j]; }, half, t);
#pragma unroll
                for (int n = 0; n < 8; ++n) v[half ? 15 - n : n][j] = t[n];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int oi = oi0 + 8 * k;
                const float2 s12 = v[oi][oj];
                const float mu = s12.x * inv_cnt[k];
                const float var = fmaxf((s12.y * inv_cnt[k] - mu * mu) * bessel[k], 0.f);
                const float xv = win[NT_R + oi][NT_R + 1 + oj];
                tgt[size_t(patch) * (3 * NT_P * NT_P) + c * NT_P * NT_P + oi * NT_P + oj] = (xv - mu) * rsqrtf(var + 1e-6f);
            }
            __syncthreads();          // win / h / v are rewritten by the next channel / patch
        }
    }
}

// =============================================================================================
// label-smoothing cross entropy fwd + bwd (timm LabelSmoothingCrossEntropy, search.py:584)
//   loss_rows[b] = (1-s)*nll + s*(-mean logp);   dlogits = (softmax - (1-s)*onehot - s/C) * gscale / B   (bf16)
// one CTA (128 threads) per row.
// =============================================================================================
__global__ void __launch_bounds__(128) ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                 float* __restrict__ loss_rows, __nv_bfloat16* __restrict__ dlogits, int C,
                                                 float smoothing, float gscale_over_B) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const int b = blockIdx.x;
    const float* lr = logits + size_t(b) * C;
    __shared__ float red[4];
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, lr[c]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float se = 0.f, sl = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { se += expf(lr[c] - mx); sl += lr[c]; }
    se = warp_sum(se);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
    __syncthreads();
    se = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    sl = warp_sum(sl);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sl;
    __syncthreads();
    sl = red[0] + red[1] + red[2] + red[3];
    const float lse = mx + logf(se);
    const int y = int(labels[b]);
    if (threadIdx.x == 0) {
        const float nll = lse - lr[y];
        const float smooth = lse - sl / C;
        loss_rows[b] = (1.f - smoothing) * nll + smoothing * smooth;
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float p = expf(lr[c] - lse);
        const float t = (c == y ? (1.f - smoothing) : 0.f) + smoothing / C;
        dlogits[size_t(b) * C + c] = __float2bfloat16((p - t) * gscale_over_B);
    }
}

// =============================================================================================
// soft-target cross entropy fwd + bwd (timm SoftTargetCrossEntropy after Mixup: finetune.py:388-389, search.py:655)
//   loss_rows[b] = sum_c -t[b,c] * logp[b,c];   dlogits = (softmax * sum_c t - t) * gscale / B   (bf16)
// =============================================================================================
__global__ void __launch_bounds__(128) soft_ce_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                      float* __restrict__ loss_rows, __nv_bfloat16* __restrict__ dlogits, int C,
                                                      float gscale_over_B) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const int b = blockIdx.x;
    const float* lr = logits + size_t(b) * C;
    const float* tr = target + size_t(b) * C;
    __shared__ float red[4];
    auto block_sum = [&](float v) {
        v = warp_sum(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        return red[0] + red[1] + red[2] + red[3];
    };
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, lr[c]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    float se = 0.f, st = 0.f, stl = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { se += expf(lr[c] - mx); st += tr[c]; stl += tr[c] * lr[c]; }
    se = block_sum(se);
    st = block_sum(st);
    stl = block_sum(stl);
    const float lse = mx + logf(se);
    if (threadIdx.x == 0) loss_rows[b] = st * lse - stl;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        dlogits[size_t(b) * C + c] = __float2bfloat16((expf(lr[c] - lse) * st - tr[c]) * gscale_over_B);
}

// =============================================================================================
// evaluation metrics (engine.evaluate, engine.py:222-257): plain cross entropy and top-1 / top-5 hits per row.
//   out_rows[b] = {nll, hit@1, hit@5}; the label is in the top k iff fewer than k logits are larger (ties: lower index wins,
//   torch.topk order)
// =============================================================================================
__global__ void __launch_bounds__(128) eval_metrics_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                           float* __restrict__ out_rows, int C) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    const int b = blockIdx.x;
    const float* lr = logits + size_t(b) * C;
    __shared__ float red[4];
    const int y = int(labels[b]);
    const float ly = lr[y];
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, lr[c]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float se = 0.f, ahead = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        se += expf(lr[c] - mx);
        ahead += (lr[c] > ly || (lr[c] == ly && c < y)) ? 1.f : 0.f;
    }
    se = warp_sum(se);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
    __syncthreads();
    se = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    ahead = warp_sum(ahead);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ahead;
    __syncthreads();
    ahead = red[0] + red[1] + red[2] + red[3];
    if (threadIdx.x == 0) {
        out_rows[3 * b + 0] = mx + logf(se) - ly;
        out_rows[3 * b + 1] = ahead < 1.f ? 1.f : 0.f;
        out_rows[3 * b + 2] = ahead < 5.f ? 1.f : 0.f;
    }
}

// =============================================================================================
// loss finalisation (engine.py:134-144): one CTA.
//   scal[0]=base CE  [1]=arch  [2]=decoder  [3]=total  [4]=w_dec=(base/dec)  [5]=decoder grad scale  [6]=#masked patches
// =============================================================================================
__global__ void __launch_bounds__(1024) loss_finalize_kernel(const float* __restrict__ loss_rows, int B, const float* __restrict__ dec_part,
                                                            int n_dec_part, const float* __restrict__ mask, int n_mask,
                                                            const float* __restrict__ arch_loss, float grad_scale, float* __restrict__ scal) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    __shared__ float red[3][32];          // one CTA of 32 warps: the kernel sits on the critical path between forward and backward
    float a = 0.f, d = 0.f, m = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) a += loss_rows[i];
    // the two long vectors (decoder partials: ~10 k, PMIM mask: B * L = 50 k floats) as 16-byte loads, four in flight per thread:
    // one CTA walking them one float at a time spent 49 dependent L2 round trips here (31 us on the critical path)
    auto vsum = [&](const float* __restrict__ p, int n) {
        float acc = 0.f;
        if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
            const float4* p4 = reinterpret_cast<const float4*>(p);
            const int n4 = n >> 2;
            int i = threadIdx.x;
            for (; i + 3 * int(blockDim.x) < n4; i += 4 * blockDim.x) {
                const float4 v0 = __ldg(p4 + i), v1 = __ldg(p4 + i + blockDim.x), v2 = __ldg(p4 + i + 2 * blockDim.x),
                             v3 = __ldg(p4 + i + 3 * blockDim.x);
                acc += ((v0.x + v0.y) + (v0.z + v0.w)) + ((v1.x + v1.y) + (v1.z + v1.w)) + ((v2.x + v2.y) + (v2.z + v2.w)) +
                       ((v3.x + v3.y) + (v3.z + v3.w));
            }
            for (; i < n4; i += blockDim.x) { const float4 v0 = __ldg(p4 + i); acc += (v0.x + v0.y) + (v0.z + v0.w); }
            for (int j = (n4 << 2) + threadIdx.x; j < n; j += blockDim.x) acc += p[j];
        } else {
            for (int j = threadIdx.x; j < n; j += blockDim.x) acc += p[j];
        }
        return acc;
    };
    d = vsum(dec_part, n_dec_part);
    m = vsum(mask, n_mask);
    a = warp_sum(a); d = warp_sum(d); m = warp_sum(m);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = d; red[2][threadIdx.x >> 5] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = d = m = 0.f;
        for (int i = 0; i < int(blockDim.x >> 5); ++i) { a += red[0][i]; d += red[1][i]; m += red[2][i]; }
        const float base = a / B;
        const float denom = (m * 256.f + 1e-5f) * 3.f;
        const float dec = d / denom;
        const float arch = arch_loss != nullptr ? arch_loss[0] : 0.f;
        float w_dec = 0.f, total = base + arch;
        if (m > 0.f && dec != 0.f) { w_dec = base / dec; total += w_dec * dec; }
        scal[0] = base; scal[1] = arch; scal[2] = dec; scal[3] = total; scal[4] = w_dec;
        scal[5] = w_dec / denom * grad_scale;
        scal[6] = m;
    }
}

// =============================================================================================
// fused multi-segment AdamW (optim.py:56-120: decay first, then Adam), bf16 shadow weights, optional grad zeroing.
// hyper[seg] = {lr, weight_decay, beta1, beta2, eps, bias_corr1, bias_corr2, unused}; seg_end[] are exclusive
// prefix ends (elements, multiples of 4) of the contiguous optimizer groups of the flat parameter arena.
// =============================================================================================
static constexpr int ADAM_MAX_SEGS = 64;      // search: 5 groups; finetune: 2 x (depth + 2) layer-decay groups (lr_decay.py)
struct AdamSegs {
    int nseg;
    long long end[ADAM_MAX_SEGS];
};
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, __nv_bfloat16* __restrict__ shadow,
                                                    const float* __restrict__ hyper, AdamSegs segs, long long n4, int zero_grad) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    int s = 0;        // segment of the current element: monotone along the grid-stride loop, so the scan is amortised
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 4;
        while (s < segs.nseg - 1 && e >= segs.end[s]) ++s;
        const float* h = hyper + s * 8;
        const float lr = h[0], wd = h[1], b1 = h[2], b2 = h[3], eps = h[4], bc1 = h[5], bc2 = h[6];
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        const float decay = 1.f - lr * wd;
        const float step = lr / bc1;
        const float rbc2 = 1.f / sqrtf(bc2);
#define OFB_ADAM1(P, G, Mm, V)                         \
        P *= decay;                                    \
        Mm = Mm * b1 + (1.f - b1) * G;                 \
        V = V * b2 + (1.f - b2) * G * G;               \
        P -= step * (Mm / (sqrtf(V) * rbc2 + eps));
        OFB_ADAM1(pp.x, gg.x, mm.x, vv.x)
        OFB_ADAM1(pp.y, gg.y, mm.y, vv.y)
        OFB_ADAM1(pp.z, gg.z, mm.z, vv.z)
        OFB_ADAM1(pp.w, gg.w, mm.w, vv.w)
#undef OFB_ADAM1
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (shadow != nullptr) {
            uint2 o;
            o.x = pack_bf16x2(pp.x, pp.y);
            o.y = pack_bf16x2(pp.z, pp.w);
            reinterpret_cast<uint2*>(shadow)[i] = o;
        }
    }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n4) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(src)[i];
        uint2 o;
        o.x = pack_bf16x2(a.x, a.y);
        o.y = pack_bf16x2(a.z, a.w);
        reinterpret_cast<uint2*>(dst)[i] = o;
    }
}

// =============================================================================================
// out[col] += scale * sum_rows x[row, col]   (x bf16 [R, N], N even).  grid (ceil(N/64), row splits), block (32, 8)
// Deterministic: the row splits of a column block park their partial sums in a device-side scratch and the split that
// arrives last (ticket per column block) adds them in ascending split order - no floating-point atomics. The scratch is
// shared by all launches: they are stream-ordered in every caller (one stream per process issues the colsum launches).
// =============================================================================================
static constexpr int COLSUM_MAX_SPLITS = 128, COLSUM_MAX_N = 4096;
__device__ float g_colsum_part[COLSUM_MAX_SPLITS * COLSUM_MAX_N];
__device__ unsigned int g_colsum_ticket[COLSUM_MAX_N / 64];
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int ld, int R, int N, float* __restrict__ out, float scale,
                                   const float* __restrict__ scale_dev) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    __shared__ float2 sm[8][32];
    __shared__ int s_last;
    const int col = (blockIdx.x * 32 + threadIdx.x) * 2;
    float2 acc = make_float2(0.f, 0.f);
    if (col < N) {
        for (int r = blockIdx.y * 8 + threadIdx.y; r < R; r += gridDim.y * 8) {
            const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + size_t(r) * ld + col));
            acc.x += v.x; acc.y += v.y;
        }
    }
    sm[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && col < N) {
        for (int i = 1; i < 8; ++i) { acc.x += sm[i][threadIdx.x].x; acc.y += sm[i][threadIdx.x].y; }
        *reinterpret_cast<float2*>(g_colsum_part + size_t(blockIdx.y) * N + col) = acc;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) s_last = atomicInc(&g_colsum_ticket[blockIdx.x], gridDim.y - 1) == gridDim.y - 1;
    __syncthreads();
    if (s_last && threadIdx.y == 0 && col < N) {
        __threadfence();
        float2 t = make_float2(0.f, 0.f);
        for (unsigned sp = 0; sp < gridDim.y; ++sp) {
            const float* pp = g_colsum_part + size_t(sp) * N + col;      // other CTAs wrote these: bypass L1
            t.x += __ldcg(pp); t.y += __ldcg(pp + 1);
        }
        const float s = scale * (scale_dev != nullptr ? *scale_dev : 1.f);
        out[col] += t.x * s;
        if (col + 1 < N) out[col + 1] += t.y * s;
    }
}

// =============================================================================================
// launchers
// =============================================================================================
// The same sum with 16-byte loads (8 columns per thread, a warp covers 512 bytes of a row, four rows in flight per thread): the
// 4-byte version keeps one small load per thread in flight and reaches 2 TB/s on the [50 432 x 768] sign matrix of the PMIM
// decoder; this one is bound by HBM. Needs N % 8 == 0, ld % 8 == 0 and a 16-byte aligned base.
__global__ void __launch_bounds__(256) colsum_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, int ld, int R, int N,
                                                            float* __restrict__ out, float scale, const float* __restrict__ scale_dev) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    __shared__ float sm[8][32][9];
    __shared__ int s_last;
    const int col = (blockIdx.x * 32 + threadIdx.x) * 8;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    auto add = [&](const uint4& v) {
        float2 f;
        f = unpack_bf16x2(v.x); acc[0] += f.x; acc[1] += f.y;
        f = unpack_bf16x2(v.y); acc[2] += f.x; acc[3] += f.y;
        f = unpack_bf16x2(v.z); acc[4] += f.x; acc[5] += f.y;
        f = unpack_bf16x2(v.w); acc[6] += f.x; acc[7] += f.y;
    };
    if (col < N) {
        const int stride = gridDim.y * 8;
        int r = blockIdx.y * 8 + threadIdx.y;
        const __nv_bfloat16* px = x + col;
        for (; r + 3 * stride < R; r += 4 * stride) {        // fixed order of the additions: r, r + stride, ...
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(px + size_t(r) * ld));
            const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(px + size_t(r + stride) * ld));
            const uint4 v2 = __ldg(reinterpret_cast<const uint4*>(px + size_t(r + 2 * stride) * ld));
            const uint4 v3 = __ldg(reinterpret_cast<const uint4*>(px + size_t(r + 3 * stride) * ld));
            add(v0); add(v1); add(v2); add(v3);
        }
        for (; r < R; r += stride) add(__ldg(reinterpret_cast<const uint4*>(px + size_t(r) * ld)));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[threadIdx.y][threadIdx.x][i] = acc[i];
    __syncthreads();
    if (threadIdx.y == 0 && col < N) {
        for (int w = 1; w < 8; ++w)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += sm[w][threadIdx.x][i];
        float* pp = g_colsum_part + size_t(blockIdx.y) * N + col;
        *reinterpret_cast<float4*>(pp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(pp + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) s_last = atomicInc(&g_colsum_ticket[blockIdx.x], gridDim.y - 1) == gridDim.y - 1;
    __syncthreads();
    if (s_last && threadIdx.y == 0 && col < N) {
        __threadfence();
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = 0.f;
        for (unsigned sp = 0; sp < gridDim.y; ++sp) {           // ascending split order; other CTAs wrote these: bypass L1
            const float4* pp = reinterpret_cast<const float4*>(g_colsum_part + size_t(sp) * N + col);
            const float4 a = __ldcg(pp), b = __ldcg(pp + 1);
            t[0] += a.x; t[1] += a.y; t[2] += a.z; t[3] += a.w; t[4] += b.x; t[5] += b.y; t[6] += b.z; t[7] += b.w;
        }
        const float s = scale * (scale_dev != nullptr ? *scale_dev : 1.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[col + i] += t[i] * s;
    }
}

int launch_colsum_bf16(const void* x, int ld, int R, int N, float* out, float scale, const float* scale_dev, cudaStream_t s) {
    if (N % 2 != 0 || ld % 2 != 0 || N > COLSUM_MAX_N) return 1014;
    int splits = (R + 255) / 256;
    if (splits > COLSUM_MAX_SPLITS) splits = COLSUM_MAX_SPLITS;
    if (splits < 1) splits = 1;
    if (N % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
        dim3 grid((N + 255) / 256, splits), block(32, 8);
        OFB_LAUNCH(colsum_bf16x8_kernel, grid, block, 0, s, reinterpret_cast<const __nv_bfloat16*>(x), ld, R, N, out, scale, scale_dev);
        return int(cudaGetLastError());
    }
    if (splits > 64) splits = 64;
    dim3 grid((N + 63) / 64, splits), block(32, 8);
    OFB_LAUNCH(colsum_bf16_kernel, grid, block, 0, s, reinterpret_cast<const __nv_bfloat16*>(x), ld, R, N, out, scale, scale_dev);
    return int(cudaGetLastError());
}


template <int MAXC>
static int ln_fwd_inst(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, int D, float eps,
                       int Dv, cudaStream_t s) {
    const int wpb = 8;
    int grid = (M + wpb - 1) / wpb;
    const int cap = num_sms() * 8;
    if (grid > cap) grid = cap;
    OFB_LAUNCH(ln_fwd_kernel<MAXC>, grid, wpb * 32, 0, s, reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta,
                                                  reinterpret_cast<__nv_bfloat16*>(y), mean, rstd, M, D, eps, Dv);
    return err();
}
template <int W>
static int ln_fwd3_inst(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, float eps,
                        cudaStream_t s) {
    const int wpb = 8;
    int grid = (M + wpb - 1) / wpb;
    const int cap = num_sms() * 8;
    if (grid > cap) grid = cap;
    if (W == 1)
        OFB_LAUNCH(ln_fwd3x2_kernel<W>, grid, wpb * 32, 0, s, reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta,
                                                      reinterpret_cast<__nv_bfloat16*>(y), mean, rstd, M, eps);
    else
        OFB_LAUNCH(ln_fwd3_kernel<W>, grid, wpb * 32, 0, s, reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta,
                                                    reinterpret_cast<__nv_bfloat16*>(y), mean, rstd, M, eps);
    return err();
}
static bool ln_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool ln_word_path() {        // OFB_LN_WORDS=0 falls back to the 16-byte-chunk kernels (A/B measurements)
    static int on = -1;
    if (on < 0) { const char* e = getenv("OFB_LN_WORDS"); on = (e == nullptr || e[0] != '0') ? 1 : 0; }
    return on == 1;
}
template <int NW>
static int ln_fwdw_inst(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, int D, int Dv,
                        float eps, cudaStream_t s) {
    const int wpb = 8;
    int grid = (M + wpb - 1) / wpb;
    const int cap = num_sms() * 8;
    if (grid > cap) grid = cap;
    OFB_LAUNCH(ln_fwdw_kernel<NW>, grid, wpb * 32, 0, s, reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta,
                                                 reinterpret_cast<__nv_bfloat16*>(y), mean, rstd, M, D, Dv, eps);
    return err();
}
int launch_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, int D, float eps,
                  int Dv, cudaStream_t s) {
    if (D % 8 != 0 || D > 1024) return 1010;
    if (Dv <= 0) Dv = D;
    if (Dv > D || Dv <= D - 8) return 1010;      // padding is at most the ragged tail of the last 8-channel chunk
    if (Dv == D && ln_aligned16(x) && ln_aligned16(y) && ln_aligned16(gamma) && ln_aligned16(beta)) {
        if (D == 192) return ln_fwd3_inst<1>(x, gamma, beta, y, mean, rstd, M, eps, s);
        if (D == 384) return ln_fwd3_inst<2>(x, gamma, beta, y, mean, rstd, M, eps, s);
        if (D == 768) return ln_fwd3_inst<4>(x, gamma, beta, y, mean, rstd, M, eps, s);
    }
    if (Dv % 2 == 0 && D <= 768 && ln_word_path()) {
        switch ((D / 2 + 31) / 32) {
            case 1: return ln_fwdw_inst<1>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            case 2: return ln_fwdw_inst<2>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            case 3: return ln_fwdw_inst<3>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            case 4: return ln_fwdw_inst<4>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            case 5: return ln_fwdw_inst<5>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            case 6: return ln_fwdw_inst<6>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            case 7: return ln_fwdw_inst<7>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            case 8: return ln_fwdw_inst<8>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            case 9: return ln_fwdw_inst<9>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            case 10: return ln_fwdw_inst<10>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            case 11: return ln_fwdw_inst<11>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
            default: return ln_fwdw_inst<12>(x, gamma, beta, y, mean, rstd, M, D, Dv, eps, s);
        }
    }
    switch ((D + 255) / 256) {
        case 1: return ln_fwd_inst<1>(x, gamma, beta, y, mean, rstd, M, D, eps, Dv, s);
        case 2: return ln_fwd_inst<2>(x, gamma, beta, y, mean, rstd, M, D, eps, Dv, s);
        case 3: return ln_fwd_inst<3>(x, gamma, beta, y, mean, rstd, M, D, eps, Dv, s);
        default: return ln_fwd_inst<4>(x, gamma, beta, y, mean, rstd, M, D, eps, Dv, s);
    }
}

// number of CTAs (= rows of the partial buffers) of the backward kernel
int ln_bwd_grid(int M) {
    int grid = (M + 7) / 8;
    const int cap = num_sms() * 2;      // two resident CTAs per SM at D <= 512 (128 registers, 48 KB ring each): one wave
    return grid > cap ? cap : grid;
}

template <int MAXC>
static int ln_bwd_inst(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                       float* part_dgamma, float* part_dbeta, float* part_dbias, const float* rowscale, int rows_per_scale, int M, int D,
                       int Dv, const void* dres, cudaStream_t s) {
    const int wpb = 8;
    const int grid = ln_bwd_grid(M);
    const size_t ring = size_t(wpb) * LN_STAGES * 2 * D * 2, red = size_t(3) * wpb * D * sizeof(float);
    const size_t smem = (ring > red ? ring : red) + wpb * LN_STAGES * 8;
    static bool configured = false;
    if (!configured) {
        const int dmax = 256 * MAXC;
        const size_t ring_max = size_t(wpb) * LN_STAGES * 2 * dmax * 2, red_max = size_t(3) * wpb * dmax * sizeof(float);
        cudaFuncSetAttribute(ln_bwd_kernel<MAXC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int((ring_max > red_max ? ring_max : red_max) + wpb * LN_STAGES * 8));
        configured = true;
    }
    OFB_LAUNCH(ln_bwd_kernel<MAXC>, grid, wpb * 32, smem, s, reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x),
                                                     mean, rstd, gamma, reinterpret_cast<__nv_bfloat16*>(dx), part_dgamma, part_dbeta,
                                                     part_dbias, rowscale, rows_per_scale > 0 ? rows_per_scale : 1, M, D, Dv,
                                                     reinterpret_cast<const __nv_bfloat16*>(dres));
    return err();
}
template <int W>
static int ln_bwd3_inst(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                        float* part_dgamma, float* part_dbeta, float* part_dbias, const float* rowscale, int rows_per_scale, int M,
                        cudaStream_t s) {
    constexpr int D = 192 * W, wpb = 8;
    const int grid = ln_bwd_grid(M);
    constexpr size_t ring = size_t(wpb) * LN_STAGES * 2 * D * 2, red = size_t(3) * wpb * D * sizeof(float);
    constexpr size_t smem = (ring > red ? ring : red) + wpb * LN_STAGES * 8;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(ln_bwd3_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        configured = true;
    }
    OFB_LAUNCH(ln_bwd3_kernel<W>, grid, wpb * 32, smem, s, reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x),
                                                   mean, rstd, gamma, reinterpret_cast<__nv_bfloat16*>(dx), part_dgamma, part_dbeta,
                                                   part_dbias, rowscale, rows_per_scale > 0 ? rows_per_scale : 1, M);
    return err();
}
template <int NW>
static int ln_bwdw_inst(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                        float* part_dgamma, float* part_dbeta, float* part_dbias, const float* rowscale, int rows_per_scale, int M, int D,
                        int Dv, const void* dres, cudaStream_t s) {
    const int wpb = 8;
    const int grid = ln_bwd_grid(M);
    const size_t ring = size_t(wpb) * LN_STAGES * 2 * D * 2, red = size_t(3) * wpb * D * sizeof(float);
    const size_t smem = (ring > red ? ring : red) + wpb * LN_STAGES * 8;
    static bool configured = false;
    if (!configured) {
        const int dmax = 64 * NW;
        const size_t ring_max = size_t(wpb) * LN_STAGES * 2 * dmax * 2, red_max = size_t(3) * wpb * dmax * sizeof(float);
        cudaFuncSetAttribute(ln_bwdw_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int((ring_max > red_max ? ring_max : red_max) + wpb * LN_STAGES * 8));
        configured = true;
    }
    OFB_LAUNCH(ln_bwdw_kernel<NW>, grid, wpb * 32, smem, s, reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x),
                                                    mean, rstd, gamma, reinterpret_cast<__nv_bfloat16*>(dx), part_dgamma, part_dbeta,
                                                    part_dbias, rowscale, rows_per_scale > 0 ? rows_per_scale : 1, M, D, Dv,
                                                    reinterpret_cast<const __nv_bfloat16*>(dres));
    return err();
}
int launch_ln_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx, float* part_dgamma,
                  float* part_dbeta, float* part_dbias, const float* rowscale, int rows_per_scale, int M, int D, int Dv, const void* dres,
                  cudaStream_t s) {
    if (D % 8 != 0 || D > 1024) return 1010;
    if (Dv <= 0) Dv = D;
    if (Dv > D || Dv <= D - 8) return 1010;
    if (dres != nullptr && !ln_aligned16(dres)) return 1010;
    if (Dv == D && dres == nullptr && ln_aligned16(dy) && ln_aligned16(x) && ln_aligned16(dx) && ln_aligned16(gamma)) {
        if (D == 192) return ln_bwd3_inst<1>(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale, rows_per_scale, M, s);
        if (D == 384) return ln_bwd3_inst<2>(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale, rows_per_scale, M, s);
        if (D == 768) return ln_bwd3_inst<4>(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale, rows_per_scale, M, s);
    }
#define OFB_LN_BWDW(n) case n: return ln_bwdw_inst<n>(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale, \
                                                      rows_per_scale, M, D, Dv, dres, s)
    if (Dv % 2 == 0 && D <= 768 && ln_word_path()) {
        switch ((D / 2 + 31) / 32) {
            OFB_LN_BWDW(1); OFB_LN_BWDW(2); OFB_LN_BWDW(3); OFB_LN_BWDW(4); OFB_LN_BWDW(5); OFB_LN_BWDW(6);
            OFB_LN_BWDW(7); OFB_LN_BWDW(8); OFB_LN_BWDW(9); OFB_LN_BWDW(10); OFB_LN_BWDW(11);
            default: return ln_bwdw_inst<12>(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale,
                                             rows_per_scale, M, D, Dv, dres, s);
        }
    }
#undef OFB_LN_BWDW
    switch ((D + 255) / 256) {
        case 1: return ln_bwd_inst<1>(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale, rows_per_scale, M, D, Dv, dres, s);
        case 2: return ln_bwd_inst<2>(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale, rows_per_scale, M, D, Dv, dres, s);
        case 3: return ln_bwd_inst<3>(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale, rows_per_scale, M, D, Dv, dres, s);
        default: return ln_bwd_inst<4>(dy, x, mean, rstd, gamma, dx, part_dgamma, part_dbeta, part_dbias, rowscale, rows_per_scale, M, D, Dv, dres, s);
    }
}

int launch_reduce_partials(const float* part, int R, int N, float* out, float scale, const float* inv_colscale, int accumulate,
                           cudaStream_t s) {
    OFB_LAUNCH(reduce_partials_kernel, (N + 31) / 32, 32 * 32, 0, s, part, R, N, out, scale, inv_colscale, accumulate);
    return err();
}

int launch_reduce_partials_multi(const void* jobs_host, int njobs, cudaStream_t s) {
    if (njobs < 1 || njobs > 12) return 1015;
    ReduceJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    const ReduceJob* src = reinterpret_cast<const ReduceJob*>(jobs_host);
    int maxn = 0;
    for (int i = 0; i < njobs; ++i) { jobs.j[i] = src[i]; if (src[i].N > maxn) maxn = src[i].N; }
    dim3 grid((maxn + 31) / 32, njobs);
    OFB_LAUNCH(reduce_partials_multi_kernel, grid, 32 * 32, 0, s, jobs);
    return err();
}

int launch_splitk_reduce(const void* jobs_host, int njobs, cudaStream_t s) {
    if (njobs < 1 || njobs > 8) return 1015;
    SplitkJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    const SplitkJob* src = reinterpret_cast<const SplitkJob*>(jobs_host);
    long long maxn = 0;
    for (int i = 0; i < njobs; ++i) {
        if (src[i].ws == nullptr || src[i].out == nullptr || src[i].splits < 1 || src[i].n4 < 1) return 1015;
        jobs.j[i] = src[i];
        if (src[i].n4 > maxn) maxn = src[i].n4;
    }
    long long gx = (maxn + 255) / 256;
    const long long cap = (long long)num_sms() * 8 / njobs + 1;
    if (gx > cap) gx = cap;
    OFB_LAUNCH(splitk_reduce_kernel, dim3(unsigned(gx), unsigned(njobs)), 256, 0, s, jobs);
    return err();
}

int launch_patchify(const float* img, void* out, int B, int HW, int P, cudaStream_t s) {
    if (HW % 8 != 0 || P % 8 != 0) return 1011;
    const size_t total = size_t(B) * 3 * HW * HW / 8;
    int grid = int((total + 255) / 256);
    const int cap = num_sms() * 16;
    if (grid > cap) grid = cap;
    OFB_LAUNCH(patchify_kernel, grid, 256, 0, s, img, reinterpret_cast<__nv_bfloat16*>(out), B, HW, P);
    return err();
}

static bool mix_params(int HW, double lam, int cutmix, int yl, int yh, int xl, int xh, MixParams& mp) {
    if (cutmix && (yl < 0 || yh > HW || xl < 0 || xh > HW || yl > yh || xl > xh)) return false;
    mp.lam = float(lam); mp.oml = float(1.0 - lam);
    mp.cutmix = cutmix; mp.yl = yl; mp.yh = yh; mp.xl = xl; mp.xh = xh;
    return true;
}
int launch_mixup_batch(const float* in, float* out, int B, int HW, double lam, int cutmix, int yl, int yh, int xl, int xh, cudaStream_t s) {
    MixParams mp;
    if (HW % 4 != 0 || B < 1 || !mix_params(HW, lam, cutmix, yl, yh, xl, xh, mp)) return 1011;
    const size_t total = size_t((B + 1) / 2) * 3 * HW * HW / 4;
    int grid = int((total + 255) / 256);
    const int cap = num_sms() * 16;
    if (grid > cap) grid = cap;
    OFB_LAUNCH(mixup_batch_kernel, grid, 256, 0, s, in, out, B, HW, mp);
    return err();
}
int launch_patchify_mixup(const float* img, void* out, int B, int HW, int P, double lam, int cutmix, int yl, int yh, int xl, int xh,
                          cudaStream_t s) {
    MixParams mp;
    if (HW % 8 != 0 || P % 8 != 0 || B < 1 || !mix_params(HW, lam, cutmix, yl, yh, xl, xh, mp)) return 1011;
    const size_t total = size_t((B + 1) / 2) * 3 * HW * HW / 8;
    int grid = int((total + 255) / 256);
    const int cap = num_sms() * 16;
    if (grid > cap) grid = cap;
    OFB_LAUNCH(patchify_mixup_kernel, grid, 256, 0, s, img, reinterpret_cast<__nv_bfloat16*>(out), B, HW, P, mp);
    return err();
}
int launch_mixup_target(const long long* labels, float* target, int B, int C, double lam, double smoothing, cudaStream_t s) {
    if (B < 1 || C < 1) return 1011;
    const double off = smoothing / C, on = 1.0 - smoothing + off;
    int grid = int((size_t(B) * C + 255) / 256);
    const int cap = num_sms() * 16;
    if (grid > cap) grid = cap;
    OFB_LAUNCH(mixup_target_kernel, grid, 256, 0, s, labels, target, B, C, float(lam), float(1.0 - lam), float(on), float(off));
    return err();
}

int launch_pmim_mask(const float* noise, float* mask, int B, int L, int keep, cudaStream_t s) {
    OFB_LAUNCH(pmim_mask_kernel, B, 256, L * sizeof(float), s, noise, mask, L, keep);
    return err();
}

int launch_droppath_scale(const float* u, const float* drop_prob, float* scale, int n_layers2, int B, cudaStream_t s) {
    const int n = n_layers2 * B;
    OFB_LAUNCH(droppath_scale_kernel, (n + 255) / 256, 256, 0, s, u, drop_prob, scale, n_layers2, B);
    return err();
}

int launch_cls_rows(const float* cls, const float* pos, const float* gate, void* x, int B, int T, int D, cudaStream_t s) {
    OFB_LAUNCH(cls_rows_kernel, (B * D + 255) / 256, 256, 0, s, cls, pos, gate, reinterpret_cast<__nv_bfloat16*>(x), B, T, D);
    return err();
}

int launch_embed_bwd(const void* g0, const void* x0, const float* gate, const float* mask, void* dconv, float* part_gx, float* part_pos,
                     float* part_mt, int B, int T, int D, cudaStream_t s) {
    if (D % 8 != 0 || D > 4096) return 1011;
    const int VC = D / 8;
    // column slices: the split (1 .. 4, a divisor of the row's 16-byte vectors, slices of at least 128 bytes) whose T x CS CTAs
    // fill whole waves best (T = 197: 591 CTAs = 3.99 per SM)
    int CS = 1;
    double best = 1e30;
    for (int cs = 1; cs <= 4; ++cs) {
        if (VC % cs != 0 || VC / cs < 8) continue;
        const double per_sm = double(T) * cs / num_sms();
        const double cost = double(long(per_sm + 0.999999)) / per_sm;
        if (cost < best - 1e-9) { best = cost; CS = cs; }
    }
    const int VCc = VC / CS;
    int G = 512 / VCc;
    if (G > 16) G = 16;
    if (G > B) G = B;
    if (G < 1) return 1011;
    const int threads = ((VCc * G + 31) / 32) * 32;
    OFB_LAUNCH(embed_bwd_kernel, dim3(T, CS), threads, size_t(G) * VCc * 8 * sizeof(float), s, reinterpret_cast<const __nv_bfloat16*>(g0),
        reinterpret_cast<const __nv_bfloat16*>(x0), gate, mask, reinterpret_cast<__nv_bfloat16*>(dconv), part_gx, part_pos, part_mt, B, T, D, G,
        VCc);
    return err();
}

int launch_norm_targets(const float* img, const float* mask, float* tgt, int B, int HW, cudaStream_t s) {
    if (HW % NT_P != 0 || (reinterpret_cast<uintptr_t>(img) & 15u) != 0) return 1012;
    const int L = (HW / NT_P) * (HW / NT_P);
    dim3 grid((B * L + NT_PPC - 1) / NT_PPC);
    OFB_LAUNCH(norm_targets_kernel, grid, NT_THREADS, 0, s, img, mask, tgt, HW, B * L);
    return err();
}

int launch_ce(const float* logits, const int64_t* labels, float* loss_rows, void* dlogits, int B, int C, float smoothing, float gscale,
              cudaStream_t s) {
    OFB_LAUNCH(ce_kernel, B, 128, 0, s, logits, labels, loss_rows, reinterpret_cast<__nv_bfloat16*>(dlogits), C, smoothing, gscale / B);
    return err();
}

int launch_soft_ce(const float* logits, const float* target, float* loss_rows, void* dlogits, int B, int C, float gscale, cudaStream_t s) {
    OFB_LAUNCH(soft_ce_kernel, B, 128, 0, s, logits, target, loss_rows, reinterpret_cast<__nv_bfloat16*>(dlogits), C, gscale / B);
    return err();
}
int launch_eval_metrics(const float* logits, const int64_t* labels, float* out_rows, int B, int C, cudaStream_t s) {
    OFB_LAUNCH(eval_metrics_kernel, B, 128, 0, s, logits, labels, out_rows, C);
    return err();
}

int launch_loss_finalize(const float* loss_rows, int B, const float* dec_part, int n_dec_part, const float* mask, int n_mask,
                         const float* arch_loss, float grad_scale, float* scal, cudaStream_t s) {
    OFB_LAUNCH(loss_finalize_kernel, 1, 1024, 0, s, loss_rows, B, dec_part, n_dec_part, mask, n_mask, arch_loss, grad_scale, scal);
    return err();
}

int launch_adamw(float* p, float* g, float* m, float* v, void* shadow, const float* hyper, int nseg, const long long* seg_end,
                 long long n, int zero_grad, cudaStream_t s) {
    if (nseg < 1 || nseg > ADAM_MAX_SEGS || n % 4 != 0) return 1013;
    AdamSegs segs;
    segs.nseg = nseg;
    for (int i = 0; i < nseg; ++i) {
        if (seg_end[i] % 4 != 0) return 1013;
        segs.end[i] = seg_end[i];
    }
    const long long n4 = n / 4;
    long long grid = (n4 + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (grid > cap) grid = cap;
    OFB_LAUNCH(adamw_kernel, int(grid), 256, 0, s, p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow), hyper, segs, n4, zero_grad);
    return err();
}

// Upload of a small fp32 vector by a kernel: src may be PINNED HOST memory (unified addressing makes it device-readable). The
// per-step hyper-parameter vector (lr, bias corrections, w_p) goes this way instead of cudaMemcpyAsync so that it never queues on
// a copy engine behind the bulk H2D copy of the next batch (measured: +2.7 ms per step when it did), and stays in stream order.
__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
    pdl_wait();          // programmatic dependent launch: inputs of the previous kernel are complete from here on (launch.cuh)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
int launch_copy_f32(const float* src, float* dst, int n, cudaStream_t s) {
    if (n < 1) return 1011;
    OFB_LAUNCH(copy_f32_kernel, ((n + 255) / 256 > 32 ? 32 : (n + 255) / 256), 256, 0, s, src, dst, n);
    return err();
}

int launch_cast_bf16(const float* src, void* dst, long long n, cudaStream_t s) {
    if (n % 4 != 0) return 1013;
    const long long n4 = n / 4;
    long long grid = (n4 + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (grid > cap) grid = cap;
    OFB_LAUNCH(cast_bf16_kernel, int(grid), 256, 0, s, src, reinterpret_cast<__nv_bfloat16*>(dst), n4);
    return err();
}

}  // namespace ofb
