// attention.cu — fused bi-masked attention forward / backward for sm_100a (tcgen05 + TMEM + TMA).
//
// Reference: MAESparseAttention.forward layers.py:507-514 — softmax(q k^T * scale) v on gated q,k,v, materialising the
// [B,H,N,N] probabilities; backward through autograd (engine.py:169).  Here one CTA owns one (image, head):
// N = 197 tokens fit one KV pass (padded to 208 columns), so there is no online-softmax rescaling.
//
//   forward : S = Q K^T (UMMA 128x208x64) -> softmax in registers (tcgen05.ld) -> P (bf16, swizzled smem)
//             -> O = P V (UMMA 128x64x208, V as MN-major operand) -> O * droppath/rowsum, LSE
//   backward: S -> P = exp(S*scale - LSE);  dP = dO V^T;  dS = scale * P (dP - delta)
//             dQ = dS K ;  dK += dS^T Q ;  dV += P^T dO   (P / dS buffers double as K-major and MN-major operands)
//             epilogue multiplies by the bi-mask gate (d pre-gate qkv), writes token-major [M, 3D] and the per-image
//             column partials of d gate (sum dY*Y) and d bias (sum dY*g).
//
// q,k,v are read straight out of the token-major qkv GEMM output [B, T, 3, H, 64] through a 5-D tensor map, so the
// reference's reshape/permute/contiguous copy (layers.py:491) never happens; TMA zero-fills tokens >= T.
#include "ptx.cuh"
#include <math.h>

namespace ofb {

int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box);

static constexpr int HD = 64;        // head dim (DeiT-T/S/B)
static constexpr int KVP = 208;      // kv columns of S (197 padded to a multiple of 16)
static constexpr int QT = 128;       // q rows per tile
static constexpr int ATT_THREADS = 160;  // warps 0-3: softmax / epilogue rows, warp 4: TMA + MMA issue
static constexpr int TILE_B = QT * 128;      // 16 KB : [128 rows][64 bf16]
static constexpr int KV_B = KVP * 128;       // 26 KB : [208 rows][64 bf16]
static constexpr int PBUF_B = 4 * TILE_B;    // 64 KB : 4 atoms of 64 kv columns x 128 q rows

struct AttnArgs {
    int B, T, H;
    float scale;                       // constant (D/H)^-0.5, SURVEY App. B-4
    const float* drop_scale;           // [B] DropPath multiplier of this block's attention branch or null
    // forward
    __nv_bfloat16* o;                  // [B, T, H*64]   (already multiplied by drop_scale)
    float* lse;                        // [B, H, T]
    // backward
    const __nv_bfloat16* qkv;          // [B, T, 3, H, 64] gated q,k,v (for the d gate products)
    const __nv_bfloat16* o_in;         // forward output (scaled)
    const __nv_bfloat16* d_o;          // [B, T, H*64]
    const float* gate;                 // [H*64]
    __nv_bfloat16* dqkv;               // [B, T, 3, H, 64]  d(pre-gate qkv)
    float* part_gate;                  // [B, H*64]   sum_t (dq*q + dk*k + dv*v)
    float* part_bias;                  // [B, 3*H*64] sum_t d(pre-gate)
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// address of 16-byte chunk `idx16` (8 bf16 columns) of row r inside a [4 atoms][128 rows][128 B] swizzled buffer
__device__ __forceinline__ uint32_t pbuf_addr(uint32_t base, int r, int idx16) {
    return base + (idx16 >> 3) * TILE_B + sw128_offset(r, idx16 & 7);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const AttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sQ = smem_u32(smem);                       // 2 tiles
    const uint32_t sK = sQ + 2 * TILE_B;
    const uint32_t sV = sK + KV_B;
    const uint32_t sP = sV + KV_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TILE_B + 2 * KV_B + PBUF_B);
    const uint32_t ld_bar = smem_u32(&bars[0]), s_bar = smem_u32(&bars[1]), p_bar = smem_u32(&bars[2]), o_bar = smem_u32(&bars[3]);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[4]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;

    if (threadIdx.x == 0) {
        mbar_init(ld_bar, 1); mbar_init(s_bar, 1); mbar_init(p_bar, 128); mbar_init(o_bar, 1);
        mbar_fence_init();
    }
    if (warp == 4) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tS = tmem, tO = tmem + 256;
    constexpr uint32_t IDESC_S = make_idesc_bf16(128, KVP, 0, 0);
    constexpr uint32_t IDESC_O = make_idesc_bf16(128, HD, 0, 1);

    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(ld_bar, 2 * TILE_B + 2 * KV_B);
            tma_load_5d(sQ, &tm_q, ld_bar, 0, 0, h, 0, b);
            tma_load_5d(sQ + TILE_B, &tm_q, ld_bar, 0, QT, h, 0, b);
            tma_load_5d(sK, &tm_kv, ld_bar, 0, 0, h, 1, b);
            tma_load_5d(sV, &tm_kv, ld_bar, 0, 0, h, 2, b);
            mbar_wait(ld_bar, 0);
            tc_fence_after();
            auto issue_s = [&](int i) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(tS, make_smem_desc_sw128(sQ + i * TILE_B + k * 32, 0, 1024), make_smem_desc_sw128(sK + k * 32, 0, 1024),
                              IDESC_S, k > 0);
                umma_commit(s_bar);
            };
            issue_s(0);
            for (int i = 0; i < 2; ++i) {
                mbar_wait(p_bar, i);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < KVP / 16; ++k)
                    umma_bf16(tO, make_smem_desc_sw128(sP + (k >> 2) * TILE_B + (k & 3) * 32, 0, 1024),
                              make_smem_desc_sw128(sV + k * 2048, 0, 1024), IDESC_O, k > 0);
                umma_commit(o_bar);
                if (i == 0) issue_s(1);
            }
        }
    } else {
        const int r = warp * 32 + lane;
        const uint32_t lane_base = uint32_t(warp * 32) << 16;
        const float sl2 = a.scale * 1.4426950408889634f;
        const float dps = a.drop_scale != nullptr ? a.drop_scale[b] : 1.f;
        for (int i = 0; i < 2; ++i) {
            const int t = i * QT + r;
            mbar_wait(s_bar, i);
            tc_fence_after();
            float v[32];
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 6; ++c) {
                tmem_ld32(tS + lane_base + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);   // columns < 192 are always valid tokens
            }
            tmem_ld16(tS + lane_base + 192, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (192 + j < a.T) mx = fmaxf(mx, v[j]);
            float sum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 7; ++c) {
                if (c < 6) tmem_ld32(tS + lane_base + c * 32, v);
                else tmem_ld16(tS + lane_base + 192, v);
                tmem_ld_wait();
                const int nj = c < 6 ? 32 : 16;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j < nj) {
                        const float p = (c * 32 + j < a.T) ? exp2f((v[j] - mx) * sl2) : 0.f;
                        v[j] = p;
                        sum += p;
                    }
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    if (q4 * 8 < nj) {
                        uint4 pk;
                        pk.x = pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]); pk.y = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                        pk.z = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); pk.w = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                        st_shared_v4(pbuf_addr(sP, r, c * 4 + q4), pk);
                    }
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_bar);

            mbar_wait(o_bar, i);
            tc_fence_after();
            const float inv = dps / sum;
            float o0[32], o1[32];
            tmem_ld32(tO + lane_base, o0);
            tmem_ld32(tO + lane_base + 32, o1);
            tmem_ld_wait();
            if (t < a.T) {
                uint4* dst = reinterpret_cast<uint4*>(a.o + (size_t(b) * a.T + t) * (a.H * HD) + h * HD);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 pk;
                    pk.x = pack_bf16x2(o0[q4 * 8 + 0] * inv, o0[q4 * 8 + 1] * inv); pk.y = pack_bf16x2(o0[q4 * 8 + 2] * inv, o0[q4 * 8 + 3] * inv);
                    pk.z = pack_bf16x2(o0[q4 * 8 + 4] * inv, o0[q4 * 8 + 5] * inv); pk.w = pack_bf16x2(o0[q4 * 8 + 6] * inv, o0[q4 * 8 + 7] * inv);
                    dst[q4] = pk;
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 pk;
                    pk.x = pack_bf16x2(o1[q4 * 8 + 0] * inv, o1[q4 * 8 + 1] * inv); pk.y = pack_bf16x2(o1[q4 * 8 + 2] * inv, o1[q4 * 8 + 3] * inv);
                    pk.z = pack_bf16x2(o1[q4 * 8 + 4] * inv, o1[q4 * 8 + 5] * inv); pk.w = pack_bf16x2(o1[q4 * 8 + 6] * inv, o1[q4 * 8 + 7] * inv);
                    dst[4 + q4] = pk;
                }
                a.lse[(size_t(b) * a.H + h) * a.T + t] = mx * a.scale + logf(sum);
            }
            tc_fence_before();   // order the O reads before the next tile's MMA (signalled through p_bar)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// TMEM columns: [0,208) S then dP then (first 64) dQ ; [256,384) dK (2 kv tiles x 64) ; [384,512) dV (2 kv tiles x 64)
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                const __grid_constant__ CUtensorMap tm_do, const AttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sDO = sQ + TILE_B;
    const uint32_t sK = sDO + TILE_B;
    const uint32_t sV = sK + KV_B;
    const uint32_t sP = sV + KV_B;
    const uint32_t sDS = sP + PBUF_B;
    uint8_t* tail = smem + 2 * TILE_B + 2 * KV_B + 2 * PBUF_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    const uint32_t ld_bar = smem_u32(&bars[0]), s_bar = smem_u32(&bars[1]), p_bar = smem_u32(&bars[2]), dp_bar = smem_u32(&bars[3]),
                   ds_bar = smem_u32(&bars[4]), dq_bar = smem_u32(&bars[5]), dqr_bar = smem_u32(&bars[6]), tile_bar = smem_u32(&bars[7]);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[8]);
    float* cs_gate = reinterpret_cast<float*>(tail + 128);   // [64]   sum dY*Y over q,k,v
    float* cs_bias = cs_gate + 64;                            // [3][64] sum d(pre-gate)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const int D = a.H * HD;

    if (threadIdx.x == 0) {
        mbar_init(ld_bar, 1); mbar_init(s_bar, 1); mbar_init(p_bar, 128); mbar_init(dp_bar, 1);
        mbar_init(ds_bar, 128); mbar_init(dq_bar, 1); mbar_init(dqr_bar, 128); mbar_init(tile_bar, 1);
        mbar_fence_init();
    }
    if (warp == 4) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
    // zero the P / dS buffers once: kv columns >= 208 (and anything not rewritten) must read as 0 for the MN-major MMAs
    for (int i = threadIdx.x; i < 2 * PBUF_B / 16; i += ATT_THREADS) st_shared_v4(sP + i * 16, make_uint4(0, 0, 0, 0));
    for (int i = threadIdx.x; i < 256; i += ATT_THREADS) cs_gate[i] = 0.f;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tA = tmem, tDK = tmem + 256, tDV = tmem + 384;
    constexpr uint32_t IDESC_S = make_idesc_bf16(128, KVP, 0, 0);     // S, dP : K-major x K-major
    constexpr uint32_t IDESC_DQ = make_idesc_bf16(128, HD, 0, 1);     // dQ     : dS K-major, K MN-major
    constexpr uint32_t IDESC_KV = make_idesc_bf16(128, HD, 1, 1);     // dK, dV : MN-major x MN-major

    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(ld_bar, 2 * TILE_B + 2 * KV_B);
            tma_load_5d(sQ, &tm_q, ld_bar, 0, 0, h, 0, b);
            tma_load_4d(sDO, &tm_do, ld_bar, h * HD, 0, b, 0);
            tma_load_5d(sK, &tm_kv, ld_bar, 0, 0, h, 1, b);
            tma_load_5d(sV, &tm_kv, ld_bar, 0, 0, h, 2, b);
            for (int i = 0; i < 2; ++i) {
                mbar_wait(ld_bar, i);
                if (i > 0) mbar_wait(dqr_bar, 0);      // dQ(0) drained from TMEM region A
                tc_fence_after();
                // S = Q K^T
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(tA, make_smem_desc_sw128(sQ + k * 32, 0, 1024), make_smem_desc_sw128(sK + k * 32, 0, 1024), IDESC_S, k > 0);
                umma_commit(s_bar);
                mbar_wait(p_bar, i);
                tc_fence_after();
                // dP = dO V^T  (region A again: S fully consumed once p_bar completes)
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(tA, make_smem_desc_sw128(sDO + k * 32, 0, 1024), make_smem_desc_sw128(sV + k * 32, 0, 1024), IDESC_S, k > 0);
                umma_commit(dp_bar);
                // dV[kv tile m] += P^T dO      (M = kv, K = q rows of this tile)
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int k = 0; k < QT / 16; ++k)
                        umma_bf16(tDV + m * HD, make_smem_desc_sw128(sP + (2 * m) * TILE_B + k * 2048, TILE_B, 1024),
                                  make_smem_desc_sw128(sDO + k * 2048, 0, 1024), IDESC_KV, (i > 0 || k > 0) ? 1u : 0u);
                mbar_wait(ds_bar, i);
                tc_fence_after();
                // dQ = dS K   (K = kv)
#pragma unroll
                for (int k = 0; k < KVP / 16; ++k)
                    umma_bf16(tA, make_smem_desc_sw128(sDS + (k >> 2) * TILE_B + (k & 3) * 32, 0, 1024),
                              make_smem_desc_sw128(sK + k * 2048, 0, 1024), IDESC_DQ, k > 0);
                umma_commit(dq_bar);
                // dK[kv tile m] += dS^T Q
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int k = 0; k < QT / 16; ++k)
                        umma_bf16(tDK + m * HD, make_smem_desc_sw128(sDS + (2 * m) * TILE_B + k * 2048, TILE_B, 1024),
                                  make_smem_desc_sw128(sQ + k * 2048, 0, 1024), IDESC_KV, (i > 0 || k > 0) ? 1u : 0u);
                umma_commit(tile_bar);
                if (i == 0) {
                    mbar_wait(tile_bar, 0);            // every MMA reading Q0 / dO0 / P / dS has completed
                    mbar_arrive_expect_tx(ld_bar, 2 * TILE_B);
                    tma_load_5d(sQ, &tm_q, ld_bar, 0, QT, h, 0, b);
                    tma_load_4d(sDO, &tm_do, ld_bar, h * HD, QT, b, 0);
                }
            }
        }
    } else {
        const int r = warp * 32 + lane;
        const uint32_t lane_base = uint32_t(warp * 32) << 16;
        const float dps = a.drop_scale != nullptr ? a.drop_scale[b] : 1.f;
        const float inv_dps = dps != 0.f ? 1.f / dps : 0.f;
        const float* gate = a.gate + h * HD;
        float v[32];
        for (int i = 0; i < 2; ++i) {
            const int t = i * QT + r;
            const bool t_ok = t < a.T;
            // per-row scalars: LSE and delta = rowsum(dO * O)
            float lse = 0.f, delta = 0.f;
            if (t_ok) {
                lse = a.lse[(size_t(b) * a.H + h) * a.T + t];
                const uint4* po = reinterpret_cast<const uint4*>(a.o_in + (size_t(b) * a.T + t) * D + h * HD);
                const uint4* pd = reinterpret_cast<const uint4*>(a.d_o + (size_t(b) * a.T + t) * D + h * HD);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const uint4 x = __ldg(po + q4), y = __ldg(pd + q4);
                    float2 f, g;
                    f = unpack_bf16x2(x.x); g = unpack_bf16x2(y.x); delta += f.x * g.x + f.y * g.y;
                    f = unpack_bf16x2(x.y); g = unpack_bf16x2(y.y); delta += f.x * g.x + f.y * g.y;
                    f = unpack_bf16x2(x.z); g = unpack_bf16x2(y.z); delta += f.x * g.x + f.y * g.y;
                    f = unpack_bf16x2(x.w); g = unpack_bf16x2(y.w); delta += f.x * g.x + f.y * g.y;
                }
                delta *= inv_dps;   // stored O carries the DropPath factor
            }
            // ---- P = exp(S*scale - LSE) ----
            mbar_wait(s_bar, i);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 7; ++c) {
                if (c < 6) tmem_ld32(tA + lane_base + c * 32, v);
                else tmem_ld16(tA + lane_base + 192, v);
                tmem_ld_wait();
                const int nj = c < 6 ? 32 : 16;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < nj) v[j] = (t_ok && c * 32 + j < a.T) ? __expf(v[j] * a.scale - lse) : 0.f;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    if (q4 * 8 < nj) {
                        uint4 pk;
                        pk.x = pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]); pk.y = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                        pk.z = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); pk.w = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                        st_shared_v4(pbuf_addr(sP, r, c * 4 + q4), pk);
                    }
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_bar);
            // ---- dS = scale * P * (dP - delta) ----
            mbar_wait(dp_bar, i);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 7; ++c) {
                if (c < 6) tmem_ld32(tA + lane_base + c * 32, v);
                else tmem_ld16(tA + lane_base + 192, v);
                tmem_ld_wait();
                const int nj = c < 6 ? 32 : 16;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    if (q4 * 8 < nj) {
                        const uint4 pp = ld_shared_v4(pbuf_addr(sP, r, c * 4 + q4));
                        float2 p0 = unpack_bf16x2(pp.x), p1 = unpack_bf16x2(pp.y), p2 = unpack_bf16x2(pp.z), p3 = unpack_bf16x2(pp.w);
                        const float* w = v + q4 * 8;
                        uint4 pk;
                        pk.x = pack_bf16x2(a.scale * p0.x * (w[0] - delta), a.scale * p0.y * (w[1] - delta));
                        pk.y = pack_bf16x2(a.scale * p1.x * (w[2] - delta), a.scale * p1.y * (w[3] - delta));
                        pk.z = pack_bf16x2(a.scale * p2.x * (w[4] - delta), a.scale * p2.y * (w[5] - delta));
                        pk.w = pack_bf16x2(a.scale * p3.x * (w[6] - delta), a.scale * p3.y * (w[7] - delta));
                        st_shared_v4(pbuf_addr(sDS, r, c * 4 + q4), pk);
                    }
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(ds_bar);
            // ---- dQ epilogue ----
            mbar_wait(dq_bar, i);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                tmem_ld32(tA + lane_base + c * 32, v);
                tmem_ld_wait();
                float gy[32];
                if (t_ok) {
                    const size_t off = ((size_t(b) * a.T + t) * 3 + 0) * D + h * HD + c * 32;
                    const uint4* pq = reinterpret_cast<const uint4*>(a.qkv + off);
                    uint4* pdq = reinterpret_cast<uint4*>(a.dqkv + off);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const uint4 x = __ldg(pq + q4);
                        float2 f;
                        f = unpack_bf16x2(x.x); gy[q4 * 8 + 0] = f.x * v[q4 * 8 + 0]; gy[q4 * 8 + 1] = f.y * v[q4 * 8 + 1];
                        f = unpack_bf16x2(x.y); gy[q4 * 8 + 2] = f.x * v[q4 * 8 + 2]; gy[q4 * 8 + 3] = f.y * v[q4 * 8 + 3];
                        f = unpack_bf16x2(x.z); gy[q4 * 8 + 4] = f.x * v[q4 * 8 + 4]; gy[q4 * 8 + 5] = f.y * v[q4 * 8 + 5];
                        f = unpack_bf16x2(x.w); gy[q4 * 8 + 6] = f.x * v[q4 * 8 + 6]; gy[q4 * 8 + 7] = f.y * v[q4 * 8 + 7];
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= __ldg(gate + c * 32 + j);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        uint4 pk;
                        pk.x = pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]); pk.y = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                        pk.z = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); pk.w = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                        pdq[q4] = pk;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) { gy[j] = 0.f; v[j] = 0.f; }
                }
                // column sums over the 32 rows of this warp, then across warps / tiles in shared memory
                float sg, sb;
                {
                    // butterfly transpose-reduce (31 shuffles each)
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool upper = (lane & o) != 0;
#pragma unroll
                        for (int j = 0; j < o; ++j) {
                            const float s0 = upper ? gy[j] : gy[j + o], k0 = upper ? gy[j + o] : gy[j];
                            gy[j] = k0 + __shfl_xor_sync(0xffffffffu, s0, o);
                            const float s1 = upper ? v[j] : v[j + o], k1 = upper ? v[j + o] : v[j];
                            v[j] = k1 + __shfl_xor_sync(0xffffffffu, s1, o);
                        }
                    }
                    sg = gy[0]; sb = v[0];
                }
                atomicAdd(&cs_gate[c * 32 + lane], sg);
                atomicAdd(&cs_bias[0 * 64 + c * 32 + lane], sb);
            }
            tc_fence_before();
            mbar_arrive(dqr_bar);
        }
        // ---- dK / dV epilogue (kv rows: tile m covers kv = m*128 + r) ----
        mbar_wait(tile_bar, 1);
        tc_fence_after();
#pragma unroll 1
        for (int which = 1; which <= 2; ++which) {
            const uint32_t tbase = which == 1 ? tDK : tDV;
#pragma unroll 1
            for (int m = 0; m < 2; ++m) {
                const int kv = m * QT + r;
                const bool kv_ok = kv < a.T;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    tmem_ld32(tbase + lane_base + m * HD + c * 32, v);
                    tmem_ld_wait();
                    float gy[32];
                    if (kv_ok) {
                        const size_t off = ((size_t(b) * a.T + kv) * 3 + which) * D + h * HD + c * 32;
                        const uint4* px = reinterpret_cast<const uint4*>(a.qkv + off);
                        uint4* pd = reinterpret_cast<uint4*>(a.dqkv + off);
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const uint4 x = __ldg(px + q4);
                            float2 f;
                            f = unpack_bf16x2(x.x); gy[q4 * 8 + 0] = f.x * v[q4 * 8 + 0]; gy[q4 * 8 + 1] = f.y * v[q4 * 8 + 1];
                            f = unpack_bf16x2(x.y); gy[q4 * 8 + 2] = f.x * v[q4 * 8 + 2]; gy[q4 * 8 + 3] = f.y * v[q4 * 8 + 3];
                            f = unpack_bf16x2(x.z); gy[q4 * 8 + 4] = f.x * v[q4 * 8 + 4]; gy[q4 * 8 + 5] = f.y * v[q4 * 8 + 5];
                            f = unpack_bf16x2(x.w); gy[q4 * 8 + 6] = f.x * v[q4 * 8 + 6]; gy[q4 * 8 + 7] = f.y * v[q4 * 8 + 7];
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= __ldg(gate + c * 32 + j);
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            uint4 pk;
                            pk.x = pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]); pk.y = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                            pk.z = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); pk.w = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                            pd[q4] = pk;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { gy[j] = 0.f; v[j] = 0.f; }
                    }
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool upper = (lane & o) != 0;
#pragma unroll
                        for (int j = 0; j < o; ++j) {
                            const float s0 = upper ? gy[j] : gy[j + o], k0 = upper ? gy[j + o] : gy[j];
                            gy[j] = k0 + __shfl_xor_sync(0xffffffffu, s0, o);
                            const float s1 = upper ? v[j] : v[j + o], k1 = upper ? v[j + o] : v[j];
                            v[j] = k1 + __shfl_xor_sync(0xffffffffu, s1, o);
                        }
                    }
                    atomicAdd(&cs_gate[c * 32 + lane], gy[0]);
                    atomicAdd(&cs_bias[which * 64 + c * 32 + lane], v[0]);
                }
            }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (r < 64) a.part_gate[size_t(b) * D + h * HD + r] = cs_gate[r];
        for (int i = r; i < 192; i += 128) a.part_bias[size_t(b) * 3 * D + (i / 64) * D + h * HD + (i % 64)] = cs_bias[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static constexpr int FWD_SMEM = 1024 + 2 * TILE_B + 2 * KV_B + PBUF_B + 256;
static constexpr int BWD_SMEM = 1024 + 2 * TILE_B + 2 * KV_B + 2 * PBUF_B + 128 + 1024 + 64;

static int make_qkv_maps(const void* qkv, int B, int T, int H, CUtensorMap* tq, CUtensorMap* tkv) {
    const uint64_t D = uint64_t(H) * HD;
    uint64_t dims[5] = {HD, uint64_t(T), uint64_t(H), 3, uint64_t(B)};
    uint64_t str[5] = {1, 3 * D, HD, D, uint64_t(T) * 3 * D};
    uint32_t boxq[5] = {HD, QT, 1, 1, 1}, boxkv[5] = {HD, KVP, 1, 1, 1};
    int r = make_tmap_bf16(tq, qkv, 5, dims, str, boxq);
    if (r) return r;
    return make_tmap_bf16(tkv, qkv, 5, dims, str, boxkv);
}

int launch_attn_fwd(const void* qkv, void* o, float* lse, const float* drop_scale, int B, int T, int H, float scale, cudaStream_t s) {
    if (T > KVP - 0 || T > 2 * QT) return 1030;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
        if (e != cudaSuccess) return int(e);
        configured = true;
    }
    CUtensorMap tq, tkv;
    int r = make_qkv_maps(qkv, B, T, H, &tq, &tkv);
    if (r) return r;
    AttnArgs a{};
    a.B = B; a.T = T; a.H = H; a.scale = scale; a.drop_scale = drop_scale;
    a.o = reinterpret_cast<__nv_bfloat16*>(o); a.lse = lse;
    attn_fwd_kernel<<<B * H, ATT_THREADS, FWD_SMEM, s>>>(tq, tkv, a);
    return int(cudaGetLastError());
}

int launch_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, const float* gate, const float* drop_scale,
                    void* dqkv, float* part_gate, float* part_bias, int B, int T, int H, float scale, cudaStream_t s) {
    if (T > KVP || T > 2 * QT) return 1030;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
        if (e != cudaSuccess) return int(e);
        configured = true;
    }
    CUtensorMap tq, tkv, tdo;
    int r = make_qkv_maps(qkv, B, T, H, &tq, &tkv);
    if (r) return r;
    {
        const uint64_t D = uint64_t(H) * HD;
        uint64_t dims[4] = {D, uint64_t(T), uint64_t(B), 1};
        uint64_t str[4] = {1, D, uint64_t(T) * D, uint64_t(B) * T * D};
        uint32_t box[4] = {HD, QT, 1, 1};
        r = make_tmap_bf16(&tdo, d_o, 4, dims, str, box);
        if (r) return r;
    }
    AttnArgs a{};
    a.B = B; a.T = T; a.H = H; a.scale = scale; a.drop_scale = drop_scale;
    a.lse = const_cast<float*>(lse);
    a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
    a.o_in = reinterpret_cast<const __nv_bfloat16*>(o);
    a.d_o = reinterpret_cast<const __nv_bfloat16*>(d_o);
    a.gate = gate;
    a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
    a.part_gate = part_gate; a.part_bias = part_bias;
    attn_bwd_kernel<<<B * H, ATT_THREADS, BWD_SMEM, s>>>(tq, tkv, tdo, a);
    return int(cudaGetLastError());
}

}  // namespace ofb
