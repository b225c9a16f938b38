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

int num_sms();
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
// forward: persistent, warp-specialised.  One CTA per SM loops over (image, head) items; the two 128-row q tiles of an
// item are two independent "slots" (own TMEM half, own P buffer, own 8 compute warps), so the S / PV MMAs of one slot
// and the TMA loads of the next item overlap the softmax of the other slot.
//   warps 0-7  : slot 0 (q rows 0..127)      warp w: TMEM lane quarter w&3, kv-column half (w>>2)&1
//   warps 8-15 : slot 1 (q rows 128..255)
//   warp 16    : TMA producer (Q0,Q1,K then V of the next item as soon as the MMAs that read them have retired)
//   warp 17    : MMA issuer (+ TMEM allocation)
// TMEM: slot s owns columns [256 s, 256 s + 208) for S; O (64 columns) overlays S once P has been written out.
// ---------------------------------------------------------------------------------------------
static constexpr int FWD_CW = 16;                         // compute warps
static constexpr int FWD_THREADS = (FWD_CW + 2) * 32;     // 576
static constexpr int KV_SPLIT = 112;                      // kv columns [0,112) -> half 0, [112,208) -> half 1

__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const AttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);                       // 2 tiles
    const uint32_t sK = sQ + 2 * TILE_B;
    const uint32_t sV = sK + KV_B;
    const uint32_t sP = sV + KV_B;                            // 2 slots x PBUF_B
    const uint32_t sX = sP + 2 * PBUF_B;                      // exchange: [max|sum][slot][half][128] floats
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TILE_B + 2 * KV_B + 2 * PBUF_B + 4096);
    const uint32_t qk_full = smem_u32(&bars[0]), qk_empty = smem_u32(&bars[1]), v_full = smem_u32(&bars[2]), v_empty = smem_u32(&bars[3]);
    // per slot: s_full, p_full, o_full, o_empty
    auto slot_bar = [&](int s, int k) { return smem_u32(&bars[4 + s * 4 + k]); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[12]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.B * a.H;

    if (threadIdx.x == 0) {
        if ((sQ & 1023u) != 0) { printf("ofb: dynamic smem base not 1024-aligned\n"); __trap(); }
        mbar_init(qk_full, 1); mbar_init(qk_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(slot_bar(s, 0), 1); mbar_init(slot_bar(s, 1), 8); mbar_init(slot_bar(s, 2), 1); mbar_init(slot_bar(s, 3), 8);
        }
        mbar_fence_init();
    }
    if (warp == FWD_CW + 1) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC_S = make_idesc_bf16(128, KVP, 0, 0);
    constexpr uint32_t IDESC_O = make_idesc_bf16(128, HD, 0, 1);

    if (warp == FWD_CW) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&tm_q);
            tma_prefetch_desc(&tm_kv);
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int b = item / a.H, h = item % a.H;
                const uint32_t ph = it & 1;
                mbar_wait(qk_empty, ph ^ 1);
                mbar_arrive_expect_tx(qk_full, 2 * TILE_B + KV_B);
                tma_load_5d(sQ, &tm_q, qk_full, 0, 0, h, 0, b);
                tma_load_5d(sQ + TILE_B, &tm_q, qk_full, 0, QT, h, 0, b);
                tma_load_5d(sK, &tm_kv, qk_full, 0, 0, h, 1, b);
                mbar_wait(v_empty, ph ^ 1);
                mbar_arrive_expect_tx(v_full, KV_B);
                tma_load_5d(sV, &tm_kv, v_full, 0, 0, h, 2, b);
            }
        }
    } else if (warp == FWD_CW + 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                mbar_wait(qk_full, ph);
                for (int s = 0; s < 2; ++s) {
                    mbar_wait(slot_bar(s, 3), ph ^ 1);          // slot's TMEM free (previous item's O read out)
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_bf16(tmem + s * 256, make_smem_desc_sw128(sQ + s * TILE_B + k * 32, 0, 1024),
                                  make_smem_desc_sw128(sK + k * 32, 0, 1024), IDESC_S, k > 0);
                    umma_commit(slot_bar(s, 0));
                }
                umma_commit(qk_empty);                          // Q / K may be overwritten by the next item's loads
                mbar_wait(v_full, ph);
                for (int s = 0; s < 2; ++s) {
                    mbar_wait(slot_bar(s, 1), ph);              // P of this slot is in shared memory, S fully consumed
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < KVP / 16; ++k)
                        umma_bf16(tmem + s * 256, make_smem_desc_sw128(sP + s * PBUF_B + (k >> 2) * TILE_B + (k & 3) * 32, 0, 1024),
                                  make_smem_desc_sw128(sV + k * 2048, 0, 1024), IDESC_O, k > 0);
                    umma_commit(slot_bar(s, 2));
                }
                umma_commit(v_empty);
            }
        }
    } else {
        // ===================== softmax / epilogue warps =====================
        const int s = warp >> 3;                 // slot = q tile
        const int q = warp & 3;                  // TMEM lane quarter
        const int hf = (warp >> 2) & 1;          // kv-column half
        const int r = q * 32 + lane;             // row inside the tile
        const uint32_t tS = tmem + s * 256 + (uint32_t(q * 32) << 16);
        const uint32_t pS = sP + s * PBUF_B;
        const uint32_t x_mine = sX + ((s * 2 + hf) * 128 + r) * 4, x_other = sX + ((s * 2 + (hf ^ 1)) * 128 + r) * 4;
        const int pair_bar = 1 + s * 4 + q;      // named barrier shared by the two warps that own this row quarter
        const float sl2 = a.scale * 1.4426950408889634f;
        const int c_begin = hf ? KV_SPLIT : 0;
        const int n32 = 3;                       // both halves: three 32-column chunks (+ one 16-column chunk in half 0)
        const int t = s * QT + r;
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int b = item / a.H, h = item % a.H;
            const uint32_t ph = it & 1;
            const float dps = a.drop_scale != nullptr ? __ldg(a.drop_scale + b) : 1.f;
            mbar_wait(slot_bar(s, 0), ph);
            tc_fence_after();
            float v[32];
            // ---- pass 1: row maximum over this warp's columns ----
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < n32; ++c) {
                const int col = c_begin + c * 32;
                tmem_ld32(tS + col, v);
                tmem_ld_wait();
                if (col + 32 <= a.T) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, col + j < a.T ? v[j] : -INFINITY);
                }
            }
            if (hf == 0) {
                tmem_ld16(tS + 96, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) mx = fmaxf(mx, 96 + j < a.T ? v[j] : -INFINITY);
            }
            st_shared_f32(x_mine, mx);
            named_bar_sync(pair_bar, 64);
            mx = fmaxf(mx, ld_shared_f32(x_other));
            const float mxs = mx * sl2;
            // ---- pass 2: P = exp(S*scale - max), row sums, P -> shared memory (bf16, swizzled K-major A operand) ----
            float sum = 0.f;
#pragma unroll 1
            for (int c = 0; c < n32 + 1; ++c) {
                const int col = c_begin + c * 32;
                const int nj = c < n32 ? 32 : (hf == 0 ? 16 : 0);
                if (nj == 0) break;
                if (nj == 32) tmem_ld32(tS + col, v);
                else tmem_ld16(tS + col, v);
                tmem_ld_wait();
                if (col + nj <= a.T) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < nj) { v[j] = fast_ex2(fmaf(v[j], sl2, -mxs)); sum += v[j]; }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < nj) { v[j] = col + j < a.T ? fast_ex2(fmaf(v[j], sl2, -mxs)) : 0.f; sum += v[j]; }
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    if (q4 * 8 < nj) {
                        uint4 pk;
                        pk.x = pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]); pk.y = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                        pk.z = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); pk.w = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                        st_shared_v4(pbuf_addr(pS, r, (col >> 3) + q4), pk);
                    }
                }
            }
            st_shared_f32(x_mine + 2048, sum);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(slot_bar(s, 1));
            named_bar_sync(pair_bar, 64);
            sum += ld_shared_f32(x_other + 2048);
            // ---- epilogue: O / rowsum * droppath; this warp owns head-dim columns [32 hf, 32 hf + 32) ----
            mbar_wait(slot_bar(s, 2), ph);
            tc_fence_after();
            tmem_ld32(tS + hf * 32, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(slot_bar(s, 3));
            if (t < a.T) {
                const float inv = dps / sum;
                uint4* dst = reinterpret_cast<uint4*>(a.o + (size_t(b) * a.T + t) * (a.H * HD) + h * HD + hf * 32);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 pk;
                    pk.x = pack_bf16x2(v[q4 * 8 + 0] * inv, v[q4 * 8 + 1] * inv); pk.y = pack_bf16x2(v[q4 * 8 + 2] * inv, v[q4 * 8 + 3] * inv);
                    pk.z = pack_bf16x2(v[q4 * 8 + 4] * inv, v[q4 * 8 + 5] * inv); pk.w = pack_bf16x2(v[q4 * 8 + 6] * inv, v[q4 * 8 + 7] * inv);
                    dst[q4] = pk;
                }
                if (hf == 0) a.lse[(size_t(b) * a.H + h) * a.T + t] = mx * a.scale + logf(sum);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FWD_CW + 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------
// backward: persistent, warp-specialised; one CTA per SM loops over (image, head) items, two q tiles per item.
//   warps 0-15 : compute.  warp w: TMEM lane quarter w&3, column group w>>2
//                (kv columns [0,64) [64,112) [112,160) [160,208) of S / dP; head-dim columns 16 cg .. 16 cg + 15 of dQ / dK / dV)
//   warp 16    : TMA producer        warp 17 : MMA issuer (+ TMEM allocation)
// TMEM columns: [0,208) S then dP then (first 64) dQ ; [256,384) dK (2 kv tiles x 64) ; [384,512) dV (2 kv tiles x 64)
// ---------------------------------------------------------------------------------------------
static constexpr int BWD_CW = 16;
static constexpr int BWD_THREADS = (BWD_CW + 2) * 32;     // 576
static constexpr int BWD_CT = BWD_CW * 32;                // compute threads

// 16 values per lane -> lanes with even index end with the sum over all 32 lanes of column (lane >> 1) & 15
__device__ __forceinline__ float bfly16(float (&v)[16]) {
    const uint32_t lane = lane_id();
#pragma unroll
    for (int cnt = 8, o = 16; cnt >= 1; cnt >>= 1, o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
            const float send = upper ? v[i] : v[i + cnt];
            const float keep = upper ? v[i + cnt] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// epilogue of one 16-column slice of dQ / dK / dV for one token row: multiply by the gate (d pre-gate) and store; the row's
// contributions to the gate / bias column sums are accumulated in REGISTERS (packed fp32x2) and reduced across rows once
// per item and quantity (reduce_cols16), not once per slice. x0/x1 = the row's 16 gated q/k/v values (prefetched by the
// caller so the global-load latency overlaps the MMAs).
__device__ __forceinline__ void dqkv_slice_acc(const float (&v)[16], bool ok, const uint4& x0, const uint4& x1, __nv_bfloat16* dy,
                                               const float2 (&g2)[8], float2 (&ag)[8], float2 (&ab)[8]) {
    if (!ok) return;
    const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    uint32_t ow[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 vv = make_float2(v[2 * j], v[2 * j + 1]);
        ag[j] = fma2(unpack_bf16x2(xw[j]), vv, ag[j]);
        const float2 o = mul2(vv, g2[j]);
        ab[j] = add2(ab[j], o);
        ow[j] = pack_bf16x2(o.x, o.y);
    }
    reinterpret_cast<uint4*>(dy)[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    reinterpret_cast<uint4*>(dy)[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
}
// column sums over the 32 rows of this warp of 16 per-thread accumulators -> shared-memory accumulators (one atomic per column
// and warp); the accumulators are cleared
__device__ __forceinline__ void reduce_cols16(float2 (&acc)[8], float* cs) {
    float t[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) { t[2 * j] = acc[j].x; t[2 * j + 1] = acc[j].y; acc[j] = make_float2(0.f, 0.f); }
    const float sum = bfly16(t);
    const uint32_t lane = lane_id();
    if ((lane & 1u) == 0) atomicAdd(cs + (lane >> 1), sum);
}
__device__ __forceinline__ float dot8(const uint4& x, const uint4& y) {
    float2 f, g;
    float d = 0.f;
    f = unpack_bf16x2(x.x); g = unpack_bf16x2(y.x); d = fmaf(f.x, g.x, fmaf(f.y, g.y, d));
    f = unpack_bf16x2(x.y); g = unpack_bf16x2(y.y); d = fmaf(f.x, g.x, fmaf(f.y, g.y, d));
    f = unpack_bf16x2(x.z); g = unpack_bf16x2(y.z); d = fmaf(f.x, g.x, fmaf(f.y, g.y, d));
    f = unpack_bf16x2(x.w); g = unpack_bf16x2(y.w); d = fmaf(f.x, g.x, fmaf(f.y, g.y, d));
    return d;
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                const __grid_constant__ CUtensorMap tm_do, const AttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sDO = sQ + TILE_B;
    const uint32_t sK = sDO + TILE_B;
    const uint32_t sV = sK + KV_B;
    const uint32_t sP = sV + KV_B;
    const uint32_t sDS = sP + PBUF_B;
    uint8_t* tail = smem + 2 * TILE_B + 2 * KV_B + 2 * PBUF_B;
    float* cs = reinterpret_cast<float*>(tail);                 // [2 parities][256]: gate[64] | bias q,k,v [3][64]
    const uint32_t sXD = smem_u32(tail + 2048);                 // [4 column groups][128 rows] delta partials
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 4096);
    enum { KV_FULL = 0, KV_EMPTY, QDO_FULL, QDO_EMPTY, S_FULL, P_FULL, DP_FULL, DS_FULL, DQ_FULL, A_EMPTY, ACC_FULL, ACC_EMPTY, NBAR };
    auto bar = [&](int k) { return smem_u32(&bars[k]); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[NBAR]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.B * a.H;
    const int D = a.H * HD;

    if (threadIdx.x == 0) {
        if ((sQ & 1023u) != 0) { printf("ofb: dynamic smem base not 1024-aligned\n"); __trap(); }
        mbar_init(bar(KV_FULL), 1); mbar_init(bar(KV_EMPTY), 1); mbar_init(bar(QDO_FULL), 1); mbar_init(bar(QDO_EMPTY), 1);
        mbar_init(bar(S_FULL), 1); mbar_init(bar(P_FULL), BWD_CW); mbar_init(bar(DP_FULL), 1); mbar_init(bar(DS_FULL), BWD_CW);
        mbar_init(bar(DQ_FULL), 1); mbar_init(bar(A_EMPTY), BWD_CW); mbar_init(bar(ACC_FULL), 1); mbar_init(bar(ACC_EMPTY), BWD_CW);
        mbar_fence_init();
    }
    if (warp == BWD_CW + 1) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
    // kv columns >= 208 of the P / dS buffers are never written; they only feed accumulator rows that are never read, but
    // keep them finite
    for (int i = threadIdx.x; i < 2 * PBUF_B / 16; i += BWD_THREADS) st_shared_v4(sP + i * 16, make_uint4(0, 0, 0, 0));
    for (int i = threadIdx.x; i < 512; i += BWD_THREADS) cs[i] = 0.f;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tA = tmem, tDK = tmem + 256, tDV = tmem + 384;
    constexpr uint32_t IDESC_S = make_idesc_bf16(128, KVP, 0, 0);     // S, dP : K-major x K-major
    constexpr uint32_t IDESC_DQ = make_idesc_bf16(128, HD, 0, 1);     // dQ     : dS K-major, K MN-major
    constexpr uint32_t IDESC_KV = make_idesc_bf16(128, HD, 1, 1);     // dK, dV : MN-major x MN-major

    if (warp == BWD_CW) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv); tma_prefetch_desc(&tm_do);
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int b = item / a.H, h = item % a.H;
                mbar_wait(bar(KV_EMPTY), (it & 1) ^ 1);
                mbar_arrive_expect_tx(bar(KV_FULL), 2 * KV_B);
                tma_load_5d(sK, &tm_kv, bar(KV_FULL), 0, 0, h, 1, b);
                tma_load_5d(sV, &tm_kv, bar(KV_FULL), 0, 0, h, 2, b);
                for (int i = 0; i < 2; ++i) {
                    mbar_wait(bar(QDO_EMPTY), i ^ 1);            // tile counter 2*it + i -> parity i
                    mbar_arrive_expect_tx(bar(QDO_FULL), 2 * TILE_B);
                    tma_load_5d(sQ, &tm_q, bar(QDO_FULL), 0, i * QT, h, 0, b);
                    tma_load_4d(sDO, &tm_do, bar(QDO_FULL), h * HD, i * QT, b, 0);
                }
            }
        }
    } else if (warp == BWD_CW + 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t pi = it & 1;
                mbar_wait(bar(KV_FULL), pi);
                for (int i = 0; i < 2; ++i) {
                    mbar_wait(bar(QDO_FULL), i);
                    mbar_wait(bar(A_EMPTY), i ^ 1);              // dQ of the previous tile drained from region A
                    tc_fence_after();
                    // S = Q K^T
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_bf16(tA, make_smem_desc_sw128(sQ + k * 32, 0, 1024), make_smem_desc_sw128(sK + k * 32, 0, 1024), IDESC_S, k > 0);
                    umma_commit(bar(S_FULL));
                    mbar_wait(bar(P_FULL), i);
                    tc_fence_after();
                    // dP = dO V^T  (region A again: S fully consumed once P_FULL completes)
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_bf16(tA, make_smem_desc_sw128(sDO + k * 32, 0, 1024), make_smem_desc_sw128(sV + k * 32, 0, 1024), IDESC_S, k > 0);
                    umma_commit(bar(DP_FULL));
                    if (i == 0) { mbar_wait(bar(ACC_EMPTY), pi ^ 1); tc_fence_after(); }   // previous item's dK / dV read out
                    // dV[kv tile m] += P^T dO      (M = kv, K = q rows of this tile)
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int k = 0; k < QT / 16; ++k)
                            umma_bf16(tDV + m * HD, make_smem_desc_sw128(sP + (2 * m) * TILE_B + k * 2048, TILE_B, 1024),
                                      make_smem_desc_sw128(sDO + k * 2048, 0, 1024), IDESC_KV, (i > 0 || k > 0) ? 1u : 0u);
                    mbar_wait(bar(DS_FULL), i);
                    tc_fence_after();
                    // dQ = dS K   (K = kv)
#pragma unroll
                    for (int k = 0; k < KVP / 16; ++k)
                        umma_bf16(tA, make_smem_desc_sw128(sDS + (k >> 2) * TILE_B + (k & 3) * 32, 0, 1024),
                                  make_smem_desc_sw128(sK + k * 2048, 0, 1024), IDESC_DQ, k > 0);
                    umma_commit(bar(DQ_FULL));
                    // dK[kv tile m] += dS^T Q
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int k = 0; k < QT / 16; ++k)
                            umma_bf16(tDK + m * HD, make_smem_desc_sw128(sDS + (2 * m) * TILE_B + k * 2048, TILE_B, 1024),
                                      make_smem_desc_sw128(sQ + k * 2048, 0, 1024), IDESC_KV, (i > 0 || k > 0) ? 1u : 0u);
                    umma_commit(bar(QDO_EMPTY));                 // every MMA reading Q_i / dO_i has been issued before this commit
                    if (i == 1) { umma_commit(bar(KV_EMPTY)); umma_commit(bar(ACC_FULL)); }
                }
            }
        }
    } else {
        // ===================== compute warps =====================
        const int q = warp & 3, cg = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_base = uint32_t(q * 32) << 16;
        const int cbase = cg == 0 ? 0 : 16 + 48 * cg;           // 0, 64, 112, 160
        const int n1 = cg == 0 ? 32 : 16;                        // second piece width
        const uint32_t xd_mine = sXD + (cg * 128 + r) * 4;       // delta partial of this warp's 16 head-dim columns
        const float sl2 = a.scale * 1.4426950408889634f;
        float v[32];
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int b = item / a.H, h = item % a.H;
            const uint32_t pi = it & 1;
            float* cs_gate = cs + pi * 256 + cg * 16;
            float* cs_bias = cs + pi * 256 + 64 + cg * 16;
            const float dps = a.drop_scale != nullptr ? __ldg(a.drop_scale + b) : 1.f;
            const float inv_dps = dps != 0.f ? 1.f / dps : 0.f;
            float2 g2[8], ag[8], ab[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                g2[j] = __ldg(reinterpret_cast<const float2*>(a.gate + h * HD + cg * 16) + j);
                ag[j] = make_float2(0.f, 0.f); ab[j] = make_float2(0.f, 0.f);
            }
            for (int i = 0; i < 2; ++i) {
                const int t = i * QT + r;
                const bool t_ok = t < a.T;
                // ---- early global loads (consumed after the S / dQ MMAs): LSE, this warp's 16-column slices of O and q ----
                float lse = INFINITY;                            // rows >= T: exp2(-inf) = 0 everywhere
                uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0, y0 = o0, y1 = o0;
                const size_t qoff = ((size_t(b) * a.T + (t_ok ? t : 0)) * 3 + 0) * D + h * HD + cg * 16;
                if (t_ok) {
                    lse = __ldg(a.lse + (size_t(b) * a.H + h) * a.T + t);
                    const uint4* po = reinterpret_cast<const uint4*>(a.o_in + (size_t(b) * a.T + t) * D + h * HD + cg * 16);
                    o0 = __ldg(po); o1 = __ldg(po + 1);
                    const uint4* py = reinterpret_cast<const uint4*>(a.qkv + qoff);
                    y0 = __ldg(py); y1 = __ldg(py + 1);
                }
                // ---- delta = rowsum(dO * O) / droppath: dO slice from the TMA-loaded tile, partials exchanged through smem ----
                mbar_wait(bar(QDO_FULL), i);
                {
                    const uint4 d0 = ld_shared_v4(sDO + sw128_offset(r, cg * 2)), d1 = ld_shared_v4(sDO + sw128_offset(r, cg * 2 + 1));
                    st_shared_f32(xd_mine, (dot8(o0, d0) + dot8(o1, d1)) * inv_dps);
                }
                named_bar_sync(2 + q, 128);
                const float delta = (ld_shared_f32(sXD + r * 4) + ld_shared_f32(sXD + (128 + r) * 4)) +
                                    (ld_shared_f32(sXD + (256 + r) * 4) + ld_shared_f32(sXD + (384 + r) * 4));
                const float2 sl22 = splat2(sl2), nlse2 = splat2(-lse * 1.4426950408889634f);
                const float2 sc2 = splat2(a.scale), ndsc2 = splat2(-delta * a.scale);
                // ---- P = exp(S*scale - LSE) over this warp's kv columns ----
                mbar_wait(bar(S_FULL), i);
                tc_fence_after();
#pragma unroll
                for (int pc = 0; pc < 2; ++pc) {
                    const int col = cbase + pc * 32;
                    const int nj = pc == 0 ? 32 : n1;
                    if (nj == 32) tmem_ld32(tA + lane_base + col, v);
                    else tmem_ld16(tA + lane_base + col, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 e = fma2(make_float2(v[2 * j], v[2 * j + 1]), sl22, nlse2);
                        v[2 * j] = fast_ex2(e.x); v[2 * j + 1] = fast_ex2(e.y);
                    }
                    if (col + nj > a.T) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = (col + j < a.T) ? v[j] : 0.f;
                    }
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        if (q4 * 8 < nj) {
                            uint4 pk;
                            pk.x = pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]); pk.y = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                            pk.z = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); pk.w = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                            st_shared_v4(pbuf_addr(sP, r, (col >> 3) + q4), pk);
                        }
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(P_FULL));
                // ---- dS = scale * P * (dP - delta) ----
                mbar_wait(bar(DP_FULL), i);
                tc_fence_after();
#pragma unroll
                for (int pc = 0; pc < 2; ++pc) {
                    const int col = cbase + pc * 32;
                    const int nj = pc == 0 ? 32 : n1;
                    if (nj == 32) tmem_ld32(tA + lane_base + col, v);
                    else tmem_ld16(tA + lane_base + col, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        if (q4 * 8 < nj) {
                            const uint4 pp = ld_shared_v4(pbuf_addr(sP, r, (col >> 3) + q4));
                            const float* w = v + q4 * 8;
                            const float2 d0 = mul2(unpack_bf16x2(pp.x), fma2(make_float2(w[0], w[1]), sc2, ndsc2));
                            const float2 d1 = mul2(unpack_bf16x2(pp.y), fma2(make_float2(w[2], w[3]), sc2, ndsc2));
                            const float2 d2 = mul2(unpack_bf16x2(pp.z), fma2(make_float2(w[4], w[5]), sc2, ndsc2));
                            const float2 d3 = mul2(unpack_bf16x2(pp.w), fma2(make_float2(w[6], w[7]), sc2, ndsc2));
                            uint4 pk;
                            pk.x = pack_bf16x2(d0.x, d0.y); pk.y = pack_bf16x2(d1.x, d1.y);
                            pk.z = pack_bf16x2(d2.x, d2.y); pk.w = pack_bf16x2(d3.x, d3.y);
                            st_shared_v4(pbuf_addr(sDS, r, (col >> 3) + q4), pk);
                        }
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(DS_FULL));
                // ---- dQ epilogue: head-dim columns [16 cg, 16 cg + 16) ----
                mbar_wait(bar(DQ_FULL), i);
                tc_fence_after();
                float w16[16];
                tmem_ld16(tA + lane_base + cg * 16, w16);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(A_EMPTY));
                dqkv_slice_acc(w16, t_ok, y0, y1, a.dqkv + qoff, g2, ag, ab);
            }
            reduce_cols16(ab, cs_bias);                      // d bias of q
            // ---- dK / dV epilogue (kv rows: tile m covers kv = m*128 + r); the gated k / v slices are prefetched one ahead ----
            auto slice_off = [&](int idx) {
                const int which = 1 + (idx >> 1), kv = (idx & 1) * QT + r;
                return ((size_t(b) * a.T + (kv < a.T ? kv : 0)) * 3 + which) * D + h * HD + cg * 16;
            };
            uint4 n0 = __ldg(reinterpret_cast<const uint4*>(a.qkv + slice_off(0))),
                  n1v = __ldg(reinterpret_cast<const uint4*>(a.qkv + slice_off(0)) + 1);
            mbar_wait(bar(ACC_FULL), pi);
            tc_fence_after();
#pragma unroll
            for (int idx = 0; idx < 4; ++idx) {
                const int which = 1 + (idx >> 1), m = idx & 1;
                const bool kv_ok = m * QT + r < a.T;
                const uint4 x0 = n0, x1 = n1v;
                if (idx < 3) {
                    n0 = __ldg(reinterpret_cast<const uint4*>(a.qkv + slice_off(idx + 1)));
                    n1v = __ldg(reinterpret_cast<const uint4*>(a.qkv + slice_off(idx + 1)) + 1);
                }
                float w16[16];
                tmem_ld16((which == 1 ? tDK : tDV) + lane_base + m * HD + cg * 16, w16);
                tmem_ld_wait();
                if (idx == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(ACC_EMPTY));
                }
                dqkv_slice_acc(w16, kv_ok, x0, x1, a.dqkv + slice_off(idx), g2, ag, ab);
                if (m == 1) reduce_cols16(ab, cs_bias + which * 64);     // d bias of k, then of v
            }
            reduce_cols16(ag, cs_gate);                      // d gate: q, k and v contributions together
            // ---- per-item column sums -> global partials; the accumulator of this parity is re-zeroed for item it+2 ----
            named_bar_sync(1, BWD_CT);
            {
                const int tid = threadIdx.x;   // 0..511
                if (tid < 256) {
                    float* src = cs + pi * 256 + tid;
                    const float val = *src;
                    *src = 0.f;
                    if (tid < 64) a.part_gate[size_t(b) * D + h * HD + tid] = val;
                    else a.part_bias[size_t(b) * 3 * D + ((tid - 64) / 64) * D + h * HD + ((tid - 64) % 64)] = val;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == BWD_CW + 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static constexpr int FWD_SMEM = 2 * TILE_B + 2 * KV_B + 2 * PBUF_B + 4096 + 256;
static constexpr int BWD_SMEM = 2 * TILE_B + 2 * KV_B + 2 * PBUF_B + 4096 + 256;

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
    const int items = B * H;
    const int grid = items < num_sms() ? items : num_sms();
    attn_fwd_kernel<<<grid, FWD_THREADS, FWD_SMEM, s>>>(tq, tkv, a);
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
    const int items = B * H;
    const int grid = items < num_sms() ? items : num_sms();
    attn_bwd_kernel<<<grid, BWD_THREADS, BWD_SMEM, s>>>(tq, tkv, tdo, a);
    return int(cudaGetLastError());
}

}  // namespace ofb
