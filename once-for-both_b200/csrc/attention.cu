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
#include "launch.cuh"
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
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                const __grid_constant__ CUtensorMap tm_o, const AttnArgs a) {
    pdl_trigger();
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
    pdl_wait();
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
        // the whole warp runs the loop (uniform operands stay in uniform registers); one elected lane issues the tcgen05 instructions
        {
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                mbar_wait(qk_full, ph);
                for (int s = 0; s < 2; ++s) {
                    mbar_wait(slot_bar(s, 3), ph ^ 1);          // slot's TMEM free (previous item's O read out)
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)
                            umma_bf16(tmem + s * 256, make_smem_desc_sw128(sQ + s * TILE_B + k * 32, 0, 1024),
                                      make_smem_desc_sw128(sK + k * 32, 0, 1024), IDESC_S, k > 0);
                        umma_commit(slot_bar(s, 0));
                        if (s == 1) umma_commit(qk_empty);      // Q / K may be overwritten by the next item's loads
                    }
                    __syncwarp();
                }
                mbar_wait(v_full, ph);
                for (int s = 0; s < 2; ++s) {
                    mbar_wait(slot_bar(s, 1), ph);              // P of this slot is in shared memory, S fully consumed
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < KVP / 16; ++k)
                            umma_bf16(tmem + s * 256, make_smem_desc_sw128(sP + s * PBUF_B + (k >> 2) * TILE_B + (k & 3) * 32, 0, 1024),
                                      make_smem_desc_sw128(sV + k * 2048, 0, 1024), IDESC_O, k > 0);
                        umma_commit(slot_bar(s, 2));
                        if (s == 1) umma_commit(v_empty);
                    }
                    __syncwarp();
                }
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
            // Softmax in ONE pass over the S accumulator: P = exp2(S * scale * log2e) without subtracting the row maximum. The
            // result of softmax does not depend on the reference point, bf16 / fp32 keep their relative precision at any
            // magnitude, and attention logits stay far inside the fp32 exponent range (|S * scale| < 88); TMEM reads run at
            // 64 B/clk per SM, so the separate maximum pass over the 128 x 208 fp32 tile cost as much as the softmax itself.
            // Guard: a row whose sum left [1e-30, 1e30] (logits beyond +-69) sends BOTH warps of its row quarter through the
            // classic two-pass path below (the vote is taken on the combined sum, identical in the two warps).
            float mx = 0.f, sum = 0.f;
            bool two_pass = false;
            while (true) {
                if (two_pass) {
                    // ---- row maximum over this warp's columns, exchanged with the warp of the other kv half ----
                    mx = -INFINITY;
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
                }
                const float mxs = mx * sl2;
                // ---- P = exp(S*scale - reference), row sums, P -> shared memory (bf16, swizzled K-major A operand) ----
                // (the first 64 kv columns of these rows double as the O staging slab of the previous item: its bulk store, issued
                //  by lane 0 of this very warp when hf == 0, must have read the slab before P overwrites it)
                if (hf == 0 && !two_pass) {
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                }
                sum = 0.f;
#pragma unroll 1
                for (int c = 0; c < n32 + 1; ++c) {
                    const int col = c_begin + c * 32;
                    const int nj = c < n32 ? 32 : (hf == 0 ? 16 : 0);
                    if (nj == 0) break;
                    if (nj == 32) tmem_ld32(tS + col, v);
                    else tmem_ld16(tS + col, v);
                    tmem_ld_wait();
                    // packed fp32x2 scale / row-sum arithmetic (half the issue slots of the scalar form: 60.3 -> 57.2 us per launch)
                    const float2 sl22 = splat2(sl2), nm2 = splat2(-mxs);
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        if (2 * k < nj) {
                            const float2 e = fma2(make_float2(v[2 * k], v[2 * k + 1]), sl22, nm2);
                            v[2 * k] = fast_ex2(e.x); v[2 * k + 1] = fast_ex2(e.y);
                        }
                    }
                    if (col + nj > a.T) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nj) v[j] = col + j < a.T ? v[j] : 0.f;
                    }
                    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (2 * k < nj) s2 = add2(s2, make_float2(v[2 * k], v[2 * k + 1]));
                    sum += s2.x + s2.y;
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
                named_bar_sync(pair_bar, 64);
                sum += ld_shared_f32(x_other + 2048);
                if (two_pass) break;
                const bool bad = !(sum > 1e-30f && sum < 1e30f) && t < a.T;
                if (!__any_sync(0xffffffffu, bad)) break;
                two_pass = true;
                named_bar_sync(pair_bar, 64);        // both warps have read the other's sum before the exchange slots are reused
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(slot_bar(s, 1));
            // ---- epilogue: O / rowsum * droppath; this warp owns head-dim columns [32 hf, 32 hf + 32) ----
            mbar_wait(slot_bar(s, 2), ph);
            tc_fence_after();
            tmem_ld32(tS + hf * 32, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(slot_bar(s, 3));
            // O rows leave through shared memory: the two warps of a row quarter fill a [32 rows][128 B] swizzled slab (in the P
            // buffer of the slot, free now that the P V MMA has retired) and one bulk tensor store writes it, clipping rows >= T -
            // a 64-byte global store per lane costs 32 LSU wavefronts per instruction
            {
                const float inv = dps / sum;
                const uint32_t slab = pS + q * 4096 + lane * 128;
                const uint32_t sw = lane & 7u;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 pk;
                    pk.x = pack_bf16x2(v[q4 * 8 + 0] * inv, v[q4 * 8 + 1] * inv); pk.y = pack_bf16x2(v[q4 * 8 + 2] * inv, v[q4 * 8 + 3] * inv);
                    pk.z = pack_bf16x2(v[q4 * 8 + 4] * inv, v[q4 * 8 + 5] * inv); pk.w = pack_bf16x2(v[q4 * 8 + 6] * inv, v[q4 * 8 + 7] * inv);
                    st_shared_v4(slab + (((hf * 4 + q4) ^ sw) << 4), pk);
                }
                if (t < a.T && hf == 0) a.lse[(size_t(b) * a.H + h) * a.T + t] = mx * a.scale + logf(sum);
                fence_proxy_async_smem();
                named_bar_sync(pair_bar, 64);
                if (hf == 0 && lane == 0 && s * QT + q * 32 < a.T) {
                    tma_store_4d(&tm_o, pS + q * 4096, h * HD, s * QT + q * 32, b, 0);
                    bulk_commit();
                }
            }
        }
        if (hf == 0 && lane == 0) bulk_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FWD_CW + 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------
// backward: persistent, warp-specialised; one CTA per SM loops over (image, head) items. An item is cut into 128 x 128
// sub-tiles (q tile i, kv half j), j outer:
//     S_ij = Q_i K_j^T -> P_ij = exp(S*scale - LSE)          dP_ij = dO_i V_j^T -> dS_ij = scale * P (dP - delta)
//     dV_j += P_ij^T dO_i      dK_j += dS_ij^T Q_i      dQ_i += dS_ij K_j
// so that S / dP need only 128 TMEM columns and TWO sub-tiles can be in flight: the S MMA of sub-tile t+1 is issued into the
// other TMEM region while the compute warps are still busy with sub-tile t (the one-tile-at-a-time version was bound by the
// MMA -> mbarrier -> compute round trips, not by instruction issue: profiles/r01). dS overwrites P in place (dV is issued
// before dP, so the dP commit also covers the MMA that reads P), which frees the shared memory to keep Q and dO of both q
// tiles resident for the whole item.
//   warps 0-15: compute.  warp w: TMEM lane quarter w&3 (q rows), column group w>>2 (32 kv columns of S / dP; 16 head-dim
//               columns of dQ / dK / dV). Every phase is a chain of long-latency steps (tcgen05.ld, MUFU, shared memory),
//               so the phases are spread over many warps rather than over long per-thread loops.
//   warp 16   : TMA producer        warp 17 : MMA issuer (+ TMEM allocation)
// TMEM columns: R0 [0,128) R1 [128,256) S then dP of even / odd sub-tiles ; dQ_0 [256,320) dQ_1 [320,384) ; dK_j [384,448) ;
//               dV_j [448,512)
// ---------------------------------------------------------------------------------------------
static constexpr int BWD_CW = 16;
static constexpr int BWD_NCG = BWD_CW / 4;                  // column groups
static constexpr int BWD_PC = 128 / BWD_NCG;              // kv columns of a sub-tile per warp (32)
static constexpr int BWD_EC = HD / BWD_NCG;               // head-dim columns per warp in the epilogues (16)
static constexpr int BWD_THREADS = (BWD_CW + 2) * 32;     // 576
static constexpr int BWD_CT = BWD_CW * 32;                // compute threads
static constexpr int PB_B = 2 * TILE_B;                   // 32 KB: one P / dS buffer = 2 atoms of [128 q rows][64 kv columns]

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
// gate products are accumulated in registers (ag, reduced once per item), the bias column sums are reduced across the 32
// rows of the warp right away and accumulated in a register. x0/x1 = the row's 16 gated q/k/v values.
// AGW: weight of this slice's sum_rows x * dx in the d gate accumulator. q, k and v share one gate; since S = scale q k^T is
// bilinear, sum_t q[t,c] dq[t,c] == sum_kv k[kv,c] dk[kv,c] (both are sum_{t,kv} q[t,c] dS[t,kv] k[kv,c]), so the k slices
// carry weight 2 and the q slices none - the dQ epilogue then needs no operand rows at all.
template <int AGW>
__device__ __forceinline__ void dqkv_slice16(float (&v)[16], bool ok, const uint4& x0, const uint4& x1, uint32_t stage_row,
                                             uint32_t chunk, const float* gate16, float2 (&ag)[8], float& bias_acc) {
    // stage_row: shared-memory address of this token row inside a [32 rows][128 B] SWIZZLE_128B slab (rows of one TMEM lane
    // quarter); the slab leaves through one TMA store, which clips rows >= T, so rows that are not `ok` may hold anything
    if (ok) {
        const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        uint32_t ow[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 vv = make_float2(v[2 * j], v[2 * j + 1]);
            if (AGW == 1) ag[j] = fma2(unpack_bf16x2(xw[j]), vv, ag[j]);
            if (AGW == 2) ag[j] = fma2(unpack_bf16x2(xw[j]), add2(vv, vv), ag[j]);
            const float2 o = mul2(vv, __ldg(reinterpret_cast<const float2*>(gate16) + j));
            v[2 * j] = o.x; v[2 * j + 1] = o.y;
            ow[j] = pack_bf16x2(o.x, o.y);
        }
        const uint32_t sw = lane_id() & 7u;
        st_shared_v4(stage_row + (((chunk) ^ sw) << 4), make_uint4(ow[0], ow[1], ow[2], ow[3]));
        st_shared_v4(stage_row + (((chunk + 1) ^ sw) << 4), make_uint4(ow[4], ow[5], ow[6], ow[7]));
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
    // column sums over the 32 rows of the warp (even lanes: column lane >> 1), accumulated in a register over the item's slices
    // and combined across warps in FIXED order at the end of the item (shared-memory atomics made d bias / d gate differ from run
    // to run)
    bias_acc += bfly16(v);
}
__device__ __forceinline__ float dot8(const uint4& x, const uint4& y) {
    float2 d = mul2(unpack_bf16x2(x.x), unpack_bf16x2(y.x));
    d = fma2(unpack_bf16x2(x.y), unpack_bf16x2(y.y), d);
    d = fma2(unpack_bf16x2(x.z), unpack_bf16x2(y.z), d);
    d = fma2(unpack_bf16x2(x.w), unpack_bf16x2(y.w), d);
    return d.x + d.y;
}

// Optional phase trace of the backward kernel (tools/attn_trace.py; compiled only with -DOFB_ATTN_TRACE into a separate debug
// library): SM clock stamps of compute warps 0 / 15 and of the MMA warp of CTA 0 for items 2..5, plus accumulated barrier waits.
#ifdef OFB_ATTN_TRACE
__device__ long long g_attn_trace[3 * 4 * 32];
#define TRC_ON(role) (blockIdx.x == 0 && lane == 0 && it >= 2 && it < 6 && (role) >= 0)
#define TRC(role, slot) do { if (TRC_ON(role)) g_attn_trace[((role) * 4 + (it - 2)) * 32 + (slot)] = clock64(); } while (0)
#define TRC_ACC(role, slot, t0) do { if (TRC_ON(role)) g_attn_trace[((role) * 4 + (it - 2)) * 32 + (slot)] += clock64() - (t0); } while (0)
#define TRC_NOW() clock64()
#else
#define TRC(role, slot) do { } while (0)
#define TRC_ACC(role, slot, t0) do { } while (0)
#define TRC_NOW() 0ll
#endif

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_out,
                const __grid_constant__ CUtensorMap tm_o, const AttnArgs a) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);                   // 2 q tiles
    const uint32_t sDO = sQ + 2 * TILE_B;                 // 2 q tiles
    const uint32_t sK = sDO + 2 * TILE_B;
    const uint32_t sV = sK + KV_B;
    const uint32_t sP = sV + KV_B;                        // 2 buffers (even / odd sub-tiles), P then dS in place
    const uint32_t sST = sP + 2 * PB_B;                   // 2 staging tiles [128 rows][64 bf16] of the d qkv epilogues (TMA stores)
    uint8_t* tail = smem + 4 * TILE_B + 2 * KV_B + 2 * PB_B + 2 * TILE_B;
    float* cs = reinterpret_cast<float*>(tail);                 // [2 parities][16 warps][4: gate | bias q,k,v][16 columns] per-warp partials
    const uint32_t sXD = smem_u32(tail + 8192);                 // [2 q tiles][4 column groups][128 rows] delta partials
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 8192 + 4096);
    enum { QK_FULL = 0, DOV_FULL, ITEM_EMPTY, S_FULL, P_FULL = S_FULL + 2, DP_FULL = P_FULL + 2, DS_FULL = DP_FULL + 2,
           PB_FREE = DS_FULL + 2, ACC_FULL = PB_FREE + 2, ACC_EMPTY, DQ_FULL, DQ_EMPTY, XREAD, NBAR };
    auto bar = [&](int k) { return smem_u32(&bars[k]); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[NBAR]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.B * a.H;
    const int D = a.H * HD;
    const int kvp = (a.T + 15) & ~15;                     // kv columns, padded to the MMA granularity
    const int nh = a.T > QT ? 2 : 1;                      // q tiles = kv halves
    const int nt = nh * nh;                               // sub-tiles per item
    const int n_kv1 = kvp - QT;                           // kv columns of half 1 (if any)

    if (threadIdx.x == 0) {
        if ((sQ & 1023u) != 0) { printf("ofb: dynamic smem base not 1024-aligned\n"); __trap(); }
        mbar_init(bar(QK_FULL), 1); mbar_init(bar(DOV_FULL), 1); mbar_init(bar(ITEM_EMPTY), 1);
        for (int s2 = 0; s2 < 2; ++s2) {
            mbar_init(bar(S_FULL + s2), 1); mbar_init(bar(P_FULL + s2), BWD_CW); mbar_init(bar(DP_FULL + s2), 1);
            mbar_init(bar(DS_FULL + s2), BWD_CW); mbar_init(bar(PB_FREE + s2), 1);
        }
        mbar_init(bar(ACC_FULL), 1); mbar_init(bar(ACC_EMPTY), BWD_CW); mbar_init(bar(DQ_FULL), 1); mbar_init(bar(DQ_EMPTY), BWD_CW);
        mbar_init(bar(XREAD), BWD_CW);
        mbar_fence_init();
    }
    if (warp == BWD_CW + 1) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
    // P / dS buffers start as zeros: the chunks of q rows >= T are skipped by the compute warps for the whole kernel
    for (int i = threadIdx.x; i < 2 * PB_B / 16; i += BWD_THREADS) st_shared_v4(sP + i * 16, make_uint4(0, 0, 0, 0));
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();
    const uint32_t tDQ = tmem + 256, tDK = tmem + 384, tDV = tmem + 448;
    constexpr uint32_t IDESC_DQ = make_idesc_bf16(128, HD, 0, 1);     // dQ     : dS K-major, K MN-major
    constexpr uint32_t IDESC_KV = make_idesc_bf16(128, HD, 1, 1);     // dK, dV : MN-major x MN-major

    if (warp == BWD_CW) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_out);
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int b = item / a.H, h = item % a.H;
                mbar_wait(bar(ITEM_EMPTY), (it & 1) ^ 1);        // every MMA of the previous item has retired
                mbar_wait(bar(XREAD), (it & 1) ^ 1);             // ... and its epilogues have read their q / k / v rows out of the tiles
                mbar_arrive_expect_tx(bar(QK_FULL), nh * TILE_B + KV_B);
                tma_load_5d(sQ, &tm_q, bar(QK_FULL), 0, 0, h, 0, b);
                tma_load_5d(sK, &tm_kv, bar(QK_FULL), 0, 0, h, 1, b);
                if (nh == 2) tma_load_5d(sQ + TILE_B, &tm_q, bar(QK_FULL), 0, QT, h, 0, b);
                mbar_arrive_expect_tx(bar(DOV_FULL), nh * TILE_B + KV_B);
                tma_load_4d(sDO, &tm_do, bar(DOV_FULL), h * HD, 0, b, 0);
                tma_load_5d(sV, &tm_kv, bar(DOV_FULL), 0, 0, h, 2, b);
                if (nh == 2) tma_load_4d(sDO + TILE_B, &tm_do, bar(DOV_FULL), h * HD, QT, b, 0);
                // pull the next item's tiles into L2 while this one is being computed
                const int nxt = item + gridDim.x;
                if (nxt < n_items) {
                    const int b2 = nxt / a.H, h2 = nxt % a.H;
                    tma_prefetch_5d(&tm_kv, 0, 0, h2, 1, b2);
                    tma_prefetch_5d(&tm_kv, 0, 0, h2, 2, b2);
                    for (int i = 0; i < nh; ++i) {
                        tma_prefetch_5d(&tm_q, 0, i * QT, h2, 0, b2);
                        tma_prefetch_4d(&tm_do, h2 * HD, i * QT, b2, 0);
                        // the forward output rows and the LSE feed the per-row statistics at the very start of the item, read by
                        // plain loads: without this they come from HBM and every compute warp idles ~4k clocks per item
                        tma_prefetch_4d(&tm_o, h2 * HD, i * QT, b2, 0);
                    }
                    const char* lse2 = reinterpret_cast<const char*>(a.lse + (size_t(b2) * a.H + h2) * a.T);
                    for (int off = 0; off < a.T * 4 + 127; off += 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(lse2 + off));
                }
            }
        }
    } else if (warp == BWD_CW + 1) {
        // ===================== MMA issuer =====================
        // Sub-tiles are processed in pairs (t0, t0+1) = the two q tiles of one kv half; the compute warps run
        //   P(t0) P(t0+1) dS(t0) dS(t0+1)
        // so that every MMA -> mbarrier -> compute round trip of one sub-tile is covered by compute work on the other one.
        // The whole warp runs the loop (uniform operands stay in uniform registers: a `lane == 0` region made the compiler
        // broadcast every descriptor through an R2UR waterfall in front of each UTCHMMA, ~130 MMAs per item); one elected lane
        // issues the tcgen05 instructions.
        {
            uint32_t use0 = 0, use1 = 0;       // completed uses of the per-slot barriers (slot = sub-tile parity)
            uint32_t n_acc = 0;                // kv halves started (ACC_EMPTY waits)
            const uint32_t n_kv0 = uint32_t(min(kvp, QT));
            auto issue_s = [&](int t) {        // (elected lane only)
                const int j = (nh == 2) ? (t >> 1) : 0, i = (nh == 2) ? (t & 1) : 0;
                const uint32_t idesc = make_idesc_bf16(128, j == 0 ? n_kv0 : uint32_t(n_kv1), 0, 0);
                const uint32_t qa = sQ + i * TILE_B, kb = sK + j * (QT * 128);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(tmem + (t & 1) * 128, make_smem_desc_sw128(qa + k * 32, 0, 1024),
                              make_smem_desc_sw128(kb + k * 32, 0, 1024), idesc, k > 0);
                umma_commit(bar(S_FULL + (t & 1)));
            };
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t pi = it & 1;
                TRC(2, 0);
                mbar_wait(bar(QK_FULL), pi);
                TRC(2, 1);
                tc_fence_after();
                if (elect_one()) {
                    issue_s(0);
                    if (nt > 1) issue_s(1);
                }
                __syncwarp();
#pragma unroll 1
                for (int t0 = 0; t0 < nt; t0 += 2) {
                    const int np = min(2, nt - t0);
                    const int j = (nh == 2) ? (t0 >> 1) : 0;
                    const uint32_t nj = (j == 0) ? n_kv0 : uint32_t(n_kv1);
                    const uint32_t vb = sV + j * (QT * 128), kb = sK + j * (QT * 128);
                    // ---- dV_j (+)= P^T dO_i, then dP = dO_i V_j^T, as soon as P of the sub-tile is in shared memory ----
#pragma unroll 1
                    for (int i = 0; i < np; ++i) {
                        const uint32_t par = (i == 0 ? use0 : use1) & 1u;
                        const uint32_t pbuf = sP + i * PB_B, doa = sDO + i * TILE_B;
                        { const long long tw = TRC_NOW(); mbar_wait(bar(P_FULL + i), par); TRC_ACC(2, 8, tw); }
                        if (t0 == 0 && i == 0) mbar_wait(bar(DOV_FULL), pi);
                        if (i == 0) { const long long tw = TRC_NOW(); mbar_wait(bar(ACC_EMPTY), (n_acc & 1u) ^ 1u); ++n_acc; TRC_ACC(2, 9, tw); }   // previous dK / dV read out
                        tc_fence_after();
                        if (elect_one()) {
                            // reduction over the q rows of tile i: k steps whose 16 rows all lie past T would add 0 x 0
                            const int kq = (min(QT, a.T - i * QT) + 15) >> 4;
                            for (int k = 0; k < kq; ++k)
                                umma_bf16(tDV, make_smem_desc_sw128(pbuf + k * 2048, TILE_B, 1024),
                                          make_smem_desc_sw128(doa + k * 2048, 0, 1024), IDESC_KV, (i > 0 || k > 0) ? 1u : 0u);
                            const uint32_t idesc = make_idesc_bf16(128, nj, 0, 0);
#pragma unroll
                            for (int k = 0; k < HD / 16; ++k)
                                umma_bf16(tmem + i * 128, make_smem_desc_sw128(doa + k * 32, 0, 1024),
                                          make_smem_desc_sw128(vb + k * 32, 0, 1024), idesc, k > 0);
                            umma_commit(bar(DP_FULL + i));
                        }
                        __syncwarp();
                    }
                    // ---- once dS of a sub-tile is in place: S of the sub-tile two ahead (its TMEM region is free now), then
                    //      dQ_i (+)= dS K_j and dK_j (+)= dS^T Q_i ----
#pragma unroll 1
                    for (int i = 0; i < np; ++i) {
                        const uint32_t par = (i == 0 ? use0 : use1) & 1u;
                        if (i == 0) ++use0; else ++use1;
                        const uint32_t pbuf = sP + i * PB_B, qa = sQ + i * TILE_B;
                        { const long long tw = TRC_NOW(); mbar_wait(bar(DS_FULL + i), par); TRC_ACC(2, 10, tw); }
                        if (t0 == 0 && i == 0) { const long long tw = TRC_NOW(); mbar_wait(bar(DQ_EMPTY), pi ^ 1u); TRC_ACC(2, 11, tw); }       // previous item's dQ read out
                        tc_fence_after();
                        if (elect_one()) {
                            if (t0 + 2 + i < nt) issue_s(t0 + 2 + i);
                            for (uint32_t k = 0; k < nj / 16; ++k)
                                umma_bf16(tDQ + i * HD, make_smem_desc_sw128(pbuf + (k >> 2) * TILE_B + (k & 3) * 32, 0, 1024),
                                          make_smem_desc_sw128(kb + k * 2048, 0, 1024), IDESC_DQ, (j > 0 || k > 0) ? 1u : 0u);
                            const int kq = (min(QT, a.T - i * QT) + 15) >> 4;
                            for (int k = 0; k < kq; ++k)
                                umma_bf16(tDK, make_smem_desc_sw128(pbuf + k * 2048, TILE_B, 1024),
                                          make_smem_desc_sw128(qa + k * 2048, 0, 1024), IDESC_KV, (i > 0 || k > 0) ? 1u : 0u);
                            umma_commit(bar(PB_FREE + i));
                            if (i == np - 1) {
                                umma_commit(bar(ACC_FULL));
                                if (t0 + 2 >= nt) { umma_commit(bar(DQ_FULL)); umma_commit(bar(ITEM_EMPTY)); }
                            }
                        }
                        __syncwarp();
                    }
                }
                TRC(2, 2);
            }
        }
    } else {
        // ===================== compute warps =====================
        const int q = warp & 3, cg = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_base = uint32_t(q * 32) << 16;
        const float sl2 = a.scale * 1.4426950408889634f;
        const float2 sl22 = splat2(sl2), sc2 = splat2(a.scale);
        float v[32];
        uint32_t use0 = 0, use1 = 0;
        uint32_t n_acc = 0;
        // per-row scalars of an item (LSE of this thread's two q rows, DropPath multiplier of the image): fetched one item ahead,
        // inside the previous item's wait for its last MMAs - read at the item start they were two serialised global-load
        // latencies (~3k clocks) that every compute warp sat through
        auto row_scalars = [&](int item, float& l0, float& l1, float& dp) {
            const int b = item / a.H, h = item % a.H;
            const float* lp = a.lse + (size_t(b) * a.H + h) * a.T;
            l0 = __ldg(lp + min(r, a.T - 1));
            l1 = __ldg(lp + min(QT + r, a.T - 1));
            dp = a.drop_scale != nullptr ? __ldg(a.drop_scale + b) : 1.f;
        };
        float lse_n0 = 0.f, lse_n1 = 0.f, dps_n = 1.f;
        if (int(blockIdx.x) < n_items) row_scalars(blockIdx.x, lse_n0, lse_n1, dps_n);
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int b = item / a.H, h = item % a.H;
            const uint32_t pi = it & 1;
#ifdef OFB_ATTN_TRACE
            const int trole = warp == 0 ? 0 : (warp == 15 ? 1 : -1);
#endif
            TRC(trole, 0);
            float* cs_warp = cs + (pi * BWD_CW + warp) * 64;     // this warp's [gate | bias q | bias k | bias v][16] partial sums
            float bs_q = 0.f, bs_k = 0.f, bs_v = 0.f;            // bias column sums of this warp's rows (even lanes)
            const float dps = dps_n;
            const float inv_dps = dps != 0.f ? 1.f / dps : 0.f;
            const float* gate16 = a.gate + h * HD + cg * BWD_EC;
            float2 ag[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) ag[j] = make_float2(0.f, 0.f);
            // ---- per-row statistics of both q tiles: LSE and delta = rowsum(dO * O) / droppath ----
            float nlse2_0 = -INFINITY, nlse2_1 = -INFINITY, ndsc_0 = 0.f, ndsc_1 = 0.f;   // rows >= T: exp2(-inf) = 0 everywhere
            // delta: 8 lanes share a row (one 16-byte chunk each), so a warp instruction reads 4 whole 128-byte rows of O - with one
            // row per lane (the TMEM mapping used everywhere else) every load instruction cost 32 LSU wavefronts. Thread tid -> row
            // (tid >> 3) + 64 * pass, chunk tid & 7. The loads are issued here and consumed only after the first P phases
            // (stats_finish): delta is not needed before the first dS phase, and the O rows take ~2.5k clocks to arrive.
            const int tid = threadIdx.x, ch = tid & 7;
            uint4 ov[2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int t = i * QT + r;
                if (i < nh && t < a.T) {
                    const float l2 = -(i == 0 ? lse_n0 : lse_n1) * 1.4426950408889634f;
                    if (i == 0) nlse2_0 = l2; else nlse2_1 = l2;
                }
#pragma unroll
                for (int ps = 0; ps < 2; ++ps) {
                    const int t2 = i * QT + ps * 64 + (tid >> 3);
                    ov[i][ps] = make_uint4(0, 0, 0, 0);
                    if (i < nh && t2 < a.T)
                        ov[i][ps] = __ldg(reinterpret_cast<const uint4*>(a.o_in + (size_t(b) * a.T + t2) * D + h * HD) + ch);
                }
            }
            TRC(trole, 15);
            TRC(trole, 1);
            auto stats_finish = [&]() {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int ps = 0; ps < 2; ++ps)   // pins the unpacking of the loaded words below this point (it is pure ALU work
                                                     // and would otherwise be hoisted above the P phases, stalling them on the loads)
                        asm volatile("" : "+r"(ov[i][ps].x), "+r"(ov[i][ps].y), "+r"(ov[i][ps].z), "+r"(ov[i][ps].w));
                mbar_wait(bar(DOV_FULL), pi);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i < nh) {
#pragma unroll
                        for (int ps = 0; ps < 2; ++ps) {
                            const int row = ps * 64 + (tid >> 3);
                            float d = dot8(ov[i][ps], ld_shared_v4(sDO + i * TILE_B + sw128_offset(row, ch)));
                            d += __shfl_xor_sync(0xffffffffu, d, 1);
                            d += __shfl_xor_sync(0xffffffffu, d, 2);
                            d += __shfl_xor_sync(0xffffffffu, d, 4);
                            if (ch == 0) st_shared_f32(sXD + (i * 128 + row) * 4, d * inv_dps);
                        }
                    }
                }
                named_bar_sync(1, BWD_CT);                       // rows are spread over all compute warps here
                ndsc_0 = -ld_shared_f32(sXD + r * 4) * a.scale;
                if (nh == 2) ndsc_1 = -ld_shared_f32(sXD + (128 + r) * 4) * a.scale;
            };
            // ---- P = exp(S*scale - LSE) of sub-tile (q tile i, kv half j) over this warp's 32 kv columns ----
            auto phase_p = [&](int i, int j) {
                const uint32_t par = (i == 0 ? use0 : use1) & 1u;
                const uint32_t tR = tmem + i * 128 + lane_base;
                const uint32_t pbuf = sP + i * PB_B;
                const float2 nl2 = splat2(i == 0 ? nlse2_0 : nlse2_1);
                { const long long tw = TRC_NOW(); mbar_wait(bar(S_FULL + i), par); TRC_ACC(trole, 20, tw); }
                { const long long tw = TRC_NOW(); mbar_wait(bar(PB_FREE + i), par ^ 1u); TRC_ACC(trole, 21, tw); }   // the MMAs that read this buffer one pair ago have retired
                tc_fence_after();
                // a chunk whose 32 q rows or 32 kv columns all lie past T is never read where it matters (its P / dS entries only
                // reach clipped output rows, or multiply zero-filled dO / Q rows - the P buffers are zeroed once at kernel start so
                // that those products are 0 x 0): 15 of the 64 chunks of an item at T = 197. Skipping them takes their TMEM reads
                // and exponentials off the column groups' schedulers.
                if (j * QT + cg * BWD_PC < a.T && i * QT + q * 32 < a.T) {
                    const int col = cg * BWD_PC;                  // column inside the sub-tile
                    const int kv0 = j * QT + col;                 // kv index of the chunk
                    tmem_ld32(tR + col, v);
                    tmem_ld_wait();
                    // (sending part of the exponentials through a polynomial on the fma / alu pipes instead of MUFU changes nothing: the
                    // phase follows the 64 B/clk tcgen05.ld rate, profiles/r02b_attention.txt)
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float2 e = fma2(make_float2(v[2 * k], v[2 * k + 1]), sl22, nl2);
                        v[2 * k] = fast_ex2(e.x); v[2 * k + 1] = fast_ex2(e.y);
                    }
                    if (kv0 + 32 > a.T) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) v[k] = (kv0 + k < a.T) ? v[k] : 0.f;
                    }
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        uint4 pk;
                        pk.x = pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]); pk.y = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                        pk.z = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); pk.w = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                        st_shared_v4(pbuf_addr(pbuf, r, (col >> 3) + q4), pk);
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(P_FULL + i));
            };
            // ---- dS = scale * P * (dP - delta), in place over P ----
            auto phase_ds = [&](int i, int j) {
                const uint32_t par = (i == 0 ? use0 : use1) & 1u;
                if (i == 0) ++use0; else ++use1;
                const uint32_t tR = tmem + i * 128 + lane_base;
                const uint32_t pbuf = sP + i * PB_B;
                const float2 nd2 = splat2(i == 0 ? ndsc_0 : ndsc_1);
                { const long long tw = TRC_NOW(); mbar_wait(bar(DP_FULL + i), par); TRC_ACC(trole, 22, tw); }
                tc_fence_after();
                if (j * QT + cg * BWD_PC < a.T && i * QT + q * 32 < a.T) {
                    const int col = cg * BWD_PC;
                    const int kv0 = j * QT + col;
                    tmem_ld32(tR + col, v);
                    tmem_ld_wait();
                    const bool ragged = kv0 + 32 > a.T;           // dP columns past T were never written by the MMA
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const uint32_t ad = pbuf_addr(pbuf, r, (col >> 3) + q4);
                        const uint4 pp = ld_shared_v4(ad);
                        const float* w = v + q4 * 8;
                        float2 d0 = mul2(unpack_bf16x2(pp.x), fma2(make_float2(w[0], w[1]), sc2, nd2));
                        float2 d1 = mul2(unpack_bf16x2(pp.y), fma2(make_float2(w[2], w[3]), sc2, nd2));
                        float2 d2 = mul2(unpack_bf16x2(pp.z), fma2(make_float2(w[4], w[5]), sc2, nd2));
                        float2 d3 = mul2(unpack_bf16x2(pp.w), fma2(make_float2(w[6], w[7]), sc2, nd2));
                        if (ragged) {
                            const int k0 = kv0 + q4 * 8;
                            d0.x = k0 + 0 < a.T ? d0.x : 0.f; d0.y = k0 + 1 < a.T ? d0.y : 0.f;
                            d1.x = k0 + 2 < a.T ? d1.x : 0.f; d1.y = k0 + 3 < a.T ? d1.y : 0.f;
                            d2.x = k0 + 4 < a.T ? d2.x : 0.f; d2.y = k0 + 5 < a.T ? d2.y : 0.f;
                            d3.x = k0 + 6 < a.T ? d3.x : 0.f; d3.y = k0 + 7 < a.T ? d3.y : 0.f;
                        }
                        uint4 pk;
                        pk.x = pack_bf16x2(d0.x, d0.y); pk.y = pack_bf16x2(d1.x, d1.y);
                        pk.z = pack_bf16x2(d2.x, d2.y); pk.w = pack_bf16x2(d3.x, d3.y);
                        st_shared_v4(ad, pk);
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(DS_FULL + i));
            };
            // ---- epilogues: d q / d k / d v rows leave through shared memory. Each TMEM lane quarter (the 4 warps with the same
            //      q) owns a [32 rows][128 B] swizzled slab of each staging tile and stores it with ONE bulk tensor copy (rows >= T
            //      are clipped by the hardware) - per-lane 32-byte global stores cost 32 LSU wavefronts per instruction and made
            //      the three epilogues half of the item time. The gated q / k / v values of the row (for the d gate products)
            //      are read from the operand tiles still in shared memory; XREAD tells the producer when the last such read is
            //      done so that the next item's tiles may land.
            const uint32_t slab0 = sST + q * 4096 + lane * 128, slab1 = slab0 + TILE_B;
            const bool issuer = cg == 0 && lane == 0;
            auto stage_open = [&]() {                            // the previous round's stores have read the slabs
                if (issuer) bulk_wait_read<0>();
                named_bar_sync(2 + q, 128);
            };
            auto stage_close = [&](int which0, int row0, int which1, int row1) {
                fence_proxy_async_smem();
                named_bar_sync(2 + q, 128);
                if (issuer) {
                    if (row0 + q * 32 < a.T) tma_store_5d(&tm_out, sST + q * 4096, 0, row0 + q * 32, h, which0, b);
                    if (row1 + q * 32 < a.T) tma_store_5d(&tm_out, sST + TILE_B + q * 4096, 0, row1 + q * 32, h, which1, b);
                    bulk_commit();
                }
            };
            // ---- dK_j / dV_j epilogue of a finished kv half (kv row = j*128 + r, head-dim columns [16 cg, 16 cg + 16)) ----
            auto epi_kv = [&](int j, bool last) {
                const int kv = j * QT + r;
                const bool kv_ok = kv < a.T;
                const uint4 k0 = ld_shared_v4(sK + sw128_offset(kv, cg * 2)), k1 = ld_shared_v4(sK + sw128_offset(kv, cg * 2 + 1));
                const uint4 v0 = ld_shared_v4(sV + sw128_offset(kv, cg * 2)), v1 = ld_shared_v4(sV + sw128_offset(kv, cg * 2 + 1));
                if (last) {          // last reads of this item's operand tiles: the producer may refill them once the MMAs retire
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(XREAD));
                    if (item + int(gridDim.x) < n_items) row_scalars(item + gridDim.x, lse_n0, lse_n1, dps_n);
                }
                { const long long tw = TRC_NOW(); mbar_wait(bar(ACC_FULL), n_acc & 1u); TRC_ACC(trole, 23, tw); }
                ++n_acc;
                tc_fence_after();
                // dK then dV, one 16-column slice in registers at a time (the kernel sits at its 96-register ceiling)
                float w1[16];
                tmem_ld16(tDK + lane_base + cg * BWD_EC, w1);
                tmem_ld_wait();
                stage_open();
                dqkv_slice16<2>(w1, kv_ok, k0, k1, slab0, cg * 2, gate16, ag, bs_k);
                tmem_ld16(tDV + lane_base + cg * BWD_EC, w1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(ACC_EMPTY));
                dqkv_slice16<1>(w1, kv_ok, v0, v1, slab1, cg * 2, gate16, ag, bs_v);
                stage_close(1, j * QT, 2, j * QT);
            };
            TRC(trole, 2);
            if (nt == 1) {
                phase_p(0, 0);
                stats_finish();
                phase_ds(0, 0);
            } else {
                phase_p(0, 0); TRC(trole, 3); phase_p(1, 0);
                stats_finish(); TRC(trole, 4);
                phase_ds(0, 0); TRC(trole, 5); phase_ds(1, 0); TRC(trole, 6);
                phase_p(0, 1); TRC(trole, 7);
                epi_kv(0, false);                                // its TMEM reads release dK / dV for the second kv half
                TRC(trole, 8);
                phase_p(1, 1); TRC(trole, 9);
                phase_ds(0, 1); TRC(trole, 10); phase_ds(1, 1); TRC(trole, 11);
            }
            epi_kv(nh - 1, true);
            TRC(trole, 12);
            // ---- dQ epilogue of both q tiles ----
            {
                { const long long tw = TRC_NOW(); mbar_wait(bar(DQ_FULL), pi); TRC_ACC(trole, 24, tw); }
                tc_fence_after();
                float w1[16];
                tmem_ld16(tDQ + lane_base + cg * BWD_EC, w1);
                tmem_ld_wait();
                stage_open();
                const uint4 none = make_uint4(0, 0, 0, 0);
                dqkv_slice16<0>(w1, r < a.T, none, none, slab0, cg * 2, gate16, ag, bs_q);
                if (nh == 2) {
                    tmem_ld16(tDQ + HD + lane_base + cg * BWD_EC, w1);
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(DQ_EMPTY));
                if (nh == 2) dqkv_slice16<0>(w1, QT + r < a.T, none, none, slab1, cg * 2, gate16, ag, bs_q);
                stage_close(0, 0, 0, nh == 2 ? QT : a.T);
            }
            // d gate: q, k and v contributions of all rows of this warp
            {
                float tsum[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) { tsum[2 * j] = ag[j].x; tsum[2 * j + 1] = ag[j].y; }
                const float sg = bfly16(tsum);
                if ((lane & 1) == 0) {
                    const int c = lane >> 1;
                    cs_warp[c] = sg; cs_warp[16 + c] = bs_q; cs_warp[32 + c] = bs_k; cs_warp[48 + c] = bs_v;
                }
            }
            TRC(trole, 13);
            // ---- per-item column sums -> global partials: the four row quarters of a column group are added in fixed order ----
            named_bar_sync(1, BWD_CT);
            TRC(trole, 14);
            {
                const int tid = threadIdx.x;   // 0..511
                if (tid < 256) {
                    const int which = tid >> 6, col = tid & 63;          // 0 = gate, 1..3 = bias of q, k, v
                    const float* src = cs + (pi * BWD_CW + (col >> 4) * 4) * 64 + which * 16 + (col & 15);   // warp = cg * 4 + quarter
                    const float val = ((src[0] + src[64]) + src[128]) + src[192];
                    if (which == 0) a.part_gate[size_t(b) * D + h * HD + col] = val;
                    else a.part_bias[size_t(b) * 3 * D + (which - 1) * D + h * HD + col] = val;
                }
            }
        }
    }
    if (warp < BWD_CW && (warp >> 2) == 0 && lane == 0) bulk_wait<0>();      // this quarter's last stores have completed
    tc_fence_before();
    __syncthreads();
    if (warp == BWD_CW + 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static constexpr int FWD_SMEM = 2 * TILE_B + 2 * KV_B + 2 * PBUF_B + 4096 + 256;
static constexpr int BWD_SMEM = 4 * TILE_B + 2 * KV_B + 2 * PB_B + 2 * TILE_B + 8192 + 4096 + 256;

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
    CUtensorMap tq, tkv, to;
    int r = make_qkv_maps(qkv, B, T, H, &tq, &tkv);
    if (r) return r;
    {
        // O leaves in [32 rows][64 columns] slabs (one per TMEM lane quarter of a q tile)
        const uint64_t D = uint64_t(H) * HD;
        uint64_t dims[4] = {D, uint64_t(T), uint64_t(B), 1};
        uint64_t str[4] = {1, D, uint64_t(T) * D, uint64_t(B) * T * D};
        uint32_t box[4] = {HD, 32, 1, 1};
        r = make_tmap_bf16(&to, o, 4, dims, str, box);
        if (r) return r;
    }
    AttnArgs a{};
    a.B = B; a.T = T; a.H = H; a.scale = scale; a.drop_scale = drop_scale;
    a.o = reinterpret_cast<__nv_bfloat16*>(o); a.lse = lse;
    const int items = B * H;
    const int grid = items < num_sms() ? items : num_sms();
    return int(launch_k(attn_fwd_kernel, dim3(grid), dim3(FWD_THREADS), size_t(FWD_SMEM), s, 1, tq, tkv, to, a));
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
    CUtensorMap tq, tkv, tdo, tout, to;
    int r = make_qkv_maps(qkv, B, T, H, &tq, &tkv);
    if (r) return r;
    {
        // d qkv leaves in [32 rows][64 columns] slabs (one per TMEM lane quarter), same 5-D geometry as the qkv input
        const uint64_t D = uint64_t(H) * HD;
        uint64_t dims[5] = {HD, uint64_t(T), uint64_t(H), 3, uint64_t(B)};
        uint64_t str[5] = {1, 3 * D, HD, D, uint64_t(T) * 3 * D};
        uint32_t box[5] = {HD, 32, 1, 1, 1};
        r = make_tmap_bf16(&tout, dqkv, 5, dims, str, box);
        if (r) return r;
    }
    {
        const uint64_t D = uint64_t(H) * HD;
        uint64_t dims[4] = {D, uint64_t(T), uint64_t(B), 1};
        uint64_t str[4] = {1, D, uint64_t(T) * D, uint64_t(B) * T * D};
        uint32_t box[4] = {HD, QT, 1, 1};
        r = make_tmap_bf16(&tdo, d_o, 4, dims, str, box);
        if (r) return r;
        r = make_tmap_bf16(&to, o, 4, dims, str, box);          // L2 prefetch of the forward output rows only
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
    return int(launch_k(attn_bwd_kernel, dim3(grid), dim3(BWD_THREADS), size_t(BWD_SMEM), s, 1, tq, tkv, tdo, tout, to, a));
}

#ifdef OFB_ATTN_TRACE
}  // namespace ofb
extern "C" int ofb_debug_attn_trace(long long* out, int n, int reset);
namespace ofb {
int attn_trace_read(long long* out, int n, int reset) {
    const size_t bytes = sizeof(long long) * size_t(n < 3 * 4 * 32 ? n : 3 * 4 * 32);
    cudaDeviceSynchronize();
    if (out != nullptr && cudaMemcpyFromSymbol(out, g_attn_trace, bytes) != cudaSuccess) return 1;
    if (reset) { static long long zeros[3 * 4 * 32]; cudaMemcpyToSymbol(g_attn_trace, zeros, sizeof(zeros)); }
    return 0;
}
#endif

}  // namespace ofb

#ifdef OFB_ATTN_TRACE
extern "C" int ofb_debug_attn_trace(long long* out, int n, int reset) { return ofb::attn_trace_read(out, n, reset); }
#endif
