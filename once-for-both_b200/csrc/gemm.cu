// gemm.cu — persistent warp-specialised bf16 GEMM for sm_100a.
//
//   D[M,N] = sum_k A[m,k] * B[n,k]       fp32 accumulation in TMEM
//
// * operands are staged by TMA (SWIZZLE_128B) into a multi-stage shared-memory ring,
// * one elected thread issues tcgen05.mma (UMMA 128 x BN x 16) with the accumulator in TMEM,
// * two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1,
// * four epilogue warps read TMEM with tcgen05.ld and apply the fused bi-mask epilogues.
//
// A and B can each be K-major (row-major [rows, K], the nn.Linear layout) or MN-major
// (row-major [K, rows]; used by the weight-gradient GEMM whose reduction runs over tokens).
//
// Replaces, on the hot path of the reference: nn.Linear in MAESparseAttention.forward (layers.py:491,515),
// MAESparseMlp.forward (layers.py:845-863), the patch-embed conv (layers.py:177), the PMIM decoder 1x1 conv
// (vision_transformer.py:723), the head (vision_transformer.py:744) and their autograd backward (engine.py:169).
#include "ptx.cuh"
#include "gemm.cuh"
#include "launch.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace ofb {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int EPI_WARPS = 8;                       // two groups of four (one warp per TMEM lane quarter each)
static constexpr int EPI_THREADS = EPI_WARPS * 32;
static constexpr int GEMM_THREADS = 64 + EPI_THREADS;     // warp0 TMA, warp1 MMA (+TMEM alloc), warps2-9 epilogue
static constexpr int SMEM_LIMIT = 232448;                 // 227 KB
static constexpr int PANEL_BYTES = BM * 128;              // one [128 rows][64 bf16] SWIZZLE_128B output panel

// Bound-finding experiments (tools/gemm_bound.py; compiled only with -DOFB_GEMM_DEBUG into a separate debug library, never the
// product): bit 0 = the producer stops issuing TMA loads once the operand ring has been filled (the MMAs then run on stale
// shared memory: main loop + epilogue without any L2 -> SM operand feed), bit 1 = the epilogue skips its global / TMA stores,
// bit 2 = the epilogue skips the TMA loads of residual / saved-activation panels, bit 3 = no epilogue at all (the accumulator
// stage is handed straight back: pure main-loop rate), bit 4 = the single-CTA kernels issue their MMAs in the TS form with the A
// operand read from (arbitrary) tensor-memory columns instead of shared memory: what a weight-stationary A operand would cost.
#ifdef OFB_GEMM_DEBUG
__device__ int g_gemm_dbg = 0;
#define GEMM_DBG(bit) ((g_gemm_dbg >> (bit)) & 1)
#else
#define GEMM_DBG(bit) 0
#endif

// CG = CTAs per tile: 1, or 2 for a CTA pair (cta_group::2) that shares one 256-row UMMA: each CTA stages its own 128 rows of A
// but only HALF of the B tile, which cuts the L2 -> shared-memory operand traffic per MAC from (128 + BN) / (128 BN) to
// (128 + BN / 2) / (128 BN). The 1-CTA kernel is bound by exactly that feed (profiles/r01_experiments.md).
template <int BN, int EPI, int TMA_OUT, int CG = 1>
struct GemmCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2 / CG;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NOUT = (EPI == EPI_FC1) ? 2 : 1;
    // Epilogue warps. The transposed-hidden epilogues (GELU / GELU' on every element) are bound by their own arithmetic, and
    // that arithmetic is latency-bound at two warps per scheduler: measured (tools/micro/epi_alu_bench.cu) 35.5 -> 47.5
    // element-pair-warps / us / scheduler going from 8 to 16 warps per SM. They therefore run 16 epilogue warps = 4 column
    // groups = 2 independent panel pipelines (groups {0,1} -> even 64-column panels, groups {2,3} -> odd panels).
    static constexpr int EW = ((EPI == EPI_FC1 || EPI == EPI_FC2_DGRAD) && BN % 128 == 0) ? 2 * EPI_WARPS : EPI_WARPS;
    static constexpr int NG = EW / 4;                 // column groups (one warp per TMEM lane quarter each)
    static constexpr int NPIPE = NG / 2;              // panel pipelines of two groups (256 threads) each
    static constexpr int THREADS = 64 + EW * 32;
    static constexpr int STAGING_BYTES = TMA_OUT ? 2 * NOUT * PANEL_BYTES : 0;          // two slots (NPIPE == 2: one per pipeline)
    // TMA-loaded residual / saved-activation panels: the ring holds about one tile of panels so that the load of a panel is
    // issued a whole tile time before its use (HBM latency is ~3 panel times). Two pipelines: two slots each.
    static constexpr int AUX_SLOTS = (TMA_OUT != 2) ? 0 : (NPIPE == 2 ? 4 : (BN == 128 ? 3 : (BN == 192 ? 4 : ((BN == 256 && EPI == EPI_FC2_DGRAD) ? 3 : 2))));
    static constexpr int AUX_BYTES = AUX_SLOTS * PANEL_BYTES;
    // weight-gradient epilogue: 2 KB per epilogue warp to regroup a [32 rows][16 columns] fp32 block so that 4 lanes share a row
    // segment of the red.global.add (it fits in what the operand ring leaves over: the stage count is unchanged for BN >= 192)
    static constexpr int SCRATCH_BYTES = (EPI == EPI_WGRAD) ? EW * 2048 : 0;
    // per-column epilogue vectors, double-buffered by tile parity (the transposed-hidden epilogues use per-thread scalars)
    static constexpr int VEC_BYTES = (EPI == EPI_FC1 || EPI == EPI_FC2_DGRAD) ? 0 : 2 * 2 * BN * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int FIXED = STAGING_BYTES + AUX_BYTES + SCRATCH_BYTES + VEC_BYTES + BAR_BYTES;
    static constexpr int STAGES_RAW = (SMEM_LIMIT - FIXED) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED;
    static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
    // the 16-warp epilogues take ~4x the main loop of a tile: two operand stages keep the (double-buffered) accumulators ahead
    static_assert(STAGES >= (NPIPE == 2 ? 2 : 3), "pipeline too shallow");
    static_assert((2 * STAGES + 4 + 4) * 8 + 8 <= BAR_BYTES, "barrier block too small");
};

// 32 values per lane -> lane L ends with the sum over lanes of v[L] (31 shuffles).
__device__ __forceinline__ float lane_transpose_sum(float (&v)[32]) {
    const uint32_t lane = lane_id();
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = upper ? v[i] : v[i + o];
            const float keep = upper ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32], int nvalid) {
    if (nvalid >= 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 p;
            p.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
            p.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
            p.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
            p.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
            reinterpret_cast<uint4*>(dst)[i] = p;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < nvalid) dst[i] = __float2bfloat16(v[i]);
    }
}
__device__ __forceinline__ void store_f32x32(float* dst, const float (&v)[32], int nvalid) {
    if (nvalid >= 32) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < nvalid) dst[i] = v[i];
    }
}
__device__ __forceinline__ void load_f32x32(const float* src, float (&v)[32], int nvalid) {
    if (nvalid >= 32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(src) + i);
            v[4 * i] = p.x; v[4 * i + 1] = p.y; v[4 * i + 2] = p.z; v[4 * i + 3] = p.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (i < nvalid) ? src[i] : 0.f;
    }
}
// 32 bf16 of one row as four packed 16-byte registers (zero-filled when the row / columns are out of range)
struct Packed32 { uint4 p[4]; };
__device__ __forceinline__ Packed32 load_packed32(const __nv_bfloat16* src, bool row_ok, int nvalid) {
    Packed32 r;
    if (row_ok && nvalid >= 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r.p[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    } else {
        // ragged tail: compile-time indexed so that the struct stays in registers (no local-memory round trip)
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t lo = (row_ok && 2 * i < nvalid) ? uint32_t(__bfloat16_as_ushort(src[2 * i])) : 0u;
            const uint32_t hi = (row_ok && 2 * i + 1 < nvalid) ? uint32_t(__bfloat16_as_ushort(src[2 * i + 1])) : 0u;
            w[i] = lo | (hi << 16);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) r.p[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
    return r;
}
__device__ __forceinline__ void unpack32(const Packed32& r, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f;
        f = unpack_bf16x2(r.p[i].x); v[8 * i + 0] = f.x; v[8 * i + 1] = f.y;
        f = unpack_bf16x2(r.p[i].y); v[8 * i + 2] = f.x; v[8 * i + 3] = f.y;
        f = unpack_bf16x2(r.p[i].z); v[8 * i + 4] = f.x; v[8 * i + 5] = f.y;
        f = unpack_bf16x2(r.p[i].w); v[8 * i + 6] = f.x; v[8 * i + 7] = f.y;
    }
}
// 32 fp32 of one row -> bf16 -> the row's 64-byte slice (chunk parity `half`) of a swizzled [128][64] output panel
__device__ __forceinline__ void stage_bf16x32(uint32_t panel, int row, int half, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t a = panel + sw128_offset(row, half * 4 + i);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(v[8 * i + 0], v[8 * i + 1])),
                     "r"(pack_bf16x2(v[8 * i + 2], v[8 * i + 3])), "r"(pack_bf16x2(v[8 * i + 4], v[8 * i + 5])),
                     "r"(pack_bf16x2(v[8 * i + 6], v[8 * i + 7]))
                     : "memory");
    }
}

// same for 16 packed pairs
__device__ __forceinline__ void stage_bf16x32_p(uint32_t panel, int row, int half, const float2 (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t a = panel + sw128_offset(row, half * 4 + i);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(v[4 * i].x, v[4 * i].y)),
                     "r"(pack_bf16x2(v[4 * i + 1].x, v[4 * i + 1].y)), "r"(pack_bf16x2(v[4 * i + 2].x, v[4 * i + 2].y)),
                     "r"(pack_bf16x2(v[4 * i + 3].x, v[4 * i + 3].y))
                     : "memory");
    }
}
// NP packed pairs (2 NP bf16 = NP / 4 16-byte chunks) of one row, starting at 16-byte chunk `chunk16` of the row's 128-byte line
template <int NP>
__device__ __forceinline__ void stage_bf16_pairs(uint32_t panel, int row, int chunk16, const float2 (&v)[NP]) {
#pragma unroll
    for (int i = 0; i < NP / 4; ++i) {
        const uint32_t a = panel + sw128_offset(row, chunk16 + i);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(v[4 * i].x, v[4 * i].y)),
                     "r"(pack_bf16x2(v[4 * i + 1].x, v[4 * i + 1].y)), "r"(pack_bf16x2(v[4 * i + 2].x, v[4 * i + 2].y)),
                     "r"(pack_bf16x2(v[4 * i + 3].x, v[4 * i + 3].y))
                     : "memory");
    }
}
__device__ __forceinline__ uint32_t packed_word(const Packed32& r, int i) {
    const uint4& q = r.p[i >> 2];
    return (i & 3) == 0 ? q.x : ((i & 3) == 1 ? q.y : ((i & 3) == 2 ? q.z : q.w));
}

// TMA_OUT: 0 = direct global stores; 1 = bf16 output panels staged in smem and written by TMA; 2 = 1 + the residual (STORE) /
// saved activation (FC2_DGRAD) tile is TMA-loaded into smem panels two panels ahead of its use.
template <int BN, int A_MN, int B_MN, int EPI, int TMA_OUT, int CG>
__global__ void __launch_bounds__((GemmCfg<BN, EPI, TMA_OUT, CG>::THREADS), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_o0, const __grid_constant__ CUtensorMap tma_o1,
            const __grid_constant__ CUtensorMap tma_aux, const GemmArgs g) {
    using Cfg = GemmCfg<BN, EPI, TMA_OUT, CG>;
    pdl_trigger();          // one CTA (pair) per SM: the next kernel's CTAs may take over each SM as soon as this one's leave it
    constexpr int STAGES = Cfg::STAGES;
    constexpr uint32_t IDESC = make_idesc_bf16(BM * CG, BN, A_MN, B_MN);
    // CTA pair: rank in the cluster (0 = leader, issues the MMAs), pair index / pair count replace the CTA index / grid size
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int cta_first = (CG == 2) ? int(blockIdx.x >> 1) : int(blockIdx.x);
    const int cta_stride = (CG == 2) ? int(gridDim.x >> 1) : int(gridDim.x);
    // transposed-hidden epilogues: rows = hidden units (few m tiles, the weight operand), columns = tokens; walking the m
    // tiles of one token tile back to back keeps that token tile in L2
    constexpr bool M_FAST = (EPI == EPI_FC1 || EPI == EPI_FC2_DGRAD);
    constexpr int NCHUNK = BN / 32;          // 32-column chunks per tile; group g handles chunks c = g, g + NG, ...
    constexpr int NPANEL = BN / 64;
    constexpr int EW = Cfg::EW, NG = Cfg::NG, NPIPE = Cfg::NPIPE;
    constexpr int EPI_THREADS = EW * 32;     // (shadows the 8-warp constant of the file scope)
    static_assert(NCHUNK % NG == 0, "column chunks must divide evenly over the groups");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;                       // 1024-aligned (stage sizes are multiples of 8 KB)
    uint8_t* auxbuf = staging + Cfg::STAGING_BYTES;                            // [2 slots][PANEL_BYTES], 1024-aligned
    float* vecs = reinterpret_cast<float*>(auxbuf + Cfg::AUX_BYTES + Cfg::SCRATCH_BYTES);   // [2 parities][2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(auxbuf + Cfg::AUX_BYTES + Cfg::SCRATCH_BYTES + Cfg::VEC_BYTES);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]
    uint64_t* aux_full = bars + 2 * STAGES + 4;   // [AUX_SLOTS <= 4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);
    constexpr uint32_t AUX_SLOTS = Cfg::AUX_SLOTS > 0 ? Cfg::AUX_SLOTS : 1;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        if ((smem_u32(smem) & 1023u) != 0) { printf("ofb: dynamic smem base not 1024-aligned\n"); __trap(); }
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        if (TMA_OUT) { tma_prefetch_desc(&tma_o0); if (Cfg::NOUT == 2) tma_prefetch_desc(&tma_o1); }
        if (TMA_OUT == 2) {
            tma_prefetch_desc(&tma_aux);
            for (uint32_t i = 0; i < AUX_SLOTS; ++i) mbar_init(smem_u32(&aux_full[i]), 1);
        }
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&tfull_bar[i]), 1);
            mbar_init(smem_u32(&tempty_bar[i]), EW * CG);             // pair: the epilogue warps of both CTAs release the leader's MMA
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        if (CG == 2) { tmem_alloc_cg2(smem_u32(tmem_slot), Cfg::TMEM_COLS); tmem_relinquish_cg2(); }
        else { tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();          // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();             // prologue done (barriers, TMEM, descriptor prefetch): from here on the previous kernel's outputs are read

    const int m_tiles = (g.M + BM - 1) / BM;              // 128-row tiles (epilogue granularity)
    const int mt = (CG == 2) ? (m_tiles + 1) / 2 : m_tiles;   // scheduled row tiles (128 CG rows each)
    const int n_tiles = (g.N + BN - 1) / BN;
    const int k_blocks = (g.K + BK - 1) / BK;
    const int splits = g.k_splits > 0 ? g.k_splits : 1;
    const int kb_per_split = (k_blocks + splits - 1) / splits;
    const int total_tiles = mt * n_tiles * splits;
    // scheduled tile index -> (split, 128-row tile of THIS CTA, column tile)
    auto decode = [&](int t, int& split, int& m_blk, int& n_blk) {
        split = t / (mt * n_tiles);
        const int mn = t % (mt * n_tiles);
        const int pm = M_FAST ? mn % mt : mn / n_tiles;
        n_blk = M_FAST ? mn / mt : mn % n_tiles;
        m_blk = (CG == 2) ? pm * 2 + int(rank) : pm;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = cta_first; t < total_tiles; t += cta_stride) {
                int split, m_blk, n_blk;
                decode(t, split, m_blk, n_blk);
                const int kb0 = split * kb_per_split;
                const int kb1 = min(k_blocks, kb0 + kb_per_split);
                const int nrow0 = n_blk * BN + int(rank) * (BN / CG);     // pair: this CTA stages its half of the B tile
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                    const uint32_t sa = smem_u32(smem_a + stage * Cfg::A_BYTES);
                    const uint32_t sb = smem_u32(smem_b + stage * Cfg::B_BYTES);
#ifdef OFB_GEMM_DEBUG
                    if (GEMM_DBG(0) && (phase != 0 || t != cta_first || kb - kb0 >= STAGES)) {
                        // stale operands: complete the stage's transaction count without moving any bytes
                        if (CG == 2) { if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), 0); }
                        else mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), 0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
#endif
                    if (CG == 2) {
                        // both CTAs' copies complete on the LEADER's full barrier, which expects the bytes of the pair
                        const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), 2 * Cfg::STAGE_BYTES);
                        if (A_MN) {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j) tma_load_2d_cg2(sa + j * (BK * 128), &tma_a, fb, m_blk * BM + j * 64, kb * BK);
                        } else {
                            tma_load_2d_cg2(sa, &tma_a, fb, kb * BK, m_blk * BM);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int j = 0; j < BN / CG / 64; ++j) tma_load_2d_cg2(sb + j * (BK * 128), &tma_b, fb, nrow0 + j * 64, kb * BK);
                        } else {
                            tma_load_2d_cg2(sb, &tma_b, fb, kb * BK, nrow0);
                        }
                    } else {
                        const uint32_t fb = smem_u32(&full_bar[stage]);
                        mbar_arrive_expect_tx(fb, Cfg::STAGE_BYTES);
                        if (A_MN) {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), &tma_a, fb, m_blk * BM + j * 64, kb * BK);
                        } else {
                            tma_load_2d(sa, &tma_a, fb, kb * BK, m_blk * BM);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &tma_b, fb, n_blk * BN + j * 64, kb * BK);
                        } else {
                            tma_load_2d(sb, &tma_b, fb, kb * BK, n_blk * BN);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair: the leader CTA only) =====================
        // The whole warp runs the loop (uniform control flow, uniform operands); one elected lane issues the tcgen05 instructions.
        if (rank == 0) {
            int stage = 0; uint32_t phase = 0;
            int it = 0;
#ifdef OFB_GEMM_DEBUG
            const bool dbg_ts = GEMM_DBG(4) != 0;
#endif
            for (int t = cta_first; t < total_tiles; t += cta_stride) {
                const int split = t / (mt * n_tiles);
                const int kb0 = split * kb_per_split;
                const int kb1 = min(k_blocks, kb0 + kb_per_split);
                if (kb1 <= kb0) continue;
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ++it;
                mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&full_bar[stage]), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem_a + stage * Cfg::A_BYTES);
                    const uint32_t sb = smem_u32(smem_b + stage * Cfg::B_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // K-major: advance 16 bf16 (32 B) inside the 128-B swizzle row.
                            // MN-major: advance 16 k-rows of 128 B; 64-wide MN atoms are BK*128 B apart (LBO).
                            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, BK * 128, 1024)
                                                     : make_smem_desc_sw128(sa + k * 32, 0, 1024);
                            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, BK * 128, 1024)
                                                     : make_smem_desc_sw128(sb + k * 32, 0, 1024);
                            if (CG == 2) umma_bf16_cg2(d_tmem, da, db, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
#ifdef OFB_GEMM_DEBUG
                            else if (dbg_ts)
                                umma_bf16_ts(d_tmem, tmem_base + uint32_t((acc ^ 1) * BN + (kb & 3) * 32 + k * 8), db,
                                             make_idesc_bf16(BM, BN, 0, B_MN), (kb > kb0 || k > 0) ? 1u : 0u);
#endif
                            else umma_bf16(d_tmem, da, db, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        if (CG == 2) {
                            umma_commit_cg2(smem_u32(&empty_bar[stage]));          // frees the stage in both CTAs
                            if (kb == kb1 - 1) umma_commit_cg2(smem_u32(&tfull_bar[acc]));
                        } else {
                            umma_commit(smem_u32(&empty_bar[stage]));
                            if (kb == kb1 - 1) umma_commit(smem_u32(&tfull_bar[acc]));
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int ew = warp - 2;           // 0..EW-1
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int grp = ew >> 2;           // column group: chunks c = grp, grp + NG, ...
        const int pp = (NPIPE == 2) ? (grp >> 1) : 0;      // panel pipeline of this warp
        const int et = q * 32 + lane;      // row inside the tile
        const int etid = ew * 32 + lane;   // 0..EPI_THREADS-1
        const bool elected = (etid == pp * 256);           // one thread per pipeline issues its TMA stores / aux loads
        int it = 0;
        uint32_t panel_ctr = 0;            // panels this pipeline has processed
        // TMA_OUT == 2: panel number P of this CTA (tile sequence P / NPT, panel P % NPT) is loaded into aux slot P % AUX_SLOTS;
        // with two pipelines P alternates between them (NPT is even), so each owns the slots of its own parity
        constexpr uint32_t NPT = NCHUNK / 2;
        constexpr uint32_t PPT = NPT / NPIPE;              // panels per tile per pipeline
        static_assert(NPIPE == 1 || (NPT % 2 == 0 && Cfg::AUX_SLOTS % 2 == 0), "two pipelines need an even panel count");
        auto issue_aux = [&](uint32_t P) {
            const long t2 = long(cta_first) + long(P / NPT) * cta_stride;
            if (t2 >= total_tiles) return;
            int s2, m2, n2;                                           // k_splits == 1 for these epilogues
            decode(int(t2), s2, m2, n2);
            const uint32_t fb = smem_u32(&aux_full[P % AUX_SLOTS]);
            if (GEMM_DBG(2)) { mbar_arrive_expect_tx(fb, 0); return; }
            mbar_arrive_expect_tx(fb, PANEL_BYTES);
            tma_load_2d(smem_u32(auxbuf) + (P % AUX_SLOTS) * PANEL_BYTES, &tma_aux, fb, n2 * BN + int(P % NPT) * 64, m2 * BM);
        };
        if (TMA_OUT == 2 && elected) {
            for (uint32_t i = 0; i < AUX_SLOTS; ++i)
                if (int(i % NPIPE) == pp) issue_aux(i);
        }
        // two pipelines share the two staging slots (one each): the slot's previous TMA store must have been read out before
        // the pipeline stages its next panel
        auto pre_stage = [&]() {
            if (NPIPE == 2) {
                if (elected) bulk_wait_read<0>();
                named_bar_sync(4 + pp, 256);
            }
        };
        // pair: the accumulator stage is handed back to the leader's MMA warp (remote arrive from the peer CTA)
        const uint32_t tempty_remote0 = (CG == 2) ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : 0u;
        for (int t = cta_first; t < total_tiles; t += cta_stride) {
            int split, m_blk, n_blk;
            decode(t, split, m_blk, n_blk);
            const int mn = m_blk * n_tiles + n_blk;       // tile id for per-tile partial outputs
            const int kb0 = split * kb_per_split;
            const int kb1 = min(k_blocks, kb0 + kb_per_split);
            if (kb1 <= kb0) continue;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            ++it;

            const int row = m_blk * BM + et;
            const bool row_ok = row < g.M;
            const int n0 = n_blk * BN;

            // ---- per-column epilogue vectors of this tile -> smem (while the MMAs of the tile are still running) ----
            float* vb = vecs + acc * 2 * BN;     // bias-like vector
            float* vs = vb + BN;                 // scale-like vector
            if (EPI == EPI_STORE || EPI == EPI_PATCH || EPI == EPI_DECODER) {
                for (int i = etid; i < BN; i += EPI_THREADS) {
                    const int col = n0 + i;
                    const bool ok = col < g.N;
                    float s = 1.f, b = 0.f;
                    if (g.colscale != nullptr) s = ok ? __ldg(g.colscale + (g.colscale_period > 0 ? col % g.colscale_period : col)) : 0.f;
                    if (g.bias != nullptr) b = ok ? __ldg(g.bias + col) : 0.f;
                    if (EPI == EPI_STORE) b *= s;          // out = ars*(acc*cs + brs*bias*cs) + res
                    vb[i] = b; vs[i] = s;
                }
            }
            float rs = 1.f;
            if (EPI == EPI_STORE) {
                if (g.rowscale != nullptr && row_ok) rs = __ldg(g.rowscale + row / g.rows_per_scale);
            }
            // transposed-hidden epilogues: this thread's row is one hidden unit -> bias / gate are per-thread scalars
            float bias_j = 0.f, gate_j = 0.f, acc_dg = 0.f, acc_db = 0.f;
            if (EPI == EPI_FC1 || EPI == EPI_FC2_DGRAD) {
                if (row_ok) {
                    gate_j = __ldg(g.colscale + row);
                    if (EPI == EPI_FC1) bias_j = __ldg(g.bias + row);
                }
            }
            // DropPath sample bookkeeping of the transposed-hidden epilogues (columns = tokens): once per tile
            int tile_b0 = 0, tile_rem = 0, tile_last = 0;
            if ((EPI == EPI_FC1 || EPI == EPI_FC2_DGRAD) && g.rowscale != nullptr) {
                tile_b0 = n0 / g.rows_per_scale;
                tile_rem = n0 - tile_b0 * g.rows_per_scale;
                tile_last = (g.N - 1) / g.rows_per_scale;
            }
            float gs = 1.f;   // global device scalar
            if (EPI == EPI_STORE || EPI == EPI_WGRAD) {
                if (g.scale_ptr != nullptr) gs = __ldg(g.scale_ptr);
            }
            // ---- prefetch this thread's residual / saved-activation slices (latency hides behind the tfull wait) ----
            Packed32 pre[(EPI == EPI_STORE && TMA_OUT == 0) ? NCHUNK / NG : 1];
            if (EPI == EPI_STORE && TMA_OUT == 0) {
                if (g.res != nullptr) {
#pragma unroll
                    for (int j = 0; j < NCHUNK / NG; ++j) {
                        const int col0 = n0 + (NG * j + grp) * 32;
                        pre[j] = load_packed32(g.res + size_t(row) * g.ldres + col0, row_ok, min(32, g.N - col0));
                    }
                }
            }
            if (Cfg::VEC_BYTES > 0) named_bar_sync(1, EPI_THREADS);       // epilogue vectors visible
            mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
            tc_fence_after();
#ifdef OFB_GEMM_DEBUG
            if (GEMM_DBG(3)) {            // no epilogue at all: hand the accumulator stage straight back (pure main-loop rate)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_cluster(tempty_remote0 + acc * 8);
                    else mbar_arrive(smem_u32(&tempty_bar[acc]));
                }
                continue;
            }
#endif

            float loss_acc = 0.f;
#pragma unroll
            for (int j = 0; j < NCHUNK / NG; ++j) {
                const int c = NG * j + grp;
                const int pj = c >> 1;            // 64-column panel of this chunk inside the tile
                float v[32];
                tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + c * 32), v);
                tmem_ld_wait();
                if (j == NCHUNK / NG - 1) {
                    // this warp's last TMEM read of the tile: hand the accumulator stage back to the MMA warp early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_cluster(tempty_remote0 + acc * 8);
                        else mbar_arrive(smem_u32(&tempty_bar[acc]));
                    }
                }
                const int col0 = n0 + c * 32;
                const int nvalid = min(32, g.N - col0);
                const float* cb = vb + c * 32;
                const float* cs = vs + c * 32;
                const uint32_t slot = (NPIPE == 2) ? uint32_t(pp) : (panel_ctr & 1u);
                const uint32_t panel0 = smem_u32(staging) + slot * (Cfg::NOUT * PANEL_BYTES);
                // global panel number of this CTA (see issue_aux)
                const uint32_t P = (panel_ctr / PPT) * NPT + (panel_ctr % PPT) * NPIPE + uint32_t(pp);
                Packed32 ax;       // this row's 32 residual / saved-activation values of the chunk (TMA-loaded panel)
                if (TMA_OUT == 2) {
                    const uint32_t aslot = P % AUX_SLOTS;
                    mbar_wait(smem_u32(&aux_full[aslot]), (P / AUX_SLOTS) & 1u);
                    const uint32_t ab = smem_u32(auxbuf) + aslot * PANEL_BYTES;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t ad = ab + sw128_offset(et, (c & 1) * 4 + i);
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(ax.p[i].x), "=r"(ax.p[i].y), "=r"(ax.p[i].z), "=r"(ax.p[i].w) : "r"(ad) : "memory");
                    }
                }

                if (EPI == EPI_STORE) {
                    const float brs = g.bias_rowscaled ? rs : 1.f;
                    const float ars = (g.bias_rowscaled ? 1.f : rs) * gs;
                    const bool has_res = (TMA_OUT == 2) || (TMA_OUT == 0 && g.res != nullptr);
                    const float2 brs2 = splat2(brs), ars2 = splat2(ars);
                    const float2* cs2 = reinterpret_cast<const float2*>(cs);
                    const float2* cb2 = reinterpret_cast<const float2*>(cb);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float2 tt = fma2(make_float2(v[2 * i], v[2 * i + 1]), cs2[i], mul2(brs2, cb2[i]));
                        float2 o;
                        if (has_res) o = fma2(ars2, tt, unpack_bf16x2(packed_word(TMA_OUT == 2 ? ax : pre[j], i)));
                        else o = mul2(ars2, tt);
                        v[2 * i] = o.x; v[2 * i + 1] = o.y;
                    }
                    if (TMA_OUT) {
                        stage_bf16x32(panel0, et, c & 1, v);
                    } else if (row_ok && nvalid > 0) {
                        if (g.out_fp32) store_f32x32(reinterpret_cast<float*>(g.out0) + size_t(row) * g.ld0 + col0, v, nvalid);
                        else store_bf16x32(reinterpret_cast<__nv_bfloat16*>(g.out0) + size_t(row) * g.ld0 + col0, v, nvalid);
                    }
                } else if (EPI == EPI_FC1 || EPI == EPI_FC2_DGRAD) {
                    // DropPath multiplier of each of the 32 tokens (columns) of this chunk: at most two samples per chunk
                    float rsA = 1.f, rsB = 1.f;
                    int nb = 32;
                    if (g.rowscale != nullptr) {
                        // sample index of the chunk's first token without a per-chunk integer division (a division is ~30
                        // dependent instructions; four per tile showed up as 12 % of the epilogue's issue slots): the tile's
                        // first sample / remainder are computed once per tile, a chunk is at most BN tokens further on
                        int bq = tile_b0, off = tile_rem + c * 32;
                        while (off >= g.rows_per_scale) { off -= g.rows_per_scale; ++bq; }
                        const int b0 = min(bq, tile_last);
                        nb = (b0 + 1) * g.rows_per_scale - col0;
                        rsA = __ldg(g.rowscale + b0);
                        rsB = __ldg(g.rowscale + min(b0 + 1, tile_last));
                    }
                    // The 16-warp variant works on 16 columns at a time (112 registers per thread), the 8-warp variant on all 32
                    constexpr int HALVES = (EW == 16) ? 2 : 1;
                    constexpr int HP = 16 / HALVES;                 // packed pairs per half
                    if (EPI == EPI_FC1) {
                        // packed fp32x2 math: the epilogue is fma-pipe bound (see ptx.cuh)
                        const float2 b2 = splat2(bias_j), g2 = splat2(gate_j);
#pragma unroll
                        for (int hh = 0; hh < HALVES; ++hh) {
                            float2 u2[HP], h2[HP];
#pragma unroll
                            for (int i = 0; i < HP; ++i) {
                                const int e = 2 * (hh * HP + i);    // column of the pair inside the chunk
                                u2[i] = add2(make_float2(v[e], v[e + 1]), b2);
                                const float2 z = mul2(u2[i], g2);
                                h2[i] = mul2(z, gelu_cdf2(z));
                            }
                            if (nb >= 32) {
                                const float2 r2 = splat2(rsA);
#pragma unroll
                                for (int i = 0; i < HP; ++i) h2[i] = mul2(h2[i], r2);
                            } else {
#pragma unroll
                                for (int i = 0; i < HP; ++i) {
                                    const int e = 2 * (hh * HP + i);
                                    h2[i] = mul2(h2[i], make_float2(e < nb ? rsA : rsB, e + 1 < nb ? rsA : rsB));
                                }
                            }
                            if (hh == 0) pre_stage();
                            stage_bf16_pairs<HP>(panel0, et, (c & 1) * 4 + hh * (HP / 4), u2);
                            stage_bf16_pairs<HP>(panel0 + PANEL_BYTES, et, (c & 1) * 4 + hh * (HP / 4), h2);
                        }
                    } else {
                        // saved pre-gate fc1 output u (zero for rows / tokens out of range: TMA fill)
                        const float2 g2 = splat2(gate_j);
                        float2 cdg2 = splat2(0.f), cdb2 = splat2(0.f);
#pragma unroll
                        for (int hh = 0; hh < HALVES; ++hh) {
                            float2 du2[HP];
                            if (nb >= 32) {
                                // one DropPath multiplier for the whole chunk: fold it into the per-chunk constants
                                const float2 rg2 = splat2(rsA * gate_j);
#pragma unroll
                                for (int i = 0; i < HP; ++i) {
                                    const int pi = hh * HP + i;
                                    const float2 u = unpack_bf16x2(packed_word(ax, pi));
                                    float2 Phi, dgelu;
                                    gelu_terms2(mul2(u, g2), Phi, dgelu);
                                    const float2 w = mul2(make_float2(v[2 * pi], v[2 * pi + 1]), dgelu);    // dh/rs * gelu'(u g)
                                    cdg2 = fma2(w, u, cdg2);
                                    du2[i] = mul2(w, rg2);                                                // du
                                    cdb2 = add2(cdb2, du2[i]);                                            // d bias[j]
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < HP; ++i) {
                                    const int pi = hh * HP + i;
                                    const float2 u = unpack_bf16x2(packed_word(ax, pi));
                                    float2 Phi, dgelu;
                                    gelu_terms2(mul2(u, g2), Phi, dgelu);
                                    const float2 rs2 = make_float2(2 * pi < nb ? rsA : rsB, 2 * pi + 1 < nb ? rsA : rsB);
                                    const float2 w = mul2(mul2(make_float2(v[2 * pi], v[2 * pi + 1]), rs2), dgelu);
                                    cdg2 = fma2(w, u, cdg2);
                                    du2[i] = mul2(w, g2);
                                    cdb2 = add2(cdb2, du2[i]);
                                }
                            }
                            if (hh == 0) pre_stage();
                            stage_bf16_pairs<HP>(panel0, et, (c & 1) * 4 + hh * (HP / 4), du2);
                        }
                        if (nb >= 32) acc_dg = fmaf(cdg2.x + cdg2.y, rsA, acc_dg);                        // d gate[j]
                        else acc_dg += cdg2.x + cdg2.y;
                        acc_db += cdb2.x + cdb2.y;
                    }
                } else if (EPI == EPI_WGRAD) {
                    if (nvalid >= 32) {
                        // One row per lane (the TMEM mapping) makes every red.global.add.v4 touch 32 different gradient rows = 32
                        // LSU wavefronts. The warp regroups its [32 rows][16 columns] half-chunk through 2 KB of shared memory
                        // (float4 slots swizzled by (row >> 1) & 3: conflict-free both ways) so that 4 lanes cover 64 contiguous
                        // bytes of a row: 8 wavefronts per instruction.
                        float* sc = reinterpret_cast<float*>(auxbuf + Cfg::AUX_BYTES) + (etid >> 5) * 512;
                        // deterministic split-K: this split's partial tile goes to its own slice of the workspace (plain stores)
                        const bool det = g.splitk_ws != nullptr;
                        float* out = det ? g.splitk_ws + size_t(split) * g.M * g.N : reinterpret_cast<float*>(g.out0);
                        const int ldo = det ? g.N : g.ld0;
                        const int rbase = m_blk * BM + q * 32;
#pragma unroll
                        for (int hv = 0; hv < 2; ++hv) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                *reinterpret_cast<float4*>(sc + lane * 16 + ((i ^ ((lane >> 1) & 3)) << 2)) =
                                    make_float4(v[hv * 16 + 4 * i] * gs, v[hv * 16 + 4 * i + 1] * gs, v[hv * 16 + 4 * i + 2] * gs,
                                                v[hv * 16 + 4 * i + 3] * gs);
                            __syncwarp();
#pragma unroll
                            for (int ps = 0; ps < 4; ++ps) {
                                const int rl = ps * 8 + (lane >> 2), idx = lane & 3;
                                const float4 w = *reinterpret_cast<const float4*>(sc + rl * 16 + ((idx ^ ((rl >> 1) & 3)) << 2));
                                if (rbase + rl < g.M) {
                                    float* dst = out + size_t(rbase + rl) * ldo + col0 + hv * 16 + idx * 4;
                                    if (det) *reinterpret_cast<float4*>(dst) = w;
                                    else asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(w.x), "f"(w.y),
                                                      "f"(w.z), "f"(w.w) : "memory");
                                }
                            }
                            __syncwarp();
                        }
                    } else if (row_ok && nvalid > 0) {
                        if (g.splitk_ws != nullptr) {
                            float* dst = g.splitk_ws + (size_t(split) * g.M + row) * g.N + col0;
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (i < nvalid) dst[i] = v[i] * gs;
                        } else {
                            float* dst = reinterpret_cast<float*>(g.out0) + size_t(row) * g.ld0 + col0;
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (i < nvalid) atomicAdd(dst + i, v[i] * gs);
                        }
                    }
                } else if (EPI == EPI_PATCH) {
                    // rows are (image b, patch l); output row skips the cls slot of each image.
                    if (row_ok && nvalid > 0) {
                        const int b = row / g.tokens, l = row % g.tokens;
                        const float mk = __ldg(g.rowmask + row);
                        float add[32];        // positional embedding of a kept patch, or the mask token of a removed one
                        if (mk != 0.f) {
                            load_f32x32(g.mask_token + col0, add, nvalid);
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = add[i] * cs[i];
                        } else {
                            load_f32x32(g.pos + size_t(1 + l) * g.N + col0, add, nvalid);
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = (v[i] + cb[i] + add[i]) * cs[i];
                        }
                        const size_t orow = size_t(b) * (g.tokens + 1) + 1 + l;
                        store_bf16x32(reinterpret_cast<__nv_bfloat16*>(g.out0) + orow * g.ld0 + col0, v, nvalid);
                    }
                } else if (EPI == EPI_DECODER) {
                    // rows are (image b, token t) incl. cls (t == 0, ignored). vision_transformer.py:720-729
                    const int tk = g.tokens + 1;
                    const int b = row / tk, tt = row % tk;
                    const bool live = row_ok && tt > 0 && nvalid > 0;
                    float mk = 0.f;
                    float tg[32];
                    if (live) {
                        const int pr = b * g.tokens + tt - 1;
                        mk = __ldg(g.rowmask + pr);
                        if (mk != 0.f) load_f32x32(g.target + size_t(pr) * g.N + col0, tg, nvalid);
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float s = 0.f;
                        if (mk != 0.f && i < nvalid) {
                            const float d = v[i] + cb[i] - tg[i];
                            loss_acc += fabsf(d);
                            s = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
                        }
                        v[i] = s;
                    }
                    // the sign matrix leaves through the staged panel + TMA store where the output allows it (one row per lane in a
                    // direct global store costs 32 LSU wavefronts per instruction); rows >= M / columns >= N are clipped there
                    if (TMA_OUT) stage_bf16x32(panel0, et, c & 1, v);
                    else if (row_ok && nvalid > 0) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(g.out0) + size_t(row) * g.ld0 + col0, v, nvalid);
                }

                if (TMA_OUT) {
                    // panel pj of this tile (chunks 2 pj, 2 pj + 1) is complete once both groups of the pipeline have staged
                    // their chunk
                    fence_proxy_async_smem();
                    if (NPIPE == 1 && elected) bulk_wait_read<0>();          // the other slot's previous store has been read out
                    named_bar_sync(2 + pp, 256);
                    if (elected && n0 + pj * 64 < g.N && !GEMM_DBG(1)) {
                        tma_store_2d(&tma_o0, panel0, n0 + pj * 64, m_blk * BM);
                        if (Cfg::NOUT == 2) tma_store_2d(&tma_o1, panel0 + PANEL_BYTES, n0 + pj * 64, m_blk * BM);
                        bulk_commit();
                    }
                    // every thread has consumed this panel's aux slot (it was read before the barrier): refill it AUX_SLOTS panels ahead
                    if (TMA_OUT == 2 && elected) { fence_proxy_async_smem(); issue_aux(P + AUX_SLOTS); }
                    ++panel_ctr;
                }
            }

            if (EPI == EPI_FC2_DGRAD) {
                // each thread owns one hidden unit: its token sums of this tile go out as one partial row per (n tile, group)
                if (row_ok) {
                    g.colpart0[size_t(n_blk * NG + grp) * g.M + row] = acc_dg;
                    g.colpart1[size_t(n_blk * NG + grp) * g.M + row] = acc_db;
                }
            }
            if (EPI == EPI_DECODER) {
                const float s = warp_sum(loss_acc);
                if (lane == 0 && m_blk < m_tiles) g.colpart0[size_t(mn) * EPI_WARPS + ew] = s;
            }
        }
        if (TMA_OUT && elected) bulk_wait<0>();    // all output panels fully written before the CTA retires
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();          // neither CTA may retire while the pair's MMAs / remote arrivals can still touch it
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || p == nullptr) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 tensor map of rank `rank`; dims/strides innermost first; strides in elements (stride[0] must be 1).
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return 1001;
    cuuint64_t gdim[5], gstride[5];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstride[i - 1] = strides_elems[i] * 2;
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstride, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1002;
}

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int BN, int A_MN, int B_MN, int EPI, int TMA_OUT, int CG>
static int launch_gemm_inst(const void* A, int lda, const void* B, int ldb, const GemmArgs& g, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, EPI, TMA_OUT, CG>;
    static bool configured = false;
    auto kfn = gemm_kernel<BN, A_MN, B_MN, EPI, TMA_OUT, CG>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return int(e);
        configured = true;
    }
    CUtensorMap ta, tb, to0, to1, tx;
    {
        // K-major: tensor [rows, K] row-major -> dims {K, rows}, box {64, 128|BN/CG}
        // MN-major: tensor [K, rows] row-major -> dims {rows, K}, box {64, 64}
        uint64_t dims[2], str[2];
        uint32_t box[2];
        if (A_MN) { dims[0] = g.M; dims[1] = g.K; box[0] = 64; box[1] = BK; }
        else      { dims[0] = g.K; dims[1] = g.M; box[0] = BK; box[1] = BM; }
        str[0] = 1; str[1] = uint64_t(lda);
        int r = make_tmap_bf16(&ta, A, 2, dims, str, box);
        if (r) return r;
        if (B_MN) { dims[0] = g.N; dims[1] = g.K; box[0] = 64; box[1] = BK; }
        else      { dims[0] = g.K; dims[1] = g.N; box[0] = BK; box[1] = BN / CG; }
        str[1] = uint64_t(ldb);
        r = make_tmap_bf16(&tb, B, 2, dims, str, box);
        if (r) return r;
        to0 = ta; to1 = ta; tx = ta;
        if (TMA_OUT == 2) {
            const void* xp = (EPI == EPI_STORE) ? static_cast<const void*>(g.res) : static_cast<const void*>(g.aux);
            dims[0] = g.N; dims[1] = g.M; box[0] = 64; box[1] = BM;
            str[1] = uint64_t(EPI == EPI_STORE ? g.ldres : g.ldaux);
            r = make_tmap_bf16(&tx, xp, 2, dims, str, box);
            if (r) return r;
        }
        if (TMA_OUT) {
            // bf16 output [M, N] with row pitch ld: box = one swizzled [128 rows][64 cols] panel; the hardware clips
            // rows >= M and columns >= N
            dims[0] = g.N; dims[1] = g.M; box[0] = 64; box[1] = BM;
            str[1] = uint64_t(g.ld0);
            r = make_tmap_bf16(&to0, g.out0, 2, dims, str, box);
            if (r) return r;
            if (Cfg::NOUT == 2) {
                str[1] = uint64_t(g.ld1);
                r = make_tmap_bf16(&to1, g.out1, 2, dims, str, box);
                if (r) return r;
            }
        }
    }
    const int m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
    const int mt = (m_tiles + CG - 1) / CG;
    const int total = mt * n_tiles * (g.k_splits > 0 ? g.k_splits : 1);
    const int slots = num_sms() / CG;
    const int grid = (total < slots ? total : slots) * CG;
    return int(launch_k(kfn, dim3(grid), dim3(Cfg::THREADS), size_t(Cfg::SMEM_BYTES), stream, CG, ta, tb, to0, to1, tx, g));
}

// pick the N tile: fewest (waves x tile width) with a penalty for narrow tiles (shared-memory bandwidth per MMA)
static int pick_bn(int M, int N, int splits_ok) {
    const int cands[4] = {256, 192, 128, 64};
    const int m_tiles = (M + BM - 1) / BM;
    int best = 128; double best_cost = 1e30;
    for (int i = 0; i < 4; ++i) {
        const int bn = cands[i];
        const int n_tiles = (N + bn - 1) / bn;
        const long tiles = long(m_tiles) * n_tiles;
        double cost;
        if (splits_ok) {
            // split-K fills the machine whatever the tile count: cost = padded work, narrow tiles are smem-bound
            cost = double(tiles) * bn * (bn <= 64 ? 2.0 : (bn <= 128 ? 1.25 : 1.0));
            if (tiles > num_sms()) cost *= double((tiles + num_sms() - 1) / num_sms()) * num_sms() / tiles;
        } else {
            // measured (tools/gemm_microbench.py): a 128-wide tile is shared-memory-bandwidth bound (A is re-read for every
            // MMA and the TMA fill competes with it), 192 / 256 reach 1.25x / 1.35x its main-loop rate
            const long waves = (tiles + num_sms() - 1) / num_sms();
            cost = double(waves) * bn * (bn <= 64 ? 1.9 : (bn <= 128 ? 1.30 : (bn <= 192 ? 1.05 : 1.0)));
        }
        if (cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

// CTA pairs (cta_group::2) per epilogue: bit e of the mask allows them for epilogue e. Default: the main-loop-bound epilogues
// (plain / residual stores, patch embed, decoder). The transposed-hidden epilogues (FC1, FC2_DGRAD) are bound by their
// epilogue arithmetic and lose a little when two CTAs have to hand their accumulators back together (measured in situ:
// +11 % / +9 %), and the split-K weight gradients gain nothing. OFB_GEMM_PAIR_MASK overrides (0 = never).
static int pair_mask() {
    static int mask = -1;
    if (mask < 0) {
        const char* e = getenv("OFB_GEMM_PAIR_MASK");
        mask = e ? atoi(e) : ((1 << EPI_STORE) | (1 << EPI_PATCH) | (1 << EPI_DECODER));
    }
    return mask;
}
static bool pair_ok(int epi, int M, int bn, bool b_mn) {
    // a pair needs at least two 128-row tiles, a B half that is a whole number of 64-wide MN atoms, and a tile width >= 128
    return ((pair_mask() >> epi) & 1) && M > BM && bn >= 128 && (!b_mn || bn % 128 == 0);
}

// split-K factor of a weight-gradient GEMM (reduction over K tokens): fills the SMs with (tiles x splits) work items and leaves
// no split without a k block. Exported (ofb_gemm_wgrad_splits) so that the caller can size the deterministic split-K workspace.
int wgrad_plan(int M, int N, int K, int b_mn, int bn_hint) {
    const int bn = bn_hint > 0 ? bn_hint : pick_bn(M, N, true);
    // same predicate as launch_gemm_bn: a CTA pair covers 256 rows and there are num_sms / 2 pairs
    const bool pair = pair_ok(EPI_WGRAD, M, bn, b_mn != 0);
    const int m_tiles = (M + BM - 1) / BM;
    const int tiles = (pair ? (m_tiles + 1) / 2 : m_tiles) * ((N + bn - 1) / bn);
    int s = (pair ? num_sms() / 2 : num_sms()) / tiles;
    const int kb = (K + BK - 1) / BK;
    if (s < 1) s = 1;
    if (s > kb) s = kb;
    const int per = (kb + s - 1) / s;
    return (kb + per - 1) / per;              // splits that actually own k blocks
}

template <int A_MN, int B_MN, int EPI, int TMA_OUT>
static int launch_gemm_bn(int bn, const void* A, int lda, const void* B, int ldb, const GemmArgs& g, cudaStream_t s) {
    const bool pair = pair_ok(EPI, g.M, bn, B_MN != 0);
    if (pair) {
        switch (bn) {
            case 256: return launch_gemm_inst<256, A_MN, B_MN, EPI, TMA_OUT, 2>(A, lda, B, ldb, g, s);
            case 192: if constexpr (!B_MN) return launch_gemm_inst<192, A_MN, B_MN, EPI, TMA_OUT, 2>(A, lda, B, ldb, g, s); else break;
            default:  return launch_gemm_inst<128, A_MN, B_MN, EPI, TMA_OUT, 2>(A, lda, B, ldb, g, s);
        }
    }
    switch (bn) {
        case 256: return launch_gemm_inst<256, A_MN, B_MN, EPI, TMA_OUT, 1>(A, lda, B, ldb, g, s);
        case 192: return launch_gemm_inst<192, A_MN, B_MN, EPI, TMA_OUT, 1>(A, lda, B, ldb, g, s);
        case 128: return launch_gemm_inst<128, A_MN, B_MN, EPI, TMA_OUT, 1>(A, lda, B, ldb, g, s);
        default:  return launch_gemm_inst<64, A_MN, B_MN, EPI, TMA_OUT, 1>(A, lda, B, ldb, g, s);
    }
}

// rows of the [rows][M] token-sum partial buffers the FC2_DGRAD epilogue writes for N tokens at column tile bn: one per
// (column tile, epilogue column group)
int mlp_partial_rows(int N, int bn) {
    if (bn != 64 && bn != 128 && bn != 192 && bn != 256) return -1;
    const int ng = (bn % 128 == 0) ? 4 : 2;
    return ng * ((N + bn - 1) / bn);
}

static bool tma_out_ok(const void* p, int ld) { return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && ld % 8 == 0; }

int launch_gemm(int epi, int a_mn, int b_mn, int bn_hint, const void* A, int lda, const void* B, int ldb, GemmArgs g,
                cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
    int bn = bn_hint > 0 ? bn_hint : pick_bn(g.M, g.N, epi == EPI_WGRAD);
    // data-gradient GEMMs read the weight as an MN-major B operand, whose half per CTA of a pair must be whole 64-wide atoms:
    // with a long reduction a 256-wide pair tile beats the 192-wide single-CTA tile even at 25 % column padding (measured in
    // situ at N = 384: K = 1536 94 -> 79 us, K = 1152 75 -> 65 us; K = 384 is epilogue-bound and stays on 192)
    // ... unless the padded 256-wide tiles also lose a wave: per k step a pair tile costs ~(128 + bn) / 2 operand-read clocks, so
    // compare waves x (128 + bn) of the 256- and 128-wide pair tiles (N = 384, M = 50 432: 6 x 384 against 8 x 256 - measured
    // tools/dgrad_bn_ab.py: K = 1536 80.6 -> 76.3 us, K = 1152 66.2 -> 60.2 us; N = 768 and N = 192 stay on 256)
    if (bn_hint <= 0 && epi == EPI_STORE && b_mn && g.K >= 768 && g.N > 128 && pair_ok(EPI_STORE, g.M, 256, true)) {
        const long mtp = ((g.M + BM - 1) / BM + 1) / 2, slots = num_sms() / 2;
        auto cost = [&](int w) { return ((mtp * ((g.N + w - 1) / w) + slots - 1) / slots) * long(128 + w); };
        bn = cost(128) < cost(256) ? 128 : 256;
    }
    if (epi == EPI_WGRAD) {
        if (g.k_splits <= 0) g.k_splits = wgrad_plan(g.M, g.N, g.K, b_mn, bn_hint);
        if (g.splitk_ws != nullptr) {
            // every split must own at least one k block (an empty split would leave its workspace slice unwritten) and the
            // vectorised partial stores need 16-byte aligned rows
            const int kb = (g.K + BK - 1) / BK;
            const int per = (kb + g.k_splits - 1) / g.k_splits;
            if ((g.k_splits - 1) * per >= kb || g.N % 4 != 0) return 1006;
        }
        // the reduction runs over tokens; operands are either token-major activations (MN-major here) or transposed
        // hidden activations [hidden, tokens] (K-major here)
        if (a_mn && b_mn) return launch_gemm_bn<1, 1, EPI_WGRAD, 0>(bn, A, lda, B, ldb, g, stream);
        if (a_mn) return launch_gemm_bn<1, 0, EPI_WGRAD, 0>(bn, A, lda, B, ldb, g, stream);
        if (b_mn) return launch_gemm_bn<0, 1, EPI_WGRAD, 0>(bn, A, lda, B, ldb, g, stream);
        return 1003;
    }
    g.k_splits = 1;
    const bool tma_out = !g.out_fp32 && tma_out_ok(g.out0, g.ld0);
    switch (epi) {
        case EPI_STORE: {
            // b_mn: data-gradient GEMMs read the nn.Linear weight [N_out, K_in] directly as an MN-major operand
            // a_mn: the A operand is a transposed hidden activation [hidden, tokens]
            // mode 2 (residual tile through TMA) needs a TMA-able residual; otherwise the direct path handles it
            const int mode = !tma_out ? 0 : (g.res == nullptr ? 1 : (tma_out_ok(g.res, g.ldres) ? 2 : 0));
#define OFB_STORE_CASE(AM, BMJ)                                                                              \
            if (mode == 2) return launch_gemm_bn<AM, BMJ, EPI_STORE, 2>(bn, A, lda, B, ldb, g, stream);          \
            if (mode == 1) return launch_gemm_bn<AM, BMJ, EPI_STORE, 1>(bn, A, lda, B, ldb, g, stream);          \
            return launch_gemm_bn<AM, BMJ, EPI_STORE, 0>(bn, A, lda, B, ldb, g, stream);
            if (a_mn && b_mn) { OFB_STORE_CASE(1, 1) }
            if (a_mn) { OFB_STORE_CASE(1, 0) }
            if (b_mn) { OFB_STORE_CASE(0, 1) }
            OFB_STORE_CASE(0, 0)
#undef OFB_STORE_CASE
        }
        case EPI_FC1:
            if (a_mn || b_mn) return 1003;
            if (!tma_out || !tma_out_ok(g.out1, g.ld1)) return 1005;
            return launch_gemm_bn<0, 0, EPI_FC1, 1>(bn, A, lda, B, ldb, g, stream);
        case EPI_FC2_DGRAD:
            if (!a_mn || b_mn) return 1003;
            if (!tma_out || !tma_out_ok(g.aux, g.ldaux)) return 1005;
            return launch_gemm_bn<1, 0, EPI_FC2_DGRAD, 2>(bn, A, lda, B, ldb, g, stream);
        case EPI_PATCH:
            if (a_mn || b_mn) return 1003;
            return launch_gemm_bn<0, 0, EPI_PATCH, 0>(bn, A, lda, B, ldb, g, stream);
        case EPI_DECODER:
            if (a_mn || b_mn) return 1003;
            if (tma_out) return launch_gemm_bn<0, 0, EPI_DECODER, 1>(bn, A, lda, B, ldb, g, stream);
            return launch_gemm_bn<0, 0, EPI_DECODER, 0>(bn, A, lda, B, ldb, g, stream);
        default: return 1004;
    }
}

#ifdef OFB_GEMM_DEBUG
int gemm_debug_flags(int flags) { return int(cudaMemcpyToSymbol(g_gemm_dbg, &flags, sizeof(int))); }
#endif

}  // namespace ofb

#ifdef OFB_GEMM_DEBUG
extern "C" int ofb_debug_gemm_flags(int flags) { return ofb::gemm_debug_flags(flags); }
#endif
