// Dense bf16 GEMM with fused epilogue on tcgen05 / TMEM / TMA (sm_100a), persistent:
//     C[M][N] = epilogue(A[M][K] * W[N][K]^T + bias[N])          (nn.Linear layout, fp32 accumulate)
// Used for the ViT patch merger (Linear 5120->5120 + GELU(erf), Linear 5120->out_hidden fused with the
// `argsort(window_index)` row un-permute) and for the dense projections of the ViT tower and of the chunk prefill
// (qkv / proj / gate-up / down), where it replaces cuBLAS + separate bias / activation / residual kernels.
//
// One CTA per SM walks output tiles (128 x BN, BN = 64 / 128 / 256) round-robin; three pipelines run side by side:
//   warp 0     TMA producer: A [128 x 64] and W [BN x 64] panels (128B swizzle) into a STAGES-deep shared-memory ring
//   warp 1     UMMA issuer: 4 x (M128, N = BN, K16) per stage into one of TWO TMEM accumulators (2 x BN columns)
//   warps 2-5  epilogue: tcgen05.ld the finished accumulator while the issuer already fills the other one
//              (+ bias, round to bf16, optional GELU / residual add / SwiGLU pairing, 128-bit stores, optional row scatter)
// Measured on B200 (benchmarks/hw_probe.cu): an SS UMMA M128 N128 K16 costs 121.6 cycles, N256 171.3 cycles -- the shared-memory
// operand fetch, not the math, bounds the instruction -- so the widest tile that still fills the SMs is chosen per shape.
#include <math.h>

#include "common.cuh"

namespace svlm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_SMEM_LIMIT = 227 * 1024;

enum { EPI_NONE = 0, EPI_GELU = 1, EPI_RESIDUAL = 2, EPI_SWIGLU = 3 };

template <int BN>
struct GemmCfg {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int W_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGES = (GEMM_SMEM_LIMIT - 2048) / (A_BYTES + W_BYTES) > 8 ? 8 : (GEMM_SMEM_LIMIT - 2048) / (A_BYTES + W_BYTES);
    static constexpr int SMEM = 1024 + STAGES * (A_BYTES + W_BYTES) + 512;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;          // two accumulators
};

struct GemmParams {
    int M, N, K, epi;
    int m_tiles, n_tiles;
    const uint16_t* bias;
    const uint16_t* residual;   // EPI_RESIDUAL: bf16 [M][N], C = residual + bf16(acc + bias)   (may alias C)
    const int32_t* out_rows;
    uint16_t* C;
    int ldc;                    // elements between rows of C (N, or N / 2 for EPI_SWIGLU)
};

// Tile order of the persistent schedulers: m fastest inside groups of 8 m-tiles, so that the ~148 tiles in flight cover a squarish
// region (8 x ~18 tiles) whose A rows and W panels both stay in L2, instead of one W panel against every row of A.
__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int& mt, int& nt) {
    constexpr int GM = 8;
    const int group = GM * n_tiles;
    const int g = tile / group, r = tile - g * group;
    const int first = g * GM;
    const int gm = min(GM, m_tiles - first);
    mt = first + r % gm;
    nt = r / gm;
}

// One accumulator row (TMEM lane) of a 128 x BN tile: + bias, round to bf16, optional GELU / residual add / SwiGLU pairing, 128-bit stores.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t taddr, int m, bool valid, int64_t orow, int n0) {
    if (p.epi == EPI_SWIGLU) {
        // W rows are interleaved in blocks of BN/2: tile columns [0, BN/2) = gate_j, [BN/2, BN) = up_j of the same j
        constexpr int HALF = BN / 2;
#pragma unroll
        for (int c = 0; c < HALF / 16; ++c) {
            uint32_t g16[16], u16[16];
            tmem_ld16(taddr + c * 16, g16);
            tmem_ld16(taddr + HALF + c * 16, u16);
            tmem_wait_ld();
            const int j = (n0 >> 1) + c * 16;               // output column
            if (!valid || j >= (p.N >> 1)) continue;
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float r2[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float gt = round_bf16(__uint_as_float(g16[2 * i + e])), up = round_bf16(__uint_as_float(u16[2 * i + e]));
                    r2[e] = round_bf16(gt / (1.0f + __expf(-gt))) * up;
                }
                o[i] = pack_bf16(r2[0], r2[1]);
            }
            uint16_t* dst = p.C + orow * p.ldc + j;
            *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
        }
    } else {
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t a32[32];
            tmem_ld32(taddr + c * 32, a32);
            tmem_wait_ld();
            if (!valid) continue;
            uint16_t* dst = p.C + orow * p.ldc + n0 + c * 32;
#pragma unroll
            for (int v8 = 0; v8 < 4; ++v8) {
                const int n = n0 + c * 32 + v8 * 8;
                if (n >= p.N) continue;
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(a32[v8 * 8 + i]);
                if (p.bias) {
                    const uint4 bv = __ldg(reinterpret_cast<const uint4*>(p.bias + n));
                    const uint32_t* pb = &bv.x;
#pragma unroll
                    for (int i = 0; i < 4; ++i) { x[2 * i] += bf16_lo(pb[i]); x[2 * i + 1] += bf16_hi(pb[i]); }
                }
                if (p.epi == EPI_GELU) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float y = round_bf16(x[i]);                 // Linear output is bf16 before nn.GELU sees it
                        x[i] = 0.5f * y * (1.0f + erff(y * 0.70710678118654752f));
                    }
                } else if (p.epi == EPI_RESIDUAL) {
                    const uint4 rv = *reinterpret_cast<const uint4*>(p.residual + (int64_t)m * p.ldc + n);
                    const uint32_t* pr = &rv.x;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                         // hidden = residual + bf16(linear)
                        x[2 * i] = round_bf16(x[2 * i]) + bf16_lo(pr[i]);
                        x[2 * i + 1] = round_bf16(x[2 * i + 1]) + bf16_hi(pr[i]);
                    }
                }
                uint4 o;
                o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]); o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
                *reinterpret_cast<uint4*>(dst + v8 * 8) = o;
            }
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sW = smem + Cfg::STAGES * Cfg::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sW + Cfg::STAGES * Cfg::W_BYTES);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tfull = empty + Cfg::STAGES;       // [2] accumulator ready for the epilogue
    uint64_t* tempty = tfull + 2;                // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_blocks = (p.K + GEMM_BK - 1) / GEMM_BK;
    const int n_tiles_total = p.m_tiles * p.n_tiles;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer (whole warp walks the loop, one elected lane issues)
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
            int mt, nt;
            tile_coords(tile, p.m_tiles, p.n_tiles, mt, nt);
            const int m0 = mt * GEMM_BM, n0 = nt * BN;
            for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                const int s = it % Cfg::STAGES;
                mbar_wait(&empty[s], ((it / Cfg::STAGES) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[s], Cfg::A_BYTES + Cfg::W_BYTES);
                    tma_load_2d(sA + s * Cfg::A_BYTES, &tmA, &full[s], kb * GEMM_BK, m0);
                    tma_load_2d(sW + s * Cfg::W_BYTES, &tmW, &full[s], kb * GEMM_BK, n0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ---- UMMA issuer: whole warp in uniform control flow, one elected lane issues (descriptors stay in uniform registers)
        constexpr uint32_t IDESC = umma_idesc_bf16(GEMM_BM, BN, 0, 0);
        int it = 0, local = 0;
        for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++local) {
            const int acc = local & 1;
            mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d = tmem + (uint32_t)acc * BN;
            for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                const int s = it % Cfg::STAGES;
                mbar_wait(&full[s], (it / Cfg::STAGES) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a = smem_u32(sA + s * Cfg::A_BYTES), w = smem_u32(sW + s * Cfg::W_BYTES);
#pragma unroll
                    for (int kk = 0; kk < GEMM_BK / 16; ++kk)
                        umma_ss(d, umma_desc_sw128(a + kk * 32, 16, 1024), umma_desc_sw128(w + kk * 32, 16, 1024), IDESC,
                                (kb > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&empty[s]);
                    if (kb == k_blocks - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue: thread = one accumulator row (TMEM lane), 32 columns per tcgen05.ld
        const int q4 = warp & 3;
        const int row = q4 * 32 + lane;
        int local = 0;
        for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++local) {
            const int acc = local & 1;
            int mt, nt;
            tile_coords(tile, p.m_tiles, p.n_tiles, mt, nt);
            const int m0 = mt * GEMM_BM, n0 = nt * BN;
            const int m = m0 + row;
            const bool valid = m < p.M;
            const int64_t orow = valid ? (p.out_rows ? (int64_t)__ldg(p.out_rows + m) : (int64_t)m) : 0;
            mbar_wait(&tfull[acc], (local >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)acc * BN;
            epilogue_tile<BN>(p, taddr, m, valid, orow, n0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, int sms, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN>::SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm_bf16_kernel) failed: %s", cudaGetErrorString(e));
            return SVLM_ECUDA;
        }
        attr_done = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = tiles < sms ? tiles : sms;
    gemm_bf16_kernel<BN><<<grid, GEMM_THREADS, GemmCfg<BN>::SMEM, st>>>(ta, tw, p);
    return launch_status("gemm_bf16_kernel");
}


// ---------------------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs computes 256 x BN tiles.  Each CTA loads its own 128 rows of A and HALF of
// the W panel (BN / 2 rows) per K-block -- half the shared-memory fetch per flop of the single-CTA kernel, which is what bounds a
// tcgen05.mma -- and keeps its 128 accumulator rows in its own TMEM.  The leader (rank 0) issues the MMAs; its `full`
// barriers collect the TMA bytes of BOTH CTAs; `empty` / `tfull` are signalled in both CTAs by multicast commits; the epilogue
// warps of both CTAs release an accumulator by arriving on the leader's `tempty`.
// ---------------------------------------------------------------------------------------------------------------------------------
template <int BN>
struct Gemm2Cfg {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int W_BYTES = (BN / 2) * GEMM_BK * 2;                 // this CTA's half of the W panel
    static constexpr int STAGES = (GEMM_SMEM_LIMIT - 2048) / (A_BYTES + W_BYTES) > 8 ? 8 : (GEMM_SMEM_LIMIT - 2048) / (A_BYTES + W_BYTES);
    static constexpr int SMEM = 1024 + STAGES * (A_BYTES + W_BYTES) + 512;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
    using Cfg = Gemm2Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sW = smem + Cfg::STAGES * Cfg::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sW + Cfg::STAGES * Cfg::W_BYTES);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tfull = empty + Cfg::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int k_blocks = (p.K + GEMM_BK - 1) / GEMM_BK;
    const int n_tiles_total = p.m_tiles * p.n_tiles;            // m_tiles counts 256-row tiles here
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8); }      // 4 epilogue warps x 2 CTAs
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();                                          // both CTAs' barriers exist before anything remote touches them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer of this CTA: its A rows, its half of W; bytes are counted on the LEADER's full barrier
        int it = 0;
        for (int tile = pair; tile < n_tiles_total; tile += n_pairs) {
            int mt, nt;
            tile_coords(tile, p.m_tiles, p.n_tiles, mt, nt);
            const int m0 = mt * 256 + (int)rank * GEMM_BM, n0 = nt * BN + (int)rank * (BN / 2);
            for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                const int s = it % Cfg::STAGES;
                mbar_wait(&empty[s], ((it / Cfg::STAGES) & 1) ^ 1);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(&full[s], 2 * (Cfg::A_BYTES + Cfg::W_BYTES));
                    const uint32_t bar = mapa_shared(smem_u32(&full[s]), 0);
                    tma_load_2d_pair(sA + s * Cfg::A_BYTES, &tmA, bar, kb * GEMM_BK, m0);
                    tma_load_2d_pair(sW + s * Cfg::W_BYTES, &tmW, bar, kb * GEMM_BK, n0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ---- UMMA issuer (leader CTA only): one instruction drives the tensor cores of both SMs
            constexpr uint32_t IDESC = umma_idesc_bf16(256, BN, 0, 0);
            int it = 0, local = 0;
            for (int tile = pair; tile < n_tiles_total; tile += n_pairs, ++local) {
                const int acc = local & 1;
                mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % Cfg::STAGES;
                    mbar_wait(&full[s], (it / Cfg::STAGES) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a = smem_u32(sA + s * Cfg::A_BYTES), w = smem_u32(sW + s * Cfg::W_BYTES);
#pragma unroll
                        for (int kk = 0; kk < GEMM_BK / 16; ++kk)
                            umma_ss_pair(d, umma_desc_sw128(a + kk * 32, 16, 1024), umma_desc_sw128(w + kk * 32, 16, 1024), IDESC,
                                         (kb > 0 || kk > 0) ? 1u : 0u);
                        umma_commit_pair(&empty[s], 3);                       // frees the slot in BOTH CTAs
                        if (kb == k_blocks - 1) umma_commit_pair(&tfull[acc], 3);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ---- epilogue of this CTA's 128 rows
        const int q4 = warp & 3;
        const int row = q4 * 32 + lane;
        const uint32_t tempty_leader[2] = {mapa_shared(smem_u32(&tempty[0]), 0), mapa_shared(smem_u32(&tempty[1]), 0)};
        int local = 0;
        for (int tile = pair; tile < n_tiles_total; tile += n_pairs, ++local) {
            const int acc = local & 1;
            int mt, nt;
            tile_coords(tile, p.m_tiles, p.n_tiles, mt, nt);
            const int m0 = mt * 256 + (int)rank * GEMM_BM, n0 = nt * BN;
            const int m = m0 + row;
            const bool valid = m < p.M;
            const int64_t orow = valid ? (p.out_rows ? (int64_t)__ldg(p.out_rows + m) : (int64_t)m) : 0;
            mbar_wait(&tfull[acc], (local >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)acc * BN;
            epilogue_tile<BN>(p, taddr, m, valid, orow, n0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader[acc]);
        }
    }
    tc_fence_before();
    cluster_sync_all();                                          // nobody leaves while the peer may still signal its barriers
    if (warp == 1) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem);
}

template <int BN>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, int sms, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg<BN>::SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm2_bf16_kernel) failed: %s", cudaGetErrorString(e));
            return SVLM_ECUDA;
        }
        attr_done = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int pairs = tiles < sms / 2 ? tiles : sms / 2;
    gemm2_bf16_kernel<BN><<<2 * pairs, GEMM_THREADS, Gemm2Cfg<BN>::SMEM, st>>>(ta, tw, p);
    return launch_status("gemm2_bf16_kernel");
}

// Tile width for a shape: modelled time = waves over the SMs x K-steps x measured cycles per (M128, N = BN, K16) UMMA
// (121.6 for N128, 171.3 for N256 on B200; ~95 for N64, whose A fetch dominates) + a fixed per-tile cost.
// Configuration for a shape: BN (64 / 128 / 256) for the single-CTA kernel, 1000 + BN for the CTA-pair kernel.  Modelled time =
// waves over the SMs (pairs: over SMs / 2) x (K-steps x measured cycles per UMMA + a per-tile cost for the un-overlapped first fill
// and last epilogue).  Cycles per K16 step on B200: single CTA M128 N64/128/256 = ~95 / 121.6 / 171.3 (hw_probe), CTA pair M256
// N64/128/256 = ~75 / 100 / 137 (fitted to benchmarks/gemm_bench.py).
static int pick_bn(int M, int N, int K, int sms, int force) {
    if (force == 64 || force == 128 || force == 256 || force == 1064 || force == 1128 || force == 1256) return force;
    const int cand[6] = {64, 128, 256, 1064, 1128, 1256};
    const double cyc[6] = {95.0, 121.6, 171.3, 75.0, 100.0, 137.0};
    int best = 128;
    double best_t = 1e30;
    for (int i = 0; i < 6; ++i) {
        const bool pair = cand[i] >= 1000;
        const int bn = pair ? cand[i] - 1000 : cand[i];
        if (pair && M <= 160) continue;                       // a pair tile is 256 rows tall: half of it would be padding
        const int mt = pair ? (M + 255) / 256 : (M + GEMM_BM - 1) / GEMM_BM;
        const int nt = (N + bn - 1) / bn;
        const int slots = pair ? sms / 2 : sms;
        const int waves = (mt * nt + slots - 1) / slots;
        const double t = waves * ((double)((K + 15) / 16) * cyc[i] + 1500.0 + 12.0 * bn);
        if (t < best_t) { best_t = t; best = cand[i]; }
    }
    return best;
}

}  // namespace svlm

using namespace svlm;

static int gemm_launch_common(const void* A, const void* W, const void* bias, const void* residual, void* C, int M, int N, int K, int epi,
                              const int32_t* out_rows, int force_bn, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(A && W && C, "NULL tensor argument");
    SVLM_REQUIRE(M > 0 && N > 0 && K > 0, "bad M/N/K");
    SVLM_REQUIRE(K % 8 == 0 && N % 8 == 0, "K (%d) and N (%d) must be multiples of 8", K, N);
    SVLM_REQUIRE(epi >= EPI_NONE && epi <= EPI_SWIGLU, "epilogue must be 0 (none), 1 (gelu), 2 (residual add) or 3 (swiglu)");
    SVLM_REQUIRE(epi != EPI_RESIDUAL || residual != nullptr, "residual is NULL");
    SVLM_REQUIRE(epi != EPI_SWIGLU || (N % 32 == 0 && bias == nullptr && out_rows == nullptr), "swiglu epilogue: N %% 32 == 0, no bias, no row scatter");
    SVLM_REQUIRE((((uintptr_t)A | (uintptr_t)W | (uintptr_t)C | (uintptr_t)bias | (uintptr_t)residual) & 15) == 0,
                 "A/W/C/bias/residual must be 16-byte aligned");
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        SVLM_CUDA(cudaGetDevice(&dev));
        SVLM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // swiglu pairs gate / up inside one tile: the weight rows are interleaved in blocks of 64 ([gate 0:64 | up 0:64 | gate 64:128 ...]),
    // which fixes the tile width at 128
    const int cfg = epi == EPI_SWIGLU ? (force_bn == 1128 ? 1128 : 128) : pick_bn(M, N, K, sms, force_bn);
    const bool pair = cfg >= 1000;
    const int bn = pair ? cfg - 1000 : cfg;
    CUtensorMap ta, tw;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {GEMM_BK, GEMM_BM};
        if ((rc = make_tensor_map_bf16(&ta, A, 2, dims, str, box, 1))) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {GEMM_BK, (uint32_t)(pair ? bn / 2 : bn)};
        if ((rc = make_tensor_map_bf16(&tw, W, 2, dims, str, box, 1))) return rc;
    }
    GemmParams p{M, N, K, epi, pair ? (M + 255) / 256 : (M + GEMM_BM - 1) / GEMM_BM, (N + bn - 1) / bn, (const uint16_t*)bias,
                 (const uint16_t*)residual, out_rows, (uint16_t*)C, epi == EPI_SWIGLU ? N / 2 : N};
    if (pair) {
        switch (bn) {
            case 256: return launch_gemm2<256>(ta, tw, p, sms, (cudaStream_t)stream);
            case 128: return launch_gemm2<128>(ta, tw, p, sms, (cudaStream_t)stream);
            default: return launch_gemm2<64>(ta, tw, p, sms, (cudaStream_t)stream);
        }
    }
    switch (bn) {
        case 256: return launch_gemm<256>(ta, tw, p, sms, (cudaStream_t)stream);
        case 128: return launch_gemm<128>(ta, tw, p, sms, (cudaStream_t)stream);
        default: return launch_gemm<64>(ta, tw, p, sms, (cudaStream_t)stream);
    }
}

extern "C" int svlm_gemm_bf16(const void* A, const void* W, const void* bias, void* C, int M, int N, int K, int act,
                              const int32_t* out_rows, void* stream) {
    SVLM_REQUIRE(act == 0 || act == 1, "act must be 0 (none) or 1 (gelu)");
    return gemm_launch_common(A, W, bias, nullptr, C, M, N, K, act, out_rows, 0, stream);
}

extern "C" int svlm_gemm_bf16_ex(const void* A, const void* W, const void* bias, const void* residual, void* C, int M, int N, int K,
                                 int epilogue, const int32_t* out_rows, int tile_n, void* stream) {
    return gemm_launch_common(A, W, bias, residual, C, M, N, K, epilogue, out_rows, tile_n, stream);
}
