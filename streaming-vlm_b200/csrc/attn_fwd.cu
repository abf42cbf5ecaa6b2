// Flash-attention forward for sm_100a, hand-written on tcgen05 / TMEM / TMA.
//
// One kernel serves the three attention shapes of the hot path:
//   * LLM chunk prefill  (Tq ~ 276, causal bottom-right, GQA: the G query heads that share a KV head are
//     packed into the 128-row UMMA M tile as (token, head-in-group) rows)       -- svlm_attn_causal
//   * LLM decode         (Tq = 1, same kernel, KV range split across CTAs + merge kernel, HBM bound)
//   * ViT windowed / full attention (non-causal, packed sequences given by cu_seqlens, head_dim 80 zero-padded
//     to 128 by TMA out-of-bounds fill)                                          -- svlm_vit_attn
//
// CTA = 192 threads, 1 CTA / SM:
//   warp 0     TMA producer: Q tile once, then a 3-stage ring of K and V tiles (128 keys x 128 dims, two
//              64-column panels each, 128B swizzle)
//   warp 1     UMMA issuer (one lane): S(j+1) = Q K_{j+1}^T is issued before O += P_j V_j, so the tensor pipe
//              works on the next score tile while the softmax warps process the current one
//   warps 2-5  online softmax, one thread per query row: tcgen05.ld the fp32 score row from TMEM, exp2 with the
//              running (lazily updated) row max, write P as bf16 back into the same TMEM columns (A operand of
//              the PV UMMA), rescale O in TMEM only when the max moved by more than 2^8; epilogue O / l.
// TMEM: S0 [0,128) | S1 [128,256) | O [256,384) columns, fp32, 128 lanes.
#include <math.h>
#include <stdlib.h>

#include "attn_common.cuh"

namespace svlm {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int ATT_STAGES = 3;
constexpr int PANEL_BYTES = 128 * 128;          // 128 rows x 64 bf16
constexpr int TILE_BYTES = 2 * PANEL_BYTES;     // 128 rows x 128 bf16
constexpr int ATT_THREADS = 320;          // warp 0 TMA, warp 1 UMMA, warps 2-5 / 6-9 softmax column halves
constexpr int ATT_SMEM = TILE_BYTES * (1 + 2 * ATT_STAGES) + 3072;   // = 232448 B, the sm_100 per-CTA maximum
constexpr float RESCALE_THRESHOLD = 8.0f;       // log2 units

constexpr int MODE_CAUSAL = 0;
constexpr int MODE_VARLEN = 1;

struct AttnParams {
    int Tq, Hq, G, D;          // D = true head dim (<= 128, multiple of 16)
    int tpt;                   // tokens per M tile = 128 / G
    int kv_len;                // causal: valid keys incl. the Tq new ones (host value)
    const int32_t* kv_len_dev; // optional device override of kv_len
    int num_splits;
    const int32_t* cu_seqlens; // varlen
    int n_seqs;
    int kv_head_base;          // coordinate offset of this layer in the K/V tensor maps (layer * Hkv)
    uint16_t* out;             // bf16 [Tq][Hq][D]
    float* part_o;             // fp32 [splits][Tq][Hq][D]   (num_splits > 1)
    float2* part_ml;           // (m, l) [splits][Tq][Hq]
    float scale_log2;          // softmax scale * log2(e)
    unsigned long long* trace; // SVLM_TRACE builds only
};

#ifdef SVLM_TRACE
#define ATRACE(ev, tile) do { if (p.trace && (tile) < 32) p.trace[((((blockIdx.z * gridDim.y) + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + (tile)) * 16 + (ev)] = clock64(); } while (0)
#else
#define ATRACE(ev, tile) do { } while (0)
#endif

struct AttnBars {
    uint64_t q_full;
    uint64_t k_full[ATT_STAGES], k_empty[ATT_STAGES];
    uint64_t v_full[ATT_STAGES], v_empty[ATT_STAGES];
    uint64_t s_full[2], p_full[2];
    uint64_t o_done[2];   // PV(j) completion alternates between the two (see the wait in the softmax warps)
    uint32_t tmem_base;
    float xmax[2][2][128];   // [tile parity][column half][row]: half-row maxima exchanged between the two softmax groups
                             // (xmax[0] doubles as the row-sum exchange buffer in the epilogue)
};
static_assert(sizeof(AttnBars) <= 3072, "AttnBars must fit behind the tiles");

__device__ __forceinline__ int find_seq(const int32_t* cu, int n_seqs, int tok) {
    int lo = 0, hi = n_seqs - 1;           // largest s with cu[s] <= tok
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(cu + mid) <= tok) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <int MODE>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];   // the kernel has no static shared memory, so the window starts 1024-aligned
    if ((smem_u32(smem) & 1023u) != 0u) __trap();       // (128B-swizzle atoms need it; checked rather than assumed)
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;
    uint8_t* sV = sK + ATT_STAGES * TILE_BYTES;
    AttnBars* bars = reinterpret_cast<AttnBars*>(sV + ATT_STAGES * TILE_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, head = blockIdx.y, split = blockIdx.z;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(&bars->q_full, 1);
        for (int s = 0; s < ATT_STAGES; ++s) {
            mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 1);
            mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->s_full[b], 1); mbar_init(&bars->p_full[b], 256); }
        mbar_init(&bars->o_done[0], 1); mbar_init(&bars->o_done[1], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(&bars->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    pdl_wait();          // barriers and TMEM are set up; q, the appended rows and the device-side length come from the predecessor

    // ---- work assignment (identical arithmetic in every warp) -------------------------------------
    const int tok0 = m_tile * p.tpt;
    const int tokL = min(tok0 + p.tpt, p.Tq) - 1;
    int kv_len = 0, off = 0, kv_lo_min, kv_hi_max;
    if (MODE == MODE_CAUSAL) {
        kv_len = p.kv_len_dev ? __ldg(p.kv_len_dev) : p.kv_len;
        off = kv_len - p.Tq;
        kv_lo_min = 0;
        kv_hi_max = min(kv_len, off + tokL + 1);
    } else {
        kv_lo_min = __ldg(p.cu_seqlens + find_seq(p.cu_seqlens, p.n_seqs, tok0));
        kv_hi_max = __ldg(p.cu_seqlens + find_seq(p.cu_seqlens, p.n_seqs, tokL) + 1);
    }
    const int t_begin = kv_lo_min / BN;
    const int t_end = (max(kv_hi_max, 0) + BN - 1) / BN;
    const int n_total = max(t_end - t_begin, 0);
    const int tps = (n_total + p.num_splits - 1) / p.num_splits;
    const int my_begin = t_begin + split * tps;
    const int n_tiles = max(min(t_end, my_begin + tps) - my_begin, 0);
    const int kv_coord = p.kv_head_base + head;

    constexpr uint32_t IDESC_QK = umma_idesc_bf16(BM, BN, 0, 0);
    constexpr uint32_t IDESC_PV = umma_idesc_bf16(BM, 128, 0, 1);
    const int ksteps = p.D / 16;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        // whole warp in uniform control flow, one elected lane issues
        if (n_tiles > 0) {
            if (elect_one()) {
                const uint32_t q_bytes = 2u * 128u * (uint32_t)(p.G * p.tpt);
                mbar_arrive_expect_tx(&bars->q_full, q_bytes);
                tma_load_4d(sQ, &tmQ, &bars->q_full, 0, 0, tok0, head);
                tma_load_4d(sQ + PANEL_BYTES, &tmQ, &bars->q_full, 64, 0, tok0, head);
            }
            __syncwarp();
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j % ATT_STAGES;
                const uint32_t ph = (j / ATT_STAGES) & 1;
                const int row = (my_begin + j) * BN;
                mbar_wait(&bars->k_empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bars->k_full[s], TILE_BYTES);
                    tma_load_3d(sK + s * TILE_BYTES, &tmK, &bars->k_full[s], 0, row, kv_coord);
                    tma_load_3d(sK + s * TILE_BYTES + PANEL_BYTES, &tmK, &bars->k_full[s], 64, row, kv_coord);
                }
                __syncwarp();
                mbar_wait(&bars->v_empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bars->v_full[s], TILE_BYTES);
                    tma_load_3d(sV + s * TILE_BYTES, &tmV, &bars->v_full[s], 0, row, kv_coord);
                    tma_load_3d(sV + s * TILE_BYTES + PANEL_BYTES, &tmV, &bars->v_full[s], 64, row, kv_coord);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // =========================== UMMA issuer ===========================
        // whole warp in uniform control flow, one elected lane issues (see attn_decode.cu for why not `if (lane == 0)`)
        if (n_tiles > 0) {
            const uint32_t q_addr = smem_u32(sQ);
            auto issue_qk = [&](int stage, int buf) {
                const uint32_t k_addr = smem_u32(sK + stage * TILE_BYTES);
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint32_t o = (uint32_t)(kk >> 2) * PANEL_BYTES + (uint32_t)(kk & 3) * 32u;
                    umma_ss(tmem + buf * 128, umma_desc_sw128(q_addr + o, 16, 1024), umma_desc_sw128(k_addr + o, 16, 1024), IDESC_QK,
                            kk > 0 ? 1u : 0u);
                }
            };
            auto issue_pv = [&](int stage, int buf, bool acc) {
                const uint32_t v_addr = smem_u32(sV + stage * TILE_BYTES);
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk) {
                    // B = V tile, MN-major: 16 keys (rows of 128 B) per step, two 64-column panels PANEL_BYTES apart
                    umma_ts(tmem + 256, tmem + buf * 128 + kk * 8, umma_desc_sw128(v_addr + kk * 2048, PANEL_BYTES, 1024), IDESC_PV,
                            (acc || kk > 0) ? 1u : 0u);
                }
            };
            mbar_wait(&bars->q_full, 0);
            mbar_wait(&bars->k_full[0], 0);
            tc_fence_after();
            if (elect_one()) {
                issue_qk(0, 0);
                umma_commit(&bars->k_empty[0]);
                umma_commit(&bars->s_full[0]);
            }
            __syncwarp();
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) {
                    const int s1 = (j + 1) % ATT_STAGES;
                    mbar_wait(&bars->k_full[s1], ((j + 1) / ATT_STAGES) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        issue_qk(s1, (j + 1) & 1);
                        umma_commit(&bars->k_empty[s1]);
                        umma_commit(&bars->s_full[(j + 1) & 1]);
                    }
                    __syncwarp();
                }
                const int s = j % ATT_STAGES;
                mbar_wait(&bars->p_full[j & 1], (j >> 1) & 1);
                mbar_wait(&bars->v_full[s], (j / ATT_STAGES) & 1);
                tc_fence_after();
                if (elect_one()) {
                    issue_pv(s, j & 1, j > 0);
                    umma_commit(&bars->v_empty[s]);
                    umma_commit(&bars->o_done[j & 1]);
                }
                __syncwarp();
            }
        }
    } else {
        // =========================== softmax + epilogue ===========================
        // Two groups of four warps split the 128 score columns of every row: thread (row, half) owns columns
        // [64*half, 64*half+64).  The MUFU pipe (16 ex2/clk/SM) and the conversion work are what bound a one-thread-per-row
        // softmax; eight warps keep them busy while the UMMA pipe works on the next tile.  The two halves of a row exchange
        // their maxima through shared memory once per tile (one named barrier), so both use the same reference max.
        const int q4 = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q4 * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const int tok = tok0 + row / p.G;
        const int g = row % p.G;
        const bool valid = (row < p.G * p.tpt) && (tok < p.Tq);
        // Rows of the tile that carry no query (row >= G * tpt, or token >= Tq) are never masked: their output is discarded, and a
        // permanently "masked" lane would drag its whole warp through the slow masked path on every tile.
        int lo = 0, hi = 0x7fffffff;
        if (valid) {
            if (MODE == MODE_CAUSAL) {
                hi = min(kv_len, off + tok + 1);
            } else {
                const int s = find_seq(p.cu_seqlens, p.n_seqs, tok);
                lo = __ldg(p.cu_seqlens + s);
                hi = __ldg(p.cu_seqlens + s + 1);
            }
        }
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int b = j & 1;
            const int kv0 = (my_begin + j) * BN + half * 64;
            mbar_wait(&bars->s_full[b], (j >> 1) & 1);
            tc_fence_after();
            uint32_t sr[2][32];
#pragma unroll
            for (int c = 0; c < 2; ++c) tmem_ld32(tmem + lane_addr + b * 128 + half * 64 + c * 32, sr[c]);
            tmem_wait_ld();
            // warp-uniform, so the warp is converged at the named barrier below
            const bool need_mask = __any_sync(0xffffffffu, (kv0 < lo) || (kv0 + 64 > hi));
            float mt_raw = -INFINITY;                      // max of the RAW scores (scale > 0 keeps the order)
            if (need_mask) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int kv = kv0 + c * 32 + i;
                        if (kv < lo || kv >= hi) sr[c][i] = 0xff800000u;   // -inf
                        mt_raw = fmaxf(mt_raw, __uint_as_float(sr[c][i]));
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mt_raw = fmaxf(mt_raw, __uint_as_float(sr[c][i]));
                }
            }
            bars->xmax[b][half][row] = mt_raw;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (lane == 0) ATRACE((warp - 2) * 2, j);
            const float mt = fmaxf(mt_raw, bars->xmax[b][half ^ 1][row]) * p.scale_log2;
            const bool upd = mt > m_run + RESCALE_THRESHOLD;      // also true for the first finite tile (m_run = -inf)
            const float m_new = upd ? mt : m_run;
            const float alpha = upd ? ex2_approx(m_run - m_new) : 1.0f; // m_run = -inf -> 0
            if (__any_sync(0xffffffffu, upd && j > 0)) {
                // rescale this thread's half of the running O row; PV(j-1) must have landed first.  s_full(j) observed =>
                // PV(j-3) and everything older completed, so barrier (j-1)&1 is in one of exactly two states and the
                // parity test is unambiguous even though earlier phases were never waited on.
                mbar_wait(&bars->o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t orow[32];
                    tmem_ld32(tmem + lane_addr + 256 + half * 64 + c * 32, orow);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * alpha);
                    tmem_st32(tmem + lane_addr + 256 + half * 64 + c * 32, orow);
                }
            }
            const float neg_m = (m_new == -INFINITY) ? 0.f : -m_new;
            // p = 2^(s * scale - m) on PAIRS of columns with packed fp32x2 math (one FFMA2 for the scale, one FADD2 for the row
            // sum, one F2FP for the bf16 pack per pair).  7 of 16 pairs take the polynomial on the FMA pipe, 9 the MUFU pipe:
            // the split that balances the two pipes (issue-slot count, not MUFU alone, bounds this loop on B200).
            const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2), negm2 = pack2(neg_m, neg_m);
            uint64_t lsum2 = pack2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint64_t x = ffma2((uint64_t)sr[c][2 * i] | ((uint64_t)sr[c][2 * i + 1] << 32), scale2, negm2);
                    const uint64_t pp = ((i & 1) && i != 15) ? ex2_poly2(x) : pack2(ex2_approx(lo2(x)), ex2_approx(hi2(x)));
                    lsum2 = fadd2(lsum2, pp);
                    pk[i] = pack_bf16(lo2(pp), hi2(pp));
                }
                tmem_st16(tmem + lane_addr + b * 128 + half * 32 + c * 16, pk);
            }
            const float lsum0 = lo2(lsum2), lsum1 = hi2(lsum2);
            l_run = l_run * alpha + (lsum0 + lsum1);
            m_run = m_new;
            if (lane == 0) ATRACE((warp - 2) * 2 + 1, j);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&bars->p_full[b]);
        }

        // ---- epilogue: the two halves add their row sums, each stores its 64 output columns ----
        pdl_trigger();
        asm volatile("bar.sync 1, 256;" ::: "memory");           // every thread is past its last xmax read
        bars->xmax[0][half][row] = l_run;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float l_tot = bars->xmax[0][0][row] + bars->xmax[0][1][row];
        if (n_tiles > 0) {
            mbar_wait(&bars->o_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
            tc_fence_after();
        }
        const int hq = head * p.G + g;
        const float inv = (p.num_splits == 1) ? (l_tot > 0.f ? 1.0f / l_tot : 0.f) : 1.0f;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = half * 2 + cc;       // 32-column chunk of the output row
            uint32_t orow[32];
            if (n_tiles > 0) {
                tmem_ld32(tmem + lane_addr + 256 + c * 32, orow);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) orow[i] = 0u;
            }
            if (valid) {
                if (p.num_splits == 1) {
                    uint16_t* dst = p.out + ((int64_t)tok * p.Hq + hq) * p.D + c * 32;
#pragma unroll
                    for (int v8 = 0; v8 < 4; ++v8) {
                        if (c * 32 + v8 * 8 < p.D) {
                            uint4 o;
                            o.x = pack_bf16(__uint_as_float(orow[v8 * 8 + 0]) * inv, __uint_as_float(orow[v8 * 8 + 1]) * inv);
                            o.y = pack_bf16(__uint_as_float(orow[v8 * 8 + 2]) * inv, __uint_as_float(orow[v8 * 8 + 3]) * inv);
                            o.z = pack_bf16(__uint_as_float(orow[v8 * 8 + 4]) * inv, __uint_as_float(orow[v8 * 8 + 5]) * inv);
                            o.w = pack_bf16(__uint_as_float(orow[v8 * 8 + 6]) * inv, __uint_as_float(orow[v8 * 8 + 7]) * inv);
                            *reinterpret_cast<uint4*>(dst + v8 * 8) = o;
                        }
                    }
                } else {
                    float* dst = p.part_o + (((int64_t)split * p.Tq + tok) * p.Hq + hq) * p.D + c * 32;
#pragma unroll
                    for (int v4 = 0; v4 < 8; ++v4) {
                        if (c * 32 + v4 * 4 < p.D)
                            *reinterpret_cast<uint4*>(dst + v4 * 4) = make_uint4(orow[v4 * 4], orow[v4 * 4 + 1], orow[v4 * 4 + 2], orow[v4 * 4 + 3]);
                    }
                }
            }
        }
        if (valid && half == 0 && p.num_splits > 1) p.part_ml[((int64_t)split * p.Tq + tok) * p.Hq + hq] = make_float2(m_run, l_tot);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// Merge of the split-KV partials: out = sum_s 2^(m_s - M) O_s / sum_s 2^(m_s - M) l_s.
// grid = Tq * Hq, block = 128 (one thread per output column).  The (m, l) pairs are fetched by all threads at once
// and the per-split weights live in shared memory, so the column loop issues independent loads.
__global__ void __launch_bounds__(128) attn_combine_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml,
                                                            uint16_t* __restrict__ out, int rows, int D, int num_splits) {
    pdl_wait();
    __shared__ float w_s[256];
    __shared__ float l_s[256];
    __shared__ float red[4];
    const int r = blockIdx.x, d = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float m_loc = -INFINITY;
    for (int s = d; s < num_splits; s += 128) {
        const float2 ml = part_ml[(int64_t)s * rows + r];
        w_s[s] = ml.x;
        l_s[s] = ml.y;
        m_loc = fmaxf(m_loc, ml.x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m_loc = fmaxf(m_loc, __shfl_xor_sync(0xffffffffu, m_loc, o));
    if (lane == 0) red[warp] = m_loc;
    __syncthreads();
    const float M = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float l_loc = 0.f;
    for (int s = d; s < num_splits; s += 128) {
        const float w = (w_s[s] == -INFINITY) ? 0.f : ex2_approx(w_s[s] - M);
        w_s[s] = w;
        l_loc += w * l_s[s];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l_loc += __shfl_xor_sync(0xffffffffu, l_loc, o);
    if (lane == 0) red[warp] = l_loc;
    __syncthreads();
    const float L = red[0] + red[1] + red[2] + red[3];
    // column sums: warp w takes splits w, w + 4, ...; a lane owns 4 adjacent columns (16-byte loads).  Up to 12 independent
    // loads in flight per thread, so 48 splits cost ONE round trip to L2; the four partial sums meet in shared memory.
    __shared__ float4 part_acc[4][32];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c4 = lane * 4;
    if (c4 < D) {
        const float* src = part_o + (int64_t)r * D + c4;
        const int64_t stride = (int64_t)rows * D;
        for (int s0 = warp; s0 < num_splits; s0 += 48) {
            float4 v[12];
#pragma unroll
            for (int u = 0; u < 12; ++u) {
                const int s = s0 + 4 * u;
                v[u] = s < num_splits ? __ldcg(reinterpret_cast<const float4*>(src + s * stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 12; ++u) {
                const int s = s0 + 4 * u;
                const float w = s < num_splits ? w_s[s] : 0.f;
                acc.x += w * v[u].x; acc.y += w * v[u].y; acc.z += w * v[u].z; acc.w += w * v[u].w;
            }
        }
    }
    part_acc[warp][lane] = acc;
    __syncthreads();
    if (d >= D) return;
    const float* pa = reinterpret_cast<const float*>(part_acc);
    const float sum = pa[0 * 128 + d] + pa[1 * 128 + d] + pa[2 * 128 + d] + pa[3 * 128 + d];
    out[(int64_t)r * D + d] = __bfloat16_as_ushort(__float2bfloat16_rn(L > 0.f ? sum / L : 0.f));
}

// Merge for FEW splits and MANY rows (chunk prefill: 2-4 splits x Tq*Hq rows): one thread per (row, 4 columns), fully
// vectorised; the one-CTA-per-row kernel above would launch 8400 tiny CTAs here (measured 21 us vs ~3 us of traffic).
__global__ void __launch_bounds__(256) attn_combine_rows_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml,
                                                                 uint16_t* __restrict__ out, int rows, int D, int num_splits) {
    pdl_wait();
    const int vec_per_row = D / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)rows * vec_per_row) return;
    const int r = (int)(idx / vec_per_row), c = (int)(idx % vec_per_row) * 4;
    float M = -INFINITY;
    for (int s = 0; s < num_splits; ++s) M = fmaxf(M, part_ml[(int64_t)s * rows + r].x);
    float L = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < num_splits; ++s) {
        const float2 ml = part_ml[(int64_t)s * rows + r];
        const float w = (ml.x == -INFINITY) ? 0.f : ex2_approx(ml.x - M);
        const float4 v = *reinterpret_cast<const float4*>(part_o + ((int64_t)s * rows + r) * D + c);
        L += w * ml.y;
        acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
    const float inv = L > 0.f ? 1.0f / L : 0.f;
    uint2 o;
    o.x = pack_bf16(acc.x * inv, acc.y * inv);
    o.y = pack_bf16(acc.z * inv, acc.w * inv);
    *reinterpret_cast<uint2*>(out + (int64_t)r * D + c) = o;
}

static int launch_attn(int mode, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, dim3 grid,
                       cudaStream_t st) {
    static bool attr_done[2] = {false, false};
    if (!attr_done[mode]) {
        cudaError_t e = mode == MODE_CAUSAL
                            ? cudaFuncSetAttribute(attn_fwd_kernel<MODE_CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM)
                            : cudaFuncSetAttribute(attn_fwd_kernel<MODE_VARLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(attn_fwd_kernel, %d B smem) failed: %s", ATT_SMEM, cudaGetErrorString(e));
            return SVLM_ECUDA;
        }
        attr_done[mode] = true;
    }
    if (mode == MODE_CAUSAL)
        launch_pdl(PDL_ELEM, attn_fwd_kernel<MODE_CAUSAL>, grid, dim3(ATT_THREADS), (size_t)ATT_SMEM, st, tq, tk, tv, p);
    else
        launch_pdl(PDL_ELEM, attn_fwd_kernel<MODE_VARLEN>, grid, dim3(ATT_THREADS), (size_t)ATT_SMEM, st, tq, tk, tv, p);
    return launch_status("attn_fwd_kernel");
}

}  // namespace svlm

using namespace svlm;

extern "C" size_t svlm_attn_workspace_bytes(int n_q_tokens, int q_heads, int head_dim, int num_splits) {
    if (num_splits <= 1) return 0;
    const size_t rows = (size_t)num_splits * n_q_tokens * q_heads;
    return rows * head_dim * sizeof(float) + rows * sizeof(float2);
}

extern "C" int svlm_attn_causal(const svlm_arena_t* arena, int buffer, int layer, const void* q, void* out, int n_q_tokens, int q_heads,
                                int kv_len, const int32_t* kv_len_dev, int num_splits, void* workspace, size_t workspace_bytes,
                                float scale, int flags, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    if ((rc = validate_arena(arena))) return rc;
    SVLM_REQUIRE(buffer == 0 || buffer == 1, "buffer must be 0 or 1");
    SVLM_REQUIRE(layer >= 0 && layer < arena->layers, "layer %d out of range", layer);
    SVLM_REQUIRE(q && out, "NULL tensor argument");
    SVLM_REQUIRE(n_q_tokens > 0 && q_heads > 0 && q_heads % arena->kv_heads == 0, "bad n_q_tokens / q_heads");
    SVLM_REQUIRE(kv_len >= n_q_tokens && kv_len <= arena->capacity, "kv_len %d must be in [Tq=%d, capacity=%d]", kv_len, n_q_tokens,
                 arena->capacity);
    SVLM_REQUIRE(num_splits >= 1 && num_splits <= 256, "num_splits %d out of range", num_splits);
    SVLM_REQUIRE((((uintptr_t)q | (uintptr_t)out) & 15) == 0, "q/out must be 16-byte aligned");
    const int D = arena->head_dim, H = arena->kv_heads, G = q_heads / H;
    SVLM_REQUIRE(G <= 128, "group size %d too large", G);
    const size_t need = svlm_attn_workspace_bytes(n_q_tokens, q_heads, D, num_splits);
    SVLM_REQUIRE(num_splits == 1 || (workspace != nullptr && workspace_bytes >= need), "workspace too small: need %zu bytes", need);
    if (n_q_tokens == 1 && G <= 8)    // decode: swap-AB kernel (attn_decode.cu)
        return attn_decode_launch(arena, buffer, layer, q, out, q_heads, kv_len, kv_len_dev, num_splits, workspace, scale, flags,
                                  (cudaStream_t)stream);

    ArenaDims ad{arena->layers, H, D, arena->capacity};
    const uint16_t* base = (const uint16_t*)arena->base;
    const int tpt = BM / G;
    CUtensorMap tq, tk, tv;
    {
        uint64_t dims[4] = {(uint64_t)D, (uint64_t)G, (uint64_t)n_q_tokens, (uint64_t)H};
        uint64_t str[3] = {(uint64_t)D * 2, (uint64_t)q_heads * D * 2, (uint64_t)G * D * 2};
        uint32_t box[4] = {64, (uint32_t)G, (uint32_t)tpt, 1};
        if ((rc = make_tensor_map_bf16(&tq, q, 4, dims, str, box, 1))) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)D, (uint64_t)ad.cap, (uint64_t)ad.L * H};
        uint64_t str[2] = {(uint64_t)D * 2, (uint64_t)ad.cap * D * 2};
        uint32_t box[3] = {64, BN, 1};
        if ((rc = make_tensor_map_bf16(&tk, base + ad.offset(buffer, SVLM_PLANE_K_ROT, 0, 0, 0), 3, dims, str, box, 1))) return rc;
        if ((rc = make_tensor_map_bf16(&tv, base + ad.offset(buffer, SVLM_PLANE_V, 0, 0, 0), 3, dims, str, box, 1))) return rc;
    }
    AttnParams p{};
    p.Tq = n_q_tokens; p.Hq = q_heads; p.G = G; p.D = D; p.tpt = tpt;
    p.kv_len = kv_len; p.kv_len_dev = kv_len_dev; p.num_splits = num_splits;
    p.kv_head_base = layer * H;
    p.out = (uint16_t*)out;
    const size_t rows = (size_t)num_splits * n_q_tokens * q_heads;
    p.part_o = (float*)workspace;
    p.part_ml = num_splits > 1 ? (float2*)((float*)workspace + rows * D) : nullptr;
    p.scale_log2 = scale * 1.4426950408889634f;
#ifdef SVLM_TRACE
    { const char* v = getenv("SVLM_TRACE_PTR"); p.trace = v ? (unsigned long long*)strtoull(v, nullptr, 0) : nullptr; }
#endif
    dim3 grid((n_q_tokens + tpt - 1) / tpt, H, num_splits);
    if ((rc = launch_attn(MODE_CAUSAL, tq, tk, tv, p, grid, (cudaStream_t)stream))) return rc;
    if (num_splits > 1) {
        const int rows = n_q_tokens * q_heads;
        if (num_splits <= 8 && rows >= 256) {
            const int64_t threads = (int64_t)rows * (D / 4);
            launch_pdl(PDL_ELEM, attn_combine_rows_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
                       (const float*)p.part_o, (const float2*)p.part_ml, p.out, rows, D, num_splits);
        } else {
            attn_combine_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(p.part_o, p.part_ml, p.out, rows, D, num_splits);
        }
        return launch_status("attn_combine_kernel");
    }
    return SVLM_OK;
}

extern "C" int svlm_vit_attn(const void* qkv, const int32_t* cu_seqlens, int n_seqs, void* out, int n_tokens, int heads, int head_dim,
                             float scale, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(qkv && cu_seqlens && out, "NULL tensor argument");
    SVLM_REQUIRE(n_tokens > 0 && heads > 0 && n_seqs > 0, "bad n_tokens / heads / n_seqs");
    SVLM_REQUIRE(head_dim % 16 == 0 && head_dim <= 128 && head_dim >= 16, "head_dim %d must be a multiple of 16 in [16,128]", head_dim);
    SVLM_REQUIRE((((uintptr_t)qkv | (uintptr_t)out) & 15) == 0, "qkv/out must be 16-byte aligned");
    const int D = head_dim;
    const uint16_t* base = (const uint16_t*)qkv;
    const uint64_t tok_stride = (uint64_t)3 * heads * D * 2;
    CUtensorMap tq, tk, tv;
    {
        uint64_t dims[4] = {(uint64_t)D, 1, (uint64_t)n_tokens, (uint64_t)heads};
        uint64_t str[3] = {(uint64_t)D * 2, tok_stride, (uint64_t)D * 2};
        uint32_t box[4] = {64, 1, BM, 1};
        if ((rc = make_tensor_map_bf16(&tq, base, 4, dims, str, box, 1))) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)D, (uint64_t)n_tokens, (uint64_t)heads};
        uint64_t str[2] = {tok_stride, (uint64_t)D * 2};
        uint32_t box[3] = {64, BN, 1};
        if ((rc = make_tensor_map_bf16(&tk, base + (size_t)heads * D, 3, dims, str, box, 1))) return rc;
        if ((rc = make_tensor_map_bf16(&tv, base + (size_t)2 * heads * D, 3, dims, str, box, 1))) return rc;
    }
    AttnParams p{};
    p.Tq = n_tokens; p.Hq = heads; p.G = 1; p.D = D; p.tpt = BM;
    p.num_splits = 1; p.cu_seqlens = cu_seqlens; p.n_seqs = n_seqs; p.kv_head_base = 0;
    p.out = (uint16_t*)out;
    p.scale_log2 = scale * 1.4426950408889634f;
    dim3 grid((n_tokens + BM - 1) / BM, heads, 1);
    return launch_attn(MODE_VARLEN, tq, tk, tv, p, grid, (cudaStream_t)stream);
}
