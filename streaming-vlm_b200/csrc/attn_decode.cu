// Split-KV decode attention (Tq = 1) for sm_100a on tcgen05 / TMEM / TMA, "swap-AB" formulation.
//
// With one query token only G (= q heads per kv head, 7 for the 7B geometry) of the 128 rows of a Q tile would be
// real, and a row-per-thread softmax would leave 121 of 128 threads idle while 7 threads walk 128 scores each.
// Here the KEYS are the M dimension instead:
//     S^T[128 keys x 16] = K_tile[128 x D] (A, K-major smem)  x  Q^T (B = Q [16 heads x D], K-major smem)
//     O^T[D=128 x 16]   += V_tile^T (A = V tile [128 keys x D] read MN-major) x P^T (B = P [16 heads x 128 keys], K-major smem)
// so every one of the 128 softmax threads owns ONE key and G scores (tcgen05.ld 32x32b.x16), the UMMAs are
// M128 N16 (8 + 8 tiny instructions per tile), and the kernel is limited by the TMA stream of K/V tiles, i.e. HBM.
// The per-head running maximum is shared by all keys of a tile (the accumulator rows are head-dim lanes), so the
// tile maximum is reduced across the 128 threads (shuffles + one named barrier); the row sums stay per-thread until
// the end.  Partial (m, l, O) per split go to the fp32 workspace and are merged by attn_combine_kernel.
//
// CTA = 192 threads (warp 0 TMA producer, warp 1 UMMA issuer, warps 2-5 softmax/epilogue), grid = (splits, H_kv).
#include <math.h>

#include "attn_common.cuh"

namespace svlm {

constexpr int DEC_BN = 128;
// K and V tiles alternate through ONE ring of 32 KB slots (use u = 2j for K(j), 2j + 1 for V(j); slot = u % SLOTS), which
// keeps a CTA at ~107 KB of shared memory: two CTAs per SM interleave their (latency-bound) per-tile chains.
// SLOTS = 3 when the grid has more CTAs than SMs (two resident CTAs hide each other's stalls), 6 (one CTA per SM, two tiles of
// look-ahead) otherwise.
constexpr int DEC_MAX_SLOTS = 6;
constexpr int DEC_PANEL = 128 * 128;           // 128 keys x 64 dims
constexpr int DEC_TILE = 2 * DEC_PANEL;        // 128 keys x 128 dims
constexpr int DEC_QP_PANEL = 16 * 128;         // 16 rows x 64 elements (Q and P^T operands)
constexpr int DEC_THREADS = 192;
constexpr int dec_smem(int slots) { return 1024 + slots * DEC_TILE + 2 * DEC_QP_PANEL + 2 * 2 * DEC_QP_PANEL + 1024; }

struct DecParams {
    int Hq, G, D;
    int kv_len;                                       // host value, used for streams whose kv_len_dev is NULL
    const int32_t* kv_len_dev[SVLM_MAX_STREAMS];      // per stream (blockIdx.z)
    int rows_total;                                   // n_streams * Hq: row count of out / the partials
    int num_splits;
    int early_kv;         // SVLM_ATTN_EARLY_KV: kv_len_dev and the K/V rows below kv_len - 1 are not outputs of the preceding launch
    int cluster;          // CTAs per cluster along the split axis: their partial (m, l, O) are merged through distributed shared memory
    int kv_head_base;
    uint16_t* out;        // bf16 [streams][Hq][D]           (num_splits == cluster: nothing left to merge)
    float* part_o;        // fp32 [splits / cluster][streams * Hq][D]
    float2* part_ml;      // (m, l) [splits / cluster][streams * Hq]
    float scale_log2;
    unsigned long long* trace;   // SVLM_TRACE builds only: [cta][tile][8] clock64 stamps
};

#ifdef SVLM_TRACE
#define TRACE(ev, tile) do { if (p.trace && (tile) < 32) p.trace[((blockIdx.y * gridDim.x + blockIdx.x) * 32 + (tile)) * 8 + (ev)] = clock64(); } while (0)
#else
#define TRACE(ev, tile) do { } while (0)
#endif

struct DecShared {
    uint64_t q_full;
    uint64_t full[DEC_MAX_SLOTS], empty[DEC_MAX_SLOTS];
    uint64_t s_full[2], p_full[2];
    uint64_t o_done[2];
    uint32_t tmem_base;
    float wmax[2][4][8];       // [tile parity][softmax warp][head]
    float lsum[4][8];
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__device__ __forceinline__ int ordered_i32(float f) {            // monotonic float -> int map (for redux.sync.max.s32)
    const int i = __float_as_int(f);
    return i >= 0 ? i : (i ^ 0x7fffffff);
}
__device__ __forceinline__ float unordered_f32(int i) { return __int_as_float(i >= 0 ? i : (i ^ 0x7fffffff)); }

__device__ __forceinline__ float ld_shared_cluster_f32(uint32_t cluster_addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_cluster_f32_(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_arrive_release_() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire_() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// One K and one V tensor map per stream (every stream owns its arena); q is one [streams][Hq][D] tensor.
struct DecMaps {
    CUtensorMap q;
    CUtensorMap k[SVLM_MAX_STREAMS];
    CUtensorMap v[SVLM_MAX_STREAMS];
};

template <int G, int DEC_SLOTS>
__global__ void __launch_bounds__(DEC_THREADS, 2)
attn_decode_kernel(const __grid_constant__ DecMaps maps, const DecParams p) {
    const int strm = blockIdx.z;
    const CUtensorMap& tmQ = maps.q;
    const CUtensorMap& tmK = maps.k[strm];
    const CUtensorMap& tmV = maps.v[strm];
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sKV = smem;                                  // ring of K / V tiles
    uint8_t* sQ = sKV + DEC_SLOTS * DEC_TILE;            // 2 panels of [16 heads][64 dims]
    uint8_t* sP = sQ + 2 * DEC_QP_PANEL;                 // 2 buffers x 2 panels of [16 heads][64 keys]
    DecShared* sh = reinterpret_cast<DecShared*>(sP + 2 * 2 * DEC_QP_PANEL);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, head = blockIdx.y;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(&sh->q_full, 1);
        for (int s = 0; s < DEC_SLOTS; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&sh->s_full[b], 1); mbar_init(&sh->p_full[b], 128); mbar_init(&sh->o_done[b], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<64>(&sh->tmem_base);
    // P^T rows that no thread writes (heads >= G) must be zero, not stale shared memory
    for (int i = threadIdx.x; i < 2 * 2 * DEC_QP_PANEL / 16; i += DEC_THREADS) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh->tmem_base;
    // everything above (barriers, TMEM, zeroed P^T) overlapped the tail of the QKV kernel; q and the new arena row are its outputs.
    // With SVLM_ATTN_EARLY_KV the caller guarantees that the device-side length and every row below kv_len - 1 are NOT: the length is
    // read, and the K/V tiles that end below the new row are requested, before the grid dependency resolves -- the first TMA round
    // trip of this latency-bound kernel overlaps the predecessor's tail.
    const bool early = p.early_kv != 0;
    if (!early) pdl_wait();

    // the host value is the documented upper bound (<= arena capacity): a device-side length beyond it never reaches the TMA
    const int kv_len = p.kv_len_dev[strm] ? min(__ldg(p.kv_len_dev[strm]), p.kv_len) : p.kv_len;
    const int n_total = (kv_len + DEC_BN - 1) / DEC_BN;
    const int tps = (n_total + p.num_splits - 1) / p.num_splits;
    const int my_begin = split * tps;
    const int n_tiles = max(min(n_total, my_begin + tps) - my_begin, 0);
    const int kv_coord = p.kv_head_base + head;

    float c_o[G], c_m[G], c_l[G];      // softmax warps: this split's (O^T lane, m, l) per head, handed to the cluster merge below
#pragma unroll
    for (int h = 0; h < G; ++h) { c_o[h] = 0.f; c_m[h] = -INFINITY; c_l[h] = 0.f; }
    constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 16, 0, 0);   // A = K tile (K-major), B = Q (K-major)
    constexpr uint32_t IDESC_O = umma_idesc_bf16(128, 16, 1, 0);   // A = V tile read MN-major, B = P^T (K-major)

    if (warp == 0) {
        // TMA producer: uniform control flow for the whole warp, one elected lane issues (same reason as the UMMA warp)
        if (n_tiles > 0) {
            int u_pre = 0;
            if (early) {
                // tiles of this CTA that end at or below the row being appended (index kv_len - 1): immutable during this launch
                const int safe_tiles = min(max((kv_len - 1) / DEC_BN - my_begin, 0), n_tiles);
                u_pre = min(2 * safe_tiles, DEC_SLOTS);
                for (int u = 0; u < u_pre; ++u) {                  // first use of every slot: nothing to wait for
                    const int row = (my_begin + (u >> 1)) * DEC_BN;
                    if (elect_one()) {
                        const CUtensorMap* tm = (u & 1) ? &tmV : &tmK;
                        mbar_arrive_expect_tx(&sh->full[u], DEC_TILE);
                        tma_load_3d(sKV + u * DEC_TILE, tm, &sh->full[u], 0, row, kv_coord);
                        tma_load_3d(sKV + u * DEC_TILE + DEC_PANEL, tm, &sh->full[u], 64, row, kv_coord);
                    }
                    __syncwarp();
                }
                pdl_wait();
            }
            if (elect_one()) {
                mbar_arrive_expect_tx(&sh->q_full, 2 * DEC_QP_PANEL);
                tma_load_3d(sQ, &tmQ, &sh->q_full, 0, head * G, strm);
                tma_load_3d(sQ + DEC_QP_PANEL, &tmQ, &sh->q_full, 64, head * G, strm);
            }
            __syncwarp();
            for (int u = u_pre; u < 2 * n_tiles; ++u) {
                const int s = u % DEC_SLOTS;
                const int row = (my_begin + (u >> 1)) * DEC_BN;
                mbar_wait(&sh->empty[s], ((u / DEC_SLOTS) & 1) ^ 1);
                if (elect_one()) {
                    const CUtensorMap* tm = (u & 1) ? &tmV : &tmK;
                    mbar_arrive_expect_tx(&sh->full[s], DEC_TILE);
                    tma_load_3d(sKV + s * DEC_TILE, tm, &sh->full[s], 0, row, kv_coord);
                    tma_load_3d(sKV + s * DEC_TILE + DEC_PANEL, tm, &sh->full[s], 64, row, kv_coord);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        if (early) pdl_wait();
        // UMMA issuer.  The whole warp walks the loop (uniform control flow and operands, so the descriptors live in
        // uniform registers); one elected lane issues.  A divergent `if (lane == 0)` around the loop makes ptxas wrap every
        // tcgen05.mma in an ELECT / R2UR.BROADCAST waterfall (~100 cycles per instruction, measured).
        if (n_tiles > 0) {
            const uint32_t q_addr = smem_u32(sQ);
            auto issue_s = [&](int stage, int buf) {
                const uint32_t k_addr = smem_u32(sKV + stage * DEC_TILE);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t ka = k_addr + (uint32_t)(kk >> 2) * DEC_PANEL + (uint32_t)(kk & 3) * 32u;
                    const uint32_t qa = q_addr + (uint32_t)(kk >> 2) * DEC_QP_PANEL + (uint32_t)(kk & 3) * 32u;
                    umma_ss(tmem + buf * 16, umma_desc_sw128(ka, 16, 1024), umma_desc_sw128(qa, 16, 1024), IDESC_S, kk > 0 ? 1u : 0u);
                }
            };
            auto issue_o = [&](int stage, int buf, bool acc) {
                const uint32_t v_addr = smem_u32(sKV + stage * DEC_TILE);
                const uint32_t p_addr = smem_u32(sP + buf * 2 * DEC_QP_PANEL);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    // A: 16 keys (rows of 128 B) per step, the two 64-dim panels DEC_PANEL apart (MN-major, LBO = panel stride)
                    const uint32_t pa = p_addr + (uint32_t)(kk >> 2) * DEC_QP_PANEL + (uint32_t)(kk & 3) * 32u;
                    umma_ss(tmem + 32, umma_desc_sw128(v_addr + kk * 2048, DEC_PANEL, 1024), umma_desc_sw128(pa, 16, 1024), IDESC_O,
                            (acc || kk > 0) ? 1u : 0u);
                }
            };
            mbar_wait(&sh->q_full, 0);
            mbar_wait(&sh->full[0], 0);
            tc_fence_after();
            if (elect_one()) {
                issue_s(0, 0);
                umma_commit(&sh->empty[0]);
                umma_commit(&sh->s_full[0]);
            }
            __syncwarp();
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) {
                    const int u1 = 2 * (j + 1), s1 = u1 % DEC_SLOTS;
                    mbar_wait(&sh->full[s1], (u1 / DEC_SLOTS) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        issue_s(s1, (j + 1) & 1);
                        umma_commit(&sh->empty[s1]);
                        umma_commit(&sh->s_full[(j + 1) & 1]);
                    }
                    __syncwarp();
                }
                const int u = 2 * j + 1, s = u % DEC_SLOTS;
                mbar_wait(&sh->p_full[j & 1], (j >> 1) & 1);
                mbar_wait(&sh->full[s], (u / DEC_SLOTS) & 1);
                tc_fence_after();
                if (elect_one()) {
                    issue_o(s, j & 1, j > 0);
                    umma_commit(&sh->empty[s]);
                    umma_commit(&sh->o_done[j & 1]);
                }
                __syncwarp();
            }
        }
    } else {
        if (early) pdl_wait();
        // ---- softmax: thread = one key of the tile; epilogue: thread = one head-dim lane of O^T
        const int q4 = warp & 3;
        const int row = q4 * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        float m_run[G], l_thr[G];
#pragma unroll
        for (int h = 0; h < G; ++h) { m_run[h] = -INFINITY; l_thr[h] = 0.f; }
        const int kcol = row & 63;                 // key column inside a 64-key P^T panel
        for (int j = 0; j < n_tiles; ++j) {
            const int b = j & 1;
            const int key = (my_begin + j) * DEC_BN + row;
            if (row == 0) TRACE(0, j);
            mbar_wait(&sh->s_full[b], (j >> 1) & 1);
            tc_fence_after();
            if (row == 0) TRACE(1, j);
            uint32_t sr[8];
            tmem_ld8(tmem + lane_addr + b * 16, sr);
            tmem_wait_ld();
            if (row == 0) TRACE(2, j);
            const bool live = key < kv_len;
            float sc[G];
#pragma unroll
            for (int h = 0; h < G; ++h) {
                sc[h] = live ? __uint_as_float(sr[h]) * p.scale_log2 : -INFINITY;
                const int mx = __reduce_max_sync(0xffffffffu, ordered_i32(sc[h]));   // one REDUX per head
                if (lane == h) sh->wmax[b][q4][h] = unordered_f32(mx);
            }
            named_bar_sync(1, 128);
            if (row == 0) TRACE(3, j);
            bool any_upd = false;
            float alpha[G], pv[G];
#pragma unroll
            for (int h = 0; h < G; ++h) {
                const float mt = fmaxf(fmaxf(sh->wmax[b][0][h], sh->wmax[b][1][h]), fmaxf(sh->wmax[b][2][h], sh->wmax[b][3][h]));
                const bool upd = mt > m_run[h];
                const float m_new = upd ? mt : m_run[h];                  // finite: every tile has at least one live key
                alpha[h] = upd ? ex2_approx(m_run[h] - m_new) : 1.0f;     // -inf -> 0 on the first tile
                any_upd |= upd;
                m_run[h] = m_new;
                pv[h] = ex2_approx(sc[h] - m_new);
                l_thr[h] = l_thr[h] * alpha[h] + pv[h];
            }
            if (row == 0) TRACE(5, j);
            if (any_upd && j > 0) {
                // rescale O^T (head-dim lane x head column); PV(j-1) must have landed (see attn_fwd.cu for the parity argument)
                mbar_wait(&sh->o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                uint32_t orow[8];
                tmem_ld8(tmem + lane_addr + 32, orow);
                tmem_wait_ld();
#pragma unroll
                for (int h = 0; h < G; ++h) orow[h] = __float_as_uint(__uint_as_float(orow[h]) * alpha[h]);
                tmem_st8(tmem + lane_addr + 32, orow);
                tmem_wait_st();
            }
            if (row == 0) TRACE(6, j);
            // P^T[h][key] -> shared memory operand (K-major, 128B swizzle: 16-byte chunk index XOR (h & 7)); rows >= G stay zero
            uint8_t* pb = sP + b * 2 * DEC_QP_PANEL + (row >> 6) * DEC_QP_PANEL;
#pragma unroll
            for (int h = 0; h < G; ++h) {
                const uint32_t off = (uint32_t)h * 128u + ((((uint32_t)kcol >> 3) ^ (uint32_t)(h & 7)) << 4) + ((uint32_t)kcol & 7u) * 2u;
                *reinterpret_cast<uint16_t*>(pb + off) = __bfloat16_as_ushort(__float2bfloat16_rn(pv[h]));
            }
            if (row == 0) TRACE(7, j);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&sh->p_full[b]);
            if (row == 0) TRACE(4, j);
        }

        // ---- epilogue: total row sums (one reduction), then O^T lanes
        pdl_trigger();
        float l_tot[G];
#pragma unroll
        for (int h = 0; h < G; ++h) {
            float v = l_thr[h];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == h) sh->lsum[q4][h] = v;
        }
        named_bar_sync(1, 128);
#pragma unroll
        for (int h = 0; h < G; ++h) l_tot[h] = sh->lsum[0][h] + sh->lsum[1][h] + sh->lsum[2][h] + sh->lsum[3][h];
        uint32_t orow[8];
        if (n_tiles > 0) {
            mbar_wait(&sh->o_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
            tc_fence_after();
            tmem_ld8(tmem + lane_addr + 32, orow);
            tmem_wait_ld();
        } else {
#pragma unroll
            for (int h = 0; h < 8; ++h) orow[h] = 0u;
        }
        const int d = row;                         // O^T lane = head-dim index
        if (p.cluster == 1) {
            if (d < p.D) {
#pragma unroll
                for (int h = 0; h < G; ++h) {
                    const int hq = strm * p.Hq + head * G + h;       // row of out / the partials
                    if (p.num_splits == 1) {
                        const float inv = l_tot[h] > 0.f ? 1.0f / l_tot[h] : 0.f;
                        p.out[(int64_t)hq * p.D + d] = __bfloat16_as_ushort(__float2bfloat16_rn(__uint_as_float(orow[h]) * inv));
                    } else {
                        p.part_o[((int64_t)split * p.rows_total + hq) * p.D + d] = __uint_as_float(orow[h]);
                    }
                }
            }
            if (p.num_splits > 1) {
#pragma unroll
                for (int h = 0; h < G; ++h)
                    if (row == h) p.part_ml[(int64_t)split * p.rows_total + strm * p.Hq + head * G + h] = make_float2(m_run[h], l_tot[h]);
            }
        } else {
#pragma unroll
            for (int h = 0; h < G; ++h) { c_o[h] = __uint_as_float(orow[h]); c_m[h] = m_run[h]; c_l[h] = l_tot[h]; }
        }
    }

    if (p.cluster > 1) {
        // ---- split merge inside the cluster, PUSH based: once every CTA's K/V ring is dead (barrier 1), the peers store their
        // (m, l, O^T) into the leader's ring (st.shared::cluster, fire and forget), barrier 2, and the leader merges out of its own
        // shared memory.  The partial sums never travel through HBM / L2 and, when cluster == num_splits, no merge kernel is launched.
        // (Round 2 first had the leader PULL its peers' results over DSMEM: one remote-load round trip per peer in sequence, 13.0 us
        // instead of 9.8 us with the merge kernel for a single stream.)
        constexpr int LAND = G * 128 + 16;                       // floats per rank: O^T [G][128], m [8], l [8]
        float* land = reinterpret_cast<float*>(sKV);
        const uint32_t rank = cluster_ctarank();
        cluster_arrive_release_();
        cluster_wait_acquire_();
        if (warp >= 2) {
            const int d = (warp & 3) * 32 + lane;
            if (rank == 0) {
#pragma unroll
                for (int h = 0; h < G; ++h) land[h * 128 + d] = c_o[h];
#pragma unroll
                for (int h = 0; h < G; ++h)
                    if (d == h) { land[G * 128 + h] = c_m[h]; land[G * 128 + 8 + h] = c_l[h]; }
            } else {
                const uint32_t base = mapa_shared(smem_u32(land + rank * LAND), 0);
#pragma unroll
                for (int h = 0; h < G; ++h) st_shared_cluster_f32_(base + (h * 128 + d) * 4, c_o[h]);
#pragma unroll
                for (int h = 0; h < G; ++h)
                    if (d == h) { st_shared_cluster_f32_(base + (G * 128 + h) * 4, c_m[h]); st_shared_cluster_f32_(base + (G * 128 + 8 + h) * 4, c_l[h]); }
            }
        }
        cluster_arrive_release_();
        cluster_wait_acquire_();
        if (rank == 0 && warp >= 2) {
            const int d = (warp & 3) * 32 + lane;
            float M[G], L[G], O[G];
#pragma unroll
            for (int h = 0; h < G; ++h) { M[h] = -INFINITY; L[h] = 0.f; O[h] = 0.f; }
            for (int r = 0; r < p.cluster; ++r) {
                const float* lr = land + r * LAND;
#pragma unroll
                for (int h = 0; h < G; ++h) {
                    const float m = lr[G * 128 + h], l = lr[G * 128 + 8 + h], o = lr[h * 128 + d];
                    const float m_new = fmaxf(M[h], m);
                    const float a = (M[h] == -INFINITY) ? 0.f : ex2_approx(M[h] - m_new);
                    const float b = (m == -INFINITY) ? 0.f : ex2_approx(m - m_new);
                    L[h] = L[h] * a + l * b;
                    O[h] = O[h] * a + o * b;
                    M[h] = m_new;
                }
            }
            const int part = split / p.cluster;                  // index of this cluster among the remaining partials
            if (d < p.D) {
#pragma unroll
                for (int h = 0; h < G; ++h) {
                    const int hq = strm * p.Hq + head * G + h;
                    if (p.num_splits == p.cluster) {
                        const float inv = L[h] > 0.f ? 1.0f / L[h] : 0.f;
                        p.out[(int64_t)hq * p.D + d] = __bfloat16_as_ushort(__float2bfloat16_rn(O[h] * inv));
                    } else {
                        p.part_o[((int64_t)part * p.rows_total + hq) * p.D + d] = O[h];
                        if (d == h) p.part_ml[(int64_t)part * p.rows_total + hq] = make_float2(M[h], L[h]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<64>(tmem);
}

template <int G, int SLOTS>
static int launch_decode_s(const DecMaps& maps, const DecParams& p, dim3 grid, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(attn_decode_kernel<G, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem(SLOTS));
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(attn_decode_kernel, %d B smem) failed: %s", dec_smem(SLOTS), cudaGetErrorString(e));
            return SVLM_ECUDA;
        }
        e = cudaFuncSetAttribute(attn_decode_kernel<G, SLOTS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(attn_decode_kernel, non-portable cluster size) failed: %s", cudaGetErrorString(e));
            return SVLM_ECUDA;
        }
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(DEC_THREADS);
    cfg.dynamicSmemBytes = (size_t)dec_smem(SLOTS);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (pdl_enabled(PDL_OTHER)) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    if (p.cluster > 1) {
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = (unsigned)p.cluster;
        attr[n_attr].val.clusterDim.y = 1;
        attr[n_attr].val.clusterDim.z = 1;
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    cudaLaunchKernelEx(&cfg, attn_decode_kernel<G, SLOTS>, maps, p);
    return launch_status("attn_decode_kernel");
}

template <int G>
static int launch_decode(const DecMaps& maps, const DecParams& p, dim3 grid, cudaStream_t st) {
    return grid.x * grid.y * grid.z > 148 ? launch_decode_s<G, 3>(maps, p, grid, st) : launch_decode_s<G, 6>(maps, p, grid, st);
}

static int launch_decode_g(int G, const DecMaps& maps, const DecParams& p, dim3 grid, cudaStream_t st) {
    switch (G) {
        case 1: return launch_decode<1>(maps, p, grid, st);
        case 2: return launch_decode<2>(maps, p, grid, st);
        case 3: return launch_decode<3>(maps, p, grid, st);
        case 4: return launch_decode<4>(maps, p, grid, st);
        case 5: return launch_decode<5>(maps, p, grid, st);
        case 6: return launch_decode<6>(maps, p, grid, st);
        case 7: return launch_decode<7>(maps, p, grid, st);
        case 8: return launch_decode<8>(maps, p, grid, st);
        default: set_error("decode attention supports 1..8 query heads per kv head (got %d)", G); return SVLM_EINVAL;
    }
}

#ifdef SVLM_TRACE
#include <stdlib.h>
static void* getenv_ptr(const char* name) {
    const char* v = getenv(name);
    return v ? (void*)strtoull(v, nullptr, 0) : nullptr;
}
#endif

// Shared by the single-stream entry (svlm_attn_causal with Tq = 1) and svlm_attn_decode_streams.
static int decode_launch_common(const svlm_arena_t* geom, const void* const* arena_base, const int* buffer, const int32_t* const* kv_len_dev,
                                int n_streams, int layer, const void* q, void* out, int q_heads, int kv_len, int num_splits,
                                void* workspace, float scale, int flags, cudaStream_t st) {
    const int D = geom->head_dim, H = geom->kv_heads, G = q_heads / H;
    ArenaDims ad{geom->layers, H, D, geom->capacity};
    static thread_local DecMaps maps;
    int rc;
    {
        uint64_t dims[3] = {(uint64_t)D, (uint64_t)q_heads, (uint64_t)n_streams};
        uint64_t str[2] = {(uint64_t)D * 2, (uint64_t)q_heads * D * 2};
        uint32_t box[3] = {64, 16, 1};
        if ((rc = make_tensor_map_bf16(&maps.q, q, 3, dims, str, box, 1))) return rc;
    }
    DecParams p{};
    for (int b = 0; b < n_streams; ++b) {
        const uint16_t* base = (const uint16_t*)arena_base[b];
        uint64_t dims[3] = {(uint64_t)D, (uint64_t)ad.cap, (uint64_t)ad.L * H};
        uint64_t str[2] = {(uint64_t)D * 2, (uint64_t)ad.cap * D * 2};
        uint32_t box[3] = {64, DEC_BN, 1};
        if ((rc = make_tensor_map_bf16(&maps.k[b], base + ad.offset(buffer[b], SVLM_PLANE_K_ROT, 0, 0, 0), 3, dims, str, box, 1))) return rc;
        if ((rc = make_tensor_map_bf16(&maps.v[b], base + ad.offset(buffer[b], SVLM_PLANE_V, 0, 0, 0), 3, dims, str, box, 1))) return rc;
        p.kv_len_dev[b] = kv_len_dev[b];
    }
    p.Hq = q_heads; p.G = G; p.D = D;
    p.kv_len = kv_len; p.num_splits = num_splits;
    p.early_kv = (flags & SVLM_ATTN_EARLY_KV) ? 1 : 0;
    // Splits of one (stream, kv head) that sit in one thread-block cluster merge through distributed shared memory (push based, see the
    // kernel): no merge launch, no fp32 round trip of the partials.  Measured on B200 (benchmarks/decode_attn_cluster_sweep.py,
    // profiles/r2k_decode_cluster_push_sweep.jsonl): for the few splits of the multi-stream launch (4 per stream and head) it wins;
    // for ONE stream at 5.8k keys 8-CTA clusters take 10.9 us (merge kernel with 8 splits: 11.8 us) and 16-CTA clusters 10.9 us
    // (merge kernel: 10.1 us), but 23 splits + the merge kernel, which overlaps the attention's tail through programmatic dependent
    // launch, remain the fastest (9.6 us) -- the single-stream launch is bound by its fixed costs, not by the merge.
    // SVLM_DECODE_CLUSTER = 1 / 2 / 4 / 8 / 16 overrides the largest cluster size (16 = "non-portable", opted in per kernel).
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("SVLM_DECODE_CLUSTER");
        forced = e ? atoi(e) : 0;
        if (forced != 1 && forced != 2 && forced != 4 && forced != 8 && forced != 16) forced = 0;
    }
    const int max_cluster = forced ? forced : (n_streams > 1 ? 8 : 1);
    p.cluster = 1;
    if (num_splits <= max_cluster && (num_splits & (num_splits - 1)) == 0) p.cluster = num_splits;
    const int n_parts = num_splits / p.cluster;
    p.rows_total = n_streams * q_heads;
    p.kv_head_base = layer * H;
    p.out = (uint16_t*)out;
    p.part_o = (float*)workspace;
    p.part_ml = n_parts > 1 ? (float2*)((float*)workspace + (size_t)n_parts * p.rows_total * D) : nullptr;
    p.scale_log2 = scale * 1.4426950408889634f;
#ifdef SVLM_TRACE
    p.trace = (unsigned long long*)getenv_ptr("SVLM_TRACE_PTR");
#endif
    if ((rc = launch_decode_g(G, maps, p, dim3(num_splits, H, n_streams), st))) return rc;
    if (n_parts > 1) {
        // What the clusters left over: a separate merge launch beats an in-kernel "last CTA merges" epilogue here (measured: 10.2 vs
        // 12.6 us per layer at 4k keys, 30.1 vs 33.7 us at 64k): one-head CTAs finish the merge in one load latency.
        launch_pdl(PDL_OTHER, attn_combine_kernel, dim3(p.rows_total), dim3(128), 0, st, (const float*)p.part_o, (const float2*)p.part_ml, p.out,
                   p.rows_total, D, n_parts);
        return launch_status("attn_combine_kernel");
    }
    return SVLM_OK;
}

int attn_decode_launch(const svlm_arena_t* arena, int buffer, int layer, const void* q, void* out, int q_heads, int kv_len,
                       const int32_t* kv_len_dev, int num_splits, void* workspace, float scale, int flags, cudaStream_t st) {
    const void* base = arena->base;
    return decode_launch_common(arena, &base, &buffer, &kv_len_dev, 1, layer, q, out, q_heads, kv_len, num_splits, workspace, scale,
                                kv_len_dev ? flags : 0, st);
}

}  // namespace svlm

using namespace svlm;

extern "C" int svlm_attn_decode_streams(const svlm_arena_t* geometry, const svlm_stream_slot_t* slots, int n_streams, int layer,
                                        const void* q, void* out, int q_heads, int max_kv_len, int num_splits, void* workspace,
                                        size_t workspace_bytes, float scale, int flags, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(geometry && slots && q && out, "NULL argument");
    SVLM_REQUIRE(geometry->layers > 0 && geometry->kv_heads > 0 && geometry->capacity > 0 && geometry->head_dim == 128, "bad arena geometry");
    SVLM_REQUIRE(n_streams >= 1 && n_streams <= SVLM_MAX_STREAMS, "n_streams %d out of range (1..%d)", n_streams, SVLM_MAX_STREAMS);
    SVLM_REQUIRE(layer >= 0 && layer < geometry->layers, "layer %d out of range", layer);
    SVLM_REQUIRE(q_heads > 0 && q_heads % geometry->kv_heads == 0 && q_heads / geometry->kv_heads <= 8,
                 "q_heads %d must be 1..8 times kv_heads %d", q_heads, geometry->kv_heads);
    SVLM_REQUIRE(max_kv_len >= 1 && max_kv_len <= geometry->capacity, "max_kv_len %d out of range", max_kv_len);
    SVLM_REQUIRE(num_splits >= 1 && num_splits <= 256, "num_splits %d out of range", num_splits);
    SVLM_REQUIRE((((uintptr_t)q | (uintptr_t)out | (uintptr_t)workspace) & 15) == 0, "q/out/workspace must be 16-byte aligned");
    const size_t rows = (size_t)num_splits * n_streams * q_heads;
    const size_t need = num_splits > 1 ? rows * geometry->head_dim * sizeof(float) + rows * sizeof(float2) : 0;
    SVLM_REQUIRE(num_splits == 1 || (workspace != nullptr && workspace_bytes >= need), "workspace too small: need %zu bytes", need);
    const void* base[SVLM_MAX_STREAMS];
    int buffer[SVLM_MAX_STREAMS];
    const int32_t* kvl[SVLM_MAX_STREAMS];
    for (int b = 0; b < n_streams; ++b) {
        SVLM_REQUIRE(slots[b].arena_base && slots[b].kv_len_dev, "slot %d: NULL pointer", b);
        SVLM_REQUIRE(slots[b].buffer == 0 || slots[b].buffer == 1, "slot %d: buffer must be 0 or 1", b);
        SVLM_REQUIRE(((uintptr_t)slots[b].arena_base & 127) == 0, "slot %d: arena must be 128-byte aligned", b);
        base[b] = slots[b].arena_base; buffer[b] = slots[b].buffer; kvl[b] = slots[b].kv_len_dev;
    }
    return decode_launch_common(geometry, base, buffer, kvl, n_streams, layer, q, out, q_heads, max_kv_len, num_splits, workspace, scale,
                                flags, (cudaStream_t)stream);
}
