// "Few tokens x whole weight matrix" GEMM on tcgen05 / TMEM / TMA (sm_100a), swap-AB, split-K over a thread-block cluster:
//     Y[T][N] = epilogue(prologue(X)[T][K] * W[N][K]^T + bias[N])          T <= 64 tokens, nn.Linear layout, fp32 accumulate
// The batched decode step of several streams (multi_stream.py: one token per stream, T = number of streams) reads every weight matrix
// once for T tokens: the op is a weight STREAM, bound by HBM exactly like the single-stream GEMVs (gemv.cu), but T x the CUDA-core
// FMAs per weight byte no longer fit under the HBM time.  So the WEIGHT rows are the UMMA M dimension (128 per tile, fed by TMA with the
// 128B swizzle straight into the canonical layout) and the tokens are N (16 / 32 / 64 accumulator columns in TMEM): per 16 KB weight
// panel the tensor core issues four M128 N16 K16 instructions (~64 cycles each -- their operand fetch, not the math), 2.7 x faster than
// the panel arrives from HBM.
//   warp 0     TMA producer: W panel [128 rows x 64 k] (two boxes of 64 rows: plain = consecutive rows, SwiGLU = 64 gate rows + the 64
//              up rows of the same columns) + X panel [TN tokens x 64 k] per stage; the weight panels of the first ring fill are
//              requested BEFORE griddepcontrol.wait (weights are immutable), i.e. while the previous kernel of the chain still drains
//   warp 1     UMMA issuer
//   warps 2-5  RMSNorm prologue (optional) while the ring streams, then the epilogue: thread = weight row, TN token columns
// RMSNorm prologue (HF Qwen2RMSNorm in front of q/k/v, gate/up and lm_head): the four otherwise idle epilogue warps compute the row
// statistics of X (same summation order as add_rmsnorm_kernel: bit-identical rstd), then normalise every X panel IN PLACE in shared
// memory as it lands (bf16(w * bf16(x * rstd)), the reference's rounding points) and hand it to the UMMA warp: no separate norm kernel,
// no normalised copy of X in global memory.
// Small N (o_proj / down_proj: 28 row tiles) cannot fill 148 SMs with one CTA per tile: the K range of a tile is split over the CTAs of
// a CLUSTER (<= 8), every CTA streams its K slice into its own TMEM accumulator, the non-leaders push their fp32 partial tiles into the
// leader's shared memory (st.shared::cluster; the leader's dead ring is the landing zone), and the leader sums them in rank order
// (deterministic) and runs the epilogue: no workspace in global memory, no second kernel, no atomics.
// Epilogues (rounding points as in the op-by-op reference, language_forward.py:175-177, 255 and HF's Qwen2MLP):
//   0  y = bf16(acc + bias)                       lm_head, plain projections
//   2  y = bf16(bf16(acc + bias) + residual)      o_proj / down_proj with the residual add
//   3  y = bf16(bf16(silu(bf16 gate)) * bf16 up)  gate/up projection, W = [gate rows ; up rows] as the engine stores it
//   4  q/k/v projection of the batched decode step: a 128-row tile is exactly one head, so the epilogue rotates it (mRoPE of a text
//      token: one position on all three axes; partner element 64 rows away through shared memory) and writes q_rot [T][Hq][128] and
//      the K_raw / K_rot / V rows of token t into the arena of stream t at the row its device-side counter names
//      (language_forward.py:47-66, 190-192; generate/streaming_cache.py:72-73) -- no qkv tensor, no append launch.
#include "common.cuh"

namespace svlm {

constexpr int TOK_BK = 64;
constexpr int TOK_THREADS = 192;
constexpr int TOK_W_BYTES = 128 * TOK_BK * 2;          // 16 KB weight panel per stage
constexpr int TOK_MAX_STAGES = 12;
constexpr int TOK_MAX_SPLITS = 8;
constexpr int TOK_MAX_STAT_TILES = 64;            // statistics hand-off: producer row tiles (hidden <= 8192)

enum { TOK_EPI_NONE = 0, TOK_EPI_RESIDUAL = 2, TOK_EPI_SWIGLU = 3, TOK_EPI_QKV_APPEND = 4, TOK_EPI_ARGMAX = 5 };

struct TokAppend {                                   // TOK_EPI_QKV_APPEND
    uint16_t* buf[SVLM_MAX_STREAMS];                 // live buffer of each stream's arena
    const int32_t* row[SVLM_MAX_STREAMS];            // device scalars: arena row of the token
    const int32_t* pos[SVLM_MAX_STREAMS];            // device scalars: its text position
    const uint16_t* cos;                             // bf16 [rows][64]
    const uint16_t* sin;
    int table_rows;
    int layer, hq, L, H, cap;
    uint16_t* q_out;                                 // bf16 [T][hq][128]
};

struct TokParams {
    int T, N, K, epi;
    int splits, stages;
    int x_slots, x_bytes;       // X ring: 2 x stages slots of x_bytes (8 or TN token rows x 64 k)
    int ring_bytes;             // max(W ring + X ring, landing zone of the split-K partial tiles + the epilogue exchange)
    int n_half;                 // SwiGLU: N / 2 (output columns); else 0
    const uint16_t* bias;
    const uint16_t* residual;   // bf16 [T][ldy]
    uint16_t* Y;
    int ldy;
    const uint16_t* x_raw;      // RMSNorm prologue: X in global memory (row statistics), norm weight [K], eps
    const uint16_t* norm_w;
    float eps;
    const float* stats_in;      // RMSNorm prologue: per-(row tile of the producer, token) sums of squares of X (stats_tiles x T), or NULL
    int stats_tiles;
    float* stats_out;           // residual epilogue: the same for the rows of Y this tile wrote ([tiles][T]), or NULL
    const uint8_t* suppress;    // TOK_EPI_ARGMAX: rows (vocabulary ids) that are never selected, or NULL
    unsigned long long* keys;   // TOK_EPI_ARGMAX: per token, max over rows of (ordered bf16 logit << 32 | ~row): first index of the maximum
};

struct TokShared {
    uint64_t full[TOK_MAX_STAGES], empty[TOK_MAX_STAGES], xfull[2 * TOK_MAX_STAGES], xready[2 * TOK_MAX_STAGES];
    uint64_t acc_full;
    uint32_t tmem_base;
    float rstd[64];
    float red[4][4];
    float stat[16][4];
    float stage_stats[TOK_MAX_STAT_TILES * 16];
    unsigned long long best[16][4];
};

__device__ __forceinline__ void st_shared_cluster_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }      // the four epilogue warps
__device__ __forceinline__ float bf16_at(const uint16_t* p) { return __uint_as_float((uint32_t)__ldg(p) << 16); }
__device__ __forceinline__ uint16_t to_bf16_bits(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t tok_ordered_f32(float f) {      // monotonic float -> uint32 map
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int TN>
__device__ __forceinline__ void tok_load_acc(uint32_t taddr, float (&v)[TN]) {
#pragma unroll
    for (int c = 0; c < TN / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(taddr + c * 16, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[c * 16 + i] = __uint_as_float(r[i]);
    }
}

template <int TN>
__global__ void __launch_bounds__(TOK_THREADS, 2)
gemm_tokens_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const TokParams p,
                   const __grid_constant__ TokAppend ap) {
    constexpr int TMEM_COLS = TN < 32 ? 32 : TN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = smem;                                                     // stages x W panel (16 KB)
    // X panels ([8 or TN tokens] x 64 k) have their own ring of 2 x stages slots and are requested a whole W-ring cycle AHEAD of the
    // weight panel they meet: they come from L2, so they land -- and, with the RMSNorm prologue, are normalised -- long before that
    // weight panel arrives from HBM, and a weight slot is never held waiting for its activations.
    uint8_t* xring = ring + p.stages * TOK_W_BYTES;
    // The leader's ring doubles as the landing zone of the peers' partial tiles ([splits - 1][TN][128] fp32) and, behind it, of the
    // epilogue's row exchange (SwiGLU: [TN][64] up values; q/k rotation: [TN][128]): both are used only after this CTA's last UMMA
    // has completed, i.e. when the ring is dead.
    float* red = reinterpret_cast<float*>(ring);
    float* xch = red + (p.splits - 1) * TN * 128;
    TokShared* sh = reinterpret_cast<TokShared*>(ring + p.ring_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int splits = p.splits;
    const int tile = blockIdx.x / splits;
    const int split = blockIdx.x - tile * splits;               // == rank in the cluster (cluster dims = (splits, 1, 1))
    const int k_blocks = (p.K + TOK_BK - 1) / TOK_BK;
    const int kb0 = (int)(((long long)k_blocks * split) / splits), kb1 = (int)(((long long)k_blocks * (split + 1)) / splits);
    const int n_kb = kb1 - kb0;
    const bool swiglu = p.epi == TOK_EPI_SWIGLU;
    const bool norm = p.norm_w != nullptr;
    // rows of W behind the two 64-row boxes of a panel
    const int row_a = swiglu ? tile * 64 : tile * 128;
    const int row_b = swiglu ? p.n_half + tile * 64 : tile * 128 + 64;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
        for (int s = 0; s < p.x_slots; ++s) { mbar_init(&sh->xfull[s], 1); mbar_init(&sh->xready[s], 4); }
        mbar_init(&sh->acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(&sh->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh->tmem_base;
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;                              // epilogue warps: weight row of the tile (TMEM lane)
    float pre_a[16], pre_b[SVLM_MAX_STREAMS];                    // epilogue operands fetched early (residual | cos, sin)
    int pre_row[SVLM_MAX_STREAMS];

    if (warp == 0) {
        // ---- TMA producer.  The weight panels of the first ring fill do not depend on the previous kernel: they are requested
        // first, the activations (its output) after griddepcontrol.wait.
        const int first = min(n_kb, p.stages);
        for (int it = 0; it < first; ++it) {
            if (elect_one()) {
                uint8_t* dst = ring + it * TOK_W_BYTES;
                mbar_arrive_expect_tx(&sh->full[it], TOK_W_BYTES);
                tma_load_2d(dst, &tmW, &sh->full[it], (kb0 + it) * TOK_BK, row_a);
                tma_load_2d(dst + TOK_W_BYTES / 2, &tmW, &sh->full[it], (kb0 + it) * TOK_BK, row_b);
            }
            __syncwarp();
        }
        pdl_wait();
        const int xfirst = min(n_kb, p.x_slots);
        for (int it = 0; it < xfirst; ++it) {
            if (elect_one()) {
                mbar_arrive_expect_tx(&sh->xfull[it], p.x_bytes);
                tma_load_2d(xring + it * p.x_bytes, &tmX, &sh->xfull[it], (kb0 + it) * TOK_BK, 0);
            }
            __syncwarp();
        }
        int s = first == p.stages ? 0 : first, ph = first == p.stages ? 1 : 0;      // slot / use count parity of iteration `first`
        for (int it = first; it < n_kb; ++it) {
            mbar_wait(&sh->empty[s], ph ^ 1);                   // UMMA (it - stages) is complete: its W slot and its X slot are free
            if (elect_one()) {
                uint8_t* dst = ring + s * TOK_W_BYTES;
                mbar_arrive_expect_tx(&sh->full[s], TOK_W_BYTES);
                tma_load_2d(dst, &tmW, &sh->full[s], (kb0 + it) * TOK_BK, row_a);
                tma_load_2d(dst + TOK_W_BYTES / 2, &tmW, &sh->full[s], (kb0 + it) * TOK_BK, row_b);
                const int xi = it + p.stages;                   // X panel one ring cycle ahead, into the slot of iteration it - stages
                if (xi < n_kb) {
                    const int xs = xi % p.x_slots;
                    mbar_arrive_expect_tx(&sh->xfull[xs], p.x_bytes);
                    tma_load_2d(xring + xs * p.x_bytes, &tmX, &sh->xfull[xs], (kb0 + xi) * TOK_BK, 0);
                }
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        pdl_trigger();                                          // the weight stream of this CTA is on its way: let the next kernel in
    } else if (warp == 1) {
        // ---- UMMA issuer: whole warp in uniform control flow, one elected lane issues
        pdl_wait();
        constexpr uint32_t IDESC = umma_idesc_bf16(128, TN, 0, 0);
        int s = 0, ph = 0, xs = 0, xph = 0;
        for (int it = 0; it < n_kb; ++it) {
            mbar_wait(norm ? &sh->xready[xs] : &sh->xfull[xs], xph);
            mbar_wait(&sh->full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t w = smem_u32(ring + s * TOK_W_BYTES), x = smem_u32(xring + xs * p.x_bytes);
#pragma unroll
                for (int kk = 0; kk < TOK_BK / 16; ++kk)
                    umma_ss(tmem, umma_desc_sw128(w + kk * 32, 16, 1024), umma_desc_sw128(x + kk * 32, 16, 1024), IDESC,
                            (it > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&sh->empty[s]);
                if (it == n_kb - 1) umma_commit(&sh->acc_full);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
            if (++xs == p.x_slots) { xs = 0; xph ^= 1; }
        }
    } else {
        pdl_wait();
        // What the epilogue of the leader will need from global memory is requested NOW, while the ring streams: at ~10 us per
        // launch an L2 round trip (two for the rotation: position, then table row) at the tail would be a tenth of the kernel.
        if (split == 0) {
            if (p.epi == TOK_EPI_QKV_APPEND) {
                if constexpr (TN == 16) {
#pragma unroll
                    for (int t = 0; t < SVLM_MAX_STREAMS; ++t) {
                        if (t < p.T) {
                            const int pp = min(max(__ldg(ap.pos[t]), 0), ap.table_rows - 1);
                            pre_a[t] = bf16_at(ap.cos + (size_t)pp * 64 + (row & 63));
                            pre_b[t] = bf16_at(ap.sin + (size_t)pp * 64 + (row & 63));
                            pre_row[t] = __ldg(ap.row[t]);
                        }
                    }
                }
            } else if (p.epi == TOK_EPI_RESIDUAL) {
                if constexpr (TN == 16) {
                    const int n = tile * 128 + row;
#pragma unroll
                    for (int t = 0; t < TN; ++t)
                        if (t < p.T && n < p.N) pre_a[t] = __uint_as_float((uint32_t)p.residual[(size_t)t * p.ldy + n] << 16);
                }
            }
        }
        if (norm) {
            // ---- RMSNorm prologue on the four epilogue warps.  Row statistics: thread e of 128 owns the 8-element vectors
            // (i * 128 + e) of a row, xor-shuffle tree, four warp sums added in order -- add_rmsnorm_kernel's order exactly.
            // Four rows per pass with all of a pass's loads in flight at once (one L2 round trip per pass; the weight ring is
            // filling from HBM meanwhile).
            const int e = threadIdx.x - 64, ew = e >> 5;
            if (p.stats_in) {
                // the producer of X (o_proj / down_proj with the residual epilogue) left the sum of squares of every 128-column slice
                // of every row: one short L2 round trip instead of reading X again in each of the grid's CTAs
                // (all entries requested at once by the 128 threads: ONE round trip -- under the weight stream an L2 hit takes
                // microseconds, and the first UMMA waits for this)
                float* stg = reinterpret_cast<float*>(sh->stage_stats);
                const int n_ent = p.stats_tiles * p.T;
                for (int i = e; i < n_ent; i += 128) stg[i] = p.stats_in[i];
                epi_bar_sync();
                if (e < p.T) {
                    float tot = 0.f;
                    for (int i = 0; i < p.stats_tiles; ++i) tot = __fadd_rn(tot, stg[i * p.T + e]);
                    sh->rstd[e] = rsqrtf(__fadd_rn(__fdiv_rn(tot, (float)p.K), p.eps));
                }
                epi_bar_sync();
            }
            for (int t0 = 0; t0 < (p.stats_in ? 0 : p.T); t0 += 4) {
                float ss[4] = {0.f, 0.f, 0.f, 0.f};
                for (int c0 = 0; c0 < p.K; c0 += 4096) {
                    uint4 buf[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = c0 + (i * 128 + e) * 8;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            buf[i][j] = (c < p.K && t0 + j < p.T) ? *reinterpret_cast<const uint4*>(p.x_raw + (size_t)(t0 + j) * p.K + c)
                                                                   : make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (c0 + (i * 128 + e) * 8 < p.K) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) ss[j] = sumsq8(ss[j], buf[i][j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], d);
                    if (lane == 0) sh->red[j][ew] = ss[j];
                }
                epi_bar_sync();
                if (e < 4 && t0 + e < p.T) sh->rstd[t0 + e] = rms_rstd(sh->red[e][0], sh->red[e][1], sh->red[e][2], sh->red[e][3], p.K, p.eps);
                epi_bar_sync();
            }
            // ---- normalise every X panel in place as it lands: thread = one 16-byte chunk (row = e / 8 (+16, +32 ...), chunk e % 8);
            // the norm-weight chunk of the next panel is fetched while this one is awaited
            const int c = e & 7;
            // The norm-weight chunks are fetched FOUR panels ahead (register ring): under the weight stream an L2 hit costs about one
            // panel period, and a converter that waited for one per panel paced the whole kernel (measured: 47.8 instead of 42.6 us).
            int xs = 0, xph = 0;
            uint4 wq[4], wn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = (kb0 + u) * TOK_BK + c * 8;
                wq[u] = (u < n_kb && k < p.K) ? __ldg(reinterpret_cast<const uint4*>(p.norm_w + k)) : make_uint4(0, 0, 0, 0);
            }
            for (int it0 = 0; it0 < n_kb; it0 += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = (kb0 + it0 + 4 + u) * TOK_BK + c * 8;
                    wn[u] = (it0 + 4 + u < n_kb && k < p.K) ? __ldg(reinterpret_cast<const uint4*>(p.norm_w + k)) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (it0 + u < n_kb) {
                        mbar_wait(&sh->xfull[xs], xph);
                        uint8_t* xp = xring + xs * p.x_bytes;
                        const uint32_t* q = &wq[u].x;
#pragma unroll
                        for (int r0 = 0; r0 < TN; r0 += 16) {
                            const int r = r0 + (e >> 3);
                            if (r < p.T) {
                                uint4* cell = reinterpret_cast<uint4*>(xp + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
                                uint4 xv = *cell;
                                uint32_t* px = &xv.x;
                                const float rs = sh->rstd[r];
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    px[i] = pack_bf16(__fmul_rn(bf16_lo(q[i]), round_bf16(__fmul_rn(bf16_lo(px[i]), rs))),
                                                      __fmul_rn(bf16_hi(q[i]), round_bf16(__fmul_rn(bf16_hi(px[i]), rs))));
                                *cell = xv;
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sh->xready[xs]);
                        if (++xs == p.x_slots) { xs = 0; xph ^= 1; }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) wq[u] = wn[u];
            }
        }
    }

    // ---- epilogue.  Thread = one weight row of the tile (TMEM lane), TN token columns.
    float v[TN];
    mbar_wait(&sh->acc_full, 0);                                // every warp: all UMMAs of this CTA are complete, its ring is dead
    tc_fence_after();
    if (warp >= 2) tok_load_acc<TN>(tmem + ((uint32_t)(q4 * 32) << 16), v);
    if (splits > 1) {
        cluster_arrive_release();                               // phase 1: "my ring is dead" (the leader's becomes the landing zone)
        cluster_wait_acquire();
        if (warp >= 2 && split > 0) {
            const uint32_t base = mapa_shared(smem_u32(red + ((split - 1) * TN) * 128 + row), 0);
#pragma unroll
            for (int t = 0; t < TN; ++t)
                if (t < p.T) st_shared_cluster_f32(base + t * 128 * 4, v[t]);
        }
        cluster_arrive_release();                               // phase 2: the partial tiles are in the leader's shared memory
        cluster_wait_acquire();
        if (warp >= 2 && split == 0) {
            for (int r = 0; r < splits - 1; ++r) {
#pragma unroll
                for (int t = 0; t < TN; ++t)
                    if (t < p.T) v[t] += red[(r * TN + t) * 128 + row];
            }
        }
    }
    if (warp >= 2 && split == 0) {
        if (swiglu) {
            // rows 0-63 of the tile are gate_j, rows 64-127 up_j of the same 64 output columns j
            if (row >= 64) {
#pragma unroll
                for (int t = 0; t < TN; ++t) xch[t * 64 + (row - 64)] = round_bf16(v[t]);
            }
            epi_bar_sync();
            const int j = tile * 64 + row;
            if (row < 64 && j < p.n_half) {
#pragma unroll
                for (int t = 0; t < TN; ++t) {
                    if (t < p.T) {
                        const float gt = round_bf16(v[t]), up = xch[t * 64 + row];
                        p.Y[(size_t)t * p.ldy + j] = to_bf16_bits(round_bf16(gt / (1.0f + __expf(-gt))) * up);
                    }
                }
            }
        } else if (p.epi == TOK_EPI_QKV_APPEND) {
            if constexpr (TN == 16) {
                // the tile is head `tile` of [q heads | k heads | v heads]; row = element of the head
                const float b = p.bias ? bf16_at(p.bias + tile * 128 + row) : 0.f;
                const bool is_q = tile < ap.hq, is_v = tile >= ap.hq + ap.H;
                const int h = is_q ? 0 : (is_v ? tile - ap.hq - ap.H : tile - ap.hq);
                const int64_t plane = (int64_t)ap.L * ap.H * ap.cap * 128;                       // ArenaDims::plane_elems
                const int64_t head_off = ((int64_t)ap.layer * ap.H + h) * ap.cap * 128 + row;
                if (!is_v) {
#pragma unroll
                    for (int t = 0; t < SVLM_MAX_STREAMS; ++t)
                        if (t < p.T) xch[t * 128 + row] = round_bf16(v[t] + b);
                    epi_bar_sync();
                }
#pragma unroll
                for (int t = 0; t < SVLM_MAX_STREAMS; ++t) {
                    if (t < p.T) {
                        const float x = round_bf16(v[t] + b);
                        float y = x;
                        if (!is_v) {
                            // out = bf16(bf16(x * cos) + bf16(rot(x) * sin)), rot(x) = (-x[64:], x[:64]): rope_chunk's arithmetic
                            const float partner = xch[t * 128 + (row ^ 64)];
                            y = __fadd_rn(round_bf16(__fmul_rn(x, pre_a[t])), round_bf16(__fmul_rn(row < 64 ? -partner : partner, pre_b[t])));
                        }
                        if (is_q) {
                            ap.q_out[((size_t)t * ap.hq + tile) * 128 + row] = to_bf16_bits(y);
                        } else {
                            const int arow = pre_row[t];
                            if (arow >= 0 && arow < ap.cap) {
                                uint16_t* dst = ap.buf[t] + head_off + (int64_t)arow * 128;
                                if (is_v) {
                                    dst[SVLM_PLANE_V * plane] = to_bf16_bits(x);
                                } else {
                                    dst[SVLM_PLANE_K_RAW * plane] = to_bf16_bits(x);
                                    dst[SVLM_PLANE_K_ROT * plane] = to_bf16_bits(y);
                                }
                            }
                        }
                    }
                }
            }
        } else if (p.epi == TOK_EPI_ARGMAX) {
            if constexpr (TN == 16) {
                // lm_head of the batched decode step: no logits tensor -- per token the first index of the maximum bf16 logit over
                // the rows that are not suppressed (torch.argmax of `lm_head(h).float().masked_fill(suppress, -inf)`): warp maximum,
                // four warp maxima through shared memory, ONE atomicMax per token and tile on a packed (value, ~row) key
                const int n = tile * 128 + row;
                const bool ok = n < p.N && !(p.suppress && p.suppress[n]);
#pragma unroll
                for (int t = 0; t < TN; ++t) {
                    unsigned long long key = 0ull;
                    if (ok && t < p.T)
                        key = ((unsigned long long)tok_ordered_f32(round_bf16(v[t])) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, d);
                        key = o > key ? o : key;
                    }
                    if (lane == 0) sh->best[t][q4] = key;
                }
                epi_bar_sync();
                if (row < p.T) {
                    unsigned long long key = sh->best[row][0];
#pragma unroll
                    for (int w = 1; w < 4; ++w) key = sh->best[row][w] > key ? sh->best[row][w] : key;
                    if (key) atomicMax(p.keys + row, key);
                }
            }
        } else {
            const int n = tile * 128 + row;
            const float b = (p.bias && n < p.N) ? bf16_at(p.bias + n) : 0.f;
#pragma unroll
            for (int t = 0; t < TN; ++t) {
                if (t < p.T && n < p.N) {
                    float y = v[t] + b;
                    if (p.epi == TOK_EPI_RESIDUAL) {
                        if constexpr (TN == 16) y = round_bf16(y) + pre_a[t];
                        else y = round_bf16(y) + __uint_as_float((uint32_t)p.residual[(size_t)t * p.ldy + n] << 16);
                    }
                    y = round_bf16(y);
                    p.Y[(size_t)t * p.ldy + n] = to_bf16_bits(y);
                    v[t] = y;
                } else {
                    v[t] = 0.f;
                }
            }
            if constexpr (TN == 16) {
                if (p.stats_out) {
                    // sum of squares of the 128 values this tile wrote per token (what the next RMSNorm needs of them): warp tree,
                    // four warp sums in order
#pragma unroll
                    for (int t = 0; t < TN; ++t) {
                        float q = __fmul_rn(v[t], v[t]);
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) q = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, d));
                        if (lane == 0) sh->stat[t][q4] = q;
                    }
                    epi_bar_sync();
                    if (row < p.T)
                        p.stats_out[tile * p.T + row] = __fadd_rn(__fadd_rn(__fadd_rn(sh->stat[row][0], sh->stat[row][1]), sh->stat[row][2]), sh->stat[row][3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
}

template <int TN>
static int launch_tokens(const CUtensorMap& tx, const CUtensorMap& tw, const TokParams& p, const TokAppend& ap, int tiles, size_t smem,
                         cudaStream_t st) {
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tokens_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm_tokens_kernel) failed: %s", cudaGetErrorString(e));
            return SVLM_ECUDA;
        }
        attr_smem = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles * p.splits);
    cfg.blockDim = dim3(TOK_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = p.splits;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    if (pdl_enabled(PDL_GEMV)) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaLaunchKernelEx(&cfg, gemm_tokens_kernel<TN>, tx, tw, p, ap);
    return launch_status("gemm_tokens_kernel");
}

// Token ids and device-side counters of the batched decode loop, from the keys the lm_head launch left: stream b takes the selected id
// and advances its kv length / next text position if it is ACTIVE (a stream whose turn has ended keeps its state); kv_after = kv_len + 1
// for every stream; the keys are cleared for the next step.
__global__ void streams_advance_kernel(unsigned long long* keys, int64_t* tok, const int32_t* active, int32_t* kv_len, int32_t* kv_after,
                                       int32_t* next_pos, int n_streams) {
    pdl_wait();
    const int b = threadIdx.x;
    if (b >= n_streams) return;
    const unsigned long long k = keys[b];
    keys[b] = 0ull;
    int n = kv_len[b];
    if (active[b]) {
        tok[b] = (int64_t)(0xffffffffu - (uint32_t)(k & 0xffffffffull));
        n += 1;
        kv_len[b] = n;
        next_pos[b] = next_pos[b] + 1;
    }
    kv_after[b] = n + 1;
}

// Split count for a shape: ~1.5 CTAs per SM (two fit; measured on B200 at T = 8: q/k/v 36 tiles -> 6 splits 8.7 us, down_proj 28 tiles
// -> 8 splits 27.5 us, 4 splits 30.5 us, no split 77 us), at least 8 K-blocks per CTA, at most 8 CTAs per cluster.
static int pick_splits(int tiles, int k_blocks, int sms) {
    if (tiles >= sms) return 1;
    int s = (3 * sms + tiles) / (2 * tiles);
    if (s > TOK_MAX_SPLITS) s = TOK_MAX_SPLITS;
    while (s > 1 && k_blocks / s < 8) --s;
    return s < 1 ? 1 : s;
}

static int tokens_common(const void* X, const void* W, const void* bias, const void* residual, void* Y, int T, int N, int K, int epilogue,
                         int splits, const void* norm_weight, float eps, const float* stats_in, int stats_tiles, float* stats_out,
                         const TokAppend& ap, void* stream, const uint8_t* suppress = nullptr, unsigned long long* keys = nullptr) {
    SVLM_REQUIRE(X && W, "NULL tensor argument");
    SVLM_REQUIRE(T >= 1 && T <= 64, "1..64 tokens per call (got %d); wider batches belong to svlm_gemm_bf16", T);
    SVLM_REQUIRE(N > 0 && K > 0 && K % 8 == 0, "bad N / K (K %d must be a multiple of 8)", K);
    SVLM_REQUIRE((((uintptr_t)X | (uintptr_t)W | (uintptr_t)norm_weight) & 15) == 0, "X / W / norm_weight must be 16-byte aligned");
    SVLM_REQUIRE(splits >= 0 && splits <= TOK_MAX_SPLITS, "splits must be 0 (auto) .. %d", TOK_MAX_SPLITS);
    SVLM_REQUIRE(stats_in == nullptr || (norm_weight != nullptr && stats_tiles > 0 && stats_tiles <= TOK_MAX_STAT_TILES && T <= 16),
                 "stats_in needs norm_weight, 1..%d producer tiles and <= 16 tokens", TOK_MAX_STAT_TILES);
    SVLM_REQUIRE(stats_out == nullptr || (epilogue != TOK_EPI_SWIGLU && epilogue != TOK_EPI_QKV_APPEND && T <= 16),
                 "stats_out: plain / residual epilogue, <= 16 tokens");
    int rc;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        SVLM_CUDA(cudaGetDevice(&dev));
        SVLM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int tn = T <= 16 ? 16 : (T <= 32 ? 32 : 64);
    const int n_half = epilogue == TOK_EPI_SWIGLU ? N / 2 : 0;
    const int tiles = epilogue == TOK_EPI_SWIGLU ? (n_half + 63) / 64 : (N + 127) / 128;
    const int k_blocks = (K + TOK_BK - 1) / TOK_BK;
    if (splits == 0) splits = pick_splits(tiles, k_blocks, sms);
    if (splits > k_blocks) splits = k_blocks;
    // Shared memory: <= ~110 KB whenever the grid is clustered or larger than the SM count, so that two CTAs share an SM (their
    // set-up and drain overlap, and clusters are placed far more freely: with one 227 KB CTA per SM, 5-CTA clusters did not tile the
    // GPCs and down_proj ran in two waves -- 41 us against 27 us); the deepest ring otherwise.
    // X panels: 8 token rows when T <= 8 (the UMMA reads 16 rows: rows 8-15 are whatever follows the slot -- the columns they
    // produce are never read), else TN rows
    const int x_rows = T <= 8 ? 8 : tn;
    const int x_bytes = x_rows * TOK_BK * 2;
    const int stage_bytes = TOK_W_BYTES + 2 * x_bytes;           // one W slot + two X slots
    const int fixed = 1024 + 1024 + (int)sizeof(TokShared) + 64;  // alignment slack + the 8 rows read beyond the last X slot
    const int budget = ((splits > 1 || tiles > sms) ? 113 : 227) * 1024;
    int stages = (budget - fixed) / stage_bytes;
    if (stages > TOK_MAX_STAGES) stages = TOK_MAX_STAGES;
    const int per_cta = (k_blocks + splits - 1) / splits;
    if (stages > per_cta) stages = per_cta;
    if (stages < 1) stages = 1;
    int ring_bytes = stages * stage_bytes + 1024;
    const int landing = (splits - 1) * tn * 128 * 4 + (epilogue == TOK_EPI_SWIGLU ? tn * 64 * 4 : (epilogue == TOK_EPI_QKV_APPEND ? tn * 128 * 4 : 0));
    if (ring_bytes < landing) ring_bytes = landing;
    ring_bytes = (ring_bytes + 1023) & ~1023;
    CUtensorMap tx, tw;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)T};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {TOK_BK, (uint32_t)x_rows};
        if ((rc = make_tensor_map_bf16(&tx, X, 2, dims, str, box, 1))) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {TOK_BK, 64};
        if ((rc = make_tensor_map_bf16(&tw, W, 2, dims, str, box, 1))) return rc;
    }
    TokParams p{T, N, K, epilogue, splits, stages, 2 * stages, x_bytes, ring_bytes, n_half, (const uint16_t*)bias, (const uint16_t*)residual, (uint16_t*)Y,
                epilogue == TOK_EPI_SWIGLU ? n_half : N, (const uint16_t*)X, (const uint16_t*)norm_weight, eps, stats_in, stats_tiles,
                stats_out, suppress, keys};
    const size_t smem = (size_t)fixed + (size_t)ring_bytes;
    switch (tn) {
        case 16: return launch_tokens<16>(tx, tw, p, ap, tiles, smem, (cudaStream_t)stream);
        case 32: return launch_tokens<32>(tx, tw, p, ap, tiles, smem, (cudaStream_t)stream);
        default: return launch_tokens<64>(tx, tw, p, ap, tiles, smem, (cudaStream_t)stream);
    }
}

}  // namespace svlm

using namespace svlm;

extern "C" int svlm_gemm_tokens(const void* X, const void* W, const void* bias, const void* residual, void* Y, int T, int N, int K,
                                int epilogue, int splits, const void* norm_weight, float eps, const float* stats_in, int stats_tiles,
                                float* stats_out, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(Y != nullptr, "Y is NULL");
    SVLM_REQUIRE(epilogue == TOK_EPI_NONE || epilogue == TOK_EPI_RESIDUAL || epilogue == TOK_EPI_SWIGLU,
                 "epilogue must be 0 (bias), 2 (residual add) or 3 (swiglu)");
    SVLM_REQUIRE(epilogue != TOK_EPI_RESIDUAL || residual != nullptr, "residual is NULL");
    SVLM_REQUIRE(epilogue != TOK_EPI_SWIGLU || (N % 2 == 0 && bias == nullptr), "swiglu epilogue: even N (gate rows, then up rows), no bias");
    TokAppend ap{};
    return tokens_common(X, W, bias, residual, Y, T, N, K, epilogue, splits, norm_weight, eps, stats_in, stats_tiles, stats_out, ap, stream);
}

extern "C" int svlm_gemm_tokens_qkv_append(const svlm_arena_t* geometry, const svlm_stream_slot_t* slots, int n_streams, int layer,
                                           const void* X, const void* W, const void* bias, const void* norm_weight, float eps,
                                           const float* stats_in, int stats_tiles, int q_heads, int hidden, void* q_rot,
                                           const void* cos_table, const void* sin_table, int table_rows, int splits, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(geometry && slots, "NULL geometry / slots");
    SVLM_REQUIRE(geometry->layers > 0 && geometry->kv_heads > 0 && geometry->capacity > 0 && geometry->head_dim == 128, "bad arena geometry");
    SVLM_REQUIRE(n_streams >= 1 && n_streams <= SVLM_MAX_STREAMS, "n_streams %d out of range (1..%d)", n_streams, SVLM_MAX_STREAMS);
    SVLM_REQUIRE(layer >= 0 && layer < geometry->layers, "layer %d out of range", layer);
    SVLM_REQUIRE(q_heads > 0 && hidden > 0, "bad q_heads / hidden");
    SVLM_REQUIRE(q_rot && cos_table && sin_table && table_rows > 0, "q_rot / rope tables missing");
    TokAppend ap{};
    const int64_t buffer_elems = 3 * (int64_t)geometry->layers * geometry->kv_heads * geometry->capacity * geometry->head_dim;
    for (int b = 0; b < n_streams; ++b) {
        SVLM_REQUIRE(slots[b].arena_base && slots[b].row_dev && slots[b].pos_dev, "slot %d: NULL pointer", b);
        SVLM_REQUIRE(slots[b].buffer == 0 || slots[b].buffer == 1, "slot %d: buffer must be 0 or 1", b);
        ap.buf[b] = (uint16_t*)slots[b].arena_base + (int64_t)slots[b].buffer * buffer_elems;
        ap.row[b] = slots[b].row_dev;
        ap.pos[b] = slots[b].pos_dev;
    }
    ap.cos = (const uint16_t*)cos_table;
    ap.sin = (const uint16_t*)sin_table;
    ap.table_rows = table_rows;
    ap.layer = layer; ap.hq = q_heads; ap.L = geometry->layers; ap.H = geometry->kv_heads; ap.cap = geometry->capacity;
    ap.q_out = (uint16_t*)q_rot;
    const int N = (q_heads + 2 * geometry->kv_heads) * 128;
    return tokens_common(X, W, bias, nullptr, nullptr, n_streams, N, hidden, TOK_EPI_QKV_APPEND, splits, norm_weight, eps, stats_in,
                         stats_tiles, nullptr, ap, stream);
}

extern "C" int svlm_gemm_tokens_lm_head_argmax(const void* X, const void* W, const void* norm_weight, float eps, const float* stats_in,
                                               int stats_tiles, const uint8_t* suppress, int T, int N, int K, unsigned long long* keys,
                                               void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(keys != nullptr && T <= 16, "keys is NULL or more than 16 tokens (%d)", T);
    SVLM_REQUIRE(((uintptr_t)keys & 7) == 0, "keys must be 8-byte aligned");
    TokAppend ap{};
    return tokens_common(X, W, nullptr, nullptr, nullptr, T, N, K, TOK_EPI_ARGMAX, 1, norm_weight, eps, stats_in, stats_tiles, nullptr, ap, stream,
                         suppress, keys);
}

extern "C" int svlm_streams_advance(unsigned long long* keys, int64_t* tok, const int32_t* active, int32_t* kv_len, int32_t* kv_after,
                                    int32_t* next_pos, int n_streams, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(keys && tok && active && kv_len && kv_after && next_pos, "NULL argument");
    SVLM_REQUIRE(n_streams >= 1 && n_streams <= SVLM_MAX_STREAMS, "n_streams %d out of range (1..%d)", n_streams, SVLM_MAX_STREAMS);
    launch_pdl(PDL_OTHER, streams_advance_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, keys, tok, active, kv_len, kv_after, next_pos,
               n_streams);
    return launch_status("streams_advance_kernel");
}
