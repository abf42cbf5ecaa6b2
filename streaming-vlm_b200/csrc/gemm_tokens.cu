// "Few tokens x whole weight matrix" GEMM on tcgen05 / TMEM / TMA (sm_100a), swap-AB, split-K over a thread-block cluster:
//     Y[T][N] = epilogue(X[T][K] * W[N][K]^T + bias[N])          T <= 64 tokens, nn.Linear layout, fp32 accumulate
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
//   warps 2-5  epilogue: thread = weight row; tcgen05.ld of the TN token columns
// Small N (o_proj / down_proj: 28 row tiles) cannot fill 148 SMs with one CTA per tile: the K range of a tile is split over the CTAs of
// a CLUSTER (<= 8), every CTA streams its K slice into its own TMEM accumulator, the non-leaders push their fp32 partial tiles into the
// leader's shared memory (st.shared::cluster), one cluster barrier, and the leader sums them in rank order (deterministic) and runs the
// epilogue: no workspace in global memory, no second kernel, no atomics.
// Epilogues (rounding points as in the op-by-op reference, language_forward.py:175-177, 255 and HF's Qwen2MLP):
//   0  y = bf16(acc + bias)                       q/k/v projection, lm_head
//   2  y = bf16(bf16(acc + bias) + residual)      o_proj / down_proj with the residual add
//   3  y = bf16(bf16(silu(bf16 gate)) * bf16 up)  gate/up projection, W = [gate rows ; up rows] as the engine stores it
#include "common.cuh"

namespace svlm {

constexpr int TOK_BK = 64;
constexpr int TOK_THREADS = 192;
constexpr int TOK_W_BYTES = 128 * TOK_BK * 2;          // 16 KB weight panel per stage
constexpr int TOK_MAX_STAGES = 12;
constexpr int TOK_MAX_SPLITS = 8;

enum { TOK_EPI_NONE = 0, TOK_EPI_RESIDUAL = 2, TOK_EPI_SWIGLU = 3 };

struct TokParams {
    int T, N, K, epi;
    int splits, stages;
    int n_half;                 // SwiGLU: N / 2 (output columns); else 0
    const uint16_t* bias;
    const uint16_t* residual;   // bf16 [T][ldy]
    uint16_t* Y;
    int ldy;
};

struct TokShared {
    uint64_t full[TOK_MAX_STAGES], empty[TOK_MAX_STAGES];
    uint64_t acc_full;
    uint32_t tmem_base;
};

__device__ __forceinline__ void st_shared_cluster_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

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
gemm_tokens_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const TokParams p) {
    constexpr int X_BYTES = TN * TOK_BK * 2;
    constexpr int STAGE_BYTES = TOK_W_BYTES + X_BYTES;
    constexpr int TMEM_COLS = TN < 32 ? 32 : TN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = smem;                                                     // stages x (W panel, X panel)
    float* red = reinterpret_cast<float*>(ring + p.stages * STAGE_BYTES);     // leader: [splits - 1][TN][128] partial tiles of the peers
    float* xch = red + (p.splits - 1) * TN * 128;                             // SwiGLU: [TN][64] up values
    TokShared* sh = reinterpret_cast<TokShared*>(xch + (p.epi == TOK_EPI_SWIGLU ? TN * 64 : 0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int splits = p.splits;
    const int tile = blockIdx.x / splits;
    const int split = blockIdx.x - tile * splits;               // == rank in the cluster (cluster dims = (splits, 1, 1))
    const int k_blocks = (p.K + TOK_BK - 1) / TOK_BK;
    const int kb0 = (int)(((long long)k_blocks * split) / splits), kb1 = (int)(((long long)k_blocks * (split + 1)) / splits);
    const int n_kb = kb1 - kb0;
    const bool swiglu = p.epi == TOK_EPI_SWIGLU;
    // rows of W behind the two 64-row boxes of a panel
    const int row_a = swiglu ? tile * 64 : tile * 128;
    const int row_b = swiglu ? p.n_half + tile * 64 : tile * 128 + 64;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
        mbar_init(&sh->acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(&sh->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh->tmem_base;
    if (splits > 1) cluster_arrive_relaxed();                   // phase 1: "this CTA runs, its shared memory exists"

    if (warp == 0) {
        // ---- TMA producer.  The weight panels of the first ring fill do not depend on the previous kernel: they are requested
        // first, the activations (its output) after griddepcontrol.wait.
        const int first = min(n_kb, p.stages);
        for (int it = 0; it < first; ++it) {
            if (elect_one()) {
                uint8_t* dst = ring + it * STAGE_BYTES;
                mbar_arrive_expect_tx(&sh->full[it], STAGE_BYTES);
                tma_load_2d(dst, &tmW, &sh->full[it], (kb0 + it) * TOK_BK, row_a);
                tma_load_2d(dst + TOK_W_BYTES / 2, &tmW, &sh->full[it], (kb0 + it) * TOK_BK, row_b);
            }
            __syncwarp();
        }
        pdl_wait();
        for (int it = 0; it < first; ++it) {
            if (elect_one()) tma_load_2d(ring + it * STAGE_BYTES + TOK_W_BYTES, &tmX, &sh->full[it], (kb0 + it) * TOK_BK, 0);
            __syncwarp();
        }
        int s = first == p.stages ? 0 : first, ph = first == p.stages ? 1 : 0;      // slot / use count parity of iteration `first`
        for (int it = first; it < n_kb; ++it) {
            mbar_wait(&sh->empty[s], ph ^ 1);
            if (elect_one()) {
                uint8_t* dst = ring + s * STAGE_BYTES;
                mbar_arrive_expect_tx(&sh->full[s], STAGE_BYTES);
                tma_load_2d(dst, &tmW, &sh->full[s], (kb0 + it) * TOK_BK, row_a);
                tma_load_2d(dst + TOK_W_BYTES / 2, &tmW, &sh->full[s], (kb0 + it) * TOK_BK, row_b);
                tma_load_2d(dst + TOK_W_BYTES, &tmX, &sh->full[s], (kb0 + it) * TOK_BK, 0);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        pdl_trigger();                                          // the weight stream of this CTA is on its way: let the next kernel in
    } else if (warp == 1) {
        // ---- UMMA issuer: whole warp in uniform control flow, one elected lane issues
        pdl_wait();
        constexpr uint32_t IDESC = umma_idesc_bf16(128, TN, 0, 0);
        int s = 0, ph = 0;
        for (int it = 0; it < n_kb; ++it) {
            mbar_wait(&sh->full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t w = smem_u32(ring + s * STAGE_BYTES), x = w + TOK_W_BYTES;
#pragma unroll
                for (int kk = 0; kk < TOK_BK / 16; ++kk)
                    umma_ss(tmem, umma_desc_sw128(w + kk * 32, 16, 1024), umma_desc_sw128(x + kk * 32, 16, 1024), IDESC,
                            (it > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&sh->empty[s]);
                if (it == n_kb - 1) umma_commit(&sh->acc_full);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
    } else {
        pdl_wait();
    }

    // ---- epilogue.  Thread = one weight row of the tile (TMEM lane), TN token columns.
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    float v[TN];
    if (warp >= 2) {
        if (n_kb > 0) {
            mbar_wait(&sh->acc_full, 0);
            tc_fence_after();
            tok_load_acc<TN>(tmem + ((uint32_t)(q4 * 32) << 16), v);
        } else {
#pragma unroll
            for (int t = 0; t < TN; ++t) v[t] = 0.f;
        }
    }
    if (splits > 1) {
        cluster_wait_acquire();                                 // phase 1 complete: every CTA of the cluster is running
        if (warp >= 2 && split > 0) {
            const uint32_t base = mapa_shared(smem_u32(red + ((split - 1) * TN) * 128 + row), 0);
#pragma unroll
            for (int t = 0; t < TN; ++t) st_shared_cluster_f32(base + t * 128 * 4, v[t]);
        }
        cluster_arrive_release();                               // phase 2: the partial tiles are in the leader's shared memory
        cluster_wait_acquire();
        if (warp >= 2 && split == 0) {
            for (int r = 0; r < splits - 1; ++r) {
#pragma unroll
                for (int t = 0; t < TN; ++t) v[t] += red[(r * TN + t) * 128 + row];
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
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int j = tile * 64 + row;
            if (row < 64 && j < p.n_half) {
#pragma unroll
                for (int t = 0; t < TN; ++t) {
                    if (t < p.T) {
                        const float gt = round_bf16(v[t]), up = xch[t * 64 + row];
                        const float a = round_bf16(gt / (1.0f + __expf(-gt))) * up;
                        p.Y[(size_t)t * p.ldy + j] = __bfloat16_as_ushort(__float2bfloat16_rn(a));
                    }
                }
            }
        } else {
            const int n = tile * 128 + row;
            if (n < p.N) {
                const float b = p.bias ? __uint_as_float((uint32_t)__ldg(p.bias + n) << 16) : 0.f;
#pragma unroll
                for (int t = 0; t < TN; ++t) {
                    if (t < p.T) {
                        float y = v[t] + b;
                        if (p.epi == TOK_EPI_RESIDUAL)
                            y = round_bf16(y) + __uint_as_float((uint32_t)p.residual[(size_t)t * p.ldy + n] << 16);
                        p.Y[(size_t)t * p.ldy + n] = __bfloat16_as_ushort(__float2bfloat16_rn(y));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
}

template <int TN>
static int launch_tokens(const CUtensorMap& tx, const CUtensorMap& tw, const TokParams& p, int tiles, size_t smem, cudaStream_t st) {
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
    cudaLaunchKernelEx(&cfg, gemm_tokens_kernel<TN>, tx, tw, p);
    return launch_status("gemm_tokens_kernel");
}

// Split count for a shape: fill the SMs (tiles x splits ~ one wave), keep >= 4 K-blocks per CTA, <= 8 CTAs per cluster.
static int pick_splits(int tiles, int k_blocks, int sms) {
    if (tiles >= sms) return 1;
    int s = sms / tiles;
    if (s > TOK_MAX_SPLITS) s = TOK_MAX_SPLITS;
    while (s > 1 && k_blocks / s < 4) --s;
    return s < 1 ? 1 : s;
}

}  // namespace svlm

using namespace svlm;

extern "C" int svlm_gemm_tokens(const void* X, const void* W, const void* bias, const void* residual, void* Y, int T, int N, int K,
                                int epilogue, int splits, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(X && W && Y, "NULL tensor argument");
    SVLM_REQUIRE(T >= 1 && T <= 64, "1..64 tokens per call (got %d); wider batches belong to svlm_gemm_bf16", T);
    SVLM_REQUIRE(N > 0 && K > 0 && K % 8 == 0, "bad N / K (K %d must be a multiple of 8)", K);
    SVLM_REQUIRE(epilogue == TOK_EPI_NONE || epilogue == TOK_EPI_RESIDUAL || epilogue == TOK_EPI_SWIGLU,
                 "epilogue must be 0 (bias), 2 (residual add) or 3 (swiglu)");
    SVLM_REQUIRE(epilogue != TOK_EPI_RESIDUAL || residual != nullptr, "residual is NULL");
    SVLM_REQUIRE(epilogue != TOK_EPI_SWIGLU || (N % 2 == 0 && bias == nullptr), "swiglu epilogue: even N (gate rows, then up rows), no bias");
    SVLM_REQUIRE((((uintptr_t)X | (uintptr_t)W) & 15) == 0, "X / W must be 16-byte aligned");
    SVLM_REQUIRE(splits >= 0 && splits <= TOK_MAX_SPLITS, "splits must be 0 (auto) .. %d", TOK_MAX_SPLITS);
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
    // shared memory: more CTAs than SMs -> <= ~110 KB so that two CTAs share an SM (their set-up and drain overlap), else the deepest ring
    const int stage_bytes = TOK_W_BYTES + tn * TOK_BK * 2;
    const int fixed = 1024 + (splits - 1) * tn * 128 * 4 + (epilogue == TOK_EPI_SWIGLU ? tn * 64 * 4 : 0) + (int)sizeof(TokShared) + 64;
    const int budget = (tiles * splits > sms ? 113 : 227) * 1024;
    int stages = (budget - fixed) / stage_bytes;
    if (stages > TOK_MAX_STAGES) stages = TOK_MAX_STAGES;
    const int per_cta = (k_blocks + splits - 1) / splits;
    if (stages > per_cta) stages = per_cta;
    SVLM_REQUIRE(stages >= 2 || per_cta < 2, "no room for a 2-stage ring (T %d, splits %d)", T, splits);
    if (stages < 1) stages = 1;
    CUtensorMap tx, tw;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)T};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {TOK_BK, (uint32_t)tn};
        if ((rc = make_tensor_map_bf16(&tx, X, 2, dims, str, box, 1))) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {TOK_BK, 64};
        if ((rc = make_tensor_map_bf16(&tw, W, 2, dims, str, box, 1))) return rc;
    }
    TokParams p{T, N, K, epilogue, splits, stages, n_half, (const uint16_t*)bias, (const uint16_t*)residual, (uint16_t*)Y,
                epilogue == TOK_EPI_SWIGLU ? n_half : N};
    const size_t smem = (size_t)fixed + (size_t)stages * stage_bytes;
    switch (tn) {
        case 16: return launch_tokens<16>(tx, tw, p, tiles, smem, (cudaStream_t)stream);
        case 32: return launch_tokens<32>(tx, tw, p, tiles, smem, (cudaStream_t)stream);
        default: return launch_tokens<64>(tx, tw, p, tiles, smem, (cudaStream_t)stream);
    }
}
