// Decode-step (M = 1) weight-streaming GEMV with fused prologue / epilogue -- HBM-bound (sm_100a).
//
// A decode forward of the 7B geometry streams 15.2 GB of weights once; everything else is noise.  One kernel
// template covers the four projections of a decoder layer and the lm_head:
//     prologue (optional)  x <- rmsnorm(x) * w_norm           (HF RMSNorm rounding: bf16(x*rsqrt) then bf16(w * .))
//     body                 y_n = sum_k W[n][k] * x[k]         bf16 weights, fp32 accumulate, 128-bit streaming loads
//     epilogue             PLAIN  : y = bf16(acc + bias)
//                          RESID  : y = bf16(residual + bf16(acc + bias))
//                          SWIGLU : W = [gate rows ; up rows], y_i = bf16(bf16(silu(bf16(g_i))) * bf16(u_i))
//                          ARGMAX : greedy token = first index of the maximum bf16 logit over non-suppressed rows
// CTA = 256 threads = 8 warps; a warp owns RW consecutive output rows (pairs for SWIGLU) and walks K with all
// 32 lanes (16 B per lane per row per step, two steps in flight); x lives in shared memory as bf16.
#include "common.cuh"

namespace svlm {

constexpr int GEMV_THREADS = 256;
enum { GEMV_PLAIN = 0, GEMV_RESID = 1, GEMV_SWIGLU = 2, GEMV_ARGMAX = 3 };

struct GemvParams {
    const uint16_t* W;
    const uint16_t* x;
    const uint16_t* norm_w;     // NULL = no RMSNorm prologue
    float eps;
    const uint16_t* bias;       // NULL ok
    const uint16_t* residual;   // RESID
    uint16_t* y;
    int N, K;                   // N = number of outputs (SWIGLU: W has 2N rows)
    const uint8_t* suppress;    // ARGMAX: 1 = never emit
    unsigned long long* key;    // ARGMAX: packed (ordered logit << 32 | ~index), reduced with atomicMax
};

__device__ __forceinline__ float dot8(const uint4& w, const uint4& x) {
    const uint32_t* a = &w.x;
    const uint32_t* b = &x.x;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += bf16_lo(a[i]) * bf16_lo(b[i]) + bf16_hi(a[i]) * bf16_hi(b[i]);
    return s;
}

__device__ __forceinline__ uint32_t ordered_f32(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int MODE, int RW>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_kernel(const GemvParams p) {
    extern __shared__ __align__(16) uint8_t gsm[];
    uint16_t* xs = reinterpret_cast<uint16_t*>(gsm);
    __shared__ float red[8];
    __shared__ unsigned long long best[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = p.K;

    // ---- prologue: stage x (optionally RMS-normalised) in shared memory
    float ss = 0.f;
    for (int k = threadIdx.x * 8; k < K; k += GEMV_THREADS * 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(p.x + k);
        *reinterpret_cast<uint4*>(xs + k) = v;
        if (p.norm_w) {
            const uint32_t* q = &v.x;
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float a = bf16_lo(q[i]), b = bf16_hi(q[i]); ss += a * a + b * b; }
        }
    }
    if (p.norm_w) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += red[i];
        const float r = rsqrtf(tot / (float)K + p.eps);
        for (int k = threadIdx.x * 8; k < K; k += GEMV_THREADS * 8) {
            uint4 v = *reinterpret_cast<uint4*>(xs + k);
            const uint4 wv = *reinterpret_cast<const uint4*>(p.norm_w + k);
            uint32_t* q = &v.x;
            const uint32_t* wq = &wv.x;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                q[i] = pack_bf16(bf16_lo(wq[i]) * round_bf16(bf16_lo(q[i]) * r), bf16_hi(wq[i]) * round_bf16(bf16_hi(q[i]) * r));
            *reinterpret_cast<uint4*>(xs + k) = v;
        }
    }
    __syncthreads();

    // ---- body: RW rows per warp (SWIGLU: RW gate rows + their RW up rows)
    constexpr int NR = (MODE == GEMV_SWIGLU) ? 2 * RW : RW;
    const int n0 = (blockIdx.x * 8 + warp) * RW;
    const uint16_t* rows[NR];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        const int n = min(n0 + r, p.N - 1);
        rows[r] = p.W + (int64_t)n * K;
        if (MODE == GEMV_SWIGLU) rows[RW + r] = p.W + (int64_t)(p.N + n) * K;
    }
    float acc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[r] = 0.f;
    int k = lane * 8;
    for (; k + 256 < K; k += 512) {
        uint4 w0[NR], w1[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) { w0[r] = ld_nc_v4(rows[r] + k); w1[r] = ld_nc_v4(rows[r] + k + 256); }
        const uint4 x0 = *reinterpret_cast<const uint4*>(xs + k);
        const uint4 x1 = *reinterpret_cast<const uint4*>(xs + k + 256);
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r] += dot8(w0[r], x0) + dot8(w1[r], x1);
    }
    for (; k < K; k += 256) {
        const uint4 x0 = *reinterpret_cast<const uint4*>(xs + k);
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r] += dot8(ld_nc_v4(rows[r] + k), x0);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], d);
    }

    // ---- epilogue
    if (MODE == GEMV_ARGMAX) {
        unsigned long long mine = 0ull;
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                const int n = n0 + r;
                if (n < p.N && !(p.suppress && p.suppress[n])) {
                    const float logit = round_bf16(acc[r]);
                    const unsigned long long key = ((unsigned long long)ordered_f32(logit) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
                    mine = key > mine ? key : mine;
                }
            }
            best[warp] = mine;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long b = 0ull;
#pragma unroll
            for (int i = 0; i < 8; ++i) b = best[i] > b ? best[i] : b;
            if (b) atomicMax(p.key, b);
        }
        return;
    }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const int n = n0 + r;
            if (n >= p.N) break;
            if (MODE == GEMV_SWIGLU) {
                const float g = round_bf16(acc[r]), u = round_bf16(acc[RW + r]);
                const float s = round_bf16(g / (1.0f + __expf(-g)));
                p.y[n] = __bfloat16_as_ushort(__float2bfloat16_rn(s * u));
            } else {
                float v = acc[r];
                if (p.bias) v += __bfloat162float(__ushort_as_bfloat16(p.bias[n]));
                if (MODE == GEMV_RESID) v = round_bf16(v) + __bfloat162float(__ushort_as_bfloat16(p.residual[n]));
                p.y[n] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
            }
        }
    }
}

// token id + device-side counters of the decode loop (kv length before / after the next step, next text position)
__global__ void decode_finalize_kernel(unsigned long long* key, int64_t* tok, int32_t* kv_len, int32_t* kv_after, int32_t* next_pos) {
    const unsigned long long k = *key;
    *tok = (int64_t)(0xffffffffu - (uint32_t)(k & 0xffffffffull));
    *key = 0ull;
    const int n = *kv_len + 1;
    *kv_len = n;
    *kv_after = n + 1;
    *next_pos = *next_pos + 1;
}

template <int MODE, int RW>
static int launch_gemv(const GemvParams& p, cudaStream_t st) {
    const size_t smem = (size_t)p.K * 2;
    static bool attr_done = false;
    if (!attr_done && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gemv_kernel<MODE, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemv_kernel) failed: %s", cudaGetErrorString(e));
            return SVLM_ECUDA;
        }
        attr_done = true;
    }
    const int grid = (p.N + 8 * RW - 1) / (8 * RW);
    gemv_kernel<MODE, RW><<<grid, GEMV_THREADS, smem, st>>>(p);
    return launch_status("gemv_kernel");
}

template <int MODE>
static int dispatch_rw(const GemvParams& p, cudaStream_t st) {
    // 4 rows per warp once that still gives every SM >= 2 CTAs; otherwise 2 (more CTAs for the small projections)
    const int rows_total = p.N;
    if (MODE != GEMV_SWIGLU && rows_total / 32 >= 2 * 148) return launch_gemv<MODE, 4>(p, st);
    return launch_gemv<MODE, 2>(p, st);
}

}  // namespace svlm

using namespace svlm;

static int gemv_common_checks(const void* W, const void* x, void* y, int N, int K) {
    SVLM_REQUIRE(W && x && y, "NULL tensor argument");
    SVLM_REQUIRE(N > 0 && K > 0 && K % 8 == 0, "bad N (%d) / K (%d): K must be a multiple of 8", N, K);
    SVLM_REQUIRE(K * 2 <= 160 * 1024, "K %d too large for the shared-memory x buffer", K);
    SVLM_REQUIRE((((uintptr_t)W | (uintptr_t)x) & 15) == 0, "W / x must be 16-byte aligned");
    return SVLM_OK;
}

extern "C" int svlm_gemv_bf16(const void* W, const void* x, const void* norm_weight, float eps, const void* bias,
                              const void* residual, void* y, int N, int K, int mode, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    if ((rc = gemv_common_checks(W, x, y, N, K))) return rc;
    SVLM_REQUIRE(mode >= GEMV_PLAIN && mode <= GEMV_SWIGLU, "mode must be 0 (plain), 1 (residual) or 2 (swiglu)");
    SVLM_REQUIRE(mode != GEMV_RESID || residual != nullptr, "residual is NULL");
    GemvParams p{(const uint16_t*)W, (const uint16_t*)x, (const uint16_t*)norm_weight, eps, (const uint16_t*)bias,
                 (const uint16_t*)residual, (uint16_t*)y, N, K, nullptr, nullptr};
    switch (mode) {
        case GEMV_PLAIN: return dispatch_rw<GEMV_PLAIN>(p, (cudaStream_t)stream);
        case GEMV_RESID: return dispatch_rw<GEMV_RESID>(p, (cudaStream_t)stream);
        default: return dispatch_rw<GEMV_SWIGLU>(p, (cudaStream_t)stream);
    }
}

extern "C" int svlm_lm_head_argmax(const void* W, const void* x, const void* norm_weight, float eps, const uint8_t* suppress, int N,
                                   int K, uint64_t* key_scratch, int64_t* token_out, int32_t* kv_len, int32_t* kv_after,
                                   int32_t* next_pos, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    if ((rc = gemv_common_checks(W, x, key_scratch, N, K))) return rc;
    SVLM_REQUIRE(token_out && kv_len && kv_after && next_pos, "NULL counter argument");
    GemvParams p{(const uint16_t*)W, (const uint16_t*)x, (const uint16_t*)norm_weight, eps, nullptr, nullptr, nullptr, N, K, suppress,
                 (unsigned long long*)key_scratch};
    if ((rc = dispatch_rw<GEMV_ARGMAX>(p, (cudaStream_t)stream))) return rc;
    decode_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)key_scratch, token_out, kv_len, kv_after, next_pos);
    return launch_status("decode_finalize_kernel");
}
