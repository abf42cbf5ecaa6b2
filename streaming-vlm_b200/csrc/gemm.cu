// Dense bf16 GEMM with fused epilogue on tcgen05 / TMEM / TMA (sm_100a):
//     C[M][N] = act(A[M][K] * W[N][K]^T + bias[N])          (nn.Linear layout, fp32 accumulate)
// Used for the ViT patch merger (Linear 5120->5120 + GELU(erf), Linear 5120->out_hidden, fused with the
// `argsort(window_index)` row un-permute).
//
// CTA = 192 threads, one 128 x BN output tile (BN = 64 or 128):
//   warp 0   TMA producer: A [128 x 64] and W [BN x 64] panels (128B swizzle) into a STAGES-deep ring
//   warp 1   UMMA issuer: 4 x (M128, N=BN, K16) per stage into one TMEM accumulator
//   warps 2-5 epilogue: tcgen05.ld the accumulator row, + bias, round to bf16, optional GELU, 128-bit stores
#include <math.h>

#include "common.cuh"

namespace svlm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN>
struct GemmCfg {
    static constexpr int STAGES = BN == 128 ? 6 : 8;
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int W_BYTES = BN * GEMM_BK * 2;
    static constexpr int SMEM = 1024 + STAGES * (A_BYTES + W_BYTES) + 256;
};

struct GemmParams {
    int M, N, K, act;
    const uint16_t* bias;
    const int32_t* out_rows;
    uint16_t* C;
};

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
    uint64_t* acc_done = empty + Cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * GEMM_BM;
    const int k_blocks = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        for (int kb = 0; kb < k_blocks; ++kb) {
            const int s = kb % Cfg::STAGES;
            mbar_wait(&empty[s], ((kb / Cfg::STAGES) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&full[s], Cfg::A_BYTES + Cfg::W_BYTES);
                tma_load_2d(sA + s * Cfg::A_BYTES, &tmA, &full[s], kb * GEMM_BK, m0);
                tma_load_2d(sW + s * Cfg::W_BYTES, &tmW, &full[s], kb * GEMM_BK, n0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // whole warp in uniform control flow, one elected lane issues (descriptors stay in uniform registers)
        constexpr uint32_t IDESC = umma_idesc_bf16(GEMM_BM, BN, 0, 0);
        for (int kb = 0; kb < k_blocks; ++kb) {
            const int s = kb % Cfg::STAGES;
            mbar_wait(&full[s], (kb / Cfg::STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a = smem_u32(sA + s * Cfg::A_BYTES), w = smem_u32(sW + s * Cfg::W_BYTES);
#pragma unroll
                for (int kk = 0; kk < GEMM_BK / 16; ++kk)
                    umma_ss(tmem, umma_desc_sw128(a + kk * 32, 16, 1024), umma_desc_sw128(w + kk * 32, 16, 1024), IDESC,
                            (kb > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&empty[s]);
                if (kb == k_blocks - 1) umma_commit(acc_done);
            }
            __syncwarp();
        }
    } else {
        const int q4 = warp & 3;
        const int row = q4 * 32 + lane;
        const int m = m0 + row;
        mbar_wait(acc_done, 0);
        tc_fence_after();
        const bool valid = m < p.M;
        const int64_t orow = valid ? (p.out_rows ? (int64_t)__ldg(p.out_rows + m) : (int64_t)m) : 0;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t acc[32];
            tmem_ld32(tmem + ((uint32_t)(q4 * 32) << 16) + c * 32, acc);
            tmem_wait_ld();
            if (!valid) continue;
            uint16_t* dst = p.C + orow * p.N + n0 + c * 32;
#pragma unroll
            for (int v8 = 0; v8 < 4; ++v8) {
                const int n = n0 + c * 32 + v8 * 8;
                if (n >= p.N) continue;
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(acc[v8 * 8 + i]);
                if (p.bias) {
                    const uint4 bv = *reinterpret_cast<const uint4*>(p.bias + n);
                    const uint32_t* pb = &bv.x;
#pragma unroll
                    for (int i = 0; i < 4; ++i) { x[2 * i] += bf16_lo(pb[i]); x[2 * i + 1] += bf16_hi(pb[i]); }
                }
                if (p.act == 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float y = round_bf16(x[i]);                 // Linear output is bf16 before nn.GELU sees it
                        x[i] = 0.5f * y * (1.0f + erff(y * 0.70710678118654752f));
                    }
                }
                uint4 o;
                o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]); o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
                *reinterpret_cast<uint4*>(dst + v8 * 8) = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<BN>(tmem);
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN>::SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm_bf16_kernel) failed: %s", cudaGetErrorString(e));
            return SVLM_ECUDA;
        }
        attr_done = true;
    }
    dim3 grid((p.N + BN - 1) / BN, (p.M + GEMM_BM - 1) / GEMM_BM);
    gemm_bf16_kernel<BN><<<grid, GEMM_THREADS, GemmCfg<BN>::SMEM, st>>>(ta, tw, p);
    return launch_status("gemm_bf16_kernel");
}

}  // namespace svlm

using namespace svlm;

extern "C" int svlm_gemm_bf16(const void* A, const void* W, const void* bias, void* C, int M, int N, int K, int act,
                              const int32_t* out_rows, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(A && W && C, "NULL tensor argument");
    SVLM_REQUIRE(M > 0 && N > 0 && K > 0, "bad M/N/K");
    SVLM_REQUIRE(K % 8 == 0 && N % 8 == 0, "K (%d) and N (%d) must be multiples of 8", K, N);
    SVLM_REQUIRE(act == 0 || act == 1, "act must be 0 (none) or 1 (gelu)");
    SVLM_REQUIRE((((uintptr_t)A | (uintptr_t)W | (uintptr_t)C | (uintptr_t)bias) & 15) == 0, "A/W/C/bias must be 16-byte aligned");
    const int tiles128 = ((N + 127) / 128) * ((M + 127) / 128);
    const int bn = tiles128 >= 148 ? 128 : 64;
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
        uint32_t box[2] = {GEMM_BK, (uint32_t)bn};
        if ((rc = make_tensor_map_bf16(&tw, W, 2, dims, str, box, 1))) return rc;
    }
    GemmParams p{M, N, K, act, (const uint16_t*)bias, out_rows, (uint16_t*)C};
    return bn == 128 ? launch_gemm<128>(ta, tw, p, (cudaStream_t)stream) : launch_gemm<64>(ta, tw, p, (cudaStream_t)stream);
}
