// Developer probe (not part of the product): issue-rate floors of the tcgen05 building blocks the attention kernels
// are made of, measured with clock64 on one B200.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I streaming-vlm_b200/csrc -I include \
//        benchmarks/hw_probe.cu -o gpurun_out/hw_probe && gpurun_out/hw_probe
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace svlm;

constexpr int SMEM_BYTES = 200 * 1024;

struct Res {
    long long cyc[8];
};


template <int VAR>
__device__ __forceinline__ float softmax_body(uint32_t tmem, uint32_t lane_addr, int half, int reps, float scale) {
    float m_run = 0.f, l_run = 0.f;
    uint32_t sr[2][32];
    for (int it = 0; it < reps; ++it) {
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_ld32(tmem + lane_addr + half * 64 + c * 32, sr[c]);
        tmem_wait_ld();
        float mt = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) mt = fmaxf(mt, __uint_as_float(sr[c][i]));
        mt *= scale;
        const float m_new = fmaxf(m_run, mt);
        const uint64_t scale2 = pack2(scale, scale), negm2 = pack2(-m_new, -m_new);
        uint64_t lsum2 = pack2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint64_t x = ffma2((uint64_t)sr[c][2 * i] | ((uint64_t)sr[c][2 * i + 1] << 32), scale2, negm2);
                uint64_t pp;
                if (VAR == 0) pp = pack2(ex2_approx(lo2(x)), ex2_approx(hi2(x)));
                else if (VAR == 2) pp = ((i & 1) && i != 15) ? ex2_poly2(x) : pack2(ex2_approx(lo2(x)), ex2_approx(hi2(x)));
                else pp = x;
                lsum2 = fadd2(lsum2, pp);
                pk[i] = pack_bf16(lo2(pp), hi2(pp));
            }
            if (VAR < 4) tmem_st16(tmem + lane_addr + 128 + half * 32 + c * 16, pk);
        }
        l_run = l_run + lo2(lsum2) + hi2(lsum2);
        m_run = m_new;
        if (VAR < 4) tmem_wait_st();
    }
    return l_run;
}

// mode bits: 1 = UMMA SS loop, 2 = UMMA TS loop, 4 = LDTM loop on warps 2-5, 8 = LDTM also on warps 6-9,
//            16 = STTM loop instead of LDTM, 32 = ld.shared stream on warps 6-9, 64 = UMMA N=256 (SS)
__global__ void __launch_bounds__(320, 1) probe_kernel(int mode, int reps, Res* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (SMEM_BYTES - 2048) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    long long t0 = 0, t1 = 0;

    if (warp == 1) {
        if (mode & (1 | 2 | 64)) {
            const uint32_t a = smem_u32(smem), b = smem_u32(smem + 65536);
            constexpr uint32_t ID128 = umma_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t ID256 = umma_idesc_bf16(128, 256, 0, 0);
            constexpr uint32_t IDTS = umma_idesc_bf16(128, 128, 0, 1);
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t off = (uint32_t)(kk >> 2) * 16384u + (uint32_t)(kk & 3) * 32u;
                        if (mode & 64)
                            umma_ss(tmem + 256, umma_desc_sw128(a + off, 16, 1024), umma_desc_sw128(b + off, 16, 1024), ID256, 1u);
                        else if (mode & 1)
                            umma_ss(tmem + 256, umma_desc_sw128(a + off, 16, 1024), umma_desc_sw128(b + off, 16, 1024), ID128, 1u);
                        if (mode & 2)
                            umma_ts(tmem + 384, tmem + kk * 8, umma_desc_sw128(b + kk * 2048, 16384, 1024), IDTS, 1u);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&done_bar);
            __syncwarp();
            mbar_wait(&done_bar, 0);
            t1 = clock64();
            if (lane == 0) { out[blockIdx.x].cyc[0] = t1 - t0; }
        }
    } else if (warp >= 2 && (mode & (128 | 256 | 512))) {
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const int half = (warp - 2) >> 2;
        t0 = clock64();
        float l = (mode & 128) ? softmax_body<2>(tmem, lane_addr, half, reps, 0.1275f)
                  : (mode & 256) ? softmax_body<3>(tmem, lane_addr, half, reps, 0.1275f)
                                 : softmax_body<0>(tmem, lane_addr, half, reps, 0.1275f);
        t1 = clock64();
        if (lane == 0 && warp == 2) out[blockIdx.x].cyc[1] = t1 - t0;
        if (lane == 0 && warp == 9) out[blockIdx.x].cyc[5] = t1 - t0;
        if (l == 1234.5f) out[blockIdx.x].cyc[7] = 1;
    } else if (warp >= 2 && warp < 6 && (mode & (4 | 16))) {
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = i;
        uint32_t acc = 0;
        t0 = clock64();
        for (int it = 0; it < reps; ++it) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (mode & 16) {
                    tmem_st32(tmem + lane_addr + c * 32, r);
                } else {
                    tmem_ld32(tmem + lane_addr + c * 32, r);
                }
            }
            if (mode & 16) tmem_wait_st(); else tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += r[i];
        }
        t1 = clock64();
        if (lane == 0) out[blockIdx.x].cyc[warp - 1] = t1 - t0;
        if (acc == 0x12345) out[blockIdx.x].cyc[7] = acc;
    } else if (warp >= 6 && (mode & 8)) {
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t r[32];
        uint32_t acc = 0;
        t0 = clock64();
        for (int it = 0; it < reps; ++it) {
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32(tmem + lane_addr + 128 + c * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += r[i];
        }
        t1 = clock64();
        if (lane == 0 && warp == 6) out[blockIdx.x].cyc[5] = t1 - t0;
        if (acc == 0x12345) out[blockIdx.x].cyc[7] = acc;
    } else if (warp >= 6 && (mode & 32)) {
        uint32_t acc = 0;
        const uint4* s4 = reinterpret_cast<const uint4*>(smem + 131072);
        t0 = clock64();
        for (int it = 0; it < reps * 8; ++it) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = s4[((it * 8 + c) * 128 + (threadIdx.x - 192)) & 2047];
                acc += v.x + v.y + v.z + v.w;
            }
        }
        t1 = clock64();
        if (lane == 0 && warp == 6) out[blockIdx.x].cyc[6] = t1 - t0;
        if (acc == 0x12345) out[blockIdx.x].cyc[7] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

static void run(const char* name, int mode, int reps, Res* d, int ctas) {
    cudaMemset(d, 0, sizeof(Res) * 148);
    probe_kernel<<<ctas, 320, SMEM_BYTES>>>(mode, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    Res h;
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    printf("%-46s ctas=%3d reps=%d", name, ctas, reps);
    if (h.cyc[0]) printf("  umma: %.1f cyc per 8-UMMA group (%.1f/UMMA)", (double)h.cyc[0] / reps, (double)h.cyc[0] / reps / ((mode & 3) == 3 ? 16 : 8));
    if (h.cyc[1]) printf("  tmem w2: %.1f cyc per 128-col row sweep (%.1f B/clk/CTA for 4 warps)", (double)h.cyc[1] / reps, 65536.0 * reps / h.cyc[1]);
    if (h.cyc[5]) printf("  tmem w6: %.1f cyc/sweep", (double)h.cyc[5] / reps);
    if (h.cyc[6]) printf("  lds w6: %.1f B/clk (4 warps)", 128.0 * 16 * 8 * 8 * reps / h.cyc[6]);
    printf("\n");
}


// ---- softmax body in isolation: 4 or 8 warps sweep a 128x128 fp32 tile in TMEM (LDTM -> max -> scale -> 2^x -> sum -> bf16 pack -> STTM),
// no barriers, no UMMA.  variant: 0 = MUFU only, 1 = polynomial only, 2 = 9/16 MUFU + 7/16 polynomial, 3 = no exponential,
// +8 = skip the TMEM traffic (registers only)
template <int NW, int VAR>
__global__ void __launch_bounds__(NW * 32, 1) softmax_probe(int reps, float scale, Res* out) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    constexpr int HALVES = NW / 4;            // 1: thread = whole row (128 cols); 2: thread = half row (64 cols)
    constexpr int COLS = 128 / HALVES;
    const int half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    float m_run = 0.f, l_run = 0.f;
    uint32_t sr[COLS / 32][32];
#pragma unroll
    for (int c = 0; c < COLS / 32; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) sr[c][i] = __float_as_uint(-(float)((lane + i + c) & 15));
    const long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
        if (!(VAR & 8)) {
#pragma unroll
            for (int c = 0; c < COLS / 32; ++c) tmem_ld32(tmem + lane_addr + half * COLS + c * 32, sr[c]);
            tmem_wait_ld();
        }
        float mt = -INFINITY;
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) mt = fmaxf(mt, __uint_as_float(sr[c][i]));
        mt *= scale;
        const float m_new = fmaxf(m_run, mt);
        const uint64_t scale2 = pack2(scale, scale), negm2 = pack2(-m_new, -m_new);
        uint64_t lsum2 = pack2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint64_t x = ffma2((uint64_t)sr[c][2 * i] | ((uint64_t)sr[c][2 * i + 1] << 32), scale2, negm2);
                uint64_t pp;
                if ((VAR & 7) == 0) pp = pack2(ex2_approx(lo2(x)), ex2_approx(hi2(x)));
                else if ((VAR & 7) == 1) pp = ex2_poly2(x);
                else if ((VAR & 7) == 2) pp = ((i & 1) && i != 15) ? ex2_poly2(x) : pack2(ex2_approx(lo2(x)), ex2_approx(hi2(x)));
                else pp = x;
                lsum2 = fadd2(lsum2, pp);
                pk[i] = pack_bf16(lo2(pp), hi2(pp));
            }
            if (!(VAR & 8)) tmem_st16(tmem + lane_addr + 256 + half * (COLS / 2) + c * 16, pk);
            else { sr[c][0] ^= pk[3] & 1; sr[c][7] ^= pk[9] & 1; }
        }
        l_run = l_run + lo2(lsum2) + hi2(lsum2);
        m_run = m_new;
        if (!(VAR & 8)) tmem_wait_st();
    }
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x].cyc[warp & 7] = t1 - t0;
    if (l_run == 1234.5f) out[blockIdx.x].cyc[7] = 1;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int NW, int VAR>
static void run_sm(const char* name, int reps, Res* d) {
    cudaMemset(d, 0, sizeof(Res) * 148);
    softmax_probe<NW, VAR><<<148, NW * 32>>>(reps, 0.1275f, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    Res h;
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 8; ++i) mx = h.cyc[i] > mx ? h.cyc[i] : mx;
    printf("softmax body %-40s warps=%d  %.1f cyc per 128x128 tile\n", name, NW, (double)mx / reps);
}

int main() {
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    Res* d;
    cudaMalloc(&d, sizeof(Res) * 148);
    const int reps = 200;
    for (int ctas : {1, 148}) {
        run("UMMA SS M128 N128 K16 x8", 1, reps, d, ctas);
        run("UMMA SS M128 N256 K16 x8", 64, reps, d, ctas);
        run("UMMA TS M128 N128 K16 x8", 2, reps, d, ctas);
        run("UMMA SS+TS interleaved x16", 3, reps, d, ctas);
        run("LDTM 4 warps", 4, reps, d, ctas);
        run("LDTM 8 warps", 4 | 8, reps, d, ctas);
        run("STTM 4 warps", 16, reps, d, ctas);
        run("UMMA SS + LDTM 4 warps", 1 | 4, reps, d, ctas);
        run("UMMA SS+TS + LDTM 8 warps", 3 | 4 | 8, reps, d, ctas);
        run("UMMA SS + ld.shared 4 warps", 1 | 32, reps, d, ctas);
        run("UMMA SS+TS + LDTM 4w + ld.shared 4w", 3 | 4 | 32, reps, d, ctas);
        run("ld.shared 4 warps alone", 32, reps, d, ctas);
        run("softmax mix 8w alone (cyc = per tile)", 128, reps, d, ctas);
        run("softmax mix 8w + UMMA SS+TS", 128 | 3, reps, d, ctas);
        run("softmax mix 8w + UMMA SS", 128 | 1, reps, d, ctas);
        run("softmax mix 8w + UMMA TS", 128 | 2, reps, d, ctas);
        run("softmax MUFU 8w + UMMA SS+TS", 512 | 3, reps, d, ctas);
        run("softmax noexp 8w + UMMA SS+TS", 256 | 3, reps, d, ctas);
    }
    run_sm<4, 0>("MUFU only", reps, d);
    run_sm<8, 0>("MUFU only", reps, d);
    run_sm<4, 1>("polynomial only", reps, d);
    run_sm<8, 1>("polynomial only", reps, d);
    run_sm<4, 2>("9/16 MUFU + 7/16 poly", reps, d);
    run_sm<8, 2>("9/16 MUFU + 7/16 poly", reps, d);
    run_sm<4, 3>("no exponential", reps, d);
    run_sm<8, 3>("no exponential", reps, d);
    run_sm<8, 8>("MUFU only, registers only", reps, d);
    run_sm<8, 9>("polynomial only, registers only", reps, d);
    run_sm<8, 10>("mix, registers only", reps, d);
    run_sm<8, 11>("no exponential, registers only", reps, d);
    return 0;
}
