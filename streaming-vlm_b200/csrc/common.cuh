// Shared device/host helpers for the sm_100a kernels: error plumbing, bf16 packing, and thin
// inline-PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor) and tcgen05 (UMMA / TMEM).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/svlm_b200.h"

namespace svlm {

// ------------------------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);
int  check_device();   // SVLM_OK if the current device is sm_100, else SVLM_ENODEV

#define SVLM_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            ::svlm::set_error(__VA_ARGS__);     \
            return SVLM_EINVAL;                 \
        }                                       \
    } while (0)

#define SVLM_CUDA(expr)                                                                     \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::svlm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SVLM_ECUDA;                                                              \
        }                                                                                   \
    } while (0)

inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return SVLM_ECUDA;
    }
    return SVLM_OK;
}

// ------------------------------------------------------------------------------------ programmatic dependent launch
// The decode step is a chain of ~170 short kernels per token.  Kernels launched through launch_pdl() may become resident
// while their predecessor in the stream is still draining: everything before pdl_wait() (barrier / TMEM set-up, row
// pointers, L2 prefetch of immutable weights) overlaps the predecessor's tail; pdl_wait() returns once the predecessor
// grid has completed and its writes are visible, so all reads of mutable data and all writes come after it.  A kernel
// calls pdl_trigger() when its own bulk work is done (after the weight stream / the key-tile loop): measured on B200,
// triggering at kernel START loses 7 ms per chunk (successor CTAs parked at the wait take SM slots from the
// weight-streaming GEMVs), triggering late gains 2.7-3 % (32.6 -> 33.4 frames/s).
// SVLM_PDL: 1 = every decode-chain kernel (default), 0 = off, 2 = only the GEMVs, 3 = only attention / merge / finalize
enum { PDL_GEMV = 1, PDL_OTHER = 2, PDL_ELEM = 4 };
// PDL_ELEM: the short kernels of the ViT tower and of the chunk prefill (add + RMSNorm, SwiGLU, ViT RoPE, append, embedding splice,
// prefill / ViT attention and its split merge).  They sit between library GEMMs; SVLM_PDL_ELEM=1 launches them with the attribute
// (every one of them starts with pdl_wait(), so the switch changes scheduling, never results).
inline bool pdl_elem_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SVLM_PDL_ELEM");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
inline bool pdl_enabled(int kind) {
    if (kind == PDL_ELEM) return pdl_elem_enabled();
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SVLM_PDL");
        v = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 1;
    }
    return v == 1 || (v == 2 && kind == PDL_GEMV) || (v == 3 && kind == PDL_OTHER);
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled(kind) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors are picked up by launch_status()
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Arena addressing (elements).  Layout: [2][3][L][H][cap][D].
struct ArenaDims {
    int L, H, D, cap;
    __host__ __device__ int64_t plane_elems() const { return (int64_t)L * H * cap * D; }
    __host__ __device__ int64_t buffer_elems() const { return 3 * plane_elems(); }
    __host__ __device__ int64_t offset(int buf, int plane, int layer, int head, int row) const {
        return buf * buffer_elems() + plane * plane_elems() + (((int64_t)layer * H + head) * cap + row) * D;
    }
};

inline int validate_arena(const svlm_arena_t* a) {
    SVLM_REQUIRE(a != nullptr && a->base != nullptr, "arena is NULL");
    SVLM_REQUIRE(a->layers > 0 && a->kv_heads > 0 && a->capacity > 0, "arena dims must be positive");
    SVLM_REQUIRE(a->head_dim == 128, "arena head_dim must be 128 (got %d)", a->head_dim);
    SVLM_REQUIRE(((uintptr_t)a->base & 1023) == 0, "arena base must be 1024-byte aligned");
    return SVLM_OK;
}

// TMA descriptor creation through the driver entry point (no link-time libcuda dependency).
// dims/strides innermost first; strides in BYTES for dims 1..rank-1; box in elements.
int make_tensor_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_128b);

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------ bf16
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x on the FMA/ALU pipes (no MUFU): Cody-Waite split x = n + r, |r| <= 0.5, degree-3 minimax polynomial for 2^r
// (relative error 7.5e-5, far below bf16's 2^-9), exponent patched in with an integer shift-add.  Valid for x <= ~100;
// x < -125 (incl. -inf) returns ~2^-125.  B200's MUFU pipe and its FMA pipe each bound a softmax that uses only one of
// them (measured: ~16 and ~2 cycles per warp instruction per SM sub-partition); evaluating a fixed share of the
// exponentials here lets the two pipes run side by side (the FlashAttention-4 idea).
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;                       // 1.5 * 2^23: low mantissa bits = round(x)
    const float r = x - (t - 12582912.0f);
    float p = fmaf(5.517165e-2f, r, 2.426111e-1f);
    p = fmaf(p, r, 6.932610e-1f);
    p = fmaf(p, r, 9.999281e-1f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 issue one instruction for two fp32 lanes of a 64-bit register pair)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float lo2(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi2(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// ex2_poly on two lanes at once: 6 FMA-pipe instructions per pair (2 FADD2 + 4 FFMA2) + 2 clamps and 2 shift-adds on the ALU pipe
__device__ __forceinline__ uint64_t ex2_poly2(uint64_t x) {
    x = pack2(fmaxf(lo2(x), -125.0f), fmaxf(hi2(x), -125.0f));
    const uint64_t t = fadd2(x, pack2(12582912.0f, 12582912.0f));
    const uint64_t u = fadd2(t, pack2(-12582912.0f, -12582912.0f));
    const uint64_t r = ffma2(u, pack2(-1.0f, -1.0f), x);
    uint64_t q = ffma2(pack2(5.517165e-2f, 5.517165e-2f), r, pack2(2.426111e-1f, 2.426111e-1f));
    q = ffma2(q, r, pack2(6.932610e-1f, 6.932610e-1f));
    q = ffma2(q, r, pack2(9.999281e-1f, 9.999281e-1f));
    const uint32_t a = (uint32_t)q + ((uint32_t)t << 23);
    const uint32_t b = (uint32_t)(q >> 32) + ((uint32_t)(t >> 32) << 23);
    return (uint64_t)a | ((uint64_t)b << 32);
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// HF RMSNorm statistics with a FIXED operation order (explicit roundings, no contraction), shared by add_rmsnorm_kernel and the
// fused RMSNorm prologue of gemm_tokens_kernel so that both produce bit-identical rstd values.
__device__ __forceinline__ float sumsq8(float ss, const uint4& v) {
    const uint32_t* p = &v.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float x0 = bf16_lo(p[k]), x1 = bf16_hi(p[k]);
        ss = __fadd_rn(ss, __fmaf_rn(x1, x1, __fmul_rn(x0, x0)));
    }
    return ss;
}
__device__ __forceinline__ float rms_rstd(float w0, float w1, float w2, float w3, int cols, float eps) {
    return rsqrtf(__fadd_rn(__fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(w0, w1), w2), w3), (float)cols), eps));
}

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_na_v4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------ smem / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Waits for the phase with the given parity to complete.  A watchdog turns a protocol bug into
// a trap (reported as a launch failure) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("svlm: mbarrier wait timed out (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
                   threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}

// ------------------------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem)),
                 "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(smem)),
                 "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ------------------------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell).
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1
//   bits [61,64) layout: 0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// UMMA instruction descriptor for kind::f16, bf16 x bf16 -> fp32 (cute::UMMA::InstrDescriptor).
//   [4,6) c_format=1 (F32)  [7,10) a_format=1 (BF16)  [10,13) b_format=1 (BF16)
//   [15] a_major (0 K, 1 MN)  [16] b_major  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All previously issued MMAs of this thread arrive on `bar` when they complete
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM <-> registers, shape 32x32b: lane i of the warp touches TMEM lane (base_lane + i), N consecutive columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------ CTA pairs (cta_group::2)
// Two CTAs of a cluster (ranks 2i, 2i + 1, same TPC) execute ONE tcgen05.mma of M = 256: each CTA supplies its 128 rows of A and
// HALF of the N rows of B from its own shared memory and receives its 128 accumulator rows in its own TMEM.  The leader (even
// rank) issues the instruction and owns the "operands have landed" barrier; both CTAs load with TMA and signal that barrier.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` (a shared::cta address) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: data into THIS CTA's shared memory, completion bytes on the mbarrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem)),
                 "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
                 : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst) {   // one whole warp of EACH CTA of the pair, same smem offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both] * B[smem of both], M = 256; issued by one thread of the leader CTA
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued pair MMAs arrive (once) on the mbarrier at this shared-memory offset in EVERY CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
#endif  // __CUDACC__

}  // namespace svlm
