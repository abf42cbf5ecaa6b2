// C-ABI plumbing: version, thread-local error text, device check, TMA descriptor creation and the
// host-buffer convenience entry point.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace svlm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
        cudaGetLastError();
        return SVLM_ENODEV;
    }
    static thread_local int cached_dev = -1, cached_major = 0;
    if (cached_dev != dev) {
        int major = 0;
        e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess) {
            set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
            return SVLM_ENODEV;
        }
        cached_dev = dev;
        cached_major = major;
    }
    if (cached_major != 10) {
        set_error("device %d has compute capability major %d; this library is built for sm_100a only", dev, cached_major);
        return SVLM_ENODEV;
    }
    return SVLM_OK;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_fn() {
    static encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    });
    return fn;
}

int make_tensor_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_128b) {
    encode_tiled_fn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable");
        return SVLM_ENODEV;
    }
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_128b ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu, box %u %u %u %u)", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return SVLM_ECUDA;
    }
    return SVLM_OK;
}

}  // namespace svlm

extern "C" {

int svlm_version(void) { return 1; }

const char* svlm_last_error(void) { return svlm::g_err; }

int svlm_kv_compact_rope_host(void* arena_host, int layers, int kv_heads, int head_dim, int capacity, int src_buffer,
                              const int32_t* runs_host, int n_runs, const int32_t* pos_host, int pos_stride,
                              const void* cos_table_host, const void* sin_table_host, int table_rows,
                              const int32_t* mrope_section_host) {
    using namespace svlm;
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(arena_host && pos_host && cos_table_host && sin_table_host, "NULL host buffer");
    SVLM_REQUIRE(src_buffer == 0 || src_buffer == 1, "src_buffer must be 0 or 1");
    ArenaDims ad{layers, kv_heads, head_dim, capacity};
    const size_t buf_bytes = (size_t)ad.buffer_elems() * 2;
    const size_t tab_bytes = (size_t)table_rows * (head_dim / 2) * 2;
    char* d_arena = nullptr;
    int32_t* d_pos = nullptr;
    char* d_tab = nullptr;
    cudaStream_t st = nullptr;
    SVLM_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    rc = SVLM_ECUDA;
    do {
        if (cudaMalloc(&d_arena, 2 * buf_bytes) != cudaSuccess) break;
        if (cudaMalloc(&d_pos, (size_t)3 * pos_stride * 4) != cudaSuccess) break;
        if (cudaMalloc(&d_tab, 2 * tab_bytes) != cudaSuccess) break;
        char* h = (char*)arena_host;
        if (cudaMemcpyAsync(d_arena + src_buffer * buf_bytes, h + src_buffer * buf_bytes, buf_bytes, cudaMemcpyHostToDevice, st)) break;
        if (cudaMemcpyAsync(d_pos, pos_host, (size_t)3 * pos_stride * 4, cudaMemcpyHostToDevice, st)) break;
        if (cudaMemcpyAsync(d_tab, cos_table_host, tab_bytes, cudaMemcpyHostToDevice, st)) break;
        if (cudaMemcpyAsync(d_tab + tab_bytes, sin_table_host, tab_bytes, cudaMemcpyHostToDevice, st)) break;
        svlm_arena_t a{d_arena, layers, kv_heads, head_dim, capacity};
        rc = svlm_kv_compact_rope(&a, src_buffer, runs_host, n_runs, d_pos, pos_stride, d_tab, d_tab + tab_bytes, table_rows,
                                  mrope_section_host, st);
        if (rc) break;
        rc = SVLM_ECUDA;
        const int dst = 1 - src_buffer;
        if (cudaMemcpyAsync(h + dst * buf_bytes, d_arena + dst * buf_bytes, buf_bytes, cudaMemcpyDeviceToHost, st)) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        rc = SVLM_OK;
    } while (0);
    if (rc == SVLM_ECUDA) set_error("svlm_kv_compact_rope_host: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_arena);
    cudaFree(d_pos);
    cudaFree(d_tab);
    cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------------------
// Host-side token-pattern range finder (no device work): the scan behind `get_qwen_range` (reference utils/get_qwen_range.py:15-86),
// which the eviction planner calls five times per chunk.  One pass over the ids: a segment starts where `pattern` matches, ends at the
// first `end_token` at or after the end of the pattern (a start without a terminator ends the scan), the next search begins behind it;
// a trailing `lf_token` is absorbed when `contain_lf`.  `index` counts from the left (>= 0) or from the right (< 0).
// ---------------------------------------------------------------------------------------------------------------------------------
extern "C" int svlm_find_segment(const int64_t* ids, int64_t n, const int64_t* pattern, int pattern_len, int64_t end_token, int index,
                                 int contain_lf, int64_t lf_token, int64_t* start_out, int64_t* end_out, int64_t* count_out) {
    SVLM_REQUIRE(ids != nullptr || n == 0, "ids is NULL");
    SVLM_REQUIRE(pattern != nullptr && pattern_len >= 1 && pattern_len <= 8, "pattern of 1..8 tokens expected");
    SVLM_REQUIRE(start_out && end_out, "NULL output pointer");
    const int m = pattern_len;
    // segments are found left to right; for index >= 0 the scan stops at that one, for index < 0 a ring of the last -index is kept
    const int keep = index < 0 ? -index : 1;
    SVLM_REQUIRE(keep <= 256, "index %d out of the supported range (-256 .. )", index);
    int64_t ring_s[256], ring_e[256];
    int64_t count = 0, i = 0;
    bool found = false;
    while (i + m <= n) {
        if (ids[i] != pattern[0]) { ++i; continue; }
        bool match = true;
        for (int j = 1; j < m; ++j)
            if (ids[i + j] != pattern[j]) { match = false; break; }
        if (!match) { ++i; continue; }
        int64_t e = i + m;
        while (e < n && ids[e] != end_token) ++e;
        if (e >= n) break;                                       // no terminator: the scan ends here (reference :80-81)
        int64_t stop = e;
        if (contain_lf && e + 1 < n && ids[e + 1] == lf_token) stop = e + 1;
        ring_s[count % keep] = i;
        ring_e[count % keep] = stop;
        ++count;
        if (index >= 0 && count == (int64_t)index + 1) { found = true; break; }
        i = e + 1;                                               // non-overlapping: the absorbed line feed may start nothing anyway
    }
    if (count_out) *count_out = count;
    if (index >= 0) {
        SVLM_REQUIRE(found, "segment index %d out of range (%lld segments)", index, (long long)count);
        *start_out = ring_s[(count - 1) % keep];
        *end_out = ring_e[(count - 1) % keep];
    } else {
        SVLM_REQUIRE(count >= keep, "segment index %d out of range (%lld segments)", index, (long long)count);
        *start_out = ring_s[(count - keep) % keep];
        *end_out = ring_e[(count - keep) % keep];
    }
    return SVLM_OK;
}
