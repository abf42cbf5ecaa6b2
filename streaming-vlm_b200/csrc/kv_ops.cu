// HBM-bound integer / byte kernels of the hot path (sm_100a):
//   kv_compact_rope   single-pass gather of the retained token runs into the other arena buffer,
//                     fused with the contiguous-RoPE re-rotation of K          (abi: svlm_kv_compact_rope)
//   mrope_positions   contiguous 3-D mRoPE position ids from the pruned id sequence  (svlm_mrope_positions)
//   qkv_rope_append   rotate q, append K_raw / K_rot / V rows of the new tokens       (svlm_qkv_rope_append)
//   vit_rope          2-D RoPE of the ViT q,k in place (fp32 arithmetic, bf16 storage) (svlm_vit_rope)
//   rmsnorm           HF RMSNorm (fp32 statistics, bf16 weight multiply)               (svlm_rmsnorm)
// All rows are 256 B (D = 128 bf16): a half-warp moves one row with one 128-bit access per lane.
#include "common.cuh"

namespace svlm {

struct RunTable {
    int n;
    int src[SVLM_MAX_RUNS];
    int dst[SVLM_MAX_RUNS];
    int len[SVLM_MAX_RUNS];
};

struct RopeTables {
    const uint4* cos;   // bf16 [rows][64] viewed as uint4 [rows][8]
    const uint4* sin;
    int rows;
    int b0, b1;         // chunk boundaries (in 8-element chunks) between the t | h | w sections
};

// mRoPE of one 8-element chunk.  `mine` = 8 bf16 of this lane's chunk, `partner` = the chunk 64
// elements away (other half of the head), `first_half` says which half this lane holds.
// out = bf16( bf16(x*cos) + bf16(rot(x)*sin) ),  rot(x) = (-x[64:], x[:64])      (language_forward.py:58-66)
__device__ __forceinline__ uint4 rope_chunk(const uint4& mine, const uint4& partner, const uint4& c, const uint4& s, bool first_half) {
    const uint32_t* m = &mine.x;
    const uint32_t* p = &partner.x;
    const uint32_t* cc = &c.x;
    const uint32_t* ss = &s.x;
    uint4 out;
    uint32_t* o = &out.x;
    const float sgn = first_half ? -1.0f : 1.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a0 = round_bf16(__fmul_rn(bf16_lo(m[i]), bf16_lo(cc[i])));
        float a1 = round_bf16(__fmul_rn(bf16_hi(m[i]), bf16_hi(cc[i])));
        float b0 = round_bf16(__fmul_rn(sgn * bf16_lo(p[i]), bf16_lo(ss[i])));
        float b1 = round_bf16(__fmul_rn(sgn * bf16_hi(p[i]), bf16_hi(ss[i])));
        o[i] = pack_bf16(__fadd_rn(a0, b0), __fadd_rn(a1, b1));
    }
    return out;
}

__device__ __forceinline__ uint4 shfl_xor_u4(const uint4& v, int mask) {
    uint4 r;
    r.x = __shfl_xor_sync(0xffffffffu, v.x, mask);
    r.y = __shfl_xor_sync(0xffffffffu, v.y, mask);
    r.z = __shfl_xor_sync(0xffffffffu, v.z, mask);
    r.w = __shfl_xor_sync(0xffffffffu, v.w, mask);
    return r;
}

__device__ __forceinline__ int axis_of_chunk(int jj, int b0, int b1) { return jj < b0 ? 0 : (jj < b1 ? 1 : 2); }

// ---------------------------------------------------------------------------------------------------
// kv_compact_rope.  grid = (ceil(n_dst / ROWS), L*H), block = 256 (16 half-warps, ROWS/16 rows each).
// Per retained token and (layer, head): read K_raw + V (512 B), write K_raw + K_rot + V (768 B).
// ---------------------------------------------------------------------------------------------------
constexpr int CR_ROWS = 64;
constexpr int CR_PER_HW = CR_ROWS / 16;

__global__ void __launch_bounds__(256) kv_compact_rope_kernel(uint16_t* __restrict__ arena, ArenaDims ad, int src_buf,
                                                               const __grid_constant__ RunTable runs, int n_dst,
                                                               const int32_t* __restrict__ pos, int pos_stride, RopeTables tab) {
    const int lane16 = threadIdx.x & 15;
    const int hw = threadIdx.x >> 4;
    const int plane_idx = blockIdx.y;                      // layer * H + head
    const int row0 = blockIdx.x * CR_ROWS;
    const int jj = lane16 & 7;
    const bool first_half = lane16 < 8;
    const int axis = axis_of_chunk(jj, tab.b0, tab.b1);

    const int64_t plane = ad.plane_elems();
    const int64_t head_base = (int64_t)plane_idx * ad.cap * ad.D;
    const uint16_t* src_b = arena + (int64_t)src_buf * ad.buffer_elems() + head_base;
    uint16_t* dst_b = arena + (int64_t)(1 - src_buf) * ad.buffer_elems() + head_base;

    int src_row[CR_PER_HW], dst_row[CR_PER_HW];
    uint4 kraw[CR_PER_HW], vv[CR_PER_HW];
#pragma unroll
    for (int i = 0; i < CR_PER_HW; ++i) {
        const int d = row0 + hw + 16 * i;
        dst_row[i] = d;
        int s = -1;
        if (d < n_dst) {
            for (int r = 0; r < runs.n; ++r) {
                const int off = d - runs.dst[r];
                if (off >= 0 && off < runs.len[r]) { s = runs.src[r] + off; break; }
            }
        }
        src_row[i] = s;
    }
#pragma unroll
    for (int i = 0; i < CR_PER_HW; ++i) {
        if (src_row[i] >= 0) {
            const int64_t so = (int64_t)src_row[i] * ad.D + lane16 * 8;
            kraw[i] = ld_nc_v4(src_b + SVLM_PLANE_K_RAW * plane + so);
            vv[i] = ld_nc_v4(src_b + SVLM_PLANE_V * plane + so);
        } else {
            kraw[i] = make_uint4(0, 0, 0, 0);
            vv[i] = make_uint4(0, 0, 0, 0);
        }
    }
#pragma unroll
    for (int i = 0; i < CR_PER_HW; ++i) {
        // the shuffle is executed by every lane of the warp (both half-warps, valid or not)
        const uint4 partner = shfl_xor_u4(kraw[i], 8);
        if (src_row[i] >= 0) {
            int p = __ldg(pos + (int64_t)axis * pos_stride + dst_row[i]);
            p = min(max(p, 0), tab.rows - 1);
            const uint4 c = __ldg(tab.cos + (int64_t)p * 8 + jj);
            const uint4 s = __ldg(tab.sin + (int64_t)p * 8 + jj);
            const uint4 krot = rope_chunk(kraw[i], partner, c, s, first_half);
            const int64_t dof = (int64_t)dst_row[i] * ad.D + lane16 * 8;
            st_na_v4(dst_b + SVLM_PLANE_K_RAW * plane + dof, kraw[i]);
            st_na_v4(dst_b + SVLM_PLANE_K_ROT * plane + dof, krot);
            st_na_v4(dst_b + SVLM_PLANE_V * plane + dof, vv[i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// mrope_positions: one CTA of 1024 threads, three block scans over the id sequence.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_scan_incl_add(int v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ int warp_scan_incl_max(int v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v = max(v, t);
    }
    return v;
}
// exclusive block scans over 1024 threads; `total` receives the block-wide result
template <bool IS_MAX>
__device__ int block_scan_excl(int v, int identity, int* smem32, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = IS_MAX ? warp_scan_incl_max(v) : warp_scan_incl_add(v);
    if (lane == 31) smem32[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = smem32[lane];
        int wi = IS_MAX ? warp_scan_incl_max(w) : warp_scan_incl_add(w);
        smem32[lane] = wi;
    }
    __syncthreads();
    int prev_warps = warp == 0 ? identity : smem32[warp - 1];
    int prev_lane = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) prev_lane = identity;
    int excl = IS_MAX ? max(prev_warps, prev_lane) : prev_warps + prev_lane;
    if (total) *total = smem32[31];
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(1024) mrope_positions_kernel(const int64_t* __restrict__ ids, int n, const int32_t* __restrict__ grids,
                                                                int n_grids, int merge, int t_step, int64_t vid, int32_t* __restrict__ pos,
                                                                int pos_stride, int32_t* __restrict__ next_pos, int32_t* __restrict__ status) {
    __shared__ int sm[32];
    const int ipt = (n + 1023) / 1024;
    const int lo = min(n, (int)threadIdx.x * ipt), hi = min(n, lo + ipt);

    // phase A: run starts (count and last start index) in this thread's slice
    int n_start = 0, last_start = -1;
    for (int i = lo; i < hi; ++i) {
        if (ids[i] == vid && (i == 0 || ids[i - 1] != vid)) { ++n_start; last_start = i; }
    }
    int ord0 = block_scan_excl<false>(n_start, 0, sm, nullptr);          // runs started before this slice
    int start0 = block_scan_excl<true>(last_start, -1, sm, nullptr);     // start of the run open at slice entry

    // phase B: position advance contributed by this slice
    int bad = 0;
    int ord = ord0 - 1, cur_start = start0, adv = 0;
    for (int i = lo; i < hi; ++i) {
        if (ids[i] != vid) { ++adv; continue; }
        if (i == 0 || ids[i - 1] != vid) { ++ord; cur_start = i; }
        if (i + 1 == n || ids[i + 1] != vid) {
            const int g = min(max(ord, 0), n_grids - 1);
            if (ord >= n_grids || n_grids <= 0) { bad |= 2; }
            const int gt = n_grids > 0 ? grids[g * 3] : 1, gh = n_grids > 0 ? grids[g * 3 + 1] / merge : 1,
                      gw = n_grids > 0 ? grids[g * 3 + 2] / merge : 1;
            if (i - cur_start + 1 != gt * gh * gw) bad |= 1;
            adv += max(max((gt - 1) * t_step, gh - 1), gw - 1) + 1;
        }
    }
    int total = 0;
    int base = block_scan_excl<false>(adv, 0, sm, &total);

    // phase C: emit
    ord = ord0 - 1; cur_start = start0;
    int p = base;
    for (int i = lo; i < hi; ++i) {
        if (ids[i] != vid) {
            pos[i] = p; pos[pos_stride + i] = p; pos[2 * pos_stride + i] = p;
            ++p;
            continue;
        }
        if (i == 0 || ids[i - 1] != vid) { ++ord; cur_start = i; }
        const int g = min(max(ord, 0), max(n_grids - 1, 0));
        const int gt = n_grids > 0 ? grids[g * 3] : 1, gh = n_grids > 0 ? max(grids[g * 3 + 1] / merge, 1) : 1,
                  gw = n_grids > 0 ? max(grids[g * 3 + 2] / merge, 1) : 1;
        const int o = i - cur_start, hw = gh * gw;
        const int tt = o / hw, rr = o % hw;
        pos[i] = p + tt * t_step;
        pos[pos_stride + i] = p + rr / gw;
        pos[2 * pos_stride + i] = p + rr % gw;
        if (i + 1 == n || ids[i + 1] != vid) p += max(max((gt - 1) * t_step, gh - 1), gw - 1) + 1;
    }
    if (bad) atomicOr(status, bad);
    if (threadIdx.x == 0) { *next_pos = total; if (!bad) atomicOr(status, 0); }
}

// ---------------------------------------------------------------------------------------------------
// qkv_rope_append.  grid = n_tokens, block = 256: half-warp per head row, rows = Hq (q) + H (k) + H (v).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qkv_rope_append_kernel(uint16_t* __restrict__ arena, ArenaDims ad, int buffer, int layer,
                                                               const uint16_t* __restrict__ q, const uint16_t* __restrict__ k,
                                                               const uint16_t* __restrict__ v, int64_t qs, int64_t ks, int64_t vs, int hq,
                                                               uint16_t* __restrict__ q_out, int start, const int32_t* __restrict__ start_dev,
                                                               const int32_t* __restrict__ pos, int pos_stride,
                                                               const int32_t* __restrict__ text_pos_dev, RopeTables tab) {
    pdl_wait();          // inputs come from the predecessor in the stream (no-op without the launch attribute)
    const int t = blockIdx.x;
    const int lane16 = threadIdx.x & 15;
    const int hw = threadIdx.x >> 4;
    const int jj = lane16 & 7;
    const bool first_half = lane16 < 8;
    const int axis = axis_of_chunk(jj, tab.b0, tab.b1);
    const int row = (start_dev ? *start_dev : start) + t;
    if (row >= ad.cap) return;                     // capacity is validated on the host; never write out of bounds
    int p = text_pos_dev ? (*text_pos_dev + t) : __ldg(pos + (int64_t)axis * pos_stride + row);
    p = min(max(p, 0), tab.rows - 1);
    const uint4 c = __ldg(tab.cos + (int64_t)p * 8 + jj);
    const uint4 s = __ldg(tab.sin + (int64_t)p * 8 + jj);
    const int64_t plane = ad.plane_elems();
    uint16_t* buf = arena + (int64_t)buffer * ad.buffer_elems();
    const int n_rows = hq + 2 * ad.H;
    const int iters = (n_rows + 15) / 16;
    for (int it = 0; it < iters; ++it) {
        const int r = it * 16 + hw;
        const bool valid = r < n_rows;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (valid) {
            const uint16_t* srcp = r < hq ? q + t * qs + (int64_t)r * ad.D
                                 : (r < hq + ad.H ? k + t * ks + (int64_t)(r - hq) * ad.D : v + t * vs + (int64_t)(r - hq - ad.H) * ad.D);
            x = *reinterpret_cast<const uint4*>(srcp + lane16 * 8);
        }
        const uint4 partner = shfl_xor_u4(x, 8);
        if (!valid) continue;
        if (r < hq) {
            const uint4 y = rope_chunk(x, partner, c, s, first_half);
            *reinterpret_cast<uint4*>(q_out + ((int64_t)t * hq + r) * ad.D + lane16 * 8) = y;
        } else if (r < hq + ad.H) {
            const int h = r - hq;
            const int64_t o = (((int64_t)layer * ad.H + h) * ad.cap + row) * ad.D + lane16 * 8;
            const uint4 y = rope_chunk(x, partner, c, s, first_half);
            *reinterpret_cast<uint4*>(buf + SVLM_PLANE_K_RAW * plane + o) = x;
            *reinterpret_cast<uint4*>(buf + SVLM_PLANE_K_ROT * plane + o) = y;
        } else {
            const int h = r - hq - ad.H;
            const int64_t o = (((int64_t)layer * ad.H + h) * ad.cap + row) * ad.D + lane16 * 8;
            *reinterpret_cast<uint4*>(buf + SVLM_PLANE_V * plane + o) = x;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// qkv_rope_append for the decode step of several streams: CTA b appends token b to the arena of stream b at the row and
// text position its device-side counters name.  Same arithmetic as qkv_rope_append_kernel.
// ---------------------------------------------------------------------------------------------------
struct AppendSlots {
    uint16_t* buf[SVLM_MAX_STREAMS];            // live buffer of each stream's arena
    const int32_t* row[SVLM_MAX_STREAMS];
    const int32_t* pos[SVLM_MAX_STREAMS];
};

__global__ void __launch_bounds__(256) qkv_rope_append_streams_kernel(const AppendSlots sl, ArenaDims ad, int layer,
                                                                       const uint16_t* __restrict__ q, const uint16_t* __restrict__ k,
                                                                       const uint16_t* __restrict__ v, int64_t qs, int64_t ks, int64_t vs,
                                                                       int hq, uint16_t* __restrict__ q_out, RopeTables tab) {
    const int t = blockIdx.x;
    const int lane16 = threadIdx.x & 15;
    const int hw = threadIdx.x >> 4;
    const int jj = lane16 & 7;
    const bool first_half = lane16 < 8;
    const int row = __ldg(sl.row[t]);
    if (row < 0 || row >= ad.cap) return;
    const int p = min(max(__ldg(sl.pos[t]), 0), tab.rows - 1);
    const uint4 c = __ldg(tab.cos + (int64_t)p * 8 + jj);
    const uint4 s = __ldg(tab.sin + (int64_t)p * 8 + jj);
    const int64_t plane = ad.plane_elems();
    uint16_t* buf = sl.buf[t];
    const int n_rows = hq + 2 * ad.H;
    const int iters = (n_rows + 15) / 16;
    for (int it = 0; it < iters; ++it) {
        const int r = it * 16 + hw;
        const bool valid = r < n_rows;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (valid) {
            const uint16_t* srcp = r < hq ? q + t * qs + (int64_t)r * ad.D
                                 : (r < hq + ad.H ? k + t * ks + (int64_t)(r - hq) * ad.D : v + t * vs + (int64_t)(r - hq - ad.H) * ad.D);
            x = *reinterpret_cast<const uint4*>(srcp + lane16 * 8);
        }
        const uint4 partner = shfl_xor_u4(x, 8);
        if (!valid) continue;
        if (r < hq) {
            *reinterpret_cast<uint4*>(q_out + ((int64_t)t * hq + r) * ad.D + lane16 * 8) = rope_chunk(x, partner, c, s, first_half);
        } else if (r < hq + ad.H) {
            const int64_t o = (((int64_t)layer * ad.H + (r - hq)) * ad.cap + row) * ad.D + lane16 * 8;
            *reinterpret_cast<uint4*>(buf + SVLM_PLANE_K_RAW * plane + o) = x;
            *reinterpret_cast<uint4*>(buf + SVLM_PLANE_K_ROT * plane + o) = rope_chunk(x, partner, c, s, first_half);
        } else {
            const int64_t o = (((int64_t)layer * ad.H + (r - hq - ad.H)) * ad.cap + row) * ad.D + lane16 * 8;
            *reinterpret_cast<uint4*>(buf + SVLM_PLANE_V * plane + o) = x;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Frame ingest: uint8 RGB frames [T][H][W][3] (the decoder's layout) -> ViT patch rows [gt*gh*gw][3*Tp*P*P] in the
// merge-unit-major order of the HF Qwen2-VL video processor, normalised with its fused arithmetic
// (x - mean') / std' in fp32 (mean' = mean * 255, std' = std * 255; IEEE sub and div, no FMA / reciprocal).
// One thread = two adjacent output elements (pw, pw + 1): one 4-byte (bf16) or 8-byte (fp32) store.
// ---------------------------------------------------------------------------------------------------
struct IngestParams {
    int T, H, W, P, TP, M;
    float mean[3], stdv[3];
    int out_bf16;
};

__global__ void __launch_bounds__(256) frames_to_patches_kernel(const uint8_t* __restrict__ frames, void* __restrict__ out, const IngestParams p) {
    const int gh = p.H / p.P, gw = p.W / p.P, gt = p.T / p.TP;
    const int cols = 3 * p.TP * p.P * p.P;
    const int64_t total2 = (int64_t)gt * gh * gw * cols / 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total2) return;
    const int col = (int)((idx * 2) % cols);
    int64_t n = (idx * 2) / cols;
    const int pw = col % p.P, ph = (col / p.P) % p.P, ti = (col / (p.P * p.P)) % p.TP, c = col / (p.P * p.P * p.TP);
    const int mw = (int)(n % p.M); n /= p.M;
    const int mh = (int)(n % p.M); n /= p.M;
    const int bw = (int)(n % (gw / p.M)); n /= (gw / p.M);
    const int bh = (int)(n % (gh / p.M)); n /= (gh / p.M);
    const int t = (int)n * p.TP + ti;
    const int y = (bh * p.M + mh) * p.P + ph, x = (bw * p.M + mw) * p.P + pw;
    const uint8_t* src = frames + (((int64_t)t * p.H + y) * p.W + x) * 3 + c;
    const float v0 = __fdiv_rn(__fsub_rn((float)src[0], p.mean[c]), p.stdv[c]);
    const float v1 = __fdiv_rn(__fsub_rn((float)src[3], p.mean[c]), p.stdv[c]);       // P is even: pw + 1 is in the same patch row
    if (p.out_bf16) reinterpret_cast<uint32_t*>(out)[idx] = pack_bf16(v0, v1);
    else reinterpret_cast<float2*>(out)[idx] = make_float2(v0, v1);
}

// ---------------------------------------------------------------------------------------------------
// Anti-aliased bicubic resize of a clip (uint8 [T][H][W][3] -> uint8 [T][H2][W2][3]), separable: a horizontal pass into
// fp32, then a vertical pass that rounds half-to-even and clamps.  Tap tables (first input index, tap count, normalised fp32
// weights per output index) are built on the host exactly as ATen's `_upsample_bicubic2d_aa` builds them; accumulation is
// sequential fp32 with separate multiply and add roundings (no FMA), so the result is a pure function of the tables.
// ---------------------------------------------------------------------------------------------------
struct ResizeTaps {
    const int32_t* start;
    const int32_t* count;
    const float* weights;     // [n_out][k]
    int k;
};

__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ in, float* __restrict__ tmp, int rows, int W, int W2,
                                                        const ResizeTaps tx) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // (row, x2); row = t * H + y
    if (idx >= (int64_t)rows * W2) return;
    const int x2 = (int)(idx % W2);
    const int64_t row = idx / W2;
    const int s0 = __ldg(tx.start + x2), n = __ldg(tx.count + x2);
    const float* w = tx.weights + (int64_t)x2 * tx.k;
    const uint8_t* src = in + (row * W + s0) * 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int j = 0; j < n; ++j) {
        const float wj = __ldg(w + j);
        a0 = __fadd_rn(a0, __fmul_rn(wj, (float)src[3 * j + 0]));
        a1 = __fadd_rn(a1, __fmul_rn(wj, (float)src[3 * j + 1]));
        a2 = __fadd_rn(a2, __fmul_rn(wj, (float)src[3 * j + 2]));
    }
    float* dst = tmp + idx * 3;
    dst[0] = a0; dst[1] = a1; dst[2] = a2;
}

__global__ void __launch_bounds__(256) resize_v_kernel(const float* __restrict__ tmp, uint8_t* __restrict__ out, int T, int H, int H2, int W2,
                                                        const ResizeTaps ty) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // (t, y2, x2)
    if (idx >= (int64_t)T * H2 * W2) return;
    const int x2 = (int)(idx % W2);
    const int y2 = (int)((idx / W2) % H2);
    const int t = (int)(idx / ((int64_t)W2 * H2));
    const int s0 = __ldg(ty.start + y2), n = __ldg(ty.count + y2);
    const float* w = ty.weights + (int64_t)y2 * ty.k;
    const float* src = tmp + (((int64_t)t * H + s0) * W2 + x2) * 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int j = 0; j < n; ++j) {
        const float wj = __ldg(w + j);
        const float* r = src + (int64_t)j * W2 * 3;
        a0 = __fadd_rn(a0, __fmul_rn(wj, r[0]));
        a1 = __fadd_rn(a1, __fmul_rn(wj, r[1]));
        a2 = __fadd_rn(a2, __fmul_rn(wj, r[2]));
    }
    uint8_t* dst = out + idx * 3;
    dst[0] = (uint8_t)fminf(fmaxf(rintf(a0), 0.f), 255.f);
    dst[1] = (uint8_t)fminf(fmaxf(rintf(a1), 0.f), 255.f);
    dst[2] = (uint8_t)fminf(fmaxf(rintf(a2), 0.f), 255.f);
}

// ---------------------------------------------------------------------------------------------------
// vit_rope: qkv [N][3][H][D] bf16, q and k rotated in place.  fp32 products and sum, no FMA contraction
// (apply_rotary_pos_emb_flashatt semantics: q.float()*cos + rotate_half(q.float())*sin, cast back).
// One thread = one 8-element chunk of the first half and its partner chunk in the second half.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vit_rope_kernel(uint16_t* __restrict__ qkv, const float* __restrict__ cosv,
                                                        const float* __restrict__ sinv, int n_tokens, int heads, int D) {
    pdl_wait();          // inputs come from the predecessor in the stream (no-op without the launch attribute)
    const int half = D / 2, chunks = half / 8;
    const int64_t total = (int64_t)n_tokens * 2 * heads * chunks;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % chunks);
    const int h = (int)((idx / chunks) % heads);
    const int which = (int)((idx / ((int64_t)chunks * heads)) % 2);
    const int n = (int)(idx / ((int64_t)chunks * heads * 2));
    uint16_t* row = qkv + (((int64_t)n * 3 + which) * heads + h) * D;
    uint4 lo = *reinterpret_cast<const uint4*>(row + c * 8);
    uint4 hi = *reinterpret_cast<const uint4*>(row + half + c * 8);
    const float* cl = cosv + (int64_t)n * D + c * 8;
    const float* ch = cl + half;
    const float* sl = sinv + (int64_t)n * D + c * 8;
    const float* sh = sl + half;
    uint32_t* pl = &lo.x;
    uint32_t* ph = &hi.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float x0 = bf16_lo(pl[i]), x1 = bf16_hi(pl[i]), y0 = bf16_lo(ph[i]), y1 = bf16_hi(ph[i]);
        // first half:  x*cos + (-y)*sin ; second half: y*cos + x*sin
        const float a0 = __fadd_rn(__fmul_rn(x0, cl[2 * i]), __fmul_rn(-y0, sl[2 * i]));
        const float a1 = __fadd_rn(__fmul_rn(x1, cl[2 * i + 1]), __fmul_rn(-y1, sl[2 * i + 1]));
        const float b0 = __fadd_rn(__fmul_rn(y0, ch[2 * i]), __fmul_rn(x0, sh[2 * i]));
        const float b1 = __fadd_rn(__fmul_rn(y1, ch[2 * i + 1]), __fmul_rn(x1, sh[2 * i + 1]));
        pl[i] = pack_bf16(a0, a1);
        ph[i] = pack_bf16(b0, b1);
    }
    *reinterpret_cast<uint4*>(row + c * 8) = lo;
    *reinterpret_cast<uint4*>(row + half + c * 8) = hi;
}

// ---------------------------------------------------------------------------------------------------
// add_rmsnorm: one CTA (128 threads) per row, the row stays in registers (single pass over global memory).
//   s = a (+ b)                          optional residual add, rounded to bf16 (the reference's `residual + x`)
//   out_norm = bf16( w * bf16( s * rsqrt(mean(s^2) + eps) ) )          (HF Qwen2_5_VLRMSNorm)
// ---------------------------------------------------------------------------------------------------
constexpr int RN_MAX_VEC = 8;     // cols <= 128 threads * 8 vectors * 8 elements = 8192
__global__ void __launch_bounds__(128) add_rmsnorm_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b,
                                                           const uint16_t* __restrict__ w, uint16_t* __restrict__ out_sum,
                                                           uint16_t* __restrict__ out_norm, int cols, float eps) {
    pdl_wait();          // inputs come from the predecessor in the stream (no-op without the launch attribute)
    __shared__ float red[4];
    const int64_t base = (int64_t)blockIdx.x * cols;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 v[RN_MAX_VEC];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < RN_MAX_VEC; ++i) {
        const int c = (i * 128 + threadIdx.x) * 8;
        if (c < cols) {
            v[i] = *reinterpret_cast<const uint4*>(a + base + c);
            if (b) {
                const uint4 y = *reinterpret_cast<const uint4*>(b + base + c);
                uint32_t* p = &v[i].x;
                const uint32_t* q = &y.x;
#pragma unroll
                for (int k = 0; k < 4; ++k) p[k] = pack_bf16(bf16_lo(p[k]) + bf16_lo(q[k]), bf16_hi(p[k]) + bf16_hi(q[k]));
                if (out_sum) *reinterpret_cast<uint4*>(out_sum + base + c) = v[i];
            }
            ss = sumsq8(ss, v[i]);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    const float r = rms_rstd(red[0], red[1], red[2], red[3], cols, eps);
#pragma unroll
    for (int i = 0; i < RN_MAX_VEC; ++i) {
        const int c = (i * 128 + threadIdx.x) * 8;
        if (c < cols) {
            const uint4 wv = *reinterpret_cast<const uint4*>(w + c);
            const uint32_t* p = &v[i].x;
            const uint32_t* q = &wv.x;
            uint4 o;
            uint32_t* po = &o.x;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                po[k] = pack_bf16(__fmul_rn(bf16_lo(q[k]), round_bf16(__fmul_rn(bf16_lo(p[k]), r))),
                                  __fmul_rn(bf16_hi(q[k]), round_bf16(__fmul_rn(bf16_hi(p[k]), r))));
            *reinterpret_cast<uint4*>(out_norm + base + c) = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// swiglu: gu [rows][2*M] (gate columns then up columns, the fused gate/up GEMM output) -> out [rows][M],
//   out = bf16( bf16(silu(gate)) * up )      (Qwen2MLP / Qwen2_5_VLMLP act_fn(gate_proj(x)) * up_proj(x))
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) swiglu_kernel(const uint16_t* __restrict__ gu, uint16_t* __restrict__ out, int rows, int M,
                                                      int64_t gu_stride, int64_t out_stride) {
    pdl_wait();          // inputs come from the predecessor in the stream (no-op without the launch attribute)
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int vec_per_row = M / 8;
    if (idx >= (int64_t)rows * vec_per_row) return;
    const int r = (int)(idx / vec_per_row), c = (int)(idx % vec_per_row) * 8;
    const uint4 gv = *reinterpret_cast<const uint4*>(gu + r * gu_stride + c);
    const uint4 uv = *reinterpret_cast<const uint4*>(gu + r * gu_stride + M + c);
    const uint32_t* g = &gv.x;
    const uint32_t* u = &uv.x;
    uint4 o;
    uint32_t* po = &o.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float g0 = bf16_lo(g[k]), g1 = bf16_hi(g[k]);
        const float s0 = round_bf16(g0 / (1.0f + __expf(-g0))), s1 = round_bf16(g1 / (1.0f + __expf(-g1)));
        po[k] = pack_bf16(s0 * bf16_lo(u[k]), s1 * bf16_hi(u[k]));
    }
    *reinterpret_cast<uint4*>(out + r * out_stride + c) = o;
}

// ---------------------------------------------------------------------------------------------------
// embed_tokens: token embedding lookup fused with the splice of the video embeddings (model_forward.py:58, :87-93
// `masked_scatter`): row t of the output is video[rank(t)] when ids[t] is the video placeholder (rank = number of
// placeholders before t among these ids), else embed[ids[t]].  One CTA per token row; the placeholder layout is read
// on the device, so a CUDA graph of the chunk prefill does not depend on it.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) embed_tokens_kernel(const int64_t* __restrict__ ids, int n, int64_t video_id, int64_t vocab,
                                                            const uint16_t* __restrict__ embed, const uint16_t* __restrict__ video,
                                                            int n_video, int hidden, uint16_t* __restrict__ out, int32_t* __restrict__ status) {
    pdl_wait();          // inputs come from the predecessor in the stream (no-op without the launch attribute)
    __shared__ int cnt[4];
    const int t = blockIdx.x;
    const int64_t id = ids[t];
    const uint16_t* src;
    if (id == video_id && video != nullptr) {
        int c = 0;
        for (int i = threadIdx.x; i < t; i += 128) c += ids[i] == video_id;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if ((threadIdx.x & 31) == 0) cnt[threadIdx.x >> 5] = c;
        __syncthreads();
        const int rank = cnt[0] + cnt[1] + cnt[2] + cnt[3];
        if (rank >= n_video) {                      // more placeholders than feature rows: flagged, row zeroed
            if (threadIdx.x == 0 && status) atomicOr(status, 8);
            for (int k = threadIdx.x * 8; k < hidden; k += 128 * 8) *reinterpret_cast<uint4*>(out + (int64_t)t * hidden + k) = make_uint4(0, 0, 0, 0);
            return;
        }
        src = video + (int64_t)rank * hidden;
    } else {
        const int64_t row = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        if (row != id && threadIdx.x == 0 && status) atomicOr(status, 16);
        src = embed + row * hidden;
    }
    for (int k = threadIdx.x * 8; k < hidden; k += 128 * 8)
        *reinterpret_cast<uint4*>(out + (int64_t)t * hidden + k) = __ldg(reinterpret_cast<const uint4*>(src + k));
}

// seen[ids[i]] = 1: the token set the repetition penalty applies to (HF RepetitionPenaltyLogitsProcessor gathers the scores at
// `input_ids`, generate/streaming_generate_qwen.py:75)
__global__ void __launch_bounds__(256) mark_seen_kernel(const int64_t* __restrict__ ids, int n, uint8_t* __restrict__ seen, int64_t vocab) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const int64_t id = ids[i];
        if (id >= 0 && id < vocab) seen[id] = 1;
    }
}

static int make_tables(RopeTables* t, const void* cos_table, const void* sin_table, int rows, const int32_t* sec, int D) {
    SVLM_REQUIRE(cos_table && sin_table && rows > 0, "rope tables missing");
    SVLM_REQUIRE(sec != nullptr, "mrope_section is NULL");
    SVLM_REQUIRE(sec[0] % 8 == 0 && sec[1] % 8 == 0 && sec[2] % 8 == 0 && sec[0] + sec[1] + sec[2] == D / 2,
                 "mrope_section {%d,%d,%d} must be multiples of 8 summing to %d", sec[0], sec[1], sec[2], D / 2);
    t->cos = reinterpret_cast<const uint4*>(cos_table);
    t->sin = reinterpret_cast<const uint4*>(sin_table);
    t->rows = rows;
    t->b0 = sec[0] / 8;
    t->b1 = (sec[0] + sec[1]) / 8;
    return SVLM_OK;
}

}  // namespace svlm

using namespace svlm;

extern "C" int svlm_kv_compact_rope(const svlm_arena_t* arena, int src_buffer, const int32_t* runs_host, int n_runs,
                                    const int32_t* pos, int pos_stride, const void* cos_table, const void* sin_table, int table_rows,
                                    const int32_t* mrope_section_host, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    if ((rc = validate_arena(arena))) return rc;
    SVLM_REQUIRE(src_buffer == 0 || src_buffer == 1, "src_buffer must be 0 or 1");
    SVLM_REQUIRE(n_runs >= 0 && n_runs <= SVLM_MAX_RUNS, "n_runs %d exceeds SVLM_MAX_RUNS %d", n_runs, SVLM_MAX_RUNS);
    SVLM_REQUIRE(n_runs == 0 || runs_host != nullptr, "runs_host is NULL");
    SVLM_REQUIRE(pos != nullptr && pos_stride > 0, "pos is NULL");
    RunTable rt;
    rt.n = n_runs;
    int n_dst = 0;
    for (int r = 0; r < n_runs; ++r) {
        const int s = runs_host[3 * r], d = runs_host[3 * r + 1], l = runs_host[3 * r + 2];
        SVLM_REQUIRE(l > 0 && s >= 0 && d >= 0 && s + l <= arena->capacity && d + l <= arena->capacity,
                     "run %d (src %d dst %d len %d) outside capacity %d", r, s, d, l, arena->capacity);
        SVLM_REQUIRE(d == n_dst, "runs must tile the destination contiguously from 0 (run %d starts at %d, expected %d)", r, d, n_dst);
        rt.src[r] = s; rt.dst[r] = d; rt.len[r] = l;
        n_dst += l;
    }
    SVLM_REQUIRE(n_dst <= pos_stride, "pos_stride %d smaller than retained length %d", pos_stride, n_dst);
    if (n_dst == 0) return SVLM_OK;
    RopeTables tab;
    if ((rc = make_tables(&tab, cos_table, sin_table, table_rows, mrope_section_host, arena->head_dim))) return rc;
    ArenaDims ad{arena->layers, arena->kv_heads, arena->head_dim, arena->capacity};
    dim3 grid((n_dst + CR_ROWS - 1) / CR_ROWS, arena->layers * arena->kv_heads);
    kv_compact_rope_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint16_t*)arena->base, ad, src_buffer, rt, n_dst, pos, pos_stride, tab);
    return launch_status("kv_compact_rope_kernel");
}

extern "C" int svlm_mrope_positions(const int64_t* ids, int n, const int32_t* grids_thw, int n_grids, int spatial_merge, int t_step,
                                    int64_t video_token_id, int32_t* pos_out, int pos_stride, int32_t* next_pos_out,
                                    int32_t* status_out, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(n >= 0 && n <= pos_stride, "n %d exceeds pos_stride %d", n, pos_stride);
    SVLM_REQUIRE(ids && pos_out && next_pos_out && status_out, "NULL pointer argument");
    SVLM_REQUIRE(n_grids == 0 || grids_thw != nullptr, "grids_thw is NULL");
    SVLM_REQUIRE(spatial_merge > 0 && t_step >= 0, "bad spatial_merge / t_step");
    SVLM_CUDA(cudaMemsetAsync(status_out, 0, sizeof(int32_t), (cudaStream_t)stream));
    mrope_positions_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ids, n, grids_thw, n_grids, spatial_merge, t_step, video_token_id,
                                                                  pos_out, pos_stride, next_pos_out, status_out);
    return launch_status("mrope_positions_kernel");
}

extern "C" int svlm_qkv_rope_append(const svlm_arena_t* arena, int buffer, int layer, const void* q, const void* k, const void* v,
                                    int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride, int n_tokens, int q_heads,
                                    void* q_out, int start, const int32_t* start_dev, const int32_t* pos, int pos_stride,
                                    const int32_t* text_pos_dev, const void* cos_table, const void* sin_table, int table_rows,
                                    const int32_t* mrope_section_host, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    if ((rc = validate_arena(arena))) return rc;
    SVLM_REQUIRE(buffer == 0 || buffer == 1, "buffer must be 0 or 1");
    SVLM_REQUIRE(layer >= 0 && layer < arena->layers, "layer %d out of range", layer);
    // q_heads == 0: append-only (K_raw / K_rot / V rows of the new tokens, no query), the route StreamingCache.update takes
    SVLM_REQUIRE(k && v && (q_heads == 0 || (q && q_out)), "NULL tensor argument");
    SVLM_REQUIRE(n_tokens >= 0 && q_heads >= 0, "bad n_tokens / q_heads");
    SVLM_REQUIRE(start >= 0 && start + n_tokens <= arena->capacity, "append of %d tokens at %d exceeds capacity %d", n_tokens, start,
                 arena->capacity);
    SVLM_REQUIRE(text_pos_dev != nullptr || (pos != nullptr && pos_stride >= start + n_tokens), "positions missing or too short");
    SVLM_REQUIRE((q_row_stride % 8) == 0 && (k_row_stride % 8) == 0 && (v_row_stride % 8) == 0, "row strides must be multiples of 8 elements");
    SVLM_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)q_out) & 15) == 0, "q/k/v/q_out must be 16-byte aligned");
    if (n_tokens == 0) return SVLM_OK;
    RopeTables tab;
    if ((rc = make_tables(&tab, cos_table, sin_table, table_rows, mrope_section_host, arena->head_dim))) return rc;
    ArenaDims ad{arena->layers, arena->kv_heads, arena->head_dim, arena->capacity};
    launch_pdl(PDL_ELEM, qkv_rope_append_kernel, dim3(n_tokens), dim3(256), 0, (cudaStream_t)stream, (uint16_t*)arena->base, ad, buffer, layer,
               (const uint16_t*)q, (const uint16_t*)k, (const uint16_t*)v, q_row_stride, k_row_stride, v_row_stride, q_heads, (uint16_t*)q_out,
               start, start_dev, pos, pos_stride, text_pos_dev, tab);
    return launch_status("qkv_rope_append_kernel");
}

extern "C" int svlm_qkv_rope_append_streams(const svlm_arena_t* geometry, const svlm_stream_slot_t* slots, int n_streams, int layer,
                                            const void* q, const void* k, const void* v, int64_t q_row_stride, int64_t k_row_stride,
                                            int64_t v_row_stride, int q_heads, void* q_out, const void* cos_table, const void* sin_table,
                                            int table_rows, const int32_t* mrope_section_host, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(geometry && slots, "NULL geometry / slots");
    SVLM_REQUIRE(geometry->layers > 0 && geometry->kv_heads > 0 && geometry->capacity > 0 && geometry->head_dim == 128, "bad arena geometry");
    SVLM_REQUIRE(n_streams >= 1 && n_streams <= SVLM_MAX_STREAMS, "n_streams %d out of range (1..%d)", n_streams, SVLM_MAX_STREAMS);
    SVLM_REQUIRE(layer >= 0 && layer < geometry->layers, "layer %d out of range", layer);
    SVLM_REQUIRE(q && k && v && q_out && cos_table && sin_table, "NULL tensor argument");
    SVLM_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)q_out) & 15) == 0, "q/k/v/q_out must be 16-byte aligned");
    SVLM_REQUIRE(q_row_stride % 8 == 0 && k_row_stride % 8 == 0 && v_row_stride % 8 == 0, "row strides must be multiples of 8 elements");
    RopeTables tab;
    if ((rc = make_tables(&tab, cos_table, sin_table, table_rows, mrope_section_host, geometry->head_dim))) return rc;
    ArenaDims ad{geometry->layers, geometry->kv_heads, geometry->head_dim, geometry->capacity};
    AppendSlots sl{};
    for (int b = 0; b < n_streams; ++b) {
        SVLM_REQUIRE(slots[b].arena_base && slots[b].row_dev && slots[b].pos_dev, "slot %d: NULL pointer", b);
        SVLM_REQUIRE(slots[b].buffer == 0 || slots[b].buffer == 1, "slot %d: buffer must be 0 or 1", b);
        sl.buf[b] = (uint16_t*)slots[b].arena_base + (int64_t)slots[b].buffer * ad.buffer_elems();
        sl.row[b] = slots[b].row_dev;
        sl.pos[b] = slots[b].pos_dev;
    }
    qkv_rope_append_streams_kernel<<<n_streams, 256, 0, (cudaStream_t)stream>>>(sl, ad, layer, (const uint16_t*)q, (const uint16_t*)k,
                                                                                (const uint16_t*)v, q_row_stride, k_row_stride,
                                                                                v_row_stride, q_heads, (uint16_t*)q_out, tab);
    return launch_status("qkv_rope_append_streams_kernel");
}

extern "C" int svlm_frames_to_patches(const void* frames_thwc_u8, int n_frames, int height, int width, int patch, int temporal_patch,
                                      int merge, const float* mean3_host, const float* std3_host, void* out, int out_is_bf16,
                                      void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(frames_thwc_u8 && out && mean3_host && std3_host, "NULL argument");
    SVLM_REQUIRE(patch > 0 && patch % 2 == 0 && temporal_patch > 0 && merge > 0, "bad patch / temporal_patch / merge");
    SVLM_REQUIRE(n_frames > 0 && n_frames % temporal_patch == 0, "n_frames %d must be a positive multiple of temporal_patch %d", n_frames,
                 temporal_patch);
    SVLM_REQUIRE(height > 0 && width > 0 && height % (patch * merge) == 0 && width % (patch * merge) == 0,
                 "frame size %dx%d must be a multiple of patch * merge = %d", height, width, patch * merge);
    SVLM_REQUIRE(((uintptr_t)out & 7) == 0, "out must be 8-byte aligned");
    IngestParams p{n_frames, height, width, patch, temporal_patch, merge, {mean3_host[0], mean3_host[1], mean3_host[2]},
                   {std3_host[0], std3_host[1], std3_host[2]}, out_is_bf16 ? 1 : 0};
    const int64_t total2 = (int64_t)n_frames * height * width * 3 / 2;
    frames_to_patches_kernel<<<(unsigned)((total2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)frames_thwc_u8, out, p);
    return launch_status("frames_to_patches_kernel");
}

extern "C" int svlm_frames_resize_bicubic(const void* in_thwc_u8, int n_frames, int height, int width, void* out_thwc_u8, int out_height,
                                          int out_width, void* tmp_f32, const int32_t* x_start, const int32_t* x_count,
                                          const float* x_weights, int kx, const int32_t* y_start, const int32_t* y_count,
                                          const float* y_weights, int ky, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(in_thwc_u8 && out_thwc_u8 && tmp_f32, "NULL frame / scratch pointer");
    SVLM_REQUIRE(x_start && x_count && x_weights && y_start && y_count && y_weights, "NULL tap table");
    SVLM_REQUIRE(n_frames > 0 && height > 0 && width > 0 && out_height > 0 && out_width > 0, "bad frame geometry");
    SVLM_REQUIRE(kx > 0 && ky > 0, "bad tap count");
    ResizeTaps tx{x_start, x_count, x_weights, kx}, ty{y_start, y_count, y_weights, ky};
    const int64_t n1 = (int64_t)n_frames * height * out_width, n2 = (int64_t)n_frames * out_height * out_width;
    resize_h_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)in_thwc_u8, (float*)tmp_f32,
                                                                                    n_frames * height, width, out_width, tx);
    if ((rc = launch_status("resize_h_kernel"))) return rc;
    resize_v_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float*)tmp_f32, (uint8_t*)out_thwc_u8, n_frames,
                                                                                    height, out_height, out_width, ty);
    return launch_status("resize_v_kernel");
}

extern "C" int svlm_vit_rope(void* qkv, const float* cos, const float* sin, int n_tokens, int heads, int head_dim, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(qkv && cos && sin, "NULL tensor argument");
    SVLM_REQUIRE(head_dim % 16 == 0 && head_dim <= 128, "head_dim %d must be a multiple of 16 and <= 128", head_dim);
    SVLM_REQUIRE(((uintptr_t)qkv & 15) == 0, "qkv must be 16-byte aligned");
    const int64_t total = (int64_t)n_tokens * 2 * heads * (head_dim / 16);
    if (total == 0) return SVLM_OK;
    launch_pdl(PDL_ELEM, vit_rope_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, (uint16_t*)qkv, cos, sin,
               n_tokens, heads, head_dim);
    return launch_status("vit_rope_kernel");
}

extern "C" int svlm_add_rmsnorm(const void* a, const void* b, const void* weight, void* out_sum, void* out_norm, int rows, int cols,
                                float eps, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(a && weight && out_norm, "NULL tensor argument");
    SVLM_REQUIRE(cols % 8 == 0 && cols <= 128 * RN_MAX_VEC * 8, "cols %d must be a multiple of 8 and <= %d", cols, 128 * RN_MAX_VEC * 8);
    SVLM_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)weight | (uintptr_t)out_sum | (uintptr_t)out_norm) & 15) == 0,
                 "tensors must be 16-byte aligned");
    if (rows == 0) return SVLM_OK;
    launch_pdl(PDL_ELEM, add_rmsnorm_kernel, dim3(rows), dim3(128), 0, (cudaStream_t)stream, (const uint16_t*)a, (const uint16_t*)b,
               (const uint16_t*)weight, (uint16_t*)out_sum, (uint16_t*)out_norm, cols, eps);
    return launch_status("add_rmsnorm_kernel");
}

extern "C" int svlm_rmsnorm(const void* x, const void* weight, void* out, int rows, int cols, float eps, void* stream) {
    return svlm_add_rmsnorm(x, nullptr, weight, nullptr, out, rows, cols, eps, stream);
}

extern "C" int svlm_swiglu(const void* gate_up, void* out, int rows, int M, int64_t gate_up_row_stride, int64_t out_row_stride, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(gate_up && out, "NULL tensor argument");
    SVLM_REQUIRE(M % 8 == 0 && gate_up_row_stride % 8 == 0 && out_row_stride % 8 == 0, "M and row strides must be multiples of 8");
    SVLM_REQUIRE((((uintptr_t)gate_up | (uintptr_t)out) & 15) == 0, "tensors must be 16-byte aligned");
    const int64_t total = (int64_t)rows * (M / 8);
    if (total == 0) return SVLM_OK;
    launch_pdl(PDL_ELEM, swiglu_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, (const uint16_t*)gate_up,
               (uint16_t*)out, rows, M, gate_up_row_stride, out_row_stride);
    return launch_status("swiglu_kernel");
}

extern "C" int svlm_embed_tokens(const int64_t* ids, int n_tokens, int64_t video_token_id, const void* embed, int64_t vocab, const void* video,
                                 int n_video, int hidden, void* out, int32_t* status, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(ids && embed && out, "NULL tensor argument");
    SVLM_REQUIRE(n_tokens >= 0 && vocab > 0 && n_video >= 0, "bad n_tokens / vocab / n_video");
    SVLM_REQUIRE(hidden > 0 && hidden % 8 == 0, "hidden %d must be a positive multiple of 8", hidden);
    SVLM_REQUIRE((((uintptr_t)embed | (uintptr_t)video | (uintptr_t)out) & 15) == 0, "embed / video / out must be 16-byte aligned");
    if (n_tokens == 0) return SVLM_OK;
    launch_pdl(PDL_ELEM, embed_tokens_kernel, dim3(n_tokens), dim3(128), 0, (cudaStream_t)stream, ids, n_tokens, (int64_t)video_token_id,
               (int64_t)vocab, (const uint16_t*)embed, (const uint16_t*)video, n_video, hidden, (uint16_t*)out, status);
    return launch_status("embed_tokens_kernel");
}

extern "C" int svlm_mark_seen(const int64_t* ids, int n, uint8_t* seen, int64_t vocab, int clear_first, void* stream) {
    int rc = check_device();
    if (rc) return rc;
    SVLM_REQUIRE(seen && vocab > 0 && n >= 0 && (n == 0 || ids), "bad argument");
    if (clear_first) SVLM_CUDA(cudaMemsetAsync(seen, 0, (size_t)vocab, (cudaStream_t)stream));
    if (n == 0) return SVLM_OK;
    mark_seen_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ids, n, seen, vocab);
    return launch_status("mark_seen_kernel");
}
