/*
 * svlm_b200.h -- C ABI of the B200-native (sm_100a) StreamingVLM hot path.
 *
 * The reference (rahim-xelpmoc/streaming-vlm) is pure Python and has no FFI: its "operator
 * interface" for this path is a set of Python call sites that invoke torch / flash-attn
 * kernels.  Each entry point below replaces one of those call sites (cited as
 * file:line relative to the reference's src/streaming_vlm/) and is what a torch custom-op
 * / ctypes layer binds.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every `const void*` / `void*` tensor argument is a DEVICE pointer unless the name ends
 *     in `_host`; `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - all kernels are asynchronous on `stream`, allocate nothing, never synchronise, and are
 *     CUDA-graph capturable; bf16 tensors are raw `uint16_t` bit patterns;
 *   - return value: 0 on success, a negative SVLM_E* code otherwise; `svlm_last_error()`
 *     returns a thread-local, human-readable message for the last failure;
 *   - there is no CPU fallback: on a machine without an sm_100 device every compute entry
 *     point returns SVLM_ENODEV (svlm_find_segment, a host-only helper of the eviction planner,
 *     is the one function that does no device work).
 *
 * KV arena (one per video stream, owned by the caller, e.g. a torch tensor):
 *   bf16 [2 buffers][3 planes][L layers][H_kv][cap tokens][D]
 *   plane 0 = K_raw (un-rotated keys, the reference's cache content in shrink mode),
 *   plane 1 = K_rot (keys rotated at their current contiguous positions),
 *   plane 2 = V.
 *   Compaction ping-pongs between the two buffers (source and destination never alias).
 */
#ifndef SVLM_B200_H_
#define SVLM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVLM_OK        0
#define SVLM_EINVAL   -1   /* bad argument (shape, alignment, capacity, too many runs ...) */
#define SVLM_ECUDA    -2   /* a CUDA runtime / driver call failed */
#define SVLM_ENODEV   -3   /* no sm_100 device / driver entry point unavailable */

#define SVLM_MAX_RUNS 32   /* maximum number of gather runs per compaction launch */
#define SVLM_ATTN_EARLY_KV 1 /* flag of svlm_attn_causal / svlm_attn_decode_streams */
#define SVLM_MAX_STREAMS 16 /* streams whose decode steps share one launch (section 8) */
#define SVLM_MAX_TP 8      /* ranks of one tensor-parallel group (one NVLink / NVSwitch node) */

#define SVLM_PLANE_K_RAW 0
#define SVLM_PLANE_K_ROT 1
#define SVLM_PLANE_V     2

int         svlm_version(void);          /* ABI version, currently 1 */
const char* svlm_last_error(void);

/* Description of one stream's arena; strides follow from the dense layout above. */
typedef struct {
    void*   base;        /* device pointer to buffer 0, plane 0 */
    int32_t layers;      /* L */
    int32_t kv_heads;    /* H_kv */
    int32_t head_dim;    /* D (128) */
    int32_t capacity;    /* cap: tokens per (layer, head) plane */
} svlm_arena_t;

/* ---------------------------------------------------------------------------------------
 * (1) KV-cache compaction fused with the contiguous-RoPE re-rotation.
 * Replaces: inference/inference.py:50-61 (prune_id_and_kv_cache: index_select per layer),
 *           :100-108 (resort_id_and_kv: torch.cat), :63-68 (contiguous_id_and_kv), i.e. every
 *           full-cache copy process_past_kv (:87-172) makes in one chunk, AND the per-forward
 *           re-rotation of the whole cached K, qwen2_5/language_forward.py:58-66, for the
 *           retained tokens.
 * The gather is given as `n_runs` runs (src_start, dst_start, length), all layers / heads /
 * planes in ONE launch: dst[plane][l][h][dst_start+i] = src[plane][l][h][src_start+i] for
 * K_raw and V (bit-exact copy), and K_rot[dst] = mrope(K_raw[src], pos[:, dst]) with the
 * reference's bf16 rounding order (x*cos and rot(x)*sin each rounded, then the sum).
 *   runs_host   : host pointer, int32 [n_runs][3]
 *   src_buffer  : 0 or 1; the destination is the other buffer
 *   pos         : int32 [3][pos_stride] contiguous mRoPE positions of the DESTINATION tokens
 *   cos_table / sin_table : bf16 [table_rows][D/2]; row p = cos/sin(p * inv_freq) rounded once
 *   mrope_section : int32 [3], e.g. {16,24,24}; every section must be a multiple of 8
 * ------------------------------------------------------------------------------------- */
int svlm_kv_compact_rope(const svlm_arena_t* arena, int src_buffer, const int32_t* runs_host, int n_runs,
                         const int32_t* pos, int pos_stride, const void* cos_table, const void* sin_table,
                         int table_rows, const int32_t* mrope_section_host, void* stream);

/* ---------------------------------------------------------------------------------------
 * (2) Contiguous 3-D mRoPE position ids of a pruned id sequence.
 * Replaces: qwen2_5/pos_emb.py:6-188 (get_rope_index) as called in shrink mode by
 *           qwen2_5/model_forward.py:101-110 (full ids, image_grid_thw=None, mask of ones).
 *   ids        : int64 [n] device
 *   grids_thw  : int32 [n_grids][3] device, consumed from index 0 in block order
 *   t_step     : (2 / FPS) * tokens_per_second as an integer (2 for the pinned versions)
 *   pos_out    : int32 [3][pos_stride]; integer-valued, equal to the reference's float32 ids
 *   next_pos_out : int32 [1] = max position + 1 (first position of a token appended later)
 *   status_out : int32 [1], set non-zero if a run of video tokens does not match its grid
 * ------------------------------------------------------------------------------------- */
int svlm_mrope_positions(const int64_t* ids, int n, const int32_t* grids_thw, int n_grids, int spatial_merge,
                         int t_step, int64_t video_token_id, int32_t* pos_out, int pos_stride,
                         int32_t* next_pos_out, int32_t* status_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * (3) Append new tokens: rotate q, write K_raw / K_rot / V rows into the arena.
 * Replaces: generate/streaming_cache.py:72-73 (torch.cat append per layer per forward) and
 *           qwen2_5/language_forward.py:47-56, :64-66 (mRoPE of q, right aligned, and of
 *           the new keys).
 *   q / k / v  : bf16, token-major; element (t, head, d) at t*row_stride + head*D + d
 *   q_out      : bf16 [T][Hq][D] rotated queries
 *   start      : arena row of the first new token; if start_dev != NULL the kernel reads it
 *                from device memory instead (graph-replayable decode)
 *   pos        : int32 [3][pos_stride]; the token at arena row r uses pos[:, r].
 *                If text_pos_dev != NULL (decode), token t instead takes the scalar
 *                *text_pos_dev + t on all three axes (plain text continuation).
 * ------------------------------------------------------------------------------------- */
int svlm_qkv_rope_append(const svlm_arena_t* arena, int buffer, int layer, const void* q, const void* k,
                         const void* v, int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride,
                         int n_tokens, int q_heads, void* q_out, int start, const int32_t* start_dev,
                         const int32_t* pos, int pos_stride, const int32_t* text_pos_dev,
                         const void* cos_table, const void* sin_table, int table_rows,
                         const int32_t* mrope_section_host, void* stream);

/* ---------------------------------------------------------------------------------------
 * (4) Causal GQA attention over the arena: chunk prefill (Tq > 1) and decode (Tq = 1).
 * Replaces: qwen2_5/language_forward.py:199-252 (repeat_kv + _flash_attention_forward ->
 *           flash_attn_func, causal, bottom-right aligned, scale 1/sqrt(D)).
 * tcgen05 (UMMA) QK^T and PV with TMEM accumulators, TMA-staged K/V tiles, online softmax.
 *   q          : bf16 [Tq][Hq][D] (rotated), out: bf16 [Tq][Hq][D]
 *   kv_len     : number of valid arena rows INCLUDING the Tq new tokens; if kv_len_dev != NULL
 *                it is read from device memory (kv_len then is an upper bound used for the grid)
 *   num_splits : >= 1; with > 1 the KV range is split across CTAs and merged by a second
 *                kernel.  workspace: fp32, svlm_attn_workspace_bytes(...) bytes.
 *   flags      : 0, or SVLM_ATTN_EARLY_KV (decode only, Tq = 1, with kv_len_dev): the caller guarantees that *kv_len_dev and
 *                the K/V rows below kv_len - 1 were NOT written by the launch immediately before this one in the stream (the
 *                decode chain: the preceding q/k/v launch appends row kv_len - 1 only; the counters were advanced at the end of
 *                the previous token).  The kernel, which is launched with programmatic stream serialization, then reads the
 *                length and requests those K/V tiles BEFORE the grid dependency resolves.
 * ------------------------------------------------------------------------------------- */
size_t svlm_attn_workspace_bytes(int n_q_tokens, int q_heads, int head_dim, int num_splits);
int svlm_attn_causal(const svlm_arena_t* arena, int buffer, int layer, const void* q, void* out, int n_q_tokens,
                     int q_heads, int kv_len, const int32_t* kv_len_dev, int num_splits, void* workspace,
                     size_t workspace_bytes, float scale, int flags, void* stream);

/* ---------------------------------------------------------------------------------------
 * (5) ViT attention: 2-D RoPE + packed variable-length non-causal attention.
 * Replaces: qwen2_5/vision_forward.py:28 (apply_rotary_pos_emb_flashatt) and :32-35
 *           (flash_attn_varlen_func with cu_seqlens = windows or whole frames).
 *   qkv        : bf16 [N][3][H][D] (output of the fused qkv Linear); q and k are rotated IN
 *                PLACE by svlm_vit_rope, then read by svlm_vit_attn through TMA
 *   cos / sin  : fp32 [N][D] (the reference's position_embeddings after window reordering)
 *   cu_seqlens : int32 [n_seqs + 1] device, ascending, last = N
 *   out        : bf16 [N][H][D]
 * ------------------------------------------------------------------------------------- */
int svlm_vit_rope(void* qkv, const float* cos, const float* sin, int n_tokens, int heads, int head_dim,
                  void* stream);
int svlm_vit_attn(const void* qkv, const int32_t* cu_seqlens, int n_seqs, void* out, int n_tokens, int heads,
                  int head_dim, float scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * (6) Dense bf16 GEMM with fused epilogue, C = act(A * W^T + bias), on tcgen05 / TMA.
 * Replaces: HF Qwen2_5_VLPatchMerger.forward as called at qwen2_5/vision_forward.py:98
 *           (Linear + GELU(erf) + Linear; the RMSNorm prologue is svlm_rmsnorm and the
 *           `argsort(window_index)` un-permute of :99 is the optional row scatter).
 *   A [M][K] bf16 row-major, W [N][K] bf16 row-major (nn.Linear layout), bias [N] bf16 or NULL
 *   act        : 0 = none, 1 = GELU(erf)
 *   out_rows   : int32 [M] or NULL; row m of the result is stored at row out_rows[m] of C
 * ------------------------------------------------------------------------------------- */
int svlm_gemm_bf16(const void* A, const void* W, const void* bias, void* C, int M, int N, int K, int act,
                   const int32_t* out_rows, void* stream);
/* The same persistent kernel with the epilogues the dense projections of the ViT blocks and of the chunk prefill need
 * (HF nn.Linear call sites qwen2_5/language_forward.py:175-177, 255, 295-297 and qwen2_5/vision_forward.py:14, 36,
 * 52), replacing cuBLAS + separate bias / activation / residual kernels:
 *   epilogue 0: C = bf16(A W^T + bias)                     1: C = GELU(bf16(A W^T + bias))
 *   epilogue 2: C = residual + bf16(A W^T + bias)          (residual bf16 [M][N], may alias C; the decoder's `hidden = residual + x`)
 *   epilogue 3: SwiGLU: W holds gate and up rows interleaved in blocks of 64 ([gate 0:64 | up 0:64 | gate 64:128 | ...], N = 2 * width);
 *               C [M][N/2] = bf16(silu(bf16(gate))) * bf16(up)  (`act_fn(gate_proj(x)) * up_proj(x)`), no bias
 *   tile_n: 0 = chosen per shape from the measured UMMA costs, or 64 / 128 / 256 to force a tile width. */
int svlm_gemm_bf16_ex(const void* A, const void* W, const void* bias, const void* residual, void* C, int M, int N, int K,
                      int epilogue, const int32_t* out_rows, int tile_n, void* stream);
/* "Few tokens x whole weight matrix": Y[T][N] = epilogue(X[T][K] W^T + bias), 1 <= T <= 64.  The dense projections of the BATCHED decode
 * step (one token per stream, T = streams; the same nn.Linear call sites as above executed for a [T, hidden] activation matrix) are
 * weight streams: the weight rows are the tensor-core M dimension (swap-AB, 128-row TMA panels), the tokens the N dimension
 * (16 / 32 / 64 TMEM columns); small N is split along K over a thread-block cluster whose partial tiles are reduced through
 * distributed shared memory in rank order (deterministic, no workspace).  One launch per projection:
 *   epilogue 0: Y = bf16(X W^T + bias)                      (q/k/v projection, lm_head)
 *   epilogue 2: Y = bf16(bf16(X W^T + bias) + residual)     (o_proj / down_proj + `hidden = residual + x`; residual [T][N], may alias Y)
 *   epilogue 3: SwiGLU: W = [gate rows ; up rows] (N = 2 * width, plain concatenation), Y [T][N/2] = bf16(silu(bf16 gate)) * bf16 up
 *   splits: 0 = chosen per shape (~1.5 CTAs per SM), else 1..8 CTAs per cluster along K.
 *   norm_weight (bf16 [K]) / eps: non-NULL = X is RMS-normalised on the fly (HF Qwen2RMSNorm in front of q/k/v, gate/up, lm_head;
 *       row statistics and rounding points identical to svlm_add_rmsnorm): the epilogue warps normalise every X panel in shared
 *       memory while the weight ring streams -- no norm launch, no normalised copy of X.
 *   stats_out (fp32 [ceil(N / 128)][T], epilogue 0 / 2, T <= 16) or NULL: the sum of squares of the 128 values of every token that each
 *       row tile wrote -- what the NEXT RMSNorm needs of Y.  stats_in / stats_tiles: such a table for X, left by the launch that
 *       produced X (o_proj -> post-attention norm, down_proj -> next layer's input norm): the row statistics then cost one short read
 *       instead of a pass over X in every CTA; NULL = computed from X in the kernel. */
int svlm_gemm_tokens(const void* X, const void* W, const void* bias, const void* residual, void* Y, int T, int N, int K,
                     int epilogue, int splits, const void* norm_weight, float eps, const float* stats_in, int stats_tiles,
                     float* stats_out, void* stream);
/* lm_head of the batched decode step without a logits tensor (final norm + `lm_head` + `suppress_tokens` + argmax, qwen2_5/
 * model_forward.py:316-317 and generate/streaming_generate_qwen.py:96-99 for greedy decoding): per token t < T <= 16, keys[t] receives
 * max over the vocabulary rows n that are not suppressed of (ordered bf16 logit << 32 | 0xffffffff - n), i.e. the FIRST index of the
 * maximum logit, exactly torch.argmax over `lm_head(h).float().masked_fill(suppress, -inf)`.  keys must be zero on entry
 * (svlm_streams_advance clears them).  norm_weight / stats_in as in svlm_gemm_tokens. */
int svlm_gemm_tokens_lm_head_argmax(const void* X, const void* W, const void* norm_weight, float eps, const float* stats_in,
                                    int stats_tiles, const uint8_t* suppress, int T, int N, int K, unsigned long long* keys,
                                    void* stream);
/* End of a batched decode step: stream b takes the id selected in keys[b] and advances its device-side counters if active[b] != 0
 * (tok, kv_len += 1, next_pos += 1); kv_after[b] = kv_len[b] + 1 for every stream; keys are cleared. */
int svlm_streams_advance(unsigned long long* keys, int64_t* tok, const int32_t* active, int32_t* kv_len, int32_t* kv_after,
                         int32_t* next_pos, int n_streams, void* stream);
int svlm_rmsnorm(const void* x, const void* weight, void* out, int rows, int cols, float eps, void* stream);

/* ---------------------------------------------------------------------------------------
 * (6a) Token embedding fused with the splice of the video embeddings.
 * Replaces: qwen2_5/model_forward.py:58 (embed_tokens) + :80-93 (count check, masked_scatter of video_embeds).
 *   ids [n_tokens] int64 device; embed bf16 [vocab][hidden]; video bf16 [n_video][hidden] or NULL
 *   out bf16 [n_tokens][hidden]: row t = video[r] if ids[t] == video_token_id (r = placeholders before t), else embed[ids[t]]
 *   status (int32 device, may be NULL): bit 8 = more placeholders than video rows, bit 16 = id outside the vocabulary
 * ------------------------------------------------------------------------------------- */
int svlm_embed_tokens(const int64_t* ids, int n_tokens, int64_t video_token_id, const void* embed, int64_t vocab,
                      const void* video, int n_video, int hidden, void* out, int32_t* status, void* stream);

/* ---------------------------------------------------------------------------------------
 * (6b) Elementwise glue of the prefill / ViT blocks, fused (bf16 in / out, fp32 arithmetic).
 * svlm_add_rmsnorm replaces `hidden = residual + x` followed by the next RMSNorm
 *   (qwen2_5/language_forward.py:290-295, qwen2_5/vision_forward.py:46-52): out_sum = bf16(a + b) (b and out_sum may
 *   be NULL = plain RMSNorm), out_norm = RMSNorm(out_sum) * weight.
 * svlm_swiglu replaces `act_fn(gate_proj(x)) * up_proj(x)` of Qwen2MLP / Qwen2_5_VLMLP on the fused [gate | up]
 *   GEMM output: gate_up [rows][2M] -> out [rows][M].
 * ------------------------------------------------------------------------------------- */
int svlm_add_rmsnorm(const void* a, const void* b, const void* weight, void* out_sum, void* out_norm, int rows, int cols,
                     float eps, void* stream);
int svlm_swiglu(const void* gate_up, void* out, int rows, int M, int64_t gate_up_row_stride, int64_t out_row_stride, void* stream);

/* ---------------------------------------------------------------------------------------
 * (7) Decode-step (one token) weight-streaming GEMV with fused RMSNorm prologue and epilogue; HBM-bound.
 * Replaces, for q_len == 1: the nn.Linear / RMSNorm / SiLU / residual-add calls of
 *           qwen2_5/language_forward.py:175-177 (q/k/v_proj after input_layernorm, :277), :255 (o_proj) + :290
 *           (residual), :295-297 (post_attention_layernorm, MLP, residual), and :497 + qwen2_5/model_forward.py:317
 *           (final norm + lm_head) followed by the greedy argmax of generate/streaming_generate_qwen.py:99.
 *   W [rows][K] bf16 row-major (nn.Linear layout), x [K] bf16, y [N] bf16
 *   norm_weight : if non-NULL, x is first replaced by HF RMSNorm(x) * norm_weight (eps)
 *   mode 0: y = W x + bias            mode 1: y = residual + (W x + bias)      (y may alias residual)
 *   mode 2: W = [gate rows ; up rows] (2N rows), y = silu(gate) * up
 * svlm_lm_head_argmax: token_out = first index of the max bf16 logit over rows with suppress[row] == 0, then advances
 * the decode loop's device counters (kv_len += 1, kv_after = kv_len + 1, next_pos += 1).  key_scratch: one zeroed
 * uint64 owned by the caller.
 * ------------------------------------------------------------------------------------- */
int svlm_gemv_bf16(const void* W, const void* x, const void* norm_weight, float eps, const void* bias,
                   const void* residual, void* y, int N, int K, int mode, void* stream);
/* Decode-step fusion of (7, mode 0 on the fused [q;k;v] weight) with (3): RMSNorm + QKV projection + bias, then the rows of
 * each head are rotated (text position *pos_dev on all three mRoPE axes) and written straight to q_rot [Hq][128] and to
 * the arena row *row_dev (K_raw, K_rot, V) of `layer` -- no intermediate qkv vector, no separate append launch. */
int svlm_gemv_qkv_rope_append(const svlm_arena_t* arena, int buffer, int layer, const void* W, const void* x,
                              const void* norm_weight, float eps, const void* bias, int q_heads, int hidden, void* q_rot,
                              const int32_t* row_dev, const int32_t* pos_dev, const void* cos_table, const void* sin_table,
                              int table_rows, void* stream);
int svlm_lm_head_argmax(const void* W, const void* x, const void* norm_weight, float eps, const uint8_t* suppress, int N,
                        int K, uint64_t* key_scratch, int64_t* token_out, int32_t* kv_len, int32_t* kv_after,
                        int32_t* next_pos, void* stream);
/* The same kernel with the reference's generation config (inference/inference.py:41-42, 426-451: do_sample=True,
 * temperature=0.9, repetition_penalty=1.05; the logits processors and the multinomial draw of
 * generate/streaming_generate_qwen.py:73-99) fused into the lm_head epilogue, per vocabulary row n:
 *     score = fp32(bf16 logit);  if seen[n]: score = score < 0 ? score * penalty : score / penalty   (HF RepetitionPenaltyLogitsProcessor)
 *     score = score / temperature (skipped when temperature == 1)                                    (HF TemperatureLogitsWarper)
 *     Both divisions are evaluated the way ATen evaluates `tensor / python_scalar` on the GPU, where the reference runs them: a
 *     multiplication with float(1.0 / double(scalar)).
 *     do_sample: token = argmax_n(score_n + Gumbel noise_n)  -- an exact sample of softmax(score) (Gumbel-max), noise from
 *                Philox4x32-10 with key (rng_state[0], rng_state[1]) and counter (n, rng_state[2]); else first argmax of score.
 * `seen` (uint8 [N], svlm_mark_seen) is the set of ids in the sequence so far; the chosen token is added to it and the step
 * word rng_state[2] is incremented on the device (rng_state: uint32 [4] in device memory: seed lo, seed hi, step, reserved),
 * so a CUDA-graph replay per token needs no host involvement and a new seed needs no new graph.  scores_out (fp32 [N] or NULL)
 * receives the processed scores (HF `output_scores`), -inf for suppressed rows. */
int svlm_lm_head_sample(const void* W, const void* x, const void* norm_weight, float eps, const uint8_t* suppress, int N,
                        int K, uint64_t* key_scratch, int64_t* token_out, int32_t* kv_len, int32_t* kv_after,
                        int32_t* next_pos, uint8_t* seen, double repetition_penalty, double temperature, int do_sample,
                        uint32_t* rng_state, float* scores_out, void* stream);
/* seen[ids[i]] = 1 for i < n (after clearing all `vocab` bytes when clear_first != 0): the repetition-penalty target set of
 * the ids already in the sequence (generate/streaming_generate_qwen.py:75 passes the full `input_ids`). */
int svlm_mark_seen(const int64_t* ids, int n, uint8_t* seen, int64_t vocab, int clear_first, void* stream);

/* ---------------------------------------------------------------------------------------
 * (8) Decode step of SEVERAL independent streams in one launch (the reference runs one stream per process,
 * eval/livesports3kcc/distributed_generate_streaming.py:53-62; streams that share a GPU share one sweep of the
 * weights here, and their split-KV attention becomes one grid that is large enough to be HBM-bound).
 * Every stream has its own arena of the SAME geometry (`geometry->base` is ignored) and its own device-side
 * counters; `slots` is a HOST array of n_streams <= SVLM_MAX_STREAMS entries, read at launch time.
 *   svlm_qkv_rope_append_streams : token b of q/k/v [n_streams][...] is rotated at text position *pos_dev (all three
 *       mRoPE axes) and appended to row *row_dev of stream b (same arithmetic as svlm_qkv_rope_append).
 *   svlm_attn_decode_streams     : q, out bf16 [n_streams][Hq][D]; stream b attends to its first *kv_len_dev rows.
 *       workspace: svlm_attn_workspace_bytes(n_streams, Hq, D, num_splits) bytes when num_splits > 1.
 * ------------------------------------------------------------------------------------- */
typedef struct {
    void*          arena_base;   /* this stream's arena (buffer 0, plane 0) */
    int32_t        buffer;       /* its live ping-pong buffer */
    const int32_t* row_dev;      /* device scalar: arena row of the token being appended */
    const int32_t* pos_dev;      /* device scalar: its text position */
    const int32_t* kv_len_dev;   /* device scalar: rows to attend to (including the appended token) */
} svlm_stream_slot_t;
int svlm_qkv_rope_append_streams(const svlm_arena_t* geometry, const svlm_stream_slot_t* slots, int n_streams, int layer,
                                 const void* q, const void* k, const void* v, int64_t q_row_stride, int64_t k_row_stride,
                                 int64_t v_row_stride, int q_heads, void* q_out, const void* cos_table, const void* sin_table,
                                 int table_rows, const int32_t* mrope_section_host, void* stream);
int svlm_attn_decode_streams(const svlm_arena_t* geometry, const svlm_stream_slot_t* slots, int n_streams, int layer,
                             const void* q, void* out, int q_heads, int max_kv_len, int num_splits, void* workspace,
                             size_t workspace_bytes, float scale, int flags, void* stream);
/* The q/k/v projection of the batched decode step with everything around it in ONE launch (replaces svlm_add_rmsnorm + GEMM + bias +
 * svlm_qkv_rope_append_streams; reference: input_layernorm + q/k/v_proj + apply_multimodal_rotary_pos_emb + cache update,
 * language_forward.py:47-66, 175-192, generate/streaming_cache.py:72-73): X [n_streams][hidden] is normalised (norm_weight, eps),
 * multiplied with W [(q_heads + 2 kv_heads) * 128][hidden] (+ bias); a 128-row output tile is one head, which the epilogue rotates
 * at text position *slots[b].pos_dev (all three mRoPE axes) and writes to q_rot [n_streams][q_heads][128] resp. to the K_raw / K_rot /
 * V planes of stream b's arena at row *slots[b].row_dev.  cos / sin tables as in svlm_qkv_rope_append; stats_in as in svlm_gemm_tokens. */
int svlm_gemm_tokens_qkv_append(const svlm_arena_t* geometry, const svlm_stream_slot_t* slots, int n_streams, int layer,
                                const void* X, const void* W, const void* bias, const void* norm_weight, float eps,
                                const float* stats_in, int stats_tiles, int q_heads, int hidden, void* q_rot,
                                const void* cos_table, const void* sin_table, int table_rows, int splits, void* stream);

/* ---------------------------------------------------------------------------------------
 * (9) Frame ingest ("next" row f3).  Replaces the pixel half of `processor(text=..., videos=clip)`
 * (inference/inference.py:390-395 -> HF Qwen2-VL video processor) for a clip that is already at its target size:
 * uint8 RGB frames [T][H][W][3] (decord's on-wire layout) -> ViT patch rows [gt*gh*gw][3*Tp*P*P] in merge-unit-major
 * order, normalised as (x - mean') / std' in fp32 with mean' = mean * 255, std' = std * 255 (HF's fused rescale +
 * normalise); out is fp32 or bf16 (the ViT's patch embedding consumes bf16).  Sizes must be multiples of patch * merge
 * (svlm_frames_resize_bicubic below brings a clip there).
 * ------------------------------------------------------------------------------------- */
int svlm_frames_to_patches(const void* frames_thwc_u8, int n_frames, int height, int width, int patch, int temporal_patch,
                           int merge, const float* mean3_host, const float* std3_host, void* out, int out_is_bf16,
                           void* stream);

/* Spatial resize in front of (9): `livecc_utils/.../video_process_patch.py:133-153` (`torchvision.transforms.functional.resize`
 * with BICUBIC, antialias=True, on the uint8 clip; target size from `smart_resize`).  Separable anti-aliased bicubic, horizontal
 * pass first into the fp32 scratch tmp_f32 [T][H][out_width][3], round-half-even + clamp at the end.  The tap tables (device
 * pointers: first input index and tap count per output index, normalised weights [n_out][k]) are built by the caller the way
 * ATen builds them (`streaming_vlm_b200/utils/frames.py::aa_bicubic_taps`). */
int svlm_frames_resize_bicubic(const void* in_thwc_u8, int n_frames, int height, int width, void* out_thwc_u8, int out_height,
                               int out_width, void* tmp_f32, const int32_t* x_start, const int32_t* x_count, const float* x_weights,
                               int kx, const int32_t* y_start, const int32_t* y_count, const float* y_weights, int ky, void* stream);

/* ---------------------------------------------------------------------------------------
 * (10) Head-sharded tensor parallelism over NVLink peer memory (BASELINE configs[4]; the reference has no such mode -- new
 * functionality).  Rank r owns a slice of the q / kv heads and of the MLP width; the two row-parallel
 * projections of a decoder layer (o_proj, down_proj) produce PARTIAL sums that have to be added over the ranks.  Instead of an
 * NCCL all-reduce + a residual-add kernel per projection, the partial sums go to a buffer in SYMMETRIC memory (same allocation
 * on every rank, every rank has every peer's copy mapped: torch.distributed._symmetric_memory or cuMem fabric handles) and the
 * reduction is a one-shot exchange over NVLink loads inside our own kernels -- no NCCL call, CUDA-graph capturable:
 *   svlm_tp_gemv_allreduce : decode step.  y = residual + bf16(sum_r W_r x_r): each CTA streams its rows of W (bf16 [N][K_local])
 *       and PUSHES its fp32 partial sums to every rank's region as 8-byte (value, epoch) words (NVLink stores; the epoch travels
 *       with the value, so there is no flag, fence or round trip); the CTA that finishes last waits until the words of all ranks
 *       carry this epoch, sums the W partial vectors in rank order, adds the residual, writes y.  (SVLM_TP_PUSH=0: the partial sums
 *       stay in the rank's own region, the last CTA meets the peers through the flag words and pulls their vectors.)
 *   svlm_tp_allreduce_rows : chunk prefill.  The partial products [rows][cols] bf16 were written into this rank's region by the
 *       GEMM (svlm_tp_ctx_t offsets below); out = residual + bf16(sum_r part_r).
 * Symmetric buffer layout (bytes, identical on every rank):  [0, 1024) flag words (uint32 per peer rank), then two regions of
 * `region_bytes` for decode partials ([source rank][N] x (fp32 value, uint32 epoch): which = 0 attention output, 1 MLP output) at
 * SVLM_TP_DECODE_OFFSET, then two
 * regions of `region_bytes` for row partials (bf16 [rows][cols]) behind them.  Two regions (`which`) are what makes reuse safe
 * without a trailing barrier: a region is rewritten only after every peer has passed the NEXT all-reduce.
 * `seq` / `done`: two zero-initialised uint32 in LOCAL device memory owned by the context. */
typedef struct {
    int32_t   world, rank;
    void*     peer_base[SVLM_MAX_TP];   /* peer_base[r]: rank r's symmetric buffer as mapped in this process (peer_base[rank] = own) */
    uint64_t  region_bytes;             /* size of each of the four partial regions */
    uint32_t* seq;                      /* local device counter: all-reduces completed (the epoch) */
    uint32_t* done;                     /* local device counter: finished CTAs of the running launch */
} svlm_tp_ctx_t;
#define SVLM_TP_DECODE_OFFSET 1024
#define SVLM_TP_ROWS_OFFSET(ctx) (SVLM_TP_DECODE_OFFSET + 2 * (size_t)(ctx)->region_bytes)
#define SVLM_TP_BUFFER_BYTES(region_bytes) (SVLM_TP_DECODE_OFFSET + 4 * (size_t)(region_bytes))
int svlm_tp_gemv_allreduce(const svlm_tp_ctx_t* ctx, int which, const void* W, const void* x, const void* residual, void* y,
                           int N, int K, void* stream);
int svlm_tp_allreduce_rows(const svlm_tp_ctx_t* ctx, int which, const void* residual, void* out, int rows, int cols, void* stream);

/* Scheduling aid of the decode loop (no reference counterpart: the reference leaves cache policy to the libraries it
 * calls): issues L2 prefetches for [ptr, ptr + bytes).  The session runner launches it on a side stream next to the
 * latency-bound decode attention so that the weights of the following projections are L2-resident when they start. */
int svlm_prefetch_l2(const void* ptr, size_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Host helper (no device work, usable without a GPU): the token-pattern scan behind
 * `get_qwen_range` (utils/get_qwen_range.py:15-86), which `process_past_kv` (inference/inference.py:91-170) runs five times per
 * chunk over the ~5.8k ids of the stream.  Segments are found left to right and never overlap: a segment starts where `pattern`
 * (1..8 tokens) matches and ends at the first `end_token` at or after the end of the pattern -- a start without a terminator ends the
 * scan --; a trailing `lf_token` is absorbed when `contain_lf`.  `index` >= 0 counts from the left, < 0 (down to -256) from the right.
 * Writes the closed interval to *start_out / *end_out and, if count_out != NULL, the number of segments seen (all of them for a
 * negative index).  SVLM_EINVAL when the index is out of range (the reference raises IndexError).
 * ------------------------------------------------------------------------------------- */
int svlm_find_segment(const int64_t* ids, int64_t n, const int64_t* pattern, int pattern_len, int64_t end_token, int index,
                      int contain_lf, int64_t lf_token, int64_t* start_out, int64_t* end_out, int64_t* count_out);

/* ---------------------------------------------------------------------------------------
 * Host-buffer convenience entry points (the reference-facing plugin call with HOST data):
 * copy in, run, copy out, synchronise.  Used by bench.py's end-to-end leg and by callers that
 * keep the cache in host memory.
 * ------------------------------------------------------------------------------------- */
int svlm_kv_compact_rope_host(void* arena_host, int layers, int kv_heads, int head_dim, int capacity,
                              int src_buffer, const int32_t* runs_host, int n_runs, const int32_t* pos_host,
                              int pos_stride, const void* cos_table_host, const void* sin_table_host,
                              int table_rows, const int32_t* mrope_section_host);

#ifdef __cplusplus
}
#endif
#endif /* SVLM_B200_H_ */
