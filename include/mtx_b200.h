/*
 * mtx_b200.h -- C ABI of the B200-native decode step for the IndexTTS2-on-MaxText
 * text+audio-token transformer (HyperBlaze456/maxtext-indextts2).
 *
 * The reference has no FFI for this path: its decode step is one jax.jit executable
 * (MaxText/maxengine.py:868-936).  Each entry point below names the reference function
 * it replaces (file:line relative to the reference tree); INTEGRATION.md shows how a
 * maintainer registers them as XLA custom calls with jax.ffi, or binds them with ctypes.
 *
 * Conventions
 *   - every pointer is DEVICE memory owned by the caller unless the comment says "host";
 *   - matrices are row-major; activations and weights are bfloat16, statistics fp32,
 *     indices int32;
 *   - `stream` is a cudaStream_t; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 = ok, otherwise an MTX_ERR_* code; mtx_last_error() gives the text;
 *   - no exceptions cross the ABI; functions are re-entrant per engine, an engine is
 *     used from one host thread at a time (as MaxEngine is, SURVEY 8b "Threading");
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 *
 * "rows": the independent sequences one call advances by one token -- the decode slots
 * of a generate step, or the consecutive prompt positions of one prefill chunk.
 */
#ifndef MTX_B200_H_
#define MTX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTX_OK 0
#define MTX_ERR_ARG 1         /* bad shape / null pointer / unsupported size           */
#define MTX_ERR_CUDA 2        /* a CUDA runtime or driver call failed                  */
#define MTX_ERR_UNSUPPORTED 3 /* config outside the hot path (e.g. head_dim not 64/128) */

/* inference_utils.py:75-84 `sampling(..., algorithm)` */
#define MTX_SAMPLE_GREEDY 0
#define MTX_SAMPLE_WEIGHTED 1
#define MTX_SAMPLE_NUCLEUS 2
#define MTX_SAMPLE_TOPK 3

typedef struct mtx_engine mtx_engine; /* opaque */
typedef void* mtx_stream;             /* cudaStream_t */

/* Model shape: the derived keys of MaxText/pyconfig.py:576-582 plus the numerics keys of
 * configs/base.yml the step reads. */
typedef struct {
  int32_t num_layers;        /* num_decoder_layers */
  int32_t emb_dim;           /* E */
  int32_t num_q_heads;       /* Hq */
  int32_t num_kv_heads;      /* Hkv */
  int32_t head_dim;          /* D: 64 or 128 */
  int32_t mlp_dim;           /* M */
  int32_t vocab_size;        /* V rows of the logits matrix held by THIS process (a shard when vocab-parallel) */
  int32_t vocab_offset;      /* global id of row 0 of that matrix */
  int32_t max_prefill_len;   /* P = max_prefill_predict_length */
  int32_t max_target_len;    /* T = max_target_length; the AR ring has T - P rows */
  int32_t num_slots;         /* KV planes: decode slots plus any prefill staging planes */
  int32_t max_rows;          /* most rows one call advances (<= 256) */
  float rms_eps;             /* normalization_layer_epsilon */
  float rope_min_timescale;  /* rope_min_timescale */
  float rope_max_timescale;  /* rope_max_timescale */
  float attn_softcap;        /* attn_logits_soft_cap, 0 = off */
  float final_softcap;       /* final_logits_soft_cap, 0 = off */
  float logits_scale;        /* 1, or 1/sqrt(E) when logits_via_embedding && normalize_embedding_logits */
  int32_t logits_round_bf16; /* 1 unless logits_dot_in_fp32 (decoders.py:557,571) */
  int32_t embedding_rows;    /* rows of the embedding table: token ids are clamped into [0, embedding_rows) as jnp indexing does
                                (embeddings.py:154); 0 = do not clamp */
  int32_t kv_quant;          /* 0: bf16 KV cache.  2: as 1, with kv_quant_axis=heads_and_dkv (the reference's default): one scale per token over
                                all kv heads, stored once per head.  3 / 4: as 1 / 2 with kv_quant_dtype=fp8: float8_e4m3fn bytes,
                                value = e4m3(x * 448 / scale) (kvcache.py:38,86-88).  1: int8 with one fp32 scale per (token, kv head) -- quantize_kvcache=True,
                                kv_quant_dtype=int8, kv_quant_axis=dkv (inference/kvcache.py:36-90): decode_state.kq_cache / vq_cache /
                                k_scale / v_scale hold the decode cache, k_cache / v_cache are ONE bf16 staging plane for prefill */
  int32_t norm_scales_folded; /* 1: wqkv / w01 already carry the per-feature RMSNorm scales of their input (W' = W * diag(scale),
                                 folded at load time) and attn_norm / mlp_norm are all ones: the step skips the scale pass */
  /* decoder block (MaxText/layers/decoders.py:352-357).  0 = llama2 (layers/llama2.py:54-165); 1 = gemma3 (layers/gemma3.py:62-197):
   * RMSNorm of q and k over head_dim before RoPE (attentions.py:2246-2248), query * query_scalar after it (:2263-2265), RMSNorm of
   * the attention and MLP outputs before their residual adds (gemma3.py:142-170), GELU-gated MLP (tanh form, linears.py:460), five
   * sliding-window layers followed by one global layer (gemma3.py:36-48), the local layers with their own RoPE base
   * (attentions.py:2085-2088).  Requires norm_scales_folded, a bf16 cache and the weights q_norm / k_norm / post_attn_norm /
   * post_ffw_norm. */
  int32_t decoder_block;
  int32_t sliding_window;          /* sliding_window_size of the local layers (gemma3), 0 = none */
  float local_rope_max_timescale;  /* RoPE base of the local layers; <= 0: rope_max_timescale */
  float query_scalar;              /* query_pre_attn_scalar (gemma3.py:51-58); 0 or 1 = none */
  /* attention=paged (configs/base.yml:737-746; inference/paged_attention.py, inference/page_manager.py).  paged_num_pages > 0: the
   * decode cache is the page pools decode_state.k_pages / v_pages, addressed through decode_state.page_map etc. (the device copy of
   * PageState); k_cache / v_cache are then ONE bf16 staging plane [L, 1, Hkv, T, D] that prefill writes (as with kv_quant) and
   * mtx_insert_prefix copies a prefix into the pages of its group.  llama2 block, bf16 cache only. */
  int32_t paged_num_pages;           /* pagedattn_num_pages, 0 = contiguous cache */
  int32_t paged_tokens_per_page;     /* pagedattn_tokens_per_page: a power of two >= 8 */
  int32_t paged_max_pages_per_group; /* pagedattn_max_pages_per_group (>= ceil(T / tokens_per_page)) */
  int32_t paged_device_state;        /* 1: the whole PageState lives in decode_state (page_status, num_pages_used, has_active_page
                                        too) and every decode step begins with PageManager.update_decode_pages ON THE DEVICE
                                        (page_manager.py:332-412; in the step's first kernel), as the reference's jitted update does;
                                        0: the caller advances the page state and refreshes page_map etc. before the step */
} mtx_model_config;

/* Weights, repacked once at load time for K-major streaming (see DESIGN.md "Data layout").
 * All bfloat16.  Relation to the reference parameter tree ("Checkpoint / resume"):
 *   wqkv[l]  [ (Hq+2Hkv)*D, E ]  rows = query|key|value kernels transposed (attentions.py:1894-2030)
 *   wo[l]    [ E, Hq*D ]         out kernel transposed
 *   w01[l]   [ 2*M, E ]          wi_0 / wi_1 transposed and interleaved in groups of 16 rows:
 *                                rows 32g..32g+15 = wi_0[:, 16g..16g+15]^T, rows 32g+16..32g+31 = wi_1[...]^T
 *   wout[l]  [ E, M ]            wo kernel transposed (linears.py:425-476)
 *   logits   [ V, E ]            logits_dense kernel transposed, or the embedding table when tied
 * gemma3: attn_norm = pre_self_attention_norm/scale (gemma3.py:86-92), mlp_norm = mlp/mlp_layer_norm/scale (use_pre_norm).
 */
typedef struct {
  const void* embedding;  /* [V_full, E]   token_embedder/embedding as bf16 (embeddings.py:154) */
  const void* attn_norm;  /* [L, E]        pre_self_attention_layer_norm/scale */
  const void* wqkv;       /* [L, (Hq+2Hkv)*D, E] */
  const void* wo;         /* [L, E, Hq*D] */
  const void* mlp_norm;   /* [L, E]        mlp/mlp_layer_norm/scale */
  const void* w01;        /* [L, 2*M, E] */
  const void* wout;       /* [L, E, M] */
  const void* final_norm; /* [E]           decoder_norm/scale */
  const void* logits;     /* [V, E] */
  /* gemma3 block only (null otherwise) */
  const void* q_norm;         /* [L, D]  self_attention/query_norm/scale */
  const void* k_norm;         /* [L, D]  self_attention/key_norm/scale */
  const void* post_attn_norm; /* [L, E]  post_self_attention_norm/scale */
  const void* post_ffw_norm;  /* [L, E]  post_ffw_norm/scale */
} mtx_weights;

/* Decode state: the device-resident fields of the reference's decode_state dict
 * (maxengine.py:930-936, 1415-1427) in this implementation's layout.
 *   k_cache/v_cache  [L, num_slots, Hkv, T, D] bf16.  Rows [0,P) of a plane are the prefill
 *                    segment (cached_prefill_key/value), rows [P,T) the AR ring
 *                    (cached_ar_key/value); the reference's axis order keys select a layout,
 *                    this is the one with the sequence contiguous per (slot, head).
 *   prefill_len[s]   number of active prefill rows of slot s (== sum(cache_prefill_segment_id[s]))
 *   ar_lengths[s]    cached_ar_lengths; ar_index[0] is the shared ring index cache_ar_index
 */
typedef struct {
  void* k_cache;
  void* v_cache;
  int32_t* tokens;       /* [B]  decode_state["tokens"]           */
  int32_t* next_pos;     /* [B]  decode_state["next_pos"]         */
  int32_t* generated;    /* [B]  decode_state["generated_tokens"] */
  int32_t* prefill_len;  /* [num_slots] */
  int32_t* ar_lengths;   /* [num_slots] */
  int32_t* ar_index;     /* [1] */
  int32_t* result;       /* [B,3] ResultTokens.data: token, valid, length (maxengine.py:916-928) */
  float* log_prob;       /* [B] or NULL (return_log_prob)         */
  float* logits;         /* [B,V] fp32 or NULL: decode_state["logits"], only written when non-NULL */
  uint32_t* rng_state;   /* [4] sampler stream: {step, seed_lo, seed_hi, 0}; step advances once per decode step */
  /* kv_quant = 1 only (else NULL): the decode cache as unsigned bytes u = q + 128, q = clip(rint(x * 127.5 / scale), -128, 127),
   * [L, num_slots, Hkv, T, D] with scale = max|x| over the D dims of the row in k_scale / v_scale [L, num_slots, Hkv, T] fp32
   * (KVQuant.quantize, kvcache.py:76-90; dequantised value = q * scale / 127.5).  k_cache / v_cache are then [L, 1, Hkv, T, D]
   * bf16: the plane prefill writes; mtx_insert_prefix quantises a prefix into a slot. */
  void* kq_cache;
  void* vq_cache;
  float* k_scale;
  float* v_scale;
  /* attention=paged only (else NULL): the page pools [L, Hkv, num_pages, tokens_per_page, D] bf16 (PagedAttentionOp.key_pages /
   * value_pages, paged_attention.py:152-160, one pool for all layers) and the device copy of the PageState fields a step reads
   * (page_manager.py:49-91), one page group per decode slot.  paged_device_state = 0: the caller runs
   * PageManager.update_decode_pages BEFORE the step (maxengine.py:847-849) and refreshes these arrays; = 1: the step does it
   * itself on the full state (the three arrays at the end of this struct included). */
  void* k_pages;
  void* v_pages;
  int32_t* page_map;        /* [num_slots, max_pages_per_group] */
  int32_t* page_lengths;    /* [num_slots] sequence_lengths (this step's token included) */
  int32_t* active_page;     /* [num_slots] */
  int32_t* active_page_pos; /* [num_slots] active_page_position */
  /* paged_device_state = 1 only: the rest of PageState, updated in place by the step */
  int32_t* page_status;     /* [num_pages] */
  int32_t* num_pages_used;  /* [num_slots] */
  int32_t* has_active_page; /* [num_slots] 0 / 1 */
} mtx_decode_state;

/* ---- engine ---------------------------------------------------------------------------- */

/* Validates the shape and plans kernels (tile counts, split-K factors, workspace layout). */
int mtx_engine_create(const mtx_model_config* cfg, mtx_engine** out);
int mtx_engine_destroy(mtx_engine* e);

/* Scratch the caller must provide to mtx_engine_bind (activations, split-K partials, ...). */
size_t mtx_engine_workspace_bytes(const mtx_engine* e);

/* Attach weights, state and scratch; builds the TMA descriptors.  Replaces
 * MaxEngine.load_params + init_decode_state (maxengine.py:218, 1370) for the device side. */
int mtx_engine_bind(mtx_engine* e, const mtx_weights* w, const mtx_decode_state* s, void* workspace, size_t workspace_bytes);

/* inference_utils.py:66-84 arguments.  The random stream is keyed by decode_state.rng_state. */
int mtx_engine_set_sampling(mtx_engine* e, int strategy, int top_k, float nucleus_p, float temperature);

/* One autoregressive step for slots [0, rows): MaxEngine._generate_jit (maxengine.py:868-936).
 * Reads tokens/next_pos, appends K/V at the shared ring index, attends over the valid rows
 * of both cache segments, samples, and advances next_pos / generated / ar_index / ar_lengths. */
int mtx_decode_step(mtx_engine* e, int rows, mtx_stream stream);

/* Same step, replayed from a CUDA graph captured on first use (one graph per `rows`). */
int mtx_decode_step_graph(mtx_engine* e, int rows, mtx_stream stream);

/* The step as a serving loop calls it, with HOST buffers (pinned memory for the copies to be asynchronous): optionally copies this
 * step's input tokens host -> device (tokens_host [rows] or NULL to keep decode_state.tokens), replays the step's CUDA graph, and
 * copies ResultTokens.data [rows, 3] (and the log-probs [rows] when log_prob_host is non-NULL and return_log_prob is on) device ->
 * host, all enqueued on `stream`: the caller synchronises the stream once and reads result_host (offline_engine.py:612-614 copies
 * the result tokens to the host the same way). */
int mtx_decode_step_host(mtx_engine* e, int rows, const int32_t* tokens_host, int32_t* result_host, float* log_prob_host, mtx_stream stream);
/* With pinned buffers the copies are nodes of the step's graph (one graph per set of buffers: reuse them), so the call above is ONE
 * driver call.  This variant also waits for the stream: result_host is valid on return. */
int mtx_decode_step_host_sync(mtx_engine* e, int rows, const int32_t* tokens_host, int32_t* result_host, float* log_prob_host, mtx_stream stream);

/* Vocab-parallel logits: this process holds `vocab_size` rows of the logits matrix starting
 * at `vocab_offset`.  mtx_decode_step_candidates runs the whole step but, instead of committing a token,
 * writes this shard's winner per row to candidates[5][rows] (fp32: score, vocab id as int bits, its
 * logit, shard max, shard sum exp) -- the payload of ONE all-gather.  mtx_commit_candidates takes the
 * gathered [n_shards][5][rows] buffer, picks the global winner (lowest id on ties) and advances the state
 * exactly as mtx_decode_step does.  Replaces the all-gather of full fp32 logits that the reference's
 * sharding rules lower to (maxengine.py:894, configs/base.yml:351). */
int mtx_decode_step_candidates(mtx_engine* e, int rows, float* candidates, mtx_stream stream);
int mtx_commit_candidates(mtx_engine* e, int rows, const float* gathered, int n_shards, mtx_stream stream);
/* With top-k / nucleus sampling the payload is each shard's 64 best logits per row instead of its single winner:
 * candidates[rows][130] = 64 logits (descending), their 64 vocabulary ids (int bits), shard max, shard sum exp
 * (decode_state.logits must be set: the shard's logits are materialised).  mtx_candidate_floats() = floats per row of the
 * payload for the engine's current sampling strategy (5 or 130); `gathered` is [n_shards][5][rows] resp. [n_shards][rows][130].
 * top-k is exact for k <= 64; nucleus is exact when no shard's 64th logit reaches the cut-off, otherwise the nucleus is
 * truncated to the candidates and the row is counted: mtx_engine_counter(e, 0, &n) (synchronising read). */
size_t mtx_candidate_floats(const mtx_engine* e);
int mtx_engine_counter(mtx_engine* e, int which, long long* value);

/* inference_utils.sampling (+ log_prob_of_chosen_token) over MATERIALISED fp32 logits [rows, ld] (vocab entries per row) with
 * the engine's strategy, temperature and random stream (row r draws the noise of row row_offset + r of the current step; the
 * step counter is not advanced; row_offset < 0: the noise row of the engine's last prefill draw): token_out [rows],
 * log_prob_out [rows] or NULL.  The decode step fuses this into the logits
 * projection; this entry point serves callers that hold logits already (vocab-parallel prefill, tests). */
int mtx_sample_logits(mtx_engine* e, const float* logits, int rows, long long ld, int vocab, int row_offset, int32_t* token_out,
                      float* log_prob_out, mtx_stream stream);

/* Measurement aid: one eager decode step with a CUDA-event pair around every kernel launch
 * (on `stream`), summed per kernel class into class_ms[10] / class_launches[10] (host arrays):
 * 0 prepare, 1 rmsnorm, 2 qkv+rope+append, 3 attention, 4 out-proj, 5 mlp up, 6 mlp down,
 * 7 logits+sampling, 8 finalize, 9 the persistent step kernel (all layers + logits in one launch; used
 * for up to 64 rows, then classes 1-7 stay empty).  Synchronises the stream; the step really advances
 * the state. */
int mtx_profile_decode_step(mtx_engine* e, int rows, mtx_stream stream, float* class_ms, int32_t* class_launches);

/* `count` consecutive prompt positions [start_pos, start_pos+count) of one sequence into the
 * prefill segment of plane `slot`: MaxEngine._prefill_jit (maxengine.py:400-530) with
 * KVCache.kv_cache_prefill (kvcache.py:584-624), processed as rows of one step with causal
 * lengths.  If `sample_last`, the last position's logits are sampled into first_token[0]
 * (and copied to logits_out [V] fp32 when non-NULL); its log-probability goes to first_log_prob[0] when non-NULL
 * (return_log_prob, maxengine.py:503-520).  The sampler's noise for a prefill draw comes from its own row namespace (one per
 * prefill call), never from the rows of a decode step. */
int mtx_prefill_chunk(mtx_engine* e, const int32_t* tokens, int count, int start_pos, int slot, int sample_last,
                      int32_t* first_token, float* logits_out, float* first_log_prob, mtx_stream stream);

/* MaxEngine.insert (maxengine.py:1166-1192, _insert_jit :1045-1164): the first n_rows cache rows of a prefix (k_src / v_src
 * [L, Hkv, n_src_rows, D] bf16, what prefill left) into the prefill segment of decode slot `slot`, every layer and head in one
 * launch, and the slot's bookkeeping: prefill_len = n_rows, ar_lengths = 0, next_pos / generated / tokens as given.  The AR ring
 * and the shared ring index are not touched. */
int mtx_insert_prefix(mtx_engine* e, const void* k_src, const void* v_src, int n_rows, int n_src_rows, int slot, int next_pos, int generated,
                      int token, mtx_stream stream);

/* ---- single fused ops (the same kernels the step uses) ------------------------------------ */

/* RMSNorm (normalizations.py:57-69): out = bf16(bf16(x*rsqrt(mean(x^2)+eps)) * scale). x,out [rows,E]. */
int mtx_rmsnorm(const void* x, const void* scale, void* out, int rows, int emb_dim, float eps, mtx_stream stream);

/* DenseGeneral (linears.py:188-232): out[rows,N] = bf16(x[rows,K] . w[N,K]^T), fp32 accumulate on
 * tcgen05 tensor cores.  x must be allocated with rows rounded up to 16/32/64/128/256 (zero
 * padded).  splits (1, 2, 4, 8 or 16) cuts K over the CTAs of one thread-block cluster per 128-row
 * weight tile; the partial tiles are reduced over distributed shared memory. */
int mtx_linear(const void* x, const void* w, void* out, int rows, int n, int k, int splits, mtx_stream stream);

/* The two remaining fused blocks of a decoder layer as single ops, on the GEMM kernels of the per-kernel path.  `rows`
 * 1..256; activation inputs must be allocated with the rows rounded up to 16/32/64/128/256, as for mtx_linear.
 *
 * Attention.out + the residual add (attentions.py:2017-2030 `out_projection`, llama2.py:139-140):
 *   out[rows, E] = bf16(x + bf16(attn[rows, Hq*D] . wo[E, Hq*D]^T)), wo = mtx_weights.wo of one layer.  x and out may alias. */
int mtx_outproj_residual(const void* attn, const void* wo, const void* x, void* out, int rows, int emb_dim, int q_dim, mtx_stream stream);
/* MlpBlock with its pre-norm + the residual add (linears.py:425-476, llama2.py:150-163):
 *   n = RMSNorm(h) * norm_scale;  out[rows, E] = bf16(h + bf16(bf16(silu(n . w0) * (n . w1)) . wout^T)),
 *   w01 [2M, E] = mtx_weights.w01 of one layer (wi_0 / wi_1 rows interleaved in groups of 16), wout [E, M].
 * scratch: mtx_mlp_scratch_bytes(); h must be padded like an activation input (it is read as n's source only up to `rows`). */
size_t mtx_mlp_scratch_bytes(int rows, int emb_dim, int mlp_dim);
int mtx_mlp(const void* h, const void* norm_scale, const void* w01, const void* wout, void* out, int rows, int emb_dim, int mlp_dim, float eps,
            void* scratch, mtx_stream stream);

/* GQA decode attention over the valid rows of both cache segments: AttentionOp.__call__ in
 * autoregressive mode (attentions.py:1399-1466: two apply_attention_dot calls merged by
 * normalize_attention).  q,out [rows, Hq*D] bf16; k_cache/v_cache one layer [num_slots,Hkv,T,D];
 * row i reads plane[i], its first len0[i] rows and ring_len[i] ring rows starting at ring
 * offset ring_first[i] (ring base row P, ring length T-P).  `scratch` needs
 * mtx_attention_scratch_bytes(); no initialisation required. */
size_t mtx_attention_scratch_bytes(int rows, int num_kv_heads, int num_q_heads, int head_dim, int max_prefill_len,
                                   int max_target_len);
int mtx_decode_attention(const void* q, const void* k_cache, const void* v_cache, const int32_t* plane, const int32_t* len0,
                         const int32_t* ring_first, const int32_t* ring_len, void* out, int rows, int num_slots,
                         int num_q_heads, int num_kv_heads, int head_dim, int max_prefill_len, int max_target_len,
                         float softcap, void* scratch, mtx_stream stream);

/* One cache segment with prefix lengths, returning the UNNORMALISED output and the softmax statistics: the contract of
 * the reference's own pluggable decode-attention kernels, AttentionOp.gpu_ragged_attention / tpu_ragged_attention
 * (attentions.py:761-845: `(out [B,1,Hq,D], max [B,1,Hq,1], sum [B,1,Hq,1])`, sm_scale = 1), so it can sit behind
 * AttentionOp.apply_attention unchanged; the caller merges the segments with normalize_attention (attentions.py:1376-1397).
 *   q [rows, Hq*D] bf16; lengths [rows] valid rows of the segment (>= 1);
 *   k, v: seq_major = 1: [rows, seq_len, Hkv, D] (the reference's logical cache layout, CACHE_BATCH/SEQUENCE/HEADS/KV),
 *         seq_major = 0: [rows, Hkv, seq_len, D] (this library's layout, ar_cache_axis_order 0,2,1,3);
 *   out [rows, Hq*D] bf16 = sum_t exp(s_t - max) v_t;  out_max, out_sum [rows, Hq] fp32. */
size_t mtx_ragged_attention_scratch_bytes(int rows, int num_kv_heads, int num_q_heads, int head_dim, int seq_len);
int mtx_ragged_attention(const void* q, const void* k, const void* v, const int32_t* lengths, void* out, float* out_max, float* out_sum,
                         int rows, int seq_len, int num_q_heads, int num_kv_heads, int head_dim, int seq_major, float softcap,
                         void* scratch, mtx_stream stream);

/* ---- paged KV cache (attention=paged) as single ops ------------------------------------------------------------------------
 * PagedAttentionOp in autoregressive mode (inference/paged_attention.py:348-401): update_decode_step_pages (:446-471) then
 * paged_attention_v1_decode (:302-346; the kernel itself is jax.experimental.pallas.ops.tpu.paged_attention, third-party:
 * softmax(q . K[:len]) V[:len] over the group's pages, no scaling of q).
 *   k_pages, v_pages [Hkv, num_pages, tokens_per_page, D] bf16 (one layer); tokens_per_page a power of two >= 8;
 *   page_map [groups, max_pages_per_group], lengths / active_page / active_pos [groups] int32: PageState.page_map,
 *   sequence_lengths, active_page, active_page_position (page_manager.py:49-91); row r is page group r. */
/* k_new, v_new [rows, Hkv, D] -> pages[h, active_page[r], active_pos[r]] for every row (inactive groups write page 0). */
int mtx_paged_append(void* k_pages, void* v_pages, const void* k_new, const void* v_new, const int32_t* active_page,
                     const int32_t* active_pos, int rows, int num_kv_heads, int head_dim, int num_pages, int tokens_per_page,
                     mtx_stream stream);
/* q, out [rows, Hq*D] bf16; a row of length 0 gets no output (its `out` row is left as it was). */
size_t mtx_paged_attention_scratch_bytes(int rows, int num_kv_heads, int num_q_heads, int head_dim, int max_tokens);
int mtx_paged_attention(const void* q, const void* k_pages, const void* v_pages, const int32_t* lengths, const int32_t* page_map,
                        void* out, int rows, int num_q_heads, int num_kv_heads, int head_dim, int num_pages, int tokens_per_page,
                        int max_pages_per_group, float softcap, void* scratch, mtx_stream stream);
/* PageManager.update_decode_pages (page_manager.py:538-563, `_update_decode_pages_global` :332-412) on device arrays, in place: one more
 * token for every active group, a new page (the lowest free index >= 1, in group order, while free pages last) for the groups that
 * crossed a page boundary.  All arrays int32 (has_active_page 0 / 1); groups <= 256. */
int mtx_page_update_decode(int32_t* page_status, int32_t* page_map, int32_t* num_pages_used, int32_t* sequence_lengths, int32_t* active_page,
                           const int32_t* has_active_page, int32_t* active_page_position, int num_pages, int groups, int max_pages_per_group,
                           int tokens_per_page, mtx_stream stream);

/* MaxEngine._insert_jit's `_copy_paged` (maxengine.py:1104-1131): the first n_tokens rows of a prefix (k_src / v_src
 * [layers, Hkv, n_src_rows, D], i.e. the prefix's pages read as rows) into the pool pages page_map_row[i] (device, the group's row
 * of PageState.page_map); pools [layers, Hkv, num_pages, tokens_per_page, D]. */
int mtx_paged_insert(void* k_pages, void* v_pages, const void* k_src, const void* v_src, const int32_t* page_map_row, int layers,
                     int num_kv_heads, int head_dim, int n_src_rows, int n_tokens, int num_pages, int tokens_per_page, mtx_stream stream);

/* Attention.query/key/value projections + RotaryEmbedding + KVCache append of one decode step (attentions.py:1894-2030,
 * 2236-2265; embeddings.py:277-315; kvcache.py:626-718) as ONE GEMM with a fused epilogue:
 *   n [rows padded to 16/32/64/128/256, E] bf16 normalised activations; wqkv [(Hq+2Hkv)*D, E] (mtx_weights.wqkv of one layer);
 *   pos [rows] RoPE positions; row i's key / value go to row write_row[i] (< 0: skip) of plane plane[i] of the layer's cache
 *   [planes, Hkv, rows_per_plane, D]; q_out [rows, Hq*D] rotated queries.  scratch: mtx_qkv_rope_append_scratch_bytes(). */
size_t mtx_qkv_rope_append_scratch_bytes(int rows, int head_dim);
int mtx_qkv_rope_append(const void* n, const void* wqkv, const int32_t* pos, const int32_t* plane, const int32_t* write_row, void* q_out,
                        void* k_cache, void* v_cache, int rows, int emb_dim, int num_q_heads, int num_kv_heads, int head_dim,
                        int rows_per_plane, float rope_min_timescale, float rope_max_timescale, void* scratch, mtx_stream stream);

/* Re-attach the decode-state buffers without repacking anything else (an XLA FFI handler receives its operand buffers anew
 * at every call; see csrc/mtx_jax_ffi.cc).  A no-op when nothing moved. */
int mtx_engine_rebind_state(mtx_engine* e, const mtx_decode_state* s);

/* ---- misc ---------------------------------------------------------------------------------- */

/* Measurement aid: grid-barrier timeline of the persistent step kernel.  While a device buffer of
 * mtx_step_trace_words() 64-bit words is installed with mtx_debug_set_trace, every CTA records
 * %globaltimer at the arrival at and the release from each grid barrier (see tools/mega_trace.py). */
size_t mtx_step_trace_words(const mtx_engine* e); /* sized for the engine's grid and layer count (NULL: 148 CTAs, 24 layers) */

/* 1 when the library was built with jaxlib's headers and exports the XLA FFI handler symbols of csrc/mtx_jax_ffi.cc
 * (MtxRaggedAttention, MtxDecodeAttention, MtxQkvRopeAppend, MtxOutprojResidual, MtxMlp, MtxDecodeStep, MtxPagedAppend,
 * MtxPagedAttention), else 0. */
int mtx_jax_ffi_available(void);

const char* mtx_last_error(void);
/* "sm_100a" build tag, so a caller can check what it loaded. */
const char* mtx_build_info(void);
/* Debug aid: when non-NULL (device memory, 256 x int64), the GEMM and attention kernels record
 * a clock64 timeline of their first CTA into it.  NULL (the default) switches it off.  The pointer is read when a step is
 * enqueued: CUDA graphs captured earlier keep the value they were captured with (use mtx_decode_step for traced steps). */
void mtx_debug_set_trace(void* device_buffer);
/* Debug aid: when non-NULL (device memory, 3004+ x uint64, zeroed), every kernel's first CTA appends
 * (kind, start ns, end ns) by %globaltimer; [0] counts the entries. */
int mtx_debug_set_timeline(void* device_buffer);
/* Kernels launched by this library since load (all engines); used by bench.py's gpu_launches. */
uint64_t mtx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MTX_B200_H_ */
