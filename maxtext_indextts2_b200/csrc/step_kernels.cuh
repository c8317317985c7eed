// Small kernels around the GEMMs and attention of one step: row descriptors, embedding
// gather + RMSNorm, sampler finalisation and the decode-state bookkeeping.
#pragma once

#include "attention.cuh"
#include "common.cuh"
#include "paged_kv.cuh"

namespace mtx {

// Per-row descriptors consumed by the QKV epilogue and the attention kernel.
struct RowDesc {
  int* token;       // [max_rows] token id fed to the embedding
  int* pos;         // [max_rows] RoPE position
  int* plane;       // [max_rows] KV plane
  int* write_row;   // [max_rows] cache row the new K/V go to
  int* len0;        // [max_rows] valid prefill rows (including a row appended there this step)
  int* ring_first;  // [max_rows]
  int* ring_len;    // [max_rows] valid ring rows (including the row appended this step)
  float2* rope_cs;  // [max_rows, D/2] (cos, sin), bf16-rounded
  int* work_items;  // attention work list
  int* work_count;
  int* iota;        // [max_rows] r            } "plane" / "write row" of a [rows, Hkv, D] scratch matrix for the QKV epilogue
  int* tmp_row;     // [max_rows] 0, or -1 when the row appends nothing } (int8 cache with one scale per token: kv_quant_rows_kernel)
  // sliding-window layers (gemma3 local attention; PrepareArgs.window > 0): the valid rows INSIDE the window's two cache
  // sub-ranges (AttnParams.skip0 / ring_off / ring_size), their work list, and the RoPE table of the local base
  int* len0_w;
  int* ring_first_w;
  int* ring_len_w;
  int* skip_w;  // per-row first visible cache row of a prefill position (mode 1: position window), 0 for decode rows
  float2* rope_cs_w;
  int* work_items_w;
  int* work_count_w;
};

struct PrepareArgs {
  // decode (mode 0): KVCache.kv_cache_autoregressive bookkeeping, inference/kvcache.py:738-795
  const int* tokens;       // [B] decode_state["tokens"]
  const int* next_pos;     // [B]
  const int* prefill_len;  // [num_slots]
  const int* ar_lengths;   // [num_slots]
  const int* ar_index;     // [1]
  // prefill chunk (mode 1): maxengine.py:444-458 positions / sequence indicator
  const int* chunk_tokens;  // [count]
  int start_pos, slot;
  int mode, rows, P, T, D;
  int tiles_per_item;  // attention work-item size (attn_tiles_per_item)
  int emb_rows;        // token ids are clamped into [0, emb_rows) (0: off)
  const float* rope_timescale;  // [D/2] min * (max/min)^(2i/D), embeddings.py:270-275
  int window;                     // sliding_window_size of the local layers (0: the model has none)
  const float* rope_timescale_w;  // [D/2] the same table for local_rope_max_timescale (attentions.py:2085-2088)
  // attention=paged, decode: [B] PageState.sequence_lengths after update_decode_pages (page_manager.py:332-412), else null.  The
  // row then attends tokens [0, length) of its page group (no prefill segment / ring).  Its key / value go to
  // pages[h, active_page, active_page_position] (update_decode_step_pages, paged_attention.py:446-471): with the pool of a layer
  // read as ONE plane of Hkv heads x (num_pages * tokens_per_page) rows that is row active_page * tokens_per_page + position of
  // plane 0, so the QKV epilogues append to a page exactly as they append to a dense cache.
  const int* page_lengths;
  const int* active_page;
  const int* active_pos;
  int tokens_per_page;
  // paged_device_state: the step itself runs PageManager.update_decode_pages first (paged_kv.cuh), on these arrays
  int page_update;
  PageStateDev page_state;
  // persistent step kernel (step_persistent.cuh): counters reset here, attention tile partition
  unsigned int* grid_bar;  // grid-barrier arrival counter
  int* tile_prefix;        // [rows + 1] exclusive prefix of the per-row 64-row tile counts
  int* attn_info;          // [0] CTAs that get attention work, [1] total tiles over all kv heads, [2] 1 = one pair per CTA
  int hkv;
  int pk_ctas;             // CTAs of the persistent grid
  int pk_max_parts;        // most attention warps of other CTAs that may hold tiles of one (row, kv head)
  int pk_warps;            // attention warps per CTA
  int pk_pair_mode_tiles;  // longest pair (in 64-row tiles) for which a CTA takes a whole pair
};

// One block of 256 threads.  Fills the row descriptors, the RoPE table
// (embeddings.py:304-307: sin/cos of position / timescale, cast to bf16) and the attention work list.
__global__ void prepare_rows_kernel(const PrepareArgs a, const RowDesc rd) {
  __shared__ int s_off[257];
  __shared__ int s_pos[256];
  __shared__ int s_nt[256];
  const int tl = timeline_begin(0);
  griddep_launch_dependents();
  griddep_wait();
  const int tid = threadIdx.x;
  const int R = a.T - a.P;
  int chunks = 0, chunks_w = 0;
  if (a.page_update) {  // (uniform)
    __shared__ int s_free[256];
    page_update_decode(a.page_state, s_off, s_free);  // (s_off: scratch until the work list below)
    __syncthreads();
  }
  if (tid < a.rows) {
    int token, pos, plane, wr, l0, rf, rl;
    if (a.mode == 0) {
      const int idx = a.ar_index[0];
      token = a.tokens[tid];
      pos = a.next_pos[tid];
      plane = tid;
      wr = a.P + idx;  // all rows append at the shared ring index (kvcache.py:696-701)
      l0 = a.prefill_len[tid];
      const int n = a.ar_lengths[tid] + 1;  // rows marked active since insert, this one included
      rl = n < R ? n : R;
      rf = ((idx + 1 - rl) % R + R) % R;
      if (a.page_lengths != nullptr) {
        l0 = a.page_lengths[tid];
        l0 = l0 < a.T ? l0 : a.T;  // (the page manager keeps counting past max_target_length; a group holds at most T tokens)
        rl = rf = 0;
        plane = 0;
        wr = a.active_page[tid] * a.tokens_per_page + a.active_pos[tid];
      }
    } else {
      token = a.chunk_tokens[tid];
      pos = a.start_pos + tid;
      plane = a.slot;
      wr = a.start_pos + tid;
      l0 = a.start_pos + tid + 1;  // causal: this position and everything before it
      rf = 0;
      rl = 0;
    }
    if (a.emb_rows > 0) token = token < 0 ? 0 : (token >= a.emb_rows ? a.emb_rows - 1 : token);
    rd.token[tid] = token;
    rd.pos[tid] = pos;
    rd.plane[tid] = plane;
    rd.write_row[tid] = wr;
    rd.len0[tid] = l0;
    rd.ring_first[tid] = rf;
    rd.ring_len[tid] = rl;
    if (rd.iota != nullptr) {
      rd.iota[tid] = tid;
      rd.tmp_row[tid] = wr >= 0 ? 0 : -1;
    }
    s_nt[tid] = attn_num_tiles(l0, rf, rl, R);
    chunks = (s_nt[tid] + a.tiles_per_item - 1) / a.tiles_per_item;
    s_pos[tid] = pos;
    if (a.window > 0) {
      // attentions.py:600-602,624-631 in AUTOREGRESSIVE mode: the window covers the last `window` CACHE INDICES of each segment
      // (next_pos = kv_seq_len - 1): prefill rows [s0, P) and ring indices [o, R).  The valid ring rows inside [o, R) are again a
      // circular range of the sub-ring of size R - o: [max(rf, o), min(rf + rl, R)) followed, when the valid range wraps past
      // index o, by [o, rf + rl - R).  (Prefill chunks, mode 1, go through prefill_attn_kernel, which masks by position.)
      const int s0 = a.P > a.window ? a.P - a.window : 0, o = R > a.window ? R - a.window : 0;
      const int l0w = l0 > s0 ? l0 - s0 : 0;
      const int a_lo = rf > o ? rf : o, a_hi = rf + rl < R ? rf + rl : R;
      const int alen = a_hi > a_lo ? a_hi - a_lo : 0;
      const int blen = rf + rl - R > o ? rf + rl - R - o : 0;
      const int rfw = alen > 0 ? a_lo - o : 0, rlw = alen + blen;
      // mode 1 (a prompt position run as a decode row: head_dim 256, which prefill_attn_kernel does not take): the window is by
      // position, keys (pos - window, pos] = cache rows [l0 - min(l0, window), l0)
      const int l0p = l0 < a.window ? l0 : a.window;
      rd.len0_w[tid] = a.mode == 0 ? l0w : l0p;
      rd.skip_w[tid] = a.mode == 0 ? 0 : l0 - l0p;
      rd.ring_first_w[tid] = rfw;
      rd.ring_len_w[tid] = rlw;
      chunks_w = (attn_num_tiles(a.mode == 0 ? l0w : l0p, rfw, rlw, R - o) + a.tiles_per_item - 1) / a.tiles_per_item;
    }
  }
  s_off[tid + 1] = chunks;
  if (tid == 0) s_off[0] = 0;
  __syncthreads();
  {  // RoPE table, all threads: (row, frequency) pairs
    const int half = a.D / 2;
    for (int idx = tid; idx < a.rows * half; idx += 256) {
      const int r = idx / half, i = idx - r * half;
      const float ang = float(s_pos[r]) / a.rope_timescale[i];
      rd.rope_cs[idx] = make_float2(bf16r(cosf(ang)), bf16r(sinf(ang)));
      if (a.window > 0) {
        const float angw = float(s_pos[r]) / a.rope_timescale_w[i];
        rd.rope_cs_w[idx] = make_float2(bf16r(cosf(angw)), bf16r(sinf(angw)));
      }
    }
  }
  if (a.grid_bar == nullptr) {  // the per-kernel path's attention work list (the persistent kernel cuts its own)
    if (tid == 0)
      for (int i = 1; i <= 256; ++i) s_off[i] += s_off[i - 1];
    __syncthreads();
    if (tid < a.rows) {
      const int base = s_off[tid];
      for (int c = 0; c < chunks; ++c) rd.work_items[base + c] = (tid << 16) | c;
    }
    if (tid == 0) *rd.work_count = s_off[a.rows];
    if (a.window > 0) {  // the same list for the sliding-window layers
      __syncthreads();
      s_off[tid + 1] = chunks_w;
      if (tid == 0) s_off[0] = 0;
      __syncthreads();
      if (tid == 0)
        for (int i = 1; i <= 256; ++i) s_off[i] += s_off[i - 1];
      __syncthreads();
      if (tid < a.rows) {
        const int base = s_off[tid];
        for (int c = 0; c < chunks_w; ++c) rd.work_items_w[base + c] = (tid << 16) | c;
      }
      if (tid == 0) *rd.work_count_w = s_off[a.rows];
    }
  }
  if (a.grid_bar != nullptr) {
    if (tid == 0) {
      *a.grid_bar = 0u;
      int acc = 0, nt_max = 1;
      for (int r = 0; r < a.rows; ++r) {
        a.tile_prefix[r] = acc;
        acc += s_nt[r];
        nt_max = s_nt[r] > nt_max ? s_nt[r] : nt_max;
      }
      a.tile_prefix[a.rows] = acc;
      // Every active CTA gets q = total / nc tiles (rounded either way) and each of its pk_warps attention warps a
      // fifth of them.  The warps that hold the tiles of a pair outside the pair's first CTA each own one slot of
      // the pair's exchange workspace: at most nt - 1 of them, and at most (nt - 1) / floor(q / warps) + 2.
      const int total = acc * a.hkv;
      int nc = total < a.pk_ctas ? total : a.pk_ctas;
      if (nt_max - 1 > a.pk_max_parts) {
        const int f = (nt_max - 1 + a.pk_max_parts - 3) / (a.pk_max_parts - 2);  // tiles per warp needed
        const int by_parts = total / (f * a.pk_warps);
        nc = by_parts < nc ? by_parts : nc;
      }
      if (nc < 1) nc = 1;
      // Few, short pairs: one whole (row, kv head) pair per CTA.  The CTAs are unevenly loaded, but no pair is cut
      // between CTAs, so no partial result crosses the L2 and the phase ends with the on-chip merge.
      const int pairs = a.rows * a.hkv;
      int mode = 0;
      if (pairs <= a.pk_ctas && nt_max <= a.pk_pair_mode_tiles) {
        mode = 1;
        nc = pairs;
      }
      a.attn_info[0] = nc;
      a.attn_info[1] = total;
      a.attn_info[2] = mode;
    }
  }
  timeline_end(tl);
}

// RMSNorm (normalizations.py:57-69), optionally fused with the embedding gather
// (embeddings.py:154: embedding.astype(bf16)[tokens]).  One CTA per row.
//   y   = bf16(x32 * rsqrt(mean(x32^2) + eps));  out = bf16(y * bf16(scale))
template <bool EMBED>
__global__ void __launch_bounds__(128)
rmsnorm_kernel(const bf16* __restrict__ x_in, const int* __restrict__ tokens, const bf16* __restrict__ embedding,
               const bf16* __restrict__ scale, bf16* __restrict__ x_out, bf16* __restrict__ n_out, int E, float eps) {
  __shared__ float s_part[4];
  const int tl = timeline_begin(1);
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x;
  const bf16* src = EMBED ? embedding + (long long)tokens[r] * E : x_in + (long long)r * E;
  const int nvec = E / 8;
  float ss = 0.0f;
  for (int i = threadIdx.x; i < nvec; i += 128) {
    const uint4 raw = *reinterpret_cast<const uint4*>(src + i * 8);
    if (EMBED) *reinterpret_cast<uint4*>(x_out + (long long)r * E + i * 8) = raw;
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(w[j]), b = bf16_hi(w[j]);
      ss += a * a + b * b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
  __syncthreads();
  const float total = s_part[0] + s_part[1] + s_part[2] + s_part[3];
  const float rstd = 1.0f / sqrtf(total / float(E) + eps);
  for (int i = threadIdx.x; i < nvec; i += 128) {
    const uint4 raw = *reinterpret_cast<const uint4*>(src + i * 8);
    const uint4 sc = *reinterpret_cast<const uint4*>(scale + i * 8);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    const uint32_t s[4] = {sc.x, sc.y, sc.z, sc.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float y0 = bf16r(bf16_lo(w[j]) * rstd), y1 = bf16r(bf16_hi(w[j]) * rstd);
      o[j] = pack_bf16x2(y0 * bf16_lo(s[j]), y1 * bf16_hi(s[j]));
    }
    *reinterpret_cast<uint4*>(n_out + (long long)r * E + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  timeline_end(tl);
}

// Embedding gather (embeddings.py:154) that also leaves the row's sum of squares as the first RMSNorm's statistic
// (ss[0 * pitch + r] = total, the other tile partials of the row zero): the fused-norm path of gemm_rows.cuh.
__global__ void __launch_bounds__(128)
embed_gather_ss_kernel(const int* __restrict__ tokens, const bf16* __restrict__ embedding, bf16* __restrict__ x_out, float* __restrict__ ss,
                       int ss_tiles, int ss_pitch, int E) {
  __shared__ float s_part[4];
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x;
  const bf16* src = embedding + (long long)tokens[r] * E;
  float acc = 0.0f;
  for (int i = threadIdx.x; i < E / 8; i += 128) {
    const uint4 raw = *reinterpret_cast<const uint4*>(src + i * 8);
    *reinterpret_cast<uint4*>(x_out + (long long)r * E + i * 8) = raw;
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(w[j]), b = bf16_hi(w[j]);
      acc += a * a + b * b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < ss_tiles) ss[threadIdx.x * ss_pitch + r] = threadIdx.x == 0 ? s_part[0] + s_part[1] + s_part[2] + s_part[3] : 0.0f;
}

// MaxEngine.insert (maxengine.py:1045-1164; KVCache prefill segment, kvcache.py:584-624): the first `n` cache rows of a prefix
// ([L, Hkv, n_src, D], as MaxEngine.prefill returned it) into plane `slot` of the decode cache [L, planes, Hkv, T, D], K and V of
// every layer and head in ONE launch of 16-byte copies, plus the slot's bookkeeping (the AR ring data and the shared ring index
// stay untouched: maxengine.py:1060-1067).
struct InsertArgs {
  const bf16* k_src;
  const bf16* v_src;
  bf16* k_cache;
  bf16* v_cache;
  int L, hkv, D, T, planes, n, n_src, slot;
  int next_pos, generated, token;
  int* prefill_len;
  int* ar_lengths;
  int* next_pos_out;
  int* generated_out;
  int* tokens_out;
};
__global__ void __launch_bounds__(256) insert_prefix_kernel(const InsertArgs a) {
  griddep_launch_dependents();
  griddep_wait();
  const int vec_per_row = a.D / 8;                              // uint4 per cache row
  const long long per_lh = (long long)a.n * vec_per_row;        // vectors per (layer, head)
  const long long total = per_lh * a.L * a.hkv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long lh = i / per_lh, rem = i - lh * per_lh;
    const int l = int(lh / a.hkv), h = int(lh - (long long)l * a.hkv);
    const long long src = ((long long)(l * a.hkv + h) * a.n_src) * a.D + rem * 8;
    const long long dst = (((long long)(l * a.planes + a.slot) * a.hkv + h) * a.T) * a.D + rem * 8;
    *reinterpret_cast<uint4*>(a.k_cache + dst) = *reinterpret_cast<const uint4*>(a.k_src + src);
    *reinterpret_cast<uint4*>(a.v_cache + dst) = *reinterpret_cast<const uint4*>(a.v_src + src);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.prefill_len[a.slot] = a.n;
    a.ar_lengths[a.slot] = 0;
    a.next_pos_out[a.slot] = a.next_pos;
    a.generated_out[a.slot] = a.generated;
    a.tokens_out[a.slot] = a.token;
  }
}

// The same insert into an int8 decode cache (mtx_model_config.kv_quant): one warp per (layer, kv head, row) quantises the 64 dims
// of the bf16 prefix row (KVQuant.quantize, kvcache.py:76-90: scale = max|x|, q = rint(x * 127.5 / scale) clipped to int8).
struct InsertQ8Args {
  InsertArgs base;
  uint8_t* kq_cache;
  uint8_t* vq_cache;
  float* k_scale;
  float* v_scale;
  int shared_scale;  // kv_quant_axis heads_and_dkv: one scale per token, over all kv heads (stored once per head)
  int fp8;           // kv_quant_dtype fp8 (float8_e4m3fn bytes) instead of int8
};
// kv_quant_axis "heads_and_dkv" (kvcache.py:69-72, the reference's default): scale = max|x| over the heads AND the head dims of a
// token.  One warp per (layer, row, K|V) walks the kv heads twice.
__global__ void __launch_bounds__(256) insert_prefix_q8_shared_kernel(const InsertQ8Args q) {
  const InsertArgs& a = q.base;
  griddep_launch_dependents();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long total = (long long)a.L * a.n * 2;
  for (long long i = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); i < total; i += warps) {
    const int which = int(i & 1);
    const long long j = i >> 1;
    const int row = int(j % a.n), l = int(j / a.n);
    const bf16* base = (which ? a.v_src : a.k_src);
    float mx = 0.0f;
    for (int h = 0; h < a.hkv; ++h) {
      const uint32_t pk = *reinterpret_cast<const uint32_t*>(base + ((long long)(l * a.hkv + h) * a.n_src + row) * 64 + 2 * lane);
      mx = fmaxf(mx, fmaxf(fabsf(bf16_lo(pk)), fabsf(bf16_hi(pk))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float inv = mx > 0.0f ? kv_quant_max(q.fp8 != 0) / mx : 0.0f;
    for (int h = 0; h < a.hkv; ++h) {
      const uint32_t pk = *reinterpret_cast<const uint32_t*>(base + ((long long)(l * a.hkv + h) * a.n_src + row) * 64 + 2 * lane);
      const long long drow = ((long long)(l * a.planes + a.slot) * a.hkv + h) * a.T + row;
      uint8_t* dst = (which ? q.vq_cache : q.kq_cache) + drow * 64;
      *reinterpret_cast<uint16_t*>(dst + 2 * lane) = uint16_t(kv_quant_pair(bf16_lo(pk), bf16_hi(pk), inv, q.fp8 != 0));
      if (lane == 0) (which ? q.v_scale : q.k_scale)[drow] = mx;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.prefill_len[a.slot] = a.n;
    a.ar_lengths[a.slot] = 0;
    a.next_pos_out[a.slot] = a.next_pos;
    a.generated_out[a.slot] = a.generated;
    a.tokens_out[a.slot] = a.token;
  }
}

// The decode step's appended rows for kv_quant_axis "heads_and_dkv": the QKV epilogue leaves the (rotated) keys and values of
// the step's rows as bf16 in [rows, Hkv, 64] scratch matrices; one CTA per (row, K|V), one warp per kv head, takes the
// token's scale over all heads, quantises (KVQuant.quantize, kvcache.py:76-90) and appends to the int8 cache.
struct KvQuantRowsArgs {
  const bf16* k_tmp;  // [rows, Hkv, 64]
  const bf16* v_tmp;
  const int* plane;
  const int* write_row;
  uint8_t* kq_cache;  // this layer's [planes, Hkv, t_alloc, 64]
  uint8_t* vq_cache;
  float* k_scale;     // this layer's [planes, Hkv, t_alloc]
  float* v_scale;
  int hkv, t_alloc;
  int fp8;
};
__global__ void __launch_bounds__(1024) kv_quant_rows_kernel(const KvQuantRowsArgs a) {
  __shared__ float s_mx[32];
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x, which = blockIdx.y, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = a.write_row[r];
  const uint32_t pk = *reinterpret_cast<const uint32_t*>((which ? a.v_tmp : a.k_tmp) + ((long long)r * a.hkv + h) * 64 + 2 * lane);
  const float x0 = bf16_lo(pk), x1 = bf16_hi(pk);
  float mx = fmaxf(fabsf(x0), fabsf(x1));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_mx[h] = mx;
  __syncthreads();
  mx = 0.0f;
  for (int w = 0; w < a.hkv; ++w) mx = fmaxf(mx, s_mx[w]);
  if (wr < 0) return;
  const float inv = mx > 0.0f ? kv_quant_max(a.fp8 != 0) / mx : 0.0f;
  const long long drow = ((long long)a.plane[r] * a.hkv + h) * a.t_alloc + wr;
  uint8_t* dst = (which ? a.vq_cache : a.kq_cache) + drow * 64;
  *reinterpret_cast<uint16_t*>(dst + 2 * lane) = uint16_t(kv_quant_pair(x0, x1, inv, a.fp8 != 0));
  if (lane == 0) (which ? a.v_scale : a.k_scale)[drow] = mx;
}

__global__ void __launch_bounds__(256) insert_prefix_q8_kernel(const InsertQ8Args q) {
  const InsertArgs& a = q.base;
  griddep_launch_dependents();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long total = (long long)a.L * a.hkv * a.n * 2;  // K and V rows
  for (long long i = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); i < total; i += warps) {
    const int which = int(i & 1);
    const long long j = i >> 1;
    const int row = int(j % a.n);
    const long long lh = j / a.n;
    const int l = int(lh / a.hkv), h = int(lh - (long long)l * a.hkv);
    const bf16* src = (which ? a.v_src : a.k_src) + ((long long)(l * a.hkv + h) * a.n_src + row) * 64;
    const uint32_t pk = *reinterpret_cast<const uint32_t*>(src + 2 * lane);
    const float x0 = bf16_lo(pk), x1 = bf16_hi(pk);
    float mx = fmaxf(fabsf(x0), fabsf(x1));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float inv = mx > 0.0f ? kv_quant_max(q.fp8 != 0) / mx : 0.0f;
    const long long drow = ((long long)(l * a.planes + a.slot) * a.hkv + h) * a.T + row;
    uint8_t* dst = (which ? q.vq_cache : q.kq_cache) + drow * 64;
    *reinterpret_cast<uint16_t*>(dst + 2 * lane) = uint16_t(kv_quant_pair(x0, x1, inv, q.fp8 != 0));
    if (lane == 0) (which ? q.v_scale : q.k_scale)[drow] = mx;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.prefill_len[a.slot] = a.n;
    a.ar_lengths[a.slot] = 0;
    a.next_pos_out[a.slot] = a.next_pos;
    a.generated_out[a.slot] = a.generated;
    a.tokens_out[a.slot] = a.token;
  }
}

struct FinalizeArgs {
  const float* part_score;  // [rows, n_tiles]
  const int* part_idx;
  const float* part_raw;
  const float* part_max;
  const float* part_sum;
  int n_tiles, rows;
  long long stride_r, stride_t;  // element (row r, tile t) of every part_* array is at [r * stride_r + t * stride_t]
  float* cand_out;               // mode 2: [5][rows] = best score, vocab id (int bits), its logit, max, sum exp
  int mode;  // 0 = decode step: advance the state; 1 = prefill: only emit the last row's token;
             // 2 = vocab-parallel shard: emit this shard's candidate per row, do not advance;
             // 3 = mtx_sample_logits: first_token[r] / log_prob[r] for every row, nothing advances
  int have_lse;  // part_max / part_sum are valid
  // decode state (maxengine.py:913-936)
  int* tokens;
  int* next_pos;
  int* generated;
  int* ar_lengths;
  int* ar_index;
  int* result;        // [rows, 3]
  float* log_prob;    // [rows] or null
  uint32_t* rng_state;
  int num_slots, R;
  // prefill
  int* first_token;
};

// One CTA (kFinalizeThreads threads) per row: reduce the per-tile candidates (first maximum wins, as
// jnp.argmax), compute log-softmax of the choice (inference_utils.py:55-63), then the
// bookkeeping of maxengine.py:913-914 and kvcache.py:778-779.  The candidates of a row are strided in memory
// (one L2 sector each): many threads and four independent candidates per thread keep the kernel at a few L2
// round trips.
constexpr int kFinalizeThreads = 512;
__global__ void __launch_bounds__(kFinalizeThreads) finalize_kernel(const FinalizeArgs a) {
  constexpr int kWarps = kFinalizeThreads / 32;
  __shared__ float s_score[kWarps], s_raw[kWarps], s_max[kWarps], s_sum[kWarps];
  __shared__ int s_idx[kWarps];
  const int tl = timeline_begin(8);
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float score = -INFINITY, raw = -INFINITY, mx = -INFINITY, sum = 0.0f;
  int idx = 0x7fffffff;
  for (int t0 = threadIdx.x; t0 < a.n_tiles; t0 += 4 * kFinalizeThreads) {
    float s4[4], r4[4], m4[4], u4[4];
    int i4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = t0 + q * kFinalizeThreads;
      const long long o = (long long)r * a.stride_r + (long long)(t < a.n_tiles ? t : t0) * a.stride_t;
      s4[q] = a.part_score[o];
      i4[q] = a.part_idx[o];
      r4[q] = a.part_raw[o];
      m4[q] = a.have_lse ? a.part_max[o] : -INFINITY;
      u4[q] = a.have_lse ? a.part_sum[o] : 0.0f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (t0 + q * kFinalizeThreads >= a.n_tiles) continue;
      if (s4[q] > score || (s4[q] == score && i4[q] < idx)) { score = s4[q]; idx = i4[q]; raw = r4[q]; }
      if (a.have_lse) {
        const float mn = fmaxf(mx, m4[q]);
        if (mn > -INFINITY) sum = sum * expf(mx - mn) + u4[q] * expf(m4[q] - mn);
        mx = mn;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float s2 = __shfl_xor_sync(0xffffffffu, score, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    const float r2 = __shfl_xor_sync(0xffffffffu, raw, o);
    if (s2 > score || (s2 == score && i2 < idx)) { score = s2; idx = i2; raw = r2; }
    const float m2 = __shfl_xor_sync(0xffffffffu, mx, o);
    const float u2 = __shfl_xor_sync(0xffffffffu, sum, o);
    const float mn = fmaxf(mx, m2);
    if (mn > -INFINITY) sum = sum * expf(mx - mn) + u2 * expf(m2 - mn);
    mx = mn;
  }
  if (lane == 0) { s_score[warp] = score; s_idx[warp] = idx; s_raw[warp] = raw; s_max[warp] = mx; s_sum[warp] = sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kWarps; ++w) {
      if (s_score[w] > score || (s_score[w] == score && s_idx[w] < idx)) { score = s_score[w]; idx = s_idx[w]; raw = s_raw[w]; }
      const float mn = fmaxf(mx, s_max[w]);
      if (mn > -INFINITY) sum = sum * expf(mx - mn) + s_sum[w] * expf(s_max[w] - mn);
      mx = mn;
    }
    const float logp = raw - (mx + logf(sum));
    if (a.mode == 2) {
      a.cand_out[0 * a.rows + r] = score;
      a.cand_out[1 * a.rows + r] = __int_as_float(idx);
      a.cand_out[2 * a.rows + r] = raw;
      a.cand_out[3 * a.rows + r] = mx;
      a.cand_out[4 * a.rows + r] = sum;
    } else if (a.mode == 0) {
      const int gen = a.generated[r] + 1;
      a.tokens[r] = idx;
      a.next_pos[r] += 1;
      a.generated[r] = gen;
      a.result[r * 3 + 0] = idx;
      a.result[r * 3 + 1] = 1;
      a.result[r * 3 + 2] = gen;
      if (a.log_prob != nullptr) a.log_prob[r] = logp;
    } else if (a.mode == 3) {
      a.first_token[r] = idx;
      if (a.log_prob != nullptr) a.log_prob[r] = logp;
    } else if (r == a.rows - 1) {
      a.first_token[0] = idx;
      if (a.log_prob != nullptr) a.log_prob[0] = logp;
    }
  }
  if (a.mode == 0 && blockIdx.x == 0) {
    // every slot's counter advances, occupied or not (kvcache.py:779 `.at[:].add(1)`)
    for (int s = threadIdx.x; s < a.num_slots; s += kFinalizeThreads) a.ar_lengths[s] += 1;
    if (threadIdx.x == 0) {
      a.ar_index[0] = (a.ar_index[0] + 1) % a.R;
      a.rng_state[0] += 1;
    }
  }
  timeline_end(tl);
}

}  // namespace mtx
