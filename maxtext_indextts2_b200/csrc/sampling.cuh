// top-k and nucleus (top-p) sampling over materialised fp32 logits, without sorting.
//
// Reference: MaxText/inference_utils.py:87-111.
//   topk    : lax.top_k keeps the k largest logits (ties towards the lower index), then
//             categorical(topk_logits / temperature).
//   nucleus : sort descending, cumsum(softmax) (temperature NOT applied), cutoff = first sorted
//             logit whose cumulative mass reaches p, logits below it -> -1e7, then
//             categorical(logits / temperature).
// categorical(x) == argmax(x + Gumbel noise); the noise is this repo's Philox stream keyed by
// (seed, step, row, vocabulary id) -- the same one the fused greedy/weighted path uses.
//
// Both cut-offs are order statistics, found by an 8-bit radix descent over the monotone
// integer image of the fp32 logits (4 passes over one row, which sits in L2): a count
// histogram for top-k, a probability-mass histogram for nucleus.  One CTA per row.
#pragma once

#include "common.cuh"

namespace mtx {

constexpr int kSampleThreads = 256;

struct SampleArgs {
  const float* logits;  // [rows, ld]
  long long ld;
  int vocab;            // entries per row
  int vocab_offset;     // global id of entry 0
  int mode;             // MTX_SAMPLE_NUCLEUS (2) or MTX_SAMPLE_TOPK (3); GREEDY (0) / WEIGHTED (1): no cut-off, every entry takes part
  int top_k;
  float nucleus_p;
  float inv_temp;
  const uint32_t* rng_state;
  int row_offset;
  // outputs, one per row (consumed by finalize_kernel with n_tiles = 1)
  float* out_score;
  int* out_idx;
  float* out_raw;
  float* out_max;
  float* out_sum;
};

// fp32 -> uint32 whose unsigned order is the float order
__device__ __forceinline__ uint32_t f32_order_key(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float order_key_f32(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < kSampleThreads / 32; ++w) r = fmaxf(r, red[w]);
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.0f;
  for (int w = 0; w < kSampleThreads / 32; ++w) r += red[w];  // fixed order
  return r;
}

__global__ void __launch_bounds__(kSampleThreads) sample_rows_kernel(const SampleArgs a) {
  __shared__ float s_hist[kSampleThreads / 32][256];  // per-warp histograms (counts or masses)
  __shared__ float s_tot[256];
  __shared__ float s_red[kSampleThreads / 32];
  __shared__ uint32_t s_prefix;
  __shared__ float s_above;
  __shared__ int s_found;
  __shared__ int s_warp_cnt[kSampleThreads / 32];
  __shared__ int s_base;
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = a.logits + (long long)r * a.ld;
  const int V = a.vocab;

  // row max and softmax denominator (needed by nucleus and by log-prob)
  float mx = -INFINITY;
  for (int i = tid; i < V; i += kSampleThreads) mx = fmaxf(mx, row[i]);
  mx = block_reduce_max(mx, s_red);
  float z = 0.0f;
  for (int i = tid; i < V; i += kSampleThreads) z += expf(row[i] - mx);
  z = block_reduce_sum(z, s_red);

  // ---- radix descent for the cut-off key ----
  const bool by_mass = a.mode == 2;
  const bool no_cut = a.mode < 2;  // greedy / weighted over materialised logits (mtx_sample_logits)
  const float target = by_mass ? a.nucleus_p * z : float(a.top_k < V ? a.top_k : V);
  if (tid == 0) { s_prefix = 0u; s_above = 0.0f; s_found = no_cut ? 0 : 1; }
  __syncthreads();
  for (int level = 3; level >= 0 && s_found; --level) {
    for (int i = tid; i < (kSampleThreads / 32) * 256; i += kSampleThreads) (&s_hist[0][0])[i] = 0.0f;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const int shift = 8 * level;
    for (int i = tid; i < V; i += kSampleThreads) {
      const float x = row[i];
      const uint32_t key = f32_order_key(x);
      if (level == 3 || (key >> (shift + 8)) == (prefix >> (shift + 8)))
        atomicAdd(&s_hist[warp][(key >> shift) & 255u], by_mass ? expf(x - mx) : 1.0f);
    }
    __syncthreads();
    {
      float t = 0.0f;
      for (int w = 0; w < kSampleThreads / 32; ++w) t += s_hist[w][tid];
      s_tot[tid] = t;
    }
    __syncthreads();
    if (tid == 0) {
      float above = s_above;
      int sel = -1;
      for (int d = 255; d >= 0; --d) {
        if (above + s_tot[d] >= target && s_tot[d] > 0.0f) { sel = d; break; }
        above += s_tot[d];
      }
      if (sel < 0) {
        s_found = 0;  // target exceeds the total (p > 1 up to rounding): nothing is cut
      } else {
        s_above = above;
        s_prefix = prefix | (uint32_t(sel) << shift);
      }
    }
    __syncthreads();
  }
  const bool found = s_found != 0;
  const uint32_t cut_key = s_prefix;
  const float cutoff = found ? order_key_f32(cut_key) : -INFINITY;
  // top-k: entries strictly above the cut-off are all kept; of those equal to it, the first `need` in index order
  const int need = by_mass ? 0 : int(target - s_above + 0.5f);

  uint32_t step = a.rng_state[0];
  const uint64_t seed = (uint64_t(a.rng_state[2]) << 32) | a.rng_state[1];
  float best = -INFINITY, raw = -INFINITY;
  int bi = 0x7fffffff;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < V; i0 += kSampleThreads) {
    const int i = i0 + tid;
    const float x = i < V ? row[i] : -INFINITY;
    bool keep;
    float val = x;
    if (by_mass) {
      keep = i < V;
      if (x < cutoff) val = -1.0e7f;  // NEG_INF of inference_utils.py:20; still takes part, as in the reference
    } else {
      const bool eq = i < V && found && f32_order_key(x) == cut_key;
      // ordered rank among the entries equal to the cut-off
      const uint32_t bal = __ballot_sync(0xffffffffu, eq);
      if (lane == 0) s_warp_cnt[warp] = __popc(bal);
      __syncthreads();
      int before = s_base, tot = 0;
      for (int w = 0; w < kSampleThreads / 32; ++w) {
        if (w < warp) before += s_warp_cnt[w];
        tot += s_warp_cnt[w];
      }
      const int rank = before + __popc(bal & ((1u << lane) - 1u));
      keep = i < V && (x > cutoff || (eq && rank < need) || !found);
      __syncthreads();  // every thread has read s_base / s_warp_cnt of this round
      if (tid == 0) s_base += tot;
    }
    if (keep) {
      const float sc = a.mode == 0 ? val : val * a.inv_temp + gumbel_noise(seed, step, uint32_t(a.row_offset + r), uint32_t(a.vocab_offset + i));
      if (sc > best) { best = sc; bi = i; raw = x; }
    }
  }
  // block arg-max, lowest index on ties
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float s2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    const float r2 = __shfl_xor_sync(0xffffffffu, raw, o);
    if (s2 > best || (s2 == best && i2 < bi)) { best = s2; bi = i2; raw = r2; }
  }
  __syncthreads();
  if (lane == 0) { s_hist[0][warp] = best; s_hist[1][warp] = __int_as_float(bi); s_hist[2][warp] = raw; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kSampleThreads / 32; ++w) {
      const float s2 = s_hist[0][w];
      const int i2 = __float_as_int(s_hist[1][w]);
      if (s2 > best || (s2 == best && i2 < bi)) { best = s2; bi = i2; raw = s_hist[2][w]; }
    }
    a.out_score[r] = best;
    a.out_idx[r] = a.vocab_offset + bi;
    a.out_raw[r] = raw;
    a.out_max[r] = mx;
    a.out_sum[r] = z;
  }
}


// =============================================================================================================
// Vocab-parallel top-k / nucleus: every shard extracts its kCandK best logits per row, ONE all-gather
// carries them (kCandFloats floats per row and shard), every rank runs the same final selection.
//
//   top-k   : exact for k <= kCandK (the global top-k is a subset of the union of the shards' top-k).
//   nucleus : the cut-off of inference_utils.py:96-99 is found on the merged candidates with the GLOBAL softmax
//             denominator (merged from the shards' (max, sum exp)); exact whenever everything at or above the cut-off is
//             among the candidates, i.e. no shard's kCandK-th logit reaches the cut-off.  Otherwise the nucleus is
//             truncated to the candidates and `truncated[0]` counts the row (peaked distributions of a trained model fit;
//             the near-uniform logits of random-init weights do not).  Entries below the cut-off take part in the
//             reference's draw with logit -1e7 (probability < e^-1e6): they are left out here.
// =============================================================================================================

constexpr int kCandK = 64;
constexpr int kCandFloats = 2 * kCandK + 2;  // per row: values[kCandK] desc, vocabulary ids[kCandK] (int bits), max, sum exp
constexpr int kCandMaxShards = 8;

struct ShardTopkArgs {
  const float* logits;  // [rows, ld] this shard's logits
  long long ld;
  int vocab;            // entries per row in this shard
  int vocab_offset;     // global id of entry 0
  float* cand;          // [rows, kCandFloats]
};

__global__ void __launch_bounds__(kSampleThreads) shard_topk_kernel(const ShardTopkArgs a) {
  __shared__ float s_hist[kSampleThreads / 32][256];
  __shared__ float s_tot[256];
  __shared__ float s_red[kSampleThreads / 32];
  __shared__ uint32_t s_prefix;
  __shared__ float s_above;
  __shared__ int s_found;
  __shared__ int s_cnt_keep[kSampleThreads / 32], s_cnt_eq[kSampleThreads / 32];
  __shared__ int s_base_keep, s_base_eq;
  __shared__ float s_val[kCandK];
  __shared__ int s_id[kCandK];
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = a.logits + (long long)r * a.ld;
  const int V = a.vocab;
  const int K = V < kCandK ? V : kCandK;

  float mx = -INFINITY;
  for (int i = tid; i < V; i += kSampleThreads) mx = fmaxf(mx, row[i]);
  mx = block_reduce_max(mx, s_red);
  float z = 0.0f;
  for (int i = tid; i < V; i += kSampleThreads) z += expf(row[i] - mx);
  z = block_reduce_sum(z, s_red);

  // radix descent (count histogram) for the key of the K-th largest logit
  if (tid == 0) { s_prefix = 0u; s_above = 0.0f; s_found = 1; }
  if (tid < kCandK) { s_val[tid] = -INFINITY; s_id[tid] = 0x7fffffff; }
  __syncthreads();
  for (int level = 3; level >= 0; --level) {
    for (int i = tid; i < (kSampleThreads / 32) * 256; i += kSampleThreads) (&s_hist[0][0])[i] = 0.0f;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const int shift = 8 * level;
    for (int i = tid; i < V; i += kSampleThreads) {
      const uint32_t key = f32_order_key(row[i]);
      if (level == 3 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[warp][(key >> shift) & 255u], 1.0f);
    }
    __syncthreads();
    {
      float t = 0.0f;
      for (int w = 0; w < kSampleThreads / 32; ++w) t += s_hist[w][tid];
      s_tot[tid] = t;
    }
    __syncthreads();
    if (tid == 0) {
      float above = s_above;
      int sel = 0;
      for (int d = 255; d >= 0; --d) {
        if (above + s_tot[d] >= float(K) && s_tot[d] > 0.0f) { sel = d; break; }
        above += s_tot[d];
      }
      s_above = above;
      s_prefix = prefix | (uint32_t(sel) << shift);
    }
    __syncthreads();
  }
  const uint32_t cut_key = s_prefix;
  const int n_above = int(s_above + 0.5f);  // entries strictly above the cut-off: all kept
  const int need = K - n_above;             // of those equal to it, the first `need` in index order
  if (tid == 0) { s_base_keep = 0; s_base_eq = 0; }
  __syncthreads();
  // ordered compaction (index order) into the candidate list
  for (int i0 = 0; i0 < V; i0 += kSampleThreads) {
    const int i = i0 + tid;
    const float x = i < V ? row[i] : -INFINITY;
    const uint32_t key = f32_order_key(x);
    const bool eq = i < V && key == cut_key;
    const uint32_t bal_eq = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) s_cnt_eq[warp] = __popc(bal_eq);
    __syncthreads();
    int eq_before = s_base_eq, eq_tot = 0;
    for (int w = 0; w < kSampleThreads / 32; ++w) {
      if (w < warp) eq_before += s_cnt_eq[w];
      eq_tot += s_cnt_eq[w];
    }
    const int eq_rank = eq_before + __popc(bal_eq & ((1u << lane) - 1u));
    const bool keep = i < V && (key > cut_key || (eq && eq_rank < need));
    const uint32_t bal_keep = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_cnt_keep[warp] = __popc(bal_keep);
    __syncthreads();
    int keep_before = s_base_keep, keep_tot = 0;
    for (int w = 0; w < kSampleThreads / 32; ++w) {
      if (w < warp) keep_before += s_cnt_keep[w];
      keep_tot += s_cnt_keep[w];
    }
    if (keep) {
      const int pos = keep_before + __popc(bal_keep & ((1u << lane) - 1u));
      if (pos < kCandK) { s_val[pos] = x; s_id[pos] = a.vocab_offset + i; }
    }
    __syncthreads();
    if (tid == 0) { s_base_eq += eq_tot; s_base_keep += keep_tot; }
    __syncthreads();
  }
  // sort the (at most 64) candidates: value descending, id ascending; rank by counting
  float* out = a.cand + (long long)r * kCandFloats;
  if (tid < kCandK) {
    const float v = s_val[tid];
    const int id = s_id[tid];
    int rank = 0;
    for (int j = 0; j < kCandK; ++j) {
      const float vj = s_val[j];
      const int idj = s_id[j];
      rank += (vj > v || (vj == v && (idj < id || (idj == id && j < tid)))) ? 1 : 0;
    }
    out[rank] = v;
    out[kCandK + rank] = __int_as_float(id);
  }
  if (tid == 0) { out[2 * kCandK] = mx; out[2 * kCandK + 1] = z; }
}

struct CommitTopkArgs {
  const float* gathered;  // [n_shards, rows, kCandFloats]
  int n_shards, rows;
  int shard_vocab;        // vocabulary entries per shard (a shard with more than kCandK entries may hide logits behind its last candidate)
  int mode;               // MTX_SAMPLE_NUCLEUS (2) or MTX_SAMPLE_TOPK (3)
  int top_k;
  float nucleus_p, inv_temp;
  int row_offset;
  // decode state (as FinalizeArgs, mode 0)
  int* tokens;
  int* next_pos;
  int* generated;
  int* ar_lengths;
  int* ar_index;
  int* result;
  float* log_prob;
  uint32_t* rng_state;
  int num_slots, R;
  int* truncated;         // [1] rows whose nucleus did not fit the candidates (since bind)
  int* ticket;            // [1] zero between launches: the last CTA to have read the step counter advances the shared state
};

__global__ void __launch_bounds__(kSampleThreads) commit_topk_kernel(const CommitTopkArgs a) {
  constexpr int kMax = kCandMaxShards * kCandK;  // 512
  __shared__ float s_v[kMax], s_sv[kMax];
  __shared__ int s_i[kMax], s_si[kMax];
  __shared__ float s_M, s_Z, s_cut;
  __shared__ float s_best[kSampleThreads / 32], s_raw[kSampleThreads / 32];
  __shared__ int s_bi[kSampleThreads / 32];
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n_shards * kCandK;
  for (int i = tid; i < n; i += kSampleThreads) {
    const int s = i / kCandK, j = i - s * kCandK;
    const float* c = a.gathered + ((long long)s * a.rows + r) * kCandFloats;
    s_v[i] = c[j];
    s_i[i] = __float_as_int(c[kCandK + j]);
  }
  if (tid == 0) {  // global softmax statistics from the shards' (max, sum exp)
    float M = -INFINITY;
    for (int s = 0; s < a.n_shards; ++s) M = fmaxf(M, a.gathered[((long long)s * a.rows + r) * kCandFloats + 2 * kCandK]);
    float Z = 0.0f;
    for (int s = 0; s < a.n_shards; ++s) {
      const float* c = a.gathered + ((long long)s * a.rows + r) * kCandFloats + 2 * kCandK;
      Z += c[1] * expf(c[0] - M);
    }
    s_M = M;
    s_Z = Z;
  }
  __syncthreads();
  // global order of the candidates (value descending, id ascending): rank by counting, scatter into sorted arrays
  for (int i = tid; i < n; i += kSampleThreads) {
    const float v = s_v[i];
    const int id = s_i[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float vj = s_v[j];
      const int idj = s_i[j];
      rank += (vj > v || (vj == v && (idj < id || (idj == id && j < i)))) ? 1 : 0;
    }
    s_sv[rank] = v;
    s_si[rank] = id;
  }
  __syncthreads();
  if (warp == 0) {
    int n_valid = 0;
    for (int i = lane; i < n; i += 32) n_valid += s_sv[i] > -INFINITY ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    float cut = -INFINITY;
    if (a.mode == 3) {  // top-k: the k-th best value and, among equals, the lower ids (lax.top_k)
      const int k = a.top_k < n_valid ? a.top_k : n_valid;
      cut = __int_as_float(k);  // top-k keeps ranks [0, k): encoded as a count, see below
    } else {
      // nucleus: cumulative softmax mass in sorted order; cutoff_idx = #(cum < p), clamped (inference_utils.py:96-99)
      const int per = (n + 31) / 32;
      float local = 0.0f;
      for (int q = 0; q < per; ++q) {
        const int i = lane * per + q;
        if (i < n && s_sv[i] > -INFINITY) local += expf(s_sv[i] - s_M) / s_Z;
      }
      float incl = local;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      float run = incl - local;
      int below = 0;
      for (int q = 0; q < per; ++q) {
        const int i = lane * per + q;
        if (i < n && s_sv[i] > -INFINITY) {
          run += expf(s_sv[i] - s_M) / s_Z;
          below += run < a.nucleus_p ? 1 : 0;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
      bool trunc = below >= n_valid;  // the nucleus reaches past the candidates
      const int idx = below < n_valid ? below : n_valid - 1;
      cut = s_sv[idx < 0 ? 0 : idx];
      // a shard whose last candidate is still inside the nucleus may hold more logits that belong to it
      if (a.shard_vocab > kCandK)
        for (int s = 0; s < a.n_shards; ++s) {  // (the candidates of a shard come in any order: its smallest one)
          float mn = INFINITY;
          for (int j = lane; j < kCandK; j += 32) mn = fminf(mn, a.gathered[((long long)s * a.rows + r) * kCandFloats + j]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          trunc = trunc || mn >= cut;
        }
      if (lane == 0 && trunc) atomicAdd(a.truncated, 1);
    }
    if (lane == 0) s_cut = cut;
  }
  __syncthreads();
  const uint32_t step = a.rng_state[0];
  const uint64_t seed = (uint64_t(a.rng_state[2]) << 32) | a.rng_state[1];
  float best = -INFINITY, raw = -INFINITY;
  int bi = 0x7fffffff;
  const int k_keep = a.mode == 3 ? __float_as_int(s_cut) : 0;
  for (int i = tid; i < n; i += kSampleThreads) {
    const float v = s_sv[i];
    const bool keep = v > -INFINITY && (a.mode == 3 ? i < k_keep : v >= s_cut);
    if (keep) {
      const float sc = v * a.inv_temp + gumbel_noise(seed, step, uint32_t(a.row_offset + r), uint32_t(s_si[i]));
      if (sc > best || (sc == best && s_si[i] < bi)) { best = sc; bi = s_si[i]; raw = v; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float s2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    const float r2 = __shfl_xor_sync(0xffffffffu, raw, o);
    if (s2 > best || (s2 == best && i2 < bi)) { best = s2; bi = i2; raw = r2; }
  }
  if (lane == 0) { s_best[warp] = best; s_bi[warp] = bi; s_raw[warp] = raw; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kSampleThreads / 32; ++w)
      if (s_best[w] > best || (s_best[w] == best && s_bi[w] < bi)) { best = s_best[w]; bi = s_bi[w]; raw = s_raw[w]; }
    const int gen = a.generated[r] + 1;
    a.tokens[r] = bi;
    a.next_pos[r] += 1;
    a.generated[r] = gen;
    a.result[r * 3 + 0] = bi;
    a.result[r * 3 + 1] = 1;
    a.result[r * 3 + 2] = gen;
    if (a.log_prob != nullptr) a.log_prob[r] = raw - (s_M + logf(s_Z));
  }
  // every CTA has read rng_state[0] above: the last one to get here advances the state shared by all rows
  __shared__ int s_last;
  if (tid == 0) {
    const int t = atomicAdd(a.ticket, 1);
    s_last = t == int(gridDim.x) - 1;
    if (s_last) *a.ticket = 0;
  }
  __syncthreads();
  if (s_last) {
    for (int s = tid; s < a.num_slots; s += kSampleThreads) a.ar_lengths[s] += 1;
    if (tid == 0) {
      a.ar_index[0] = (a.ar_index[0] + 1) % a.R;
      a.rng_state[0] += 1;
    }
  }
}

}  // namespace mtx

namespace mtx {

// =============================================================================================================
// top-k / nucleus over materialised logits on ALL SMs (round 2).  sample_rows_kernel above walks a row with one CTA
// (64 CTAs at batch 64, seven passes over 1 MB each: ~1.4 ms per step); here the same radix descent is cut along the
// vocabulary: slices x ceil(rows / row_group) CTAs histogram their slice in shared memory and add the non-empty bins to
// the row's global histogram, one small CTA per row picks the bin; three levels of 11 + 11 + 10 key bits, one sweep counts
// the entries equal to the cut-off (top-k ties) and one more draws the Gumbel-max winner of every (row, slice, warp);
// finalize_kernel merges those like the tiles of the fused path.  Every sweep reads 16 bytes per lane, keeps four such
// loads in flight per thread and issues the next batch's loads before it works on the current one: a sweep is a handful of
// dependent DRAM/L2 round trips per CTA, so what it costs is how many of them are serialised.
// Why 11-bit digits: the top BYTE of a float key is sign + 7 exponent bits, and the logits of a row share a few exponents,
// so an 8-bit first level piles every element onto ~6 shared-memory addresses (measured: 0.8 ms of serialised atomics;
// warp-aggregating them with match.any was slower still); two mantissa bits more spread a row over ~80 bins.
// Mass histograms are 2^-40 fixed point in integers: exact, and independent of the order in which threads and slices
// add -- the cut-off is a deterministic function of the logits.
// =============================================================================================================

constexpr int kParMaxSlices = 128;
constexpr int kParThreads = 256;
constexpr int kParWarps = kParThreads / 32;
constexpr int kParBatch = 4 * kParThreads;  // 16-byte units one CTA has in flight
constexpr int kParLevels = 3;
constexpr int kParBins = 2048;
constexpr float kParFix = 1099511627776.0f;  // 2^40
__device__ __forceinline__ int par_shift(int level) { return level == 0 ? 21 : level == 1 ? 10 : 0; }  // level 0 = most significant digit
__device__ __forceinline__ int par_width(int level) { return level == 2 ? 10 : 11; }

// Host: the cut of a row -- one batch per slice when the vocabulary allows it -- and the rows a CTA walks.
inline int par_slices(int vocab) {
  const int n4 = (vocab + 3) / 4;
  const int s = (n4 + kParBatch - 1) / kParBatch;
  return s < 1 ? 1 : s > kParMaxSlices ? kParMaxSlices : s;
}
inline int par_row_group(int vocab, int rows) {
  const int g = par_slices(vocab) * rows / 1024;  // ~7 CTAs of 256 threads per SM
  return g < 1 ? 1 : g > 16 ? 16 : g;
}

struct ParSampleArgs {
  const float* logits;  // [rows, ld]
  long long ld;
  int vocab, vocab_offset, rows;
  int slices, row_group;
  int mode;             // MTX_SAMPLE_NUCLEUS (2) or MTX_SAMPLE_TOPK (3)
  int top_k;
  float nucleus_p, inv_temp;
  const uint32_t* rng_state;
  int row_offset;
  const float* part_max;  // [rows, n_tiles] per-tile (max, sum exp) the logits epilogue left
  const float* part_sum;
  int n_tiles;
  float* row_M;                    // [rows]
  float* row_Z;                    // [rows]
  unsigned long long* target;      // [rows] mass (2^-40 fixed point of exp(x - M)) or count to reach
  unsigned long long* above;       // [rows] mass / count strictly above the selected prefix
  uint32_t* prefix;                // [rows] radix prefix of the cut-off key
  int* found;                      // [rows] 0: the target exceeds the total, nothing is cut; 1: cut; 2: cut, and only some of the
                                   // entries equal to the cut-off key are kept (top-k ties: rank them by index)
  unsigned long long* hist;        // [rows, kParBins] zero between levels (par_select_kernel clears what it reads)
  int* eq_count;                   // [rows, slices * kParWarps] entries equal to the cut-off per warp piece (rows with found == 2)
  float* cand;                     // vocab-parallel (par_collect_kernel): [rows, kCandFloats] this shard's candidates, or null
  int* cand_count;                 // [rows] candidates written so far (reset by par_stats_kernel)
  float* out_score;                // [rows, slices * kParWarps] partials for finalize_kernel
  int* out_idx;
  float* out_raw;
  float* out_max;
  float* out_sum;
};

// Slice s of a row in units of 4 consecutive entries (16-byte loads): vec4 indices [lo4, hi4).
__device__ __forceinline__ void par_slice4(int vocab, int slices, int s, int& lo4, int& hi4) {
  const int n4 = (vocab + 3) >> 2;
  lo4 = int((long long)s * n4 / slices);
  hi4 = int((long long)(s + 1) * n4 / slices);
}
// Warp w's contiguous piece of a slice.
__device__ __forceinline__ void par_piece4(int lo4, int hi4, int w, int& wlo, int& whi) {
  wlo = lo4 + int((long long)w * (hi4 - lo4) / kParWarps);
  whi = lo4 + int((long long)(w + 1) * (hi4 - lo4) / kParWarps);
}
// Entries 4 q .. 4 q + 3 of a row; entries at or past `vocab` read as -inf.  `vec` = rows are 16-byte aligned and whole.
__device__ __forceinline__ float4 par_load4(const float* row, int q, int vocab, bool vec) {
  if (vec) return __ldg(reinterpret_cast<const float4*>(row) + q);
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = 4 * q + j < vocab ? __ldg(row + 4 * q + j) : -INFINITY;
  return make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ float4 par_neg_inf4() { return make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); }
__device__ __forceinline__ int par_count_eq(const float4& x, int q, int vocab, uint32_t key) {
  return (4 * q < vocab && f32_order_key(x.x) == key) + (4 * q + 1 < vocab && f32_order_key(x.y) == key) +
         (4 * q + 2 < vocab && f32_order_key(x.z) == key) + (4 * q + 3 < vocab && f32_order_key(x.w) == key);
}

// One CTA per row: softmax statistics from the per-tile partials, the descent's target, state reset.
__global__ void __launch_bounds__(kParThreads) par_stats_kernel(const ParSampleArgs a) {
  __shared__ float s_red[kParWarps];
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x, tid = threadIdx.x;
  float mx = -INFINITY;
  for (int t = tid; t < a.n_tiles; t += kParThreads) mx = fmaxf(mx, a.part_max[(long long)r * a.n_tiles + t]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = mx;
  __syncthreads();
  mx = s_red[0];
  for (int w = 1; w < kParWarps; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  float z = 0.0f;
  for (int t = tid; t < a.n_tiles; t += kParThreads) {
    const float m = a.part_max[(long long)r * a.n_tiles + t];
    if (m > -INFINITY) z += a.part_sum[(long long)r * a.n_tiles + t] * expf(m - mx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = z;
  __syncthreads();
  for (int i = tid; i < kParBins; i += kParThreads) a.hist[(long long)r * kParBins + i] = 0ull;
  if (tid == 0) {
    float tot = 0.0f;
    for (int w = 0; w < kParWarps; ++w) tot += s_red[w];  // fixed order
    a.row_M[r] = mx;
    a.row_Z[r] = tot;
    a.target[r] = a.mode == 2 ? (unsigned long long)(double(a.nucleus_p) * double(tot) * double(kParFix))
                              : (unsigned long long)(a.top_k < a.vocab ? a.top_k : a.vocab);
    a.above[r] = 0ull;
    a.prefix[r] = 0u;
    a.found[r] = 1;
    if (a.cand_count != nullptr) a.cand_count[r] = 0;
  }
  if (a.cand != nullptr) {  // empty candidate slots: -inf with an id past every vocabulary
    for (int i = tid; i < kCandK; i += kParThreads) {
      a.cand[(long long)r * kCandFloats + i] = -INFINITY;
      a.cand[(long long)r * kCandFloats + kCandK + i] = __int_as_float(0x7fffffff);
    }
  }
}

// Histogram of one radix level over one vocabulary slice for a group of rows, added to the rows' global histograms.
// Shared-memory counters are 32-bit (native atomics; the 64-bit ones are compare-and-swap loops): a count is one counter,
// a 2^-40 fixed-point weight is split into 14 + 14 + 13 bits over three (a slice would need 2^18 entries to overflow one).
__global__ void __launch_bounds__(kParThreads) par_hist_kernel(const ParSampleArgs a, int level) {
  __shared__ uint32_t s_hist[3][kParBins];
  griddep_launch_dependents();
  const int tid = threadIdx.x;
  for (int i = tid; i < 3 * kParBins; i += kParThreads) (&s_hist[0][0])[i] = 0u;
  __syncthreads();
  griddep_wait();
  int lo4, hi4;
  par_slice4(a.vocab, a.slices, blockIdx.x, lo4, hi4);
  const bool vec = (a.ld & 3) == 0 && (a.vocab & 3) == 0;
  const bool mass = a.mode == 2;
  const int r0 = blockIdx.y * a.row_group;
  const int n_rows = min(a.row_group, a.rows - r0);
  const int nb = (hi4 - lo4 + kParBatch - 1) / kParBatch;  // batches per row
  const int shift = par_shift(level), above_bits = shift + par_width(level);
  const uint32_t mask = (1u << par_width(level)) - 1u;
  auto load = [&](int it, float4 (&x)[4]) {
    const float* row = a.logits + (long long)(r0 + it / nb) * a.ld;
    const int q0 = lo4 + (it % nb) * kParBatch;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * kParThreads + tid;
      x[u] = q < hi4 ? par_load4(row, q, a.vocab, vec) : par_neg_inf4();
    }
  };
  const int total = n_rows * nb;
  float4 cur[4], nxt[4];
  if (total > 0) load(0, nxt);
  for (int it = 0; it < total; ++it) {
    const int r = r0 + it / nb, b = it % nb;
    const bool live = a.found[r] != 0;  // (block-uniform)
    const uint32_t prefix = a.prefix[r];
    const float M = a.row_M[r];
#pragma unroll
    for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
    if (it + 1 < total) load(it + 1, nxt);
    if (live) {
      const int q0 = lo4 + b * kParBatch;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = q0 + u * kParThreads + tid;
        const float xs[4] = {cur[u].x, cur[u].y, cur[u].z, cur[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (q >= hi4 || 4 * q + j >= a.vocab) continue;
          const uint32_t key = f32_order_key(xs[j]);
          if (level == 0 || (key >> above_bits) == (prefix >> above_bits)) {
            const uint32_t bin = (key >> shift) & mask;
            if (level > 0) {
              // the entries under the selected prefix are few (a row spreads over ~80 first-level bins): no staging
              atomicAdd(a.hist + (long long)r * kParBins + bin, mass ? __float2ull_rz(expf(xs[j] - M) * kParFix) : 1ull);
            } else if (mass) {
              const unsigned long long w = __float2ull_rz(expf(xs[j] - M) * kParFix);
              atomicAdd(&s_hist[0][bin], uint32_t(w & 0x3FFFu));
              atomicAdd(&s_hist[1][bin], uint32_t((w >> 14) & 0x3FFFu));
              atomicAdd(&s_hist[2][bin], uint32_t(w >> 28));
            } else {
              atomicAdd(&s_hist[0][bin], 1u);
            }
          }
        }
      }
    }
    if (level == 0 && b == nb - 1 && live) {
      __syncthreads();
      for (int i = tid; i < kParBins; i += kParThreads) {
        const unsigned long long v = mass ? (unsigned long long)s_hist[0][i] + ((unsigned long long)s_hist[1][i] << 14) + ((unsigned long long)s_hist[2][i] << 28)
                                          : (unsigned long long)s_hist[0][i];
        if (v != 0ull) {
          atomicAdd(a.hist + (long long)r * kParBins + i, v);
          s_hist[0][i] = s_hist[1][i] = s_hist[2][i] = 0u;
        }
      }
      __syncthreads();
    }
  }
}

// One CTA per row: descend one level on the row's histogram (and clear it for the next level).  Thread t owns bins
// 8 t .. 8 t + 7; a block-wide suffix sum over the 256 chunk totals finds the chunk the target falls in.
__global__ void __launch_bounds__(256) par_select_kernel(const ParSampleArgs a, int level) {
  __shared__ unsigned long long s_warp[8];
  __shared__ int s_sel;
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (!a.found[r]) return;
  constexpr int kPer = kParBins / 256;
  unsigned long long* h = a.hist + (long long)r * kParBins;
  unsigned long long bins[kPer];
  unsigned long long t = 0ull;
  {
    const ulonglong2* h2 = reinterpret_cast<const ulonglong2*>(h + tid * kPer);
#pragma unroll
    for (int j = 0; j < kPer / 2; ++j) {
      const ulonglong2 v = h2[j];
      bins[2 * j] = v.x;
      bins[2 * j + 1] = v.y;
    }
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      t += bins[j];
      if (bins[j] != 0ull) h[tid * kPer + j] = 0ull;
    }
  }
  if (tid == 0) s_sel = -1;
  // inclusive suffix sum over threads (thread 255 = the highest bins)
  unsigned long long suf = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += v;
  }
  if (lane == 0) s_warp[warp] = suf;
  __syncthreads();
  for (int w = warp + 1; w < 8; ++w) suf += s_warp[w];
  const unsigned long long above0 = a.above[r], target = a.target[r];
  // the highest chunk whose inclusive suffix reaches the target owns the bin (its own total is then non-zero)
  const bool hit = above0 + suf >= target && t > 0ull;
  if (hit) atomicMax(&s_sel, tid);
  __syncthreads();
  const int sel_chunk = s_sel;
  if (sel_chunk < 0) {
    if (tid == 0) a.found[r] = 0;  // the target exceeds the total (p >= 1 up to rounding): nothing is cut
    return;
  }
  if (tid == sel_chunk) {
    unsigned long long above = above0 + suf - t;
    int sel = tid * kPer;
#pragma unroll
    for (int j = kPer - 1; j >= 0; --j) {
      if (above + bins[j] >= target && bins[j] > 0ull) { sel = tid * kPer + j; break; }
      above += bins[j];
    }
    a.above[r] = above;
    a.prefix[r] |= uint32_t(sel) << par_shift(level);
    // top-k, last level: bins[sel] entries carry the cut-off key and target - above of them are kept
    if (a.mode != 2 && level == kParLevels - 1 && bins[sel - tid * kPer] != target - above) a.found[r] = 2;
  }
}

// top-k only: entries equal to the cut-off key per (slice, warp piece, row) -- the tie rule of lax.top_k is "lowest index
// first", and par_scan_kernel's warps need the count in the pieces before theirs.
__global__ void __launch_bounds__(kParThreads) par_eqcount_kernel(const ParSampleArgs a) {
  griddep_launch_dependents();
  griddep_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int lo4, hi4, wlo, whi;
  par_slice4(a.vocab, a.slices, blockIdx.x, lo4, hi4);
  par_piece4(lo4, hi4, warp, wlo, whi);
  const bool vec = (a.ld & 3) == 0 && (a.vocab & 3) == 0;
  const int r0 = blockIdx.y * a.row_group;
  const int n_rows = min(a.row_group, a.rows - r0);
  const int nb = (whi - wlo + 127) / 128;  // batches of 4 x 32 lanes per row
  const int total = n_rows * nb;
  const int piece = blockIdx.x * kParWarps + warp, pieces = a.slices * kParWarps;
  auto load = [&](int it, float4 (&x)[4]) {
    const int r = r0 + it / nb;
    const float* row = a.logits + (long long)r * a.ld;
    const int q0 = wlo + (it % nb) * 128;
    const bool tied = a.found[r] == 2;  // the other rows keep every entry that equals the cut-off: nothing to count
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * 32 + lane;
      x[u] = tied && q < whi ? par_load4(row, q, a.vocab, vec) : par_neg_inf4();
    }
  };
  float4 cur[4], nxt[4];
  if (total > 0) load(0, nxt);
  int mine = 0;
  for (int it = 0; it < total; ++it) {
    const int r = r0 + it / nb, b = it % nb;
    const uint32_t cut_key = a.prefix[r];
#pragma unroll
    for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
    if (it + 1 < total) load(it + 1, nxt);
    const int q0 = wlo + b * 128;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * 32 + lane;
      if (q < whi) mine += par_count_eq(cur[u], q, a.vocab, cut_key);
    }
    if (b == nb - 1) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
      if (lane == 0) a.eq_count[(long long)r * pieces + piece] = mine;
      mine = 0;
    }
  }
  if (total == 0 && lane == 0)
    for (int rr = 0; rr < n_rows; ++rr) a.eq_count[(long long)(r0 + rr) * pieces + piece] = 0;
}

// The draw: for every (row, slice, warp piece) the best Gumbel-perturbed kept entry (inference_utils.py:87-111), lowest
// index on ties.  A warp owns a contiguous piece and a lane four consecutive entries, so "the first `need` entries equal to
// the cut-off in index order" needs the counts of the pieces before this one (par_eqcount_kernel) and four ballots per 128
// entries -- no block-wide scan, no barrier; the noise of a lane's four entries is one Philox block.  Nucleus: the
// reference leaves the entries below the cut-off in the draw with logit -1e7; their score is below every kept entry's by
// ~1e7 / temperature (Gumbel noise is < 17), they cannot win, and they are skipped here (no noise is generated for them).
__global__ void __launch_bounds__(kParThreads, 4) par_scan_kernel(const ParSampleArgs a) {
  griddep_launch_dependents();
  griddep_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int piece = blockIdx.x * kParWarps + warp, pieces = a.slices * kParWarps;
  int lo4, hi4, wlo, whi;
  par_slice4(a.vocab, a.slices, blockIdx.x, lo4, hi4);
  par_piece4(lo4, hi4, warp, wlo, whi);
  const bool vec = (a.ld & 3) == 0 && (a.vocab & 3) == 0;
  const bool noise4 = (a.vocab_offset & 3) == 0;
  const uint32_t step = a.rng_state[0];
  const uint64_t seed = (uint64_t(a.rng_state[2]) << 32) | a.rng_state[1];
  const uint32_t lt = (1u << lane) - 1u;
  const int r0 = blockIdx.y * a.row_group;
  const int n_rows = min(a.row_group, a.rows - r0);
  const int nb = (whi - wlo + 127) / 128;
  const int total = n_rows * nb;
  auto load = [&](int it, float4 (&x)[4]) {
    const float* row = a.logits + (long long)(r0 + it / nb) * a.ld;
    const int q0 = wlo + (it % nb) * 128;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * 32 + lane;
      x[u] = q < whi ? par_load4(row, q, a.vocab, vec) : par_neg_inf4();
    }
  };
  float4 cur[4], nxt[4];
  if (total > 0) load(0, nxt);
  float best = -INFINITY, raw = -INFINITY;
  int bi = 0x7fffffff, need = 0, before = 0;
  uint32_t cut_key = 0u;
  bool found = false, tied = false;
  for (int it = 0; it < total; ++it) {
    const int r = r0 + it / nb, b = it % nb;
    if (b == 0) {
      const int f = a.found[r];
      found = f != 0;
      tied = f == 2;
      cut_key = found ? a.prefix[r] : 0u;
      best = raw = -INFINITY;
      bi = 0x7fffffff;
      before = 0;
      if (tied) {
        need = int(a.target[r] - a.above[r]);
        for (int p2 = lane; p2 < piece; p2 += 32) before += a.eq_count[(long long)r * pieces + p2];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
    if (it + 1 < total) load(it + 1, nxt);
    const int q0 = wlo + b * 128;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * 32 + lane;
      const bool in = q < whi;
      const float xs[4] = {cur[u].x, cur[u].y, cur[u].z, cur[u].w};
      bool keep[4];
      if (!tied) {  // (warp-uniform)
#pragma unroll
        for (int j = 0; j < 4; ++j) keep[j] = in && 4 * q + j < a.vocab && f32_order_key(xs[j]) >= cut_key;  // (by key: -0.0 < +0.0)
      } else {
        bool eq[4];
        uint32_t bal[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          eq[j] = in && 4 * q + j < a.vocab && f32_order_key(xs[j]) == cut_key;
          bal[j] = __ballot_sync(0xffffffffu, eq[j]);
        }
        // index order inside the 128 entries: lane-major, then the lane's four entries
        int rank = before + __popc(bal[0] & lt) + __popc(bal[1] & lt) + __popc(bal[2] & lt) + __popc(bal[3] & lt);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          keep[j] = in && 4 * q + j < a.vocab && (f32_order_key(xs[j]) > cut_key || (eq[j] && rank < need));
          rank += eq[j] ? 1 : 0;
        }
        before += __popc(bal[0]) + __popc(bal[1]) + __popc(bal[2]) + __popc(bal[3]);
      }
      if (keep[0] || keep[1] || keep[2] || keep[3]) {
        float g[4];
        if (noise4) {
          const float4 g4 = gumbel_noise4(seed, step, uint32_t(a.row_offset + r), uint32_t(a.vocab_offset + 4 * q));
          g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) g[j] = keep[j] ? gumbel_noise(seed, step, uint32_t(a.row_offset + r), uint32_t(a.vocab_offset + 4 * q + j)) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!keep[j]) continue;
          const float sc = xs[j] * a.inv_temp + g[j];
          if (sc > best) { best = sc; bi = 4 * q + j; raw = xs[j]; }
        }
      }
    }
    if (b == nb - 1) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float s2 = __shfl_xor_sync(0xffffffffu, best, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
        const float r2 = __shfl_xor_sync(0xffffffffu, raw, o);
        if (s2 > best || (s2 == best && i2 < bi)) { best = s2; bi = i2; raw = r2; }
      }
      if (lane == 0) {
        const long long o = (long long)r * pieces + piece;
        a.out_score[o] = best;
        a.out_idx[o] = bi == 0x7fffffff ? 0x7fffffff : a.vocab_offset + bi;
        a.out_raw[o] = raw;
        a.out_max[o] = piece == 0 ? a.row_M[r] : -INFINITY;
        a.out_sum[o] = piece == 0 ? a.row_Z[r] : 0.0f;
      }
    }
  }
  if (total == 0 && lane == 0) {
    for (int rr = 0; rr < n_rows; ++rr) {
      const long long o = (long long)(r0 + rr) * pieces + piece;
      a.out_score[o] = -INFINITY;
      a.out_idx[o] = 0x7fffffff;
      a.out_raw[o] = -INFINITY;
      a.out_max[o] = piece == 0 ? a.row_M[r0 + rr] : -INFINITY;
      a.out_sum[o] = piece == 0 ? a.row_Z[r0 + rr] : 0.0f;
    }
  }
}

// Vocab-parallel top-k / nucleus: this shard's kCandK best logits of every row as (value, global id) pairs in
// any order, plus the shard's (max, sum exp) -- what shard_topk_kernel produces with one CTA per row, here from the radix
// descent above with top_k = kCandK.  The kept set is exactly par_scan_kernel's (cut-off key, ties by index); a kept entry
// takes the next free slot of its row (the merge on the gathered candidates, commit_topk_kernel, orders by value and id).
__global__ void __launch_bounds__(kParThreads) par_collect_kernel(const ParSampleArgs a) {
  griddep_launch_dependents();
  griddep_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int piece = blockIdx.x * kParWarps + warp, pieces = a.slices * kParWarps;
  int lo4, hi4, wlo, whi;
  par_slice4(a.vocab, a.slices, blockIdx.x, lo4, hi4);
  par_piece4(lo4, hi4, warp, wlo, whi);
  const bool vec = (a.ld & 3) == 0 && (a.vocab & 3) == 0;
  const uint32_t lt = (1u << lane) - 1u;
  const int r0 = blockIdx.y * a.row_group;
  const int n_rows = min(a.row_group, a.rows - r0);
  const int nb = (whi - wlo + 127) / 128;
  const int total = n_rows * nb;
  auto load = [&](int it, float4 (&x)[4]) {
    const float* row = a.logits + (long long)(r0 + it / nb) * a.ld;
    const int q0 = wlo + (it % nb) * 128;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * 32 + lane;
      x[u] = q < whi ? par_load4(row, q, a.vocab, vec) : par_neg_inf4();
    }
  };
  float4 cur[4], nxt[4];
  if (total > 0) load(0, nxt);
  int need = 0, before = 0;
  uint32_t cut_key = 0u;
  bool tied = false;
  for (int it = 0; it < total; ++it) {
    const int r = r0 + it / nb, b = it % nb;
    if (b == 0) {
      const int f = a.found[r];
      tied = f == 2;
      cut_key = f != 0 ? a.prefix[r] : 0u;
      before = 0;
      if (tied) {
        need = int(a.target[r] - a.above[r]);
        for (int p2 = lane; p2 < piece; p2 += 32) before += a.eq_count[(long long)r * pieces + p2];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
      }
      if (piece == 0 && lane == 0) {
        a.cand[(long long)r * kCandFloats + 2 * kCandK] = a.row_M[r];
        a.cand[(long long)r * kCandFloats + 2 * kCandK + 1] = a.row_Z[r];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
    if (it + 1 < total) load(it + 1, nxt);
    const int q0 = wlo + b * 128;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * 32 + lane;
      const bool in = q < whi;
      const float xs[4] = {cur[u].x, cur[u].y, cur[u].z, cur[u].w};
      bool keep[4];
      if (!tied) {  // (warp-uniform)
#pragma unroll
        for (int j = 0; j < 4; ++j) keep[j] = in && 4 * q + j < a.vocab && f32_order_key(xs[j]) >= cut_key;
      } else {
        bool eq[4];
        uint32_t bal[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          eq[j] = in && 4 * q + j < a.vocab && f32_order_key(xs[j]) == cut_key;
          bal[j] = __ballot_sync(0xffffffffu, eq[j]);
        }
        int rank = before + __popc(bal[0] & lt) + __popc(bal[1] & lt) + __popc(bal[2] & lt) + __popc(bal[3] & lt);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          keep[j] = in && 4 * q + j < a.vocab && (f32_order_key(xs[j]) > cut_key || (eq[j] && rank < need));
          rank += eq[j] ? 1 : 0;
        }
        before += __popc(bal[0]) + __popc(bal[1]) + __popc(bal[2]) + __popc(bal[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!keep[j]) continue;
        const int slot = atomicAdd(a.cand_count + r, 1);
        if (slot < kCandK) {
          a.cand[(long long)r * kCandFloats + slot] = xs[j];
          a.cand[(long long)r * kCandFloats + kCandK + slot] = __int_as_float(a.vocab_offset + 4 * q + j);
        }
      }
    }
  }
  if (total == 0 && piece == 0 && lane == 0)
    for (int rr = 0; rr < n_rows; ++rr) {
      a.cand[(long long)(r0 + rr) * kCandFloats + 2 * kCandK] = a.row_M[r0 + rr];
      a.cand[(long long)(r0 + rr) * kCandFloats + 2 * kCandK + 1] = a.row_Z[r0 + rr];
    }
}

}  // namespace mtx
