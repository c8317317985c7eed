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
        for (int s = 0; s < a.n_shards; ++s)
          trunc = trunc || a.gathered[((long long)s * a.rows + r) * kCandFloats + kCandK - 1] >= cut;
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
