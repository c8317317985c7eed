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
  int mode;             // MTX_SAMPLE_NUCLEUS (2) or MTX_SAMPLE_TOPK (3)
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
  const float target = by_mass ? a.nucleus_p * z : float(a.top_k < V ? a.top_k : V);
  if (tid == 0) { s_prefix = 0u; s_above = 0.0f; s_found = 1; }
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
      const float sc = val * a.inv_temp + gumbel_noise(seed, step, uint32_t(a.row_offset + r), uint32_t(a.vocab_offset + i));
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

}  // namespace mtx
