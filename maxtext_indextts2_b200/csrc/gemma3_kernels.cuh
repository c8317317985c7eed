// Element-wise kernels of the Gemma-3 decoder block (MaxText/layers/gemma3.py:62-197) around the tensor-core GEMMs of
// gemm_rows.cuh.  The block differs from llama2 in what happens BETWEEN the projections: q and k are RMS-normalised over
// head_dim before RoPE and q is scaled after it (attentions.py:2246-2265), and the attention / MLP outputs are
// RMS-normalised before their residual adds (gemma3.py:142-170) -- a statistic over the whole output row, which a GEMM
// epilogue that sees one 128-feature tile cannot produce.  So the QKV, out-proj and MLP-down GEMMs store plain bf16
// (EPI_STORE_BF16, with the fused pre-norm's rstd) and these two kernels finish the job; both are one pass over a
// [rows, features] bf16 matrix that was just written (L2-resident).
#pragma once

#include "common.cuh"

namespace mtx {

struct QkNormArgs {
  const bf16* qkv;       // [rows, (Hq + 2 Hkv) D]  the QKV projection (bf16, pre-norm already applied)
  const bf16* q_scale;   // [D] self_attention/query_norm/scale of this layer
  const bf16* k_scale;   // [D] self_attention/key_norm/scale
  const float2* rope_cs; // [rows, D/2] (cos, sin), bf16-rounded, for this layer's RoPE base
  const int* plane;      // [rows]
  const int* write_row;  // [rows] cache row of the new K/V (< 0: not appended)
  bf16* q_out;           // [rows, Hq D]
  bf16* k_cache;         // this layer's [planes, Hkv, t_alloc, D]
  bf16* v_cache;
  int hq, hkv, d, t_alloc;
  float eps, q_scalar;   // normalization_layer_epsilon; query_pre_attn_scalar (0 / 1: none)
};

// One CTA per (row, head), D / 2 threads: thread i owns dims i and i + D/2 (the RoPE pair, embeddings.py:304-315).
//   q, k:  y = bf16(x * rsqrt(mean(x^2) + eps)); y = bf16(y * scale)          normalizations.py:57-69 over head_dim
//          RoPE with every product rounded to bf16 (bf16 arrays in the reference)
//   q:     bf16(q * query_pre_attn_scalar)                                      attentions.py:2263-2265
//   k, v:  appended to the cache row (kvcache.py:626-718)
__global__ void __launch_bounds__(128) qk_norm_rope_append_kernel(const QkNormArgs a) {
  __shared__ float s_part[4];
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x, head = blockIdx.y, i = threadIdx.x;
  const int D = a.d, half = D >> 1;
  const bool is_q = head < a.hq, is_k = !is_q && head < a.hq + a.hkv;
  const bf16* src = a.qkv + (long long)r * (a.hq + 2 * a.hkv) * D + (long long)head * D;
  float x0 = __bfloat162float(src[i]), x1 = __bfloat162float(src[i + half]);
  if (is_q || is_k) {
    float ss = x0 * x0 + x1 * x1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((i & 31) == 0) s_part[i >> 5] = ss;
    __syncthreads();
    float tot = 0.0f;
    for (int w = 0; w < (half + 31) / 32; ++w) tot += s_part[w];
    const float rstd = 1.0f / sqrtf(tot / float(D) + a.eps);
    const bf16* sc = is_q ? a.q_scale : a.k_scale;
    x0 = bf16r(bf16r(x0 * rstd) * __bfloat162float(sc[i]));
    x1 = bf16r(bf16r(x1 * rstd) * __bfloat162float(sc[i + half]));
    const float2 c = a.rope_cs[(long long)r * half + i];
    const float f = bf16r(bf16r(x0 * c.x) - bf16r(x1 * c.y));
    const float s = bf16r(bf16r(x1 * c.x) + bf16r(x0 * c.y));
    x0 = f;
    x1 = s;
    if (is_q && a.q_scalar != 0.0f && a.q_scalar != 1.0f) {
      x0 = bf16r(x0 * a.q_scalar);
      x1 = bf16r(x1 * a.q_scalar);
    }
  }
  if (is_q) {
    bf16* dst = a.q_out + (long long)r * a.hq * D + (long long)head * D;
    dst[i] = __float2bfloat16_rn(x0);
    dst[i + half] = __float2bfloat16_rn(x1);
  } else {
    const int wr = a.write_row[r];
    if (wr < 0) return;
    const int kvh = is_k ? head - a.hq : head - a.hq - a.hkv;
    bf16* dst = (is_k ? a.k_cache : a.v_cache) + (((long long)a.plane[r] * a.hkv + kvh) * a.t_alloc + wr) * D;
    dst[i] = __float2bfloat16_rn(x0);
    dst[i + half] = __float2bfloat16_rn(x1);
  }
}

// out[r, :] = bf16(resid[r, :] + RMSNorm(y[r, :]; scale))   (gemma3.py:142-153 / 161-170), and the per-128-feature-tile sums of
// squares of `out` that the next GEMM's fused pre-norm reads (ss[tile * ss_pitch + r], as gemm_rows.cuh's residual epilogue
// leaves them).  One CTA of 128 threads per row, two passes over the row (the first one for the statistic).
__global__ void __launch_bounds__(128)
post_norm_residual_kernel(const bf16* __restrict__ y, const bf16* __restrict__ resid, const bf16* __restrict__ scale, bf16* __restrict__ out,
                          float* __restrict__ ss, int ss_pitch, int E, float eps) {
  __shared__ float s_part[4];
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x;
  const int nvec = E / 8;
  const bf16* src = y + (long long)r * E;
  float acc = 0.0f;
  for (int i = threadIdx.x; i < nvec; i += 128) {
    const uint4 raw = *reinterpret_cast<const uint4*>(src + i * 8);
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
  const float rstd = 1.0f / sqrtf((s_part[0] + s_part[1] + s_part[2] + s_part[3]) / float(E) + eps);
  // vector i covers features 8 i .. 8 i + 7: 16 consecutive vectors (16 consecutive lanes) make one 128-feature tile
  for (int i0 = 0; i0 < nvec; i0 += 128) {
    const int i = i0 + threadIdx.x;
    float sq = 0.0f;
    if (i < nvec) {
      const uint4 raw = *reinterpret_cast<const uint4*>(src + i * 8);
      const uint4 sc = *reinterpret_cast<const uint4*>(scale + i * 8);
      const uint4 rs = *reinterpret_cast<const uint4*>(resid + (long long)r * E + i * 8);
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w}, s[4] = {sc.x, sc.y, sc.z, sc.w}, q[4] = {rs.x, rs.y, rs.z, rs.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float n0 = bf16r(bf16r(bf16_lo(w[j]) * rstd) * bf16_lo(s[j])), n1 = bf16r(bf16r(bf16_hi(w[j]) * rstd) * bf16_hi(s[j]));
        o[j] = pack_bf16x2(n0 + bf16_lo(q[j]), n1 + bf16_hi(q[j]));
        sq += bf16_lo(o[j]) * bf16_lo(o[j]) + bf16_hi(o[j]) * bf16_hi(o[j]);
      }
      *reinterpret_cast<uint4*>(out + (long long)r * E + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (ss != nullptr && (threadIdx.x & 15) == 0 && i < nvec) ss[(i >> 4) * ss_pitch + r] = sq;
  }
}

}  // namespace mtx
