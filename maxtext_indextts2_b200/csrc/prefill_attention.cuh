// Causal GQA attention for a chunk of prompt positions (MaxEngine._prefill_jit, MaxText/maxengine.py:400-530;
// AttentionOp.apply_attention_dot with the causal + segment mask of attentions.py:606-643).
//
// A prefill chunk is `rows` consecutive prompt positions [start_pos, start_pos + rows) of ONE sequence whose keys
// and values (the rows of earlier chunks and, written by the QKV epilogue just before, this chunk's own) sit in the
// prefill segment of one cache plane.  Round 1 ran every position as an independent decode row, so each of them
// streamed the whole prefix again (1 GB per layer for the last 256 positions of a 4k prompt).  Here the positions
// share the key / value tiles:
//
//   CTA = (query head, 16 query positions): a 16 x D query fragment (mma.sync m16n8k16 A operand: rows = positions).  The
//   four warps of the CTA split the visible 64-key tiles between them (tile t belongs to warp t % 4), each streaming its own
//   K and V tiles by TMA (the K buffer is refilled while the softmax and P V of the tile run, the V buffer after them) and
//   computing S = Q K^T (fp32), the causal mask, an online softmax (exp2, quad shuffles) and O += P V with the probabilities
//   rounded to bf16 (kernels/ragged_attention.py:156): the tile body of the decode kernel (attention.cuh).  The four partial
//   (O, max, sum) are merged in shared memory.  Tiles past the last visible key of the CTA's positions are never loaded.
//
// Hq x ceil(rows / 16) CTAs per layer and chunk (320 for a 256-row chunk of the IndexTTS2-scale model): the kernel is bound by
// the mma.sync rate of the SM sub-partitions, so what matters is that every sub-partition has a warp (the first version,
// one CTA per (kv head, 16 positions) with the group's heads as warps, kept 64 SMs busy and ran 2.5x longer).
#pragma once

#include "attention.cuh"

namespace mtx {

constexpr int kPfWarps = 4;   // key tiles in flight per CTA (one per warp)
constexpr int kPfQRows = 16;  // query positions per CTA

struct PrefillAttnParams {
  const bf16* q;   // [rows, Hq*D] rotated queries of the chunk
  bf16* out;       // [rows, Hq*D]
  int rows;        // positions in the chunk
  int start_pos;   // position of row 0
  int hq, hkv, T;
  int plane_row0;  // cache row of (layer, plane, kv head 0, position 0) in the K/V tensor maps; kv head h adds h * T
  float softcap;
  int window;      // sliding_window_size of a local layer (attentions.py:624-631: key in (position - window, position]); 0 = none
};

__host__ inline size_t prefill_attn_smem_bytes(int D) {
  return 1024 + size_t(kPfWarps) * 2 * (64 * D * 2) + size_t(kPfWarps) * kPfQRows * D * 4 + 2 * kPfWarps * 16 * 4 + 2 * kPfWarps * 8 + 64;
}

template <int D>
__global__ void __launch_bounds__(kPfWarps * 32)
prefill_attn_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const PrefillAttnParams p) {
  constexpr int kSub = D / 64;
  constexpr int kTileBytes = 64 * D * 2;
  constexpr float kLog2e = 1.4426950408889634f;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sm_o = reinterpret_cast<float*>(smem + size_t(kPfWarps) * 2 * kTileBytes);  // [warp][16][D]
  float* sm_m = sm_o + kPfWarps * kPfQRows * D;                                       // [warp][16]
  float* sm_l = sm_m + kPfWarps * 16;                                                 // [warp][16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_l + kPfWarps * 16);                 // [warp][K, V]

  const int G = p.hq / p.hkv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x;               // query head
  const int h = head / G;                    // its kv head
  const int q0 = blockIdx.y * kPfQRows;      // first chunk row of this CTA
  // keys visible to the CTA's last valid position: [0, start_pos + min(q0 + 16, rows))
  const int last_row = (q0 + kPfQRows < p.rows ? q0 + kPfQRows : p.rows) - 1;
  const int n_keys = p.start_pos + last_row + 1;
  const int n_tiles = (n_keys + 63) / 64;
  // sliding window: the lowest key the CTA's first position still sees
  const int first_key = p.window > 0 ? max(0, p.start_pos + q0 - p.window + 1) : 0;
  const int t_first = first_key / 64;
  const int plane_row = p.plane_row0 + h * p.T;
  uint8_t* k_tile = smem + size_t(warp) * 2 * kTileBytes;
  uint8_t* v_tile = k_tile + kTileBytes;
  uint64_t* bar_k = bars + 2 * warp;
  uint64_t* bar_v = bar_k + 1;
  const uint32_t kb = smem_u32(k_tile), vb = smem_u32(v_tile);

  const int tl = timeline_begin(40);
  griddep_launch_dependents();
  if (lane == 0) {
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    fence_barrier_init();
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  __syncthreads();
  griddep_wait();  // q, and this chunk's K / V rows, come from the QKV kernel before us

  const int gid = lane >> 2, tid4 = lane & 3;
  const int mtx_i = lane >> 3, lrow = lane & 7;
  const int r_lo = q0 + gid, r_hi = q0 + gid + 8;  // the two chunk rows of this thread's fragment rows
  const bool v_lo = r_lo < p.rows, v_hi = r_hi < p.rows;
  const int pos_lo = p.start_pos + r_lo, pos_hi = p.start_pos + r_hi;  // a key at index j is visible to a row iff j <= its position

  int t = t_first + warp;
  if (t < n_tiles && lane == 0) {
    mbar_expect_tx(bar_k, kTileBytes);
#pragma unroll
    for (int ss = 0; ss < kSub; ++ss) tma_load_2d(k_tile + ss * 8192, &tm_k, ss * 64, plane_row + t * 64, bar_k, kEvictLast);
    mbar_expect_tx(bar_v, kTileBytes);
#pragma unroll
    for (int ss = 0; ss < kSub; ++ss) tma_load_2d(v_tile + ss * 8192, &tm_v, ss * 64, plane_row + t * 64, bar_v, kEvictLast);
  }

  uint32_t qf[D / 16][4];
  {
    const bf16* qa = p.q + (long long)r_lo * p.hq * D + (long long)head * D;
    const bf16* qb = p.q + (long long)r_hi * p.hq * D + (long long)head * D;
#pragma unroll
    for (int tt = 0; tt < D / 16; ++tt) {
      const int d = tt * 16 + tid4 * 2;
      qf[tt][0] = v_lo ? *reinterpret_cast<const uint32_t*>(qa + d) : 0u;
      qf[tt][1] = v_hi ? *reinterpret_cast<const uint32_t*>(qb + d) : 0u;
      qf[tt][2] = v_lo ? *reinterpret_cast<const uint32_t*>(qa + d + 8) : 0u;
      qf[tt][3] = v_hi ? *reinterpret_cast<const uint32_t*>(qb + d + 8) : 0u;
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
  float o[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f;
  uint32_t phase = 0;

  for (; t < n_tiles; t += kPfWarps) {
    const int key0 = t * 64;
    const int tn = t + kPfWarps;
    mbar_wait(bar_k, phase);
    // ---- S = Q K^T ----
    float sc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.0f;
#pragma unroll
    for (int tt = 0; tt < D / 16; ++tt) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        const int row = 8 * (2 * jp + (mtx_i >> 1)) + lrow;
        const int c = 2 * tt + (mtx_i & 1);
        const uint32_t addr = kb + (c >> 3) * 8192 + row * 128 + (((c & 7) ^ (row & 7)) << 4);
        uint32_t b00, b01, b10, b11;
        ldmatrix_x4(addr, b00, b01, b10, b11);
        mma_m16n8k16_bf16(sc[2 * jp], qf[tt], b00, b01);
        mma_m16n8k16_bf16(sc[2 * jp + 1], qf[tt], b10, b11);
      }
    }
    // K tile consumed: refill it with this warp's next tile while the softmax and P V run
    fence_proxy_async();
    __syncwarp();
    if (tn < n_tiles && lane == 0) {
      mbar_expect_tx(bar_k, kTileBytes);
#pragma unroll
      for (int ss = 0; ss < kSub; ++ss) tma_load_2d(k_tile + ss * 8192, &tm_k, ss * 64, plane_row + tn * 64, bar_k, kEvictLast);
    }
    // ---- causal mask + online softmax ----
    const bool need_mask = key0 + 63 > p.start_pos + q0 || p.window > 0;  // some key of the tile lies beyond the CTA's first position
    float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float x = sc[j][e];
        if (p.softcap != 0.0f) x = tanhf(x / p.softcap) * p.softcap;
        if (need_mask) {
          const int key = key0 + 8 * j + tid4 * 2 + (e & 1);
          const int pos = (e & 2) ? pos_hi : pos_lo;
          if (key > pos || (p.window > 0 && key <= pos - p.window)) x = -INFINITY;
        }
        sc[j][e] = x;
      }
      tm0 = fmaxf(tm0, fmaxf(sc[j][0], sc[j][1]));
      tm1 = fmaxf(tm1, fmaxf(sc[j][2], sc[j][3]));
    }
    tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1));
    tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
    tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1));
    tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
    // A warp's tiles may lie entirely beyond an early position of the block (this warp's first tile is tile `warp`): the
    // row's running maximum then stays -inf, and (-inf) - (-inf) must not reach exp2: such a row keeps l = 0, O = 0 and
    // drops out of the merge below.
    const float nm0 = fmaxf(m0, tm0), nm1 = fmaxf(m1, tm1);
    const float a0 = nm0 == -INFINITY ? 1.0f : exp2f((m0 - nm0) * kLog2e), a1 = nm1 == -INFINITY ? 1.0f : exp2f((m1 - nm1) * kLog2e);
    m0 = nm0;
    m1 = nm1;
    l0 *= a0;
    l1 *= a1;
    const float ms0 = m0 == -INFINITY ? 0.0f : m0, ms1 = m1 == -INFINITY ? 0.0f : m1;
    uint32_t pa[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f((sc[j][0] - ms0) * kLog2e), p1 = exp2f((sc[j][1] - ms0) * kLog2e);
      const float p2 = exp2f((sc[j][2] - ms1) * kLog2e), p3 = exp2f((sc[j][3] - ms1) * kLog2e);
      l0 += p0 + p1;
      l1 += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int j = 0; j < D / 8; ++j) {
      o[j][0] *= a0;
      o[j][1] *= a0;
      o[j][2] *= a1;
      o[j][3] *= a1;
    }
    // ---- O += P V ----  (rows of the tile past the written keys hold whatever the cache held: their probabilities are
    // exactly 0, but 0 * NaN is NaN, so the cache must not hold NaNs: MaxEngine zeroes it at init_decode_state)
    mbar_wait(bar_v, phase);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int jp = 0; jp < D / 16; ++jp) {
        const int row = 16 * u + 8 * (mtx_i & 1) + lrow;
        const int c = 2 * jp + (mtx_i >> 1);
        const uint32_t addr = vb + (c >> 3) * 8192 + row * 128 + (((c & 7) ^ (row & 7)) << 4);
        uint32_t b00, b01, b10, b11;
        ldmatrix_x4_trans(addr, b00, b01, b10, b11);
        mma_m16n8k16_bf16(o[2 * jp], pa[u], b00, b01);
        mma_m16n8k16_bf16(o[2 * jp + 1], pa[u], b10, b11);
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (tn < n_tiles && lane == 0) {
      mbar_expect_tx(bar_v, kTileBytes);
#pragma unroll
      for (int ss = 0; ss < kSub; ++ss) tma_load_2d(v_tile + ss * 8192, &tm_v, ss * 64, plane_row + tn * 64, bar_v, kEvictLast);
    }
    phase ^= 1;
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // ---- merge the four warps' partials (a warp without tiles contributes max = -inf) ----
  {
    float* my_o = sm_o + warp * kPfQRows * D;
#pragma unroll
    for (int j = 0; j < D / 8; ++j) {
      const int d = 8 * j + tid4 * 2;
      *reinterpret_cast<float2*>(my_o + gid * D + d) = make_float2(o[j][0], o[j][1]);
      *reinterpret_cast<float2*>(my_o + (gid + 8) * D + d) = make_float2(o[j][2], o[j][3]);
    }
    if (tid4 == 0) {
      sm_m[warp * 16 + gid] = m0;
      sm_m[warp * 16 + gid + 8] = m1;
      sm_l[warp * 16 + gid] = l0;
      sm_l[warp * 16 + gid + 8] = l1;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kPfQRows * (D / 2); e += kPfWarps * 32) {
    const int row = e / (D / 2), d = (e - row * (D / 2)) * 2;
    const int r = q0 + row;
    if (r >= p.rows) continue;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kPfWarps; ++w) M = fmaxf(M, sm_m[w * 16 + row]);
    float L = 0.0f, O0 = 0.0f, O1 = 0.0f;
#pragma unroll
    for (int w = 0; w < kPfWarps; ++w) {
      const float mw = sm_m[w * 16 + row];
      if (mw > -INFINITY) {
        const float scl = exp2f((mw - M) * kLog2e);
        L += sm_l[w * 16 + row] * scl;
        const float2 ov = *reinterpret_cast<const float2*>(sm_o + (w * kPfQRows + row) * D + d);
        O0 += ov.x * scl;
        O1 += ov.y * scl;
      }
    }
    const float inv = 1.0f / L;
    *reinterpret_cast<uint32_t*>(p.out + (long long)r * p.hq * D + (long long)head * D + d) = pack_bf16x2(O0 * inv, O1 * inv);
  }
  timeline_end(tl);
}

}  // namespace mtx
