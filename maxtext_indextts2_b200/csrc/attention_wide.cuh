// GQA decode attention for WIDE heads (head_dim 256: gemma3-1b/4b/12b), same contract, work list and cross-item merge as
// decode_attn_kernel (attention.cuh), different arithmetic layout.
//
// decode_attn_kernel keeps the group's query heads in the 16-row dimension of the MMA and D/2 accumulator registers per
// thread: 128 at D = 256, which does not fit beside the score and query fragments.  Here the products are TRANSPOSED, as in
// the persistent kernel (step_persistent.cuh): S^T = K Q^T and O^T = V^T P^T, so the 16-row dimension carries KV rows /
// head dims and the 8-column dimension the (up to 8) query heads of the group: D/4 accumulator registers, D/8 for the
// queries, no padding of the heads to 16.  A K or V tile of 64 rows is 32 KB at D = 256, so a CTA holds THREE warps' tiles
// (192 KB): kWideWarps.
#pragma once

#include "attention.cuh"

namespace mtx {

constexpr int kWideWarps = 3;
constexpr int kWideThreads = kWideWarps * 32;

__device__ __forceinline__ uint32_t aw_movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ void aw_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kWideThreads) : "memory"); }

__host__ inline size_t attn_wide_smem_bytes(int D) {
  return 1024 + size_t(kWideWarps) * 2 * (64 * D * 2) + size_t(kWideWarps) * 8 * D * 4 + 2 * kWideWarps * 8 + 2 * kWideWarps * 16 * 4 + 32;
}

template <int D>
__global__ void __launch_bounds__(kWideThreads)
decode_attn_wide_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const AttnParams p) {
  constexpr int kSub = D / 64;
  constexpr int kTileBytes = 64 * D * 2;
  constexpr float kLog2e = 1.4426950408889634f;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int G = p.hq / p.hkv;  // <= 8 (checked by the host)
  float* sm_o_all = reinterpret_cast<float*>(smem + kWideWarps * 2 * kTileBytes);  // [warp][8][D]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_o_all + kWideWarps * 8 * D);
  float* sm_m = reinterpret_cast<float*>(bars + 2 * kWideWarps);  // [warp][16]
  float* sm_l = sm_m + kWideWarps * 16;
  volatile int* s_last_p = reinterpret_cast<volatile int*>(sm_l + kWideWarps * 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  griddep_launch_dependents();
  if (lane == 0) {
    mbar_init(bars + 2 * warp, 1);
    mbar_init(bars + 2 * warp + 1, 1);
    fence_barrier_init();
  }
  if (tid == 0) {
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  __syncthreads();
  griddep_wait();

  const int gid = lane >> 2, tid4 = lane & 3;
  const int a_row = (lane & 7) + 8 * ((lane >> 3) & 1), a_chunk = lane >> 4;  // ldmatrix addressing, K tile as A
  const int v_row = (lane & 7) + 8 * (lane >> 4), v_chunk = (lane >> 3) & 1;   // ldmatrix.trans addressing, V^T as A
  const int R = p.ring_size > 0 ? p.ring_size : p.T - p.P;
  const int TPI = p.tiles_per_item;
  uint8_t* k_tile = smem + warp * 2 * kTileBytes;
  uint8_t* v_tile = k_tile + kTileBytes;
  float* sm_o = sm_o_all + warp * 8 * D;
  uint64_t* bar_k = bars + 2 * warp;
  uint64_t* bar_v = bar_k + 1;
  const uint32_t kb = smem_u32(k_tile), vb = smem_u32(v_tile);
  uint32_t phase = 0;
  const int n_items = *p.work_count * p.hkv;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int packed = p.work_items[item / p.hkv];
    const int h = item % p.hkv;
    const int r = packed >> 16, chunk = packed & 0xffff;
    const int len0 = p.len0[r], rf = p.ring_first[r], rl = p.ring_len[r];
    const int skip0 = p.skip0 + (p.skip0_rows != nullptr ? p.skip0_rows[r] : 0);
    const int nt = attn_num_tiles(len0, rf, rl, R);
    const int n_chunks = (nt + TPI - 1) / TPI;
    const int t_begin = chunk * TPI, t_end = min(nt, t_begin + TPI);
    const int plane_row = ((p.plane_base + p.plane[r]) * p.hkv + h) * p.T;

    int t = t_begin + warp;
    TileLoc loc = attn_tile(t < t_end ? t : nt, len0, rf, rl, p.P, R, skip0, p.ring_off);
    if (t < t_end && lane == 0) {
      mbar_expect_tx(bar_k, kTileBytes);
#pragma unroll
      for (int s = 0; s < kSub; ++s) tma_load_2d(k_tile + s * 8192, &tm_k, s * 64, plane_row + loc.p0, bar_k, kEvictFirst);
      mbar_expect_tx(bar_v, kTileBytes);
#pragma unroll
      for (int s = 0; s < kSub; ++s) tma_load_2d(v_tile + s * 8192, &tm_v, s * 64, plane_row + loc.p0, bar_v, kEvictFirst);
    }
    // Q^T fragments (B operand): thread (gid, tid4) holds head gid, dims 16 kk + 2 tid4 (+8)
    uint32_t qb[D / 16][2];
    {
      const bf16* qrow = p.q + (long long)r * p.hq * D + (long long)h * G * D + gid * D + tid4 * 2;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        qb[kk][0] = gid < G ? *reinterpret_cast<const uint32_t*>(qrow + kk * 16) : 0u;
        qb[kk][1] = gid < G ? *reinterpret_cast<const uint32_t*>(qrow + kk * 16 + 8) : 0u;
      }
    }
    // thread (gid, tid4) owns heads h0 = 2 tid4, h1 = h0 + 1:
    //   s[mb][0..3] = S^T[kv = 16 mb + gid (+8 for 2,3)][h0, h1];  o[db][0..3] = O^T[d = 16 db + gid (+8 for 2,3)][h0, h1]
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    float o[D / 16][4];
#pragma unroll
    for (int db = 0; db < D / 16; ++db) o[db][0] = o[db][1] = o[db][2] = o[db][3] = 0.0f;

    for (; t < t_end; t += kWideWarps) {
      const int cnt = loc.cnt;
      const int tn = t + kWideWarps;
      const TileLoc nloc = attn_tile(tn < t_end ? tn : nt, len0, rf, rl, p.P, R, skip0, p.ring_off);
      float s[4][4];
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) s[mb][0] = s[mb][1] = s[mb][2] = s[mb][3] = 0.0f;
      mbar_wait(bar_k, phase);
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
          const int row = 16 * mb + a_row;
          const int c = 2 * kk + a_chunk;
          const uint32_t addr = kb + (c >> 3) * 8192 + row * 128 + (((c & 7) ^ (row & 7)) << 4);
          uint32_t a[4];
          ldmatrix_x4(addr, a[0], a[1], a[2], a[3]);
          mma_m16n8k16_bf16(s[mb], a, qb[kk][0], qb[kk][1]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (tn < t_end && lane == 0) {
        mbar_expect_tx(bar_k, kTileBytes);
#pragma unroll
        for (int ss = 0; ss < kSub; ++ss) tma_load_2d(k_tile + ss * 8192, &tm_k, ss * 64, plane_row + nloc.p0, bar_k, kEvictFirst);
      }
      // ---- mask + online softmax: a head's scores live in the 8 threads of equal tid4 ----
      float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float x = s[mb][q];
          if (p.softcap != 0.0f) x = tanhf(x / p.softcap) * p.softcap;
          s[mb][q] = 16 * mb + gid + (q >> 1) * 8 < cnt ? x : -INFINITY;
        }
        tm0 = fmaxf(tm0, fmaxf(s[mb][0], s[mb][2]));
        tm1 = fmaxf(tm1, fmaxf(s[mb][1], s[mb][3]));
      }
#pragma unroll
      for (int sh = 4; sh < 32; sh <<= 1) {
        tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, sh));
        tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, sh));
      }
      const float nm0 = fmaxf(m0, tm0), nm1 = fmaxf(m1, tm1);  // (a tile holds at least one valid row: finite)
      const float a0 = exp2f((m0 - nm0) * kLog2e), a1 = exp2f((m1 - nm1) * kLog2e);  // exp2(-inf) = 0 on the first tile
      m0 = nm0;
      m1 = nm1;
      l0 *= a0;
      l1 *= a1;
      uint32_t pb[4][2];  // P^T as the B operand of O^T += V^T P^T, one k-step per 16 KV rows
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {
        const float p0 = exp2f((s[mb][0] - m0) * kLog2e), p1 = exp2f((s[mb][1] - m1) * kLog2e);
        const float p2 = exp2f((s[mb][2] - m0) * kLog2e), p3 = exp2f((s[mb][3] - m1) * kLog2e);
        l0 += p0 + p2;
        l1 += p1 + p3;
        // probabilities are cast to the value dtype before the PV product (kernels/ragged_attention.py:156)
        pb[mb][0] = aw_movmatrix_trans(pack_bf16x2(p0, p1));
        pb[mb][1] = aw_movmatrix_trans(pack_bf16x2(p2, p3));
      }
#pragma unroll
      for (int db = 0; db < D / 16; ++db) {
        o[db][0] *= a0;
        o[db][1] *= a1;
        o[db][2] *= a0;
        o[db][3] *= a1;
      }
      mbar_wait(bar_v, phase);
      if (cnt < 64) {  // rows past the valid count may hold anything: zero them (0 * NaN != 0)
        const int nvec = (64 - cnt) * 8;
        for (int i = lane; i < nvec * kSub; i += 32) {
          const int sub = i / nvec, w = i % nvec;
          *reinterpret_cast<uint4*>(v_tile + sub * 8192 + (cnt + w / 8) * 128 + (w & 7) * 16) = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
      }
#pragma unroll
      for (int db = 0; db < D / 16; ++db) {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
          const int row = 16 * mb + v_row;
          const int c = 2 * db + v_chunk;
          const uint32_t addr = vb + (c >> 3) * 8192 + row * 128 + (((c & 7) ^ (row & 7)) << 4);
          uint32_t a[4];
          ldmatrix_x4_trans(addr, a[0], a[1], a[2], a[3]);
          mma_m16n8k16_bf16(o[db], a, pb[mb][0], pb[mb][1]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (tn < t_end && lane == 0) {
        mbar_expect_tx(bar_v, kTileBytes);
#pragma unroll
        for (int ss = 0; ss < kSub; ++ss) tma_load_2d(v_tile + ss * 8192, &tm_v, ss * 64, plane_row + nloc.p0, bar_v, kEvictFirst);
      }
      phase ^= 1;
      loc = nloc;
    }
#pragma unroll
    for (int sh = 4; sh < 32; sh <<= 1) {
      l0 += __shfl_xor_sync(0xffffffffu, l0, sh);
      l1 += __shfl_xor_sync(0xffffffffu, l1, sh);
    }

    // ---- merge the warps of the item in shared memory: sm_o [head][D] ----
    const int h0 = tid4 * 2;
    if (m0 > -INFINITY) {
#pragma unroll
      for (int db = 0; db < D / 16; ++db) {
        const int d = 16 * db + gid;
        if (h0 < G) {
          sm_o[h0 * D + d] = o[db][0];
          sm_o[h0 * D + d + 8] = o[db][2];
        }
        if (h0 + 1 < G) {
          sm_o[(h0 + 1) * D + d] = o[db][1];
          sm_o[(h0 + 1) * D + d + 8] = o[db][3];
        }
      }
    }
    if (gid == 0) {
      sm_m[warp * 16 + h0] = m0;
      sm_m[warp * 16 + h0 + 1] = m1;
      sm_l[warp * 16 + h0] = l0;
      sm_l[warp * 16 + h0 + 1] = l1;
    }
    aw_bar_sync();

    const long long out_base = (long long)r * p.hq * D + (long long)h * G * D;
    const long long part_base = ((long long)(r * p.hkv + h) * p.max_chunks + chunk) * G;
    for (int e = tid; e < G * D; e += kWideThreads) {
      const int g = e / D, d = e - g * D;
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < kWideWarps; ++w) M = fmaxf(M, sm_m[w * 16 + g]);
      float L = 0.0f, O = 0.0f;
#pragma unroll
      for (int w = 0; w < kWideWarps; ++w) {
        const float mw = sm_m[w * 16 + g];
        if (mw > -INFINITY) {
          const float sc = exp2f((mw - M) * kLog2e);
          L += sm_l[w * 16 + g] * sc;
          O += sm_o_all[(w * 8 + g) * D + d] * sc;
        }
      }
      if (n_chunks == 1) {
        p.out[out_base + e] = __float2bfloat16_rn(O / L);
      } else {
        __stcg(p.part_o + (part_base + g) * D + d, O);
        if (d == 0) {
          __stcg(p.part_ml + (part_base + g) * 2, M);
          __stcg(p.part_ml + (part_base + g) * 2 + 1, L);
        }
      }
    }
    if (n_chunks > 1) {
      __threadfence();
      aw_bar_sync();
      if (tid == 0) {
        const int old = atomicAdd(p.tickets + r * p.hkv + h, 1);
        const int last = old == n_chunks - 1;
        if (last) p.tickets[r * p.hkv + h] = 0;
        *s_last_p = last;
      }
      aw_bar_sync();
      if (*s_last_p) {
        __threadfence();
        const long long pbase = (long long)(r * p.hkv + h) * p.max_chunks * G;
        for (int e = tid; e < G * D; e += kWideThreads) {
          const int g = e / D, d = e - g * D;
          float M = -INFINITY;
          for (int c = 0; c < n_chunks; ++c) M = fmaxf(M, __ldcg(p.part_ml + (pbase + c * G + g) * 2));
          float L = 0.0f, O = 0.0f;
          for (int c = 0; c < n_chunks; ++c) {
            const float sc = exp2f((__ldcg(p.part_ml + (pbase + c * G + g) * 2) - M) * kLog2e);
            L += __ldcg(p.part_ml + (pbase + c * G + g) * 2 + 1) * sc;
            O += __ldcg(p.part_o + (pbase + c * G + g) * D + d) * sc;
          }
          p.out[out_base + e] = __float2bfloat16_rn(O / L);
        }
      }
    }
    aw_bar_sync();  // the merge buffer is rewritten by the next item
  }
}

}  // namespace mtx
