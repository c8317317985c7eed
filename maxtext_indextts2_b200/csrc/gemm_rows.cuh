// Dense GEMM for steps of 65..256 rows (batch 128 / 256 decode, prefill chunks): the tensor-core regime of
// BASELINE configs[2].
//
//   acc[r, n] = sum_k X[r, k] * W[n, k]        r in a 128-row block of the step's rows (UMMA M = 128),
//                                               n in a 128-row tile of the weight matrix (UMMA N = 128)
//
// The skinny kernel (gemm_umma.cuh) keeps the weights in the UMMA M dimension and the batch in N; its epilogue
// thread owns ONE output feature and walks the rows, which is the right shape while a step has few rows and
// costs a dependent chain per row when it has hundreds (76 us for the MLP-up GEMM at 256 rows, round 1).  Here the
// roles are swapped: the ACTIVATIONS are the A operand, so a TMEM lane is a batch row and an epilogue thread owns
// one row with its output features in consecutive columns:
//   * RoPE partners (d +- D/2) and SwiGLU (gate, value) pairs are columns of the same thread: no shuffles, no
//     shared-memory exchange;
//   * every store is a 32-byte run of one output row; the logits scan (arg-max, log-sum-exp) is a register loop.
//
// Two modes.
//   persistent (splits == 1): grid = min(tiles, SMs); a CTA walks weight tiles cta, cta + grid, ...; the two
//     128-row blocks of a tile accumulate in TMEM (2 x 128 columns) and the accumulators are double-buffered
//     (512 columns), so the eight epilogue warps drain tile i while the MMAs of tile i + 1 run and the TMA ring
//     never stops at a tile boundary.
//   split-K (splits 2..8): one (tile, 128-row block, k-range) unit per CTA (grid = tiles x splits x row blocks), the
//     CTAs of one (tile, row block) form a thread-block cluster; partial accumulators are parked in shared memory and
//     CTA s reduces rows [s 128/S, (s+1) 128/S) of all S partials over distributed shared memory in fixed order
//     (deterministic), then runs the epilogue on its slice.  A unit needs < 113 KB of shared memory and 128 TMEM
//     columns, so the CTAs of the NEXT kernel become resident beside it and its prologue and weight loads (issued
//     before griddepcontrol.wait) overlap this kernel's reduction.
//
// Per k-block a CTA ingests 16 KB of weights and rows x 128 B of activations (TMA, SWIZZLE_128B); an SM ingests
// at most ~64 GB/s (profiles/r1f_tma_stream_subset.txt), which bounds the kernel at 256 rows.
#pragma once

#include "gemm_umma.cuh"

namespace mtx {

constexpr int kRowsThreads = 320;      // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-9: epilogue
constexpr int kRowsEpiWarps = 8;
constexpr int kRowsMaxStages = 6;
constexpr int kRowsPitch = 132;        // fp32 words per row of a parked partial tile

struct RowsSmemTail {
  uint64_t full[kRowsMaxStages];
  uint64_t empty[kRowsMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__host__ __device__ inline size_t rows_main_region_bytes(int stages, int r_tile, int splits) {
  size_t pipe = size_t(stages) * (kWTileBytes + r_tile * kBlockK * 2);
  size_t scratch = 0;
  if (splits > 1) scratch = size_t(r_tile) * kRowsPitch * 4 + size_t(r_tile / splits) * kRowsPitch * 4;
  size_t m = pipe > scratch ? pipe : scratch;
  if (splits == 1 && m < 120 * 1024) m = 120 * 1024;  // one CTA per SM: two could not both hold 512 TMEM columns
  return (m + 1023) / 1024 * 1024;
}
__host__ __device__ inline size_t rows_smem_bytes(int stages, int r_tile, int splits) {
  return 1024 + rows_main_region_bytes(stages, r_tile, splits) + sizeof(RowsSmemTail) + 16;
}

__device__ __forceinline__ void tmem_ld_x16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- row-major epilogue fragments: 16 consecutive output features n0 .. n0+15 of row r ------------------

__device__ __forceinline__ void frag_store_bf16(bf16* dst, const float (&v)[16]) {
  uint4 a, b;
  a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]); a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
  b.x = pack_bf16x2(v[8], v[9]); b.y = pack_bf16x2(v[10], v[11]); b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
}

// out[r, n] = bf16(acc)
__device__ __forceinline__ void rows_epi_store(const EpiArgs& e, int N, int r, int n0, const float (&v)[16]) {
  bf16* dst = e.out + (long long)r * e.ld_out + n0;
  if (n0 + 16 <= N && (e.ld_out & 7) == 0) {
    frag_store_bf16(dst, v);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < N) dst[j] = __float2bfloat16_rn(v[j]);
  }
}

// out[r, n] = bf16(resid[r, n] + bf16(acc))      (llama2.py:139-165 residual adds on bf16 arrays)
// Returns the sum of squares of the 16 stored (bf16) values: the next RMSNorm's statistic.
__device__ __forceinline__ float rows_epi_residual(const EpiArgs& e, int N, int r, int n0, const float (&v)[16]) {
  const long long o = (long long)r * e.ld_out + n0;
  float sq = 0.0f;
  if (n0 + 16 <= N && (e.ld_out & 7) == 0) {
    const uint4 ra = reinterpret_cast<const uint4*>(e.resid + o)[0], rb = reinterpret_cast<const uint4*>(e.resid + o)[1];
    const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
    float s[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t pk = pack_bf16x2(v[2 * j], v[2 * j + 1]);  // bf16(acc), one packed conversion
      s[2 * j] = bf16_lo(rw[j]) + bf16_lo(pk);
      s[2 * j + 1] = bf16_hi(rw[j]) + bf16_hi(pk);
    }
    frag_store_bf16(e.out + o, s);
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const uint32_t pk = pack_bf16x2(s[j], s[j + 1]);
      sq += bf16_lo(pk) * bf16_lo(pk) + bf16_hi(pk) * bf16_hi(pk);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < N) {
        const float y = bf16r(__bfloat162float(e.resid[o + j]) + bf16r(v[j]));
        e.out[o + j] = __float2bfloat16_rn(y);
        sq += y * y;
      }
  }
  return sq;
}

// rstd of row r from the per-tile partial sums of squares a residual epilogue left (normalizations.py:57-69).
__device__ __forceinline__ float rows_rstd(const EpiArgs& e, int r) {
  if (e.ss_in == nullptr) return 1.0f;
  float tot = 0.0f;
  for (int t = 0; t < e.ss_tiles; ++t) tot += __ldcg(e.ss_in + t * e.ss_pitch + r);
  return 1.0f / sqrtf(tot / float(e.ss_dim) + e.ss_eps);
}
__device__ __forceinline__ void rows_scale16(float (&v)[16], float s) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] *= s;
}

// linears.py:425-476, mlp_activations [silu, linear]; w01 rows are interleaved 16 at a time (gate, value), so the
// fragment at n0 (a multiple of 32) is the gate and the one at n0 + 16 the value of MLP features n0/2 .. n0/2+15.
__device__ __forceinline__ void rows_epi_swiglu(const EpiArgs& e, int N2, int r, int n0, const float (&g)[16], const float (&u)[16]) {
  if (n0 >= N2) return;
  float o[16];
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const uint32_t ga = pack_bf16x2(g[j], g[j + 1]), ub = pack_bf16x2(u[j], u[j + 1]);
    const float a0 = bf16_lo(ga), a1 = bf16_hi(ga);
    // gate: silu(a) = a * sigmoid(a), or (gemma3, linears.py:460 with "gelu" = flax nn.gelu) a * 0.5 (1 + tanh(sqrt(2/pi) (a + 0.044715 a^3)))
    const uint32_t sg = e.act_gelu ? pack_bf16x2(0.5f * (1.0f + tanhf(0.7978845608028654f * (a0 + 0.044715f * a0 * a0 * a0))),
                                                 0.5f * (1.0f + tanhf(0.7978845608028654f * (a1 + 0.044715f * a1 * a1 * a1))))
                                   : pack_bf16x2(__fdividef(1.0f, 1.0f + __expf(-a0)), __fdividef(1.0f, 1.0f + __expf(-a1)));
    const uint32_t act = pack_bf16x2(a0 * bf16_lo(sg), a1 * bf16_hi(sg));
    o[j] = bf16_lo(act) * bf16_lo(ub);
    o[j + 1] = bf16_hi(act) * bf16_hi(ub);
  }
  frag_store_bf16(e.out + (long long)r * e.ld_out + (n0 >> 1), o);
}

// embeddings.py:304-315 + kvcache.py:626-718: `lo` = features d0 .. d0+15 of a head's first half, `hi` = the same
// dims of its second half (d0 + D/2).  Query / key heads are rotated, value heads stored as they are.
__device__ __forceinline__ void rows_epi_qkv(const EpiArgs& e, int N, int r, int head, int d0, const float (&lo)[16], const float (&hi)[16]) {
  const int D = e.d, half = D >> 1;
  if (head * D >= N) return;
  const bool is_q = head < e.hq, is_k = !is_q && head < e.hq + e.hkv;
  float a[16], b[16];
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const uint32_t pa = pack_bf16x2(lo[j], lo[j + 1]), pb = pack_bf16x2(hi[j], hi[j + 1]);
    a[j] = bf16_lo(pa); a[j + 1] = bf16_hi(pa);
    b[j] = bf16_lo(pb); b[j + 1] = bf16_hi(pb);
  }
  if (is_q || is_k) {
    const float2* cs = e.rope_cs + (long long)r * half + d0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float2 c = cs[j];
      // first half: a cos - b sin ; second half: b cos + a sin, every product rounded to bf16 (bf16 arrays in the reference)
      const float ac = bf16r(a[j] * c.x), bs = bf16r(b[j] * c.y), bc = bf16r(b[j] * c.x), as = bf16r(a[j] * c.y);
      a[j] = ac - bs;
      b[j] = bc + as;
    }
  }
  if (is_q) {
    bf16* dst = e.q_out + (long long)r * (e.hq * D) + head * D + d0;
    frag_store_bf16(dst, a);
    frag_store_bf16(dst + half, b);
  } else {
    const int wr = e.write_row[r];
    if (wr < 0) return;
    const int kvh = is_k ? head - e.hq : head - e.hq - e.hkv;
    bf16* dst = (is_k ? e.k_cache : e.v_cache) + (((long long)e.plane[r] * e.hkv + kvh) * e.t_alloc + wr) * D + d0;
    frag_store_bf16(dst, a);
    frag_store_bf16(dst + half, b);
  }
}

// One head (64 dims) of one row for the int8 cache: RoPE (key heads), then KVQuant.quantize over the head's dims.
// x0 = dims 0..15, x1 = 16..31 (first half), y0 = 32..47, y1 = 48..63 (second half), accumulator values.
__device__ __forceinline__ void rows_epi_qkv_q8(const EpiArgs& e, int N, int r, int head, float (&x0)[16], float (&x1)[16], float (&y0)[16], float (&y1)[16]) {
  if (head * 64 >= N) return;
  const bool is_q = head < e.hq, is_k = !is_q && head < e.hq + e.hkv;
  if (is_q) {  // queries are not cached: the bf16 path
    rows_epi_qkv(e, N, r, head, 0, x0, y0);
    rows_epi_qkv(e, N, r, head, 16, x1, y1);
    return;
  }
  float* part[4] = {x0, x1, y0, y1};
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 16; ++j) part[q][j] = bf16r(part[q][j]);
  if (is_k) {
    const float2* cs = e.rope_cs + (long long)r * 32;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 c = cs[16 * q + j];
        const float a = part[q][j], b = part[q + 2][j];
        part[q][j] = bf16r(bf16r(a * c.x) - bf16r(b * c.y));      // the cached key is a bf16 array (embeddings.py:304-315)
        part[q + 2][j] = bf16r(bf16r(b * c.x) + bf16r(a * c.y));
      }
  }
  const int wr = e.write_row[r];
  if (wr < 0) return;
  float mx = 0.0f;
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fabsf(part[q][j]));
  const bool fp8 = e.kv_fp8 != 0;
  const float inv = mx > 0.0f ? kv_quant_max(fp8) / mx : 0.0f;
  const int kvh = is_k ? head - e.hq : head - e.hq - e.hkv;
  const long long row = ((long long)e.plane[r] * e.hkv + kvh) * e.t_alloc + wr;
  uint8_t* dst = (is_k ? e.kq_cache : e.vq_cache) + row * 64;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      w[i] = kv_quant_pair(part[q][4 * i], part[q][4 * i + 1], inv, fp8) | (kv_quant_pair(part[q][4 * i + 2], part[q][4 * i + 3], inv, fp8) << 16);
    *reinterpret_cast<uint4*>(dst + 16 * q) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  (is_k ? e.k_scale : e.v_scale)[row] = mx;
}

// Running sampling partials of one row over the columns of one weight tile (decoders.py:537-589 transform, then
// inference_utils.py:55-84): best (optionally Gumbel-perturbed) score with the lowest index on ties, its raw logit,
// and the (max, sum exp) pair.
struct RowsLogitScan {
  float best, raw, mx, sum;
  int bi;
};
__device__ __forceinline__ void rows_logits_frag(const EpiArgs& e, int V, int r, bool row_valid, int n0, const float (&v)[16], RowsLogitScan& s,
                                                 uint32_t step, uint64_t seed) {
  float lg[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) lg[j] = logit_transform(e, v[j]);
  if (row_valid && e.logits_out != nullptr) {
    float* dst = e.logits_only_row < 0 ? e.logits_out + (long long)r * e.ld_logits + n0 : (r == e.logits_only_row ? e.logits_out + n0 : nullptr);
    if (dst != nullptr) {
      if (n0 + 16 <= V && (e.ld_logits & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(lg[j], lg[j + 1], lg[j + 2], lg[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + j < V) dst[j] = lg[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (n0 + j < V) {
      const float sc = e.gumbel ? lg[j] * e.inv_temp + gumbel_noise(seed, step, uint32_t(e.row_offset + r), uint32_t(e.vocab_offset + n0 + j)) : lg[j];
      if (sc > s.best) { s.best = sc; s.raw = lg[j]; s.bi = n0 + j; }  // ascending scan: strict > keeps the lowest index
      if (e.want_lse) {
        const float mn = fmaxf(s.mx, lg[j]);
        s.sum = s.sum * __expf(s.mx - mn) + __expf(lg[j] - mn);
        s.mx = mn;
      }
    }
  }
}

// ---- the kernel -------------------------------------------------------------------------------------------

template <int EPI>
__global__ void __launch_bounds__(kRowsThreads, 1)
gemm_rows_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x, const GemmParams p, const EpiArgs e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int x_bytes = p.r_tile * kBlockK * 2;
  const int stage_bytes = kWTileBytes + x_bytes;
  RowsSmemTail* tail = reinterpret_cast<RowsSmemTail*>(smem + rows_main_region_bytes(p.stages, p.r_tile, p.splits));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.splits > 1;
  const int n_tiles = (p.n + kTileN - 1) / kTileN;
  const int kb_total = p.k / kBlockK;
  const int mblocks = p.r_tile / 128;  // 1 or 2 blocks of 128 rows (split mode: p.r_tile = 128, the row block is blockIdx.z)
  const int row_base = split ? int(blockIdx.z) * 128 : 0;
  // work of this CTA: split mode = one (tile, k-range) unit; persistent mode = tiles first, first + stride, ...
  const int first_tile = int(blockIdx.x), tile_stride = split ? n_tiles : int(gridDim.x);
  const int kb0 = split ? int((long long)blockIdx.y * kb_total / p.splits) : 0;
  const int kb1 = split ? int((long long)(blockIdx.y + 1) * kb_total / p.splits) : kb_total;
  const int nkb = kb1 - kb0;
  const int my_tiles = first_tile < n_tiles ? (n_tiles - first_tile + tile_stride - 1) / tile_stride : 0;
  const int acc_cols = mblocks * 128;                    // TMEM columns of one tile's accumulators
  const uint32_t tmem_cols = uint32_t(split ? acc_cols : 512);

  const int tl = timeline_begin(30 + EPI);
  griddep_launch_dependents();
  // debug: clock64 timeline of CTA (0,0,0): [0] setup done, [1] producer past griddep wait, [2] first stage landed, [3] MMAs committed,
  // [4] accumulator ready, [5] parked, [6] after cluster barrier, [7] reduced, [8] epilogue done, [9] after 2nd cluster barrier
  long long* trace = (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? p.trace : nullptr;
  const long long t_start = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->tmem_full[b], 1);
      mbar_init(&tail->tmem_empty[b], uint32_t(4 * mblocks));  // one arrival per epilogue warp that reads the buffer
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tail->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tail->tmem_base;
  if (trace && threadIdx.x == 0) trace[0] = clock64() - t_start;

  const int F = my_tiles * nkb;  // ring fills of this CTA: fill f = (tile f / nkb, k-block kb0 + f % nkb)

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      auto w_load = [&](int f) {
        const int s = f % p.stages, n0 = (first_tile + (f / nkb) * tile_stride) * kTileN, kb = kb0 + f % nkb;
        mbar_expect_tx(&tail->full[s], uint32_t(stage_bytes));
        tma_load_2d(smem + size_t(s) * stage_bytes, &tm_w, kb * kBlockK, n0, &tail->full[s], kEvictFirst);
      };
      auto x_load = [&](int f) {
        const int s = f % p.stages, kb = kb0 + f % nkb;
        tma_load_2d(smem + size_t(s) * stage_bytes + kWTileBytes, &tm_x, kb * kBlockK, row_base, &tail->full[s], kEvictLast);
      };
      const int pre = F < p.stages ? F : p.stages;
      for (int f = 0; f < pre; ++f) w_load(f);  // weights do not depend on the previous kernel
      griddep_wait();
      if (trace) trace[1] = clock64() - t_start;
      for (int f = 0; f < pre; ++f) x_load(f);
      for (int f = pre; f < F; ++f) {
        mbar_wait(&tail->empty[f % p.stages], ((f / p.stages) & 1) ^ 1);
        w_load(f);
        x_load(f);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, kTileN);
      int f = 0;
      for (int j = 0; j < my_tiles; ++j) {
        const uint32_t buf = split ? 0u : uint32_t(j & 1);
        if (j >= 2) mbar_wait(&tail->tmem_empty[buf], ((j >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        for (int i = 0; i < nkb; ++i, ++f) {
          const int s = f % p.stages;
          mbar_wait(&tail->full[s], (f / p.stages) & 1);
          tcgen05_fence_after();
          if (trace && f == 0) trace[2] = clock64() - t_start;
          const uint64_t dw = umma_desc_sw128(smem + size_t(s) * stage_bytes);
          for (int mb = 0; mb < mblocks; ++mb) {
            const uint64_t dx = umma_desc_sw128(smem + size_t(s) * stage_bytes + kWTileBytes + size_t(mb) * 128 * kBlockK * 2);
            const uint32_t acc = tmem_base + buf * uint32_t(acc_cols) + uint32_t(mb) * 128u;
#pragma unroll
            for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) umma_bf16(acc, dx + uint64_t(kk * 2), dw + uint64_t(kk * 2), idesc, uint32_t((i | kk) != 0));
          }
          umma_commit(&tail->empty[s]);
        }
        umma_commit(&tail->tmem_full[buf]);
      }
      if (trace) trace[3] = clock64() - t_start;
    }
    __syncwarp();
  } else {
    // ===== epilogue: eight warps; warp w reads TMEM lanes 32 (w % 4) .., warps of quarter pairs split the row blocks =====
    const int ew = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // the TMEM lane quarter this warp may read
    const int mb = ew >> 2;                  // row block of this warp
    const int r = mb * 128 + quarter * 32 + lane;
    const bool row_valid = r < p.rows;
    const bool active = mb < mblocks;
    uint32_t step = 0;
    uint64_t seed = 0;
    if (EPI == EPI_LOGITS && e.gumbel) {
      step = e.rng_state[0];
      seed = (uint64_t(e.rng_state[2]) << 32) | e.rng_state[1];
    }

    // One tile of this thread's row, columns read through `ld16(col, dst)` (TMEM or the reduced slice in shared memory).
    auto run_tile = [&](int tile, int row, bool valid, float rs, auto&& ld16, int c_begin, int c_end) {
      const int nbase = tile * kTileN;
      if (EPI == EPI_STORE_BF16 || EPI == EPI_RESIDUAL) {
        float sq = 0.0f;
        for (int c = c_begin; c < c_end; c += 16) {
          float v[16];
          ld16(c, v);
          if (!valid || nbase + c >= p.n) continue;
          if (EPI == EPI_STORE_BF16) {
            rows_scale16(v, rs);  // (1 unless the input's RMSNorm is fused: ss_in)
            rows_epi_store(e, p.n, row, nbase + c, v);
          } else {
            sq += rows_epi_residual(e, p.n, row, nbase + c, v);
          }
        }
        if (EPI == EPI_RESIDUAL && e.ss_out != nullptr && valid) e.ss_out[tile * e.ss_pitch + row] = sq;
      } else if (EPI == EPI_SWIGLU) {
        for (int c = c_begin; c < c_end; c += 32) {
          float g[16], u[16];
          ld16(c, g);
          ld16(c + 16, u);
          rows_scale16(g, rs);
          rows_scale16(u, rs);
          if (valid) rows_epi_swiglu(e, p.n, row, nbase + c, g, u);
        }
      } else if (EPI == EPI_QKV_ROPE) {
        const int D = e.d, half = D >> 1;
        if (e.kq_cache != nullptr) {  // int8 cache (head_dim 64): a whole head per step, so that its scale is one thread's business
          for (int c = c_begin; c < c_end; c += 64) {
            float x0[16], x1[16], y0[16], y1[16];
            ld16(c, x0);
            ld16(c + 16, x1);
            ld16(c + 32, y0);
            ld16(c + 48, y1);
            rows_scale16(x0, rs);
            rows_scale16(x1, rs);
            rows_scale16(y0, rs);
            rows_scale16(y1, rs);
            if (valid) rows_epi_qkv_q8(e, p.n, row, (nbase + c) / 64, x0, x1, y0, y1);
          }
          return;
        }
        for (int c = c_begin; c < c_end; c += 16) {
          const int d = (nbase + c) % D;
          if (d >= half) continue;  // second halves are produced together with their first halves
          float lo[16], hi[16];
          ld16(c, lo);
          ld16(c + half, hi);
          rows_scale16(lo, rs);
          rows_scale16(hi, rs);
          if (valid) rows_epi_qkv(e, p.n, row, (nbase + c) / D, d, lo, hi);
        }
      }
    };

    if (!split) {
      float rs = 1.0f;
      for (int j = 0; j < my_tiles; ++j) {
        const int tile = first_tile + j * tile_stride;
        const uint32_t buf = uint32_t(j & 1);
        if (active) {
          mbar_wait(&tail->tmem_full[buf], (j >> 1) & 1);
          tcgen05_fence_after();
          if (trace && threadIdx.x == 64 && j == 0) trace[4] = clock64() - t_start;
          if (j == 0) {
            griddep_wait();  // side inputs (residual, row descriptors, row statistics) come from earlier kernels
            if (row_valid) rs = rows_rstd(e, r);
          }
          const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + buf * uint32_t(acc_cols) + uint32_t(mb) * 128u;
          auto ld16 = [&](int c, float (&v)[16]) { tmem_ld_x16(taddr + uint32_t(c), v); };
          if (EPI == EPI_LOGITS) {
            RowsLogitScan sc;
            sc.best = -INFINITY; sc.raw = -INFINITY; sc.mx = -INFINITY; sc.sum = 0.0f; sc.bi = 0x7fffffff;
            for (int c = 0; c < kTileN; c += 16) {
              float v[16];
              ld16(c, v);
              rows_scale16(v, rs);
              rows_logits_frag(e, p.n, r, row_valid, tile * kTileN + c, v, sc, step, seed);
            }
            if (row_valid) {
              const long long o = (long long)r * e.n_tiles + tile;
              e.part_score[o] = sc.best;
              e.part_idx[o] = e.vocab_offset + sc.bi;
              e.part_raw[o] = sc.raw;
              if (e.want_lse) { e.part_max[o] = sc.mx; e.part_sum[o] = sc.sum; }
            }
          } else {
            run_tile(tile, r, row_valid, rs, ld16, 0, kTileN);
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tail->tmem_empty[buf]);
          if (trace && threadIdx.x == 64 && j == 0) trace[8] = clock64() - t_start;
        }
      }
    } else {
      // ---- split-K: park the partial accumulators, cluster barrier, reduce-scatter over DSMEM ----
      float* partial = reinterpret_cast<float*>(smem);         // [r_tile][kRowsPitch]; the ring is idle once tmem_full fires
      if (active && my_tiles > 0) {
        mbar_wait(&tail->tmem_full[0], 0);
        tcgen05_fence_after();
        if (trace && threadIdx.x == 64) trace[4] = clock64() - t_start;
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(mb) * 128u;
        for (int c = 0; c < kTileN; c += 16) {
          float v[16];
          tmem_ld_x16(taddr + uint32_t(c), v);
#pragma unroll
          for (int q = 0; q < 16; q += 4) *reinterpret_cast<float4*>(partial + r * kRowsPitch + c + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
        }
      }
      tcgen05_fence_before();
      if (trace && threadIdx.x == 64) trace[5] = clock64() - t_start;
    }
  }
  if (split) {
    cluster_sync_all();
    if (trace && threadIdx.x == 64) trace[6] = clock64() - t_start;
    if (warp >= 2 && my_tiles > 0) {
      float* partial = reinterpret_cast<float*>(smem);
      float* reduced = partial + p.r_tile * kRowsPitch;
      const int epi_tid = threadIdx.x - 64;
      const int rpc = p.r_tile / p.splits;  // rows reduced and finished by this CTA
      const int r_begin = int(blockIdx.y) * rpc;  // within the row block
      const uint32_t my = smem_u32(partial);
      // pull: one float4 column group of one row per thread and peer, all peers in flight; summed in split order
      for (int u = epi_tid; u < rpc * 32; u += kRowsEpiWarps * 32) {
        const int rr = u >> 5, c4 = u & 31;
        const uint32_t off = uint32_t(((r_begin + rr) * kRowsPitch + c4 * 4) * 4);
        float4 t[8];
#pragma unroll
        for (int ss = 0; ss < 8; ++ss) {
          t[ss] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ss < p.splits) t[ss] = ld_dsmem_f32x4(dsmem_addr(my, uint32_t(ss)) + off);
        }
        float4 acc = t[0];
#pragma unroll
        for (int ss = 1; ss < 8; ++ss) { acc.x += t[ss].x; acc.y += t[ss].y; acc.z += t[ss].z; acc.w += t[ss].w; }
        *reinterpret_cast<float4*>(reduced + rr * kRowsPitch + c4 * 4) = acc;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (trace && threadIdx.x == 64) trace[7] = clock64() - t_start;
      griddep_wait();
      // epilogue: thread -> (row of the slice, 16-column chunk); RoPE / SwiGLU partners are read from the same row
      const int tile = first_tile;
      if (EPI == EPI_QKV_ROPE && e.kq_cache != nullptr) {  // int8 cache: thread -> (row of the slice, one of the tile's two heads)
        for (int u = epi_tid; u < rpc * 2; u += kRowsEpiWarps * 32) {
          const int rr = u >> 1, c0 = (u & 1) * 64;
          const int row = row_base + r_begin + rr;
          if (row >= p.rows) continue;
          const float* src = reduced + rr * kRowsPitch + c0;
          const float rs = rows_rstd(e, row);
          float x[4][16];
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 16; ++j) x[q][j] = src[16 * q + j] * rs;
          rows_epi_qkv_q8(e, p.n, row, (tile * kTileN + c0) / 64, x[0], x[1], x[2], x[3]);
        }
      } else
      for (int u = epi_tid; u < rpc * 8; u += kRowsEpiWarps * 32) {
        const int rr = u >> 3, c0 = (u & 7) * 16;
        const int row = row_base + r_begin + rr;
        const float* src = reduced + rr * kRowsPitch;
        auto ld16 = [&](int c, float (&v)[16]) {
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const float4 x = *reinterpret_cast<const float4*>(src + c + q);
            v[q] = x.x; v[q + 1] = x.y; v[q + 2] = x.z; v[q + 3] = x.w;
          }
        };
        const bool valid = row < p.rows;
        const float rs = (valid && (EPI == EPI_SWIGLU || EPI == EPI_QKV_ROPE || EPI == EPI_STORE_BF16)) ? rows_rstd(e, row) : 1.0f;
        if (EPI == EPI_SWIGLU) {
          if ((c0 & 16) == 0) {
            float g[16], uu[16];
            ld16(c0, g);
            ld16(c0 + 16, uu);
            rows_scale16(g, rs);
            rows_scale16(uu, rs);
            if (valid) rows_epi_swiglu(e, p.n, row, tile * kTileN + c0, g, uu);
          }
        } else if (EPI == EPI_QKV_ROPE) {
          const int D = e.d, half = D >> 1;
          const int d = (tile * kTileN + c0) % D;
          if (d < half) {
            float lo[16], hi[16];
            ld16(c0, lo);
            ld16(c0 + half, hi);
            rows_scale16(lo, rs);
            rows_scale16(hi, rs);
            if (valid) rows_epi_qkv(e, p.n, row, (tile * kTileN + c0) / D, d, lo, hi);
          }
        } else if (EPI == EPI_STORE_BF16 || EPI == EPI_RESIDUAL) {
          float v[16];
          ld16(c0, v);
          float sq = 0.0f;
          if (valid && tile * kTileN + c0 < p.n) {
            if (EPI == EPI_STORE_BF16) {
              rows_scale16(v, rs);
              rows_epi_store(e, p.n, row, tile * kTileN + c0, v);
            } else {
              sq = rows_epi_residual(e, p.n, row, tile * kTileN + c0, v);
            }
          }
          if (EPI == EPI_RESIDUAL && e.ss_out != nullptr) {
            // the eight chunks of a row sit in eight consecutive lanes (the trip count is warp-uniform: rpc * 8 is a multiple of 32)
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            sq += __shfl_xor_sync(0xffffffffu, sq, 4);
            if ((u & 7) == 0 && valid) e.ss_out[tile * e.ss_pitch + row] = sq;
          }
        }
      }
    }
    if (trace && threadIdx.x == 64) trace[8] = clock64() - t_start;
    cluster_sync_all();  // keep every CTA's partial alive until all peers have read it
    if (trace && threadIdx.x == 64) trace[9] = clock64() - t_start;
  }

  tcgen05_fence_before();
  __syncthreads();
  if (trace && threadIdx.x == 0) trace[10] = clock64() - t_start;
  timeline_end(tl);
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mtx
