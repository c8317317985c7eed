// Weight-streaming "skinny" GEMM on tcgen05 tensor cores with fused epilogues.
//
//   acc[n, r] = sum_k W[n, k] * X[r, k]        n in a 128-row tile of the weight matrix,
//                                               r over all rows of the step (UMMA N = r_tile)
//
// The weight matrix is the UMMA "A" operand (M = 128 output features per CTA), the
// activations are the "B" operand (N = rows, 16..256), so every weight byte is read from
// HBM exactly once per step and the batch rides along in the MMA N dimension.  Both
// operands are K-major bf16, staged by TMA (SWIZZLE_128B, 64-element k-blocks) into an
// mbarrier ring; one elected thread issues tcgen05.mma; the fp32 accumulator lives in
// TMEM and is read back by four epilogue warps with tcgen05.ld.
//
// Small matrices are split along K over several CTAs; partial tiles go to an L2-resident
// workspace and the last CTA to arrive (atomic ticket) sums them in split order -- the
// result is deterministic -- and runs the epilogue.
//
// Programmatic dependent launch: the producer prefetches weight tiles (which no other
// kernel writes) before griddepcontrol.wait, so the HBM stream of kernel n+1 starts while
// kernel n is still draining.
#pragma once

#include "common.cuh"

namespace mtx {

constexpr int kTileN = 128;    // weight rows (output features) per CTA = UMMA M
constexpr int kBlockK = 64;    // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;     // K of one tcgen05.mma for 16-bit inputs
constexpr int kWTileBytes = kTileN * kBlockK * 2;
constexpr int kMaxStages = 8;
constexpr int kGemmThreads = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int kEpiThreads = 128;

enum Epilogue : int {
  EPI_STORE_BF16 = 0,  // out[r, n] = bf16(acc)                               (mtx_linear, tests)
  EPI_QKV_ROPE = 1,    // RoPE on q,k; q -> buffer, k/v -> cache append
  EPI_RESIDUAL = 2,    // out[r, n] = bf16(resid[r, n] + bf16(acc))
  EPI_SWIGLU = 3,      // act[r, m] = bf16(bf16(silu(a)) * b), a/b rows interleaved by 16
  EPI_LOGITS = 4,      // logits (+optional store) and per-tile arg-max / log-sum-exp partials
};

struct GemmParams {
  int n;        // weight rows
  int k;        // reduction length (multiple of 64)
  int rows;     // valid activation rows
  int r_tile;   // UMMA N: rows rounded up to 16/32/64/128/256
  int splits;   // K splits (grid.y)
  int stages;   // smem ring depth
  float* ws;    // [n_tiles * splits][r_tile][128] fp32 partials (splits > 1)
  int* tickets; // [n_tiles] arrival counters, zero between launches
};

struct EpiArgs {
  // EPI_STORE_BF16 / EPI_RESIDUAL / EPI_SWIGLU
  bf16* out;          // [rows, ld_out]
  const bf16* resid;  // [rows, ld_out]
  int ld_out;
  // EPI_QKV_ROPE
  bf16* q_out;             // [rows, Hq*D]
  bf16* k_cache;           // layer base: [num_slots, Hkv, T, D]
  bf16* v_cache;
  const int* plane;        // [rows] KV plane of each row
  const int* write_row;    // [rows] cache row to append at, < 0 = skip
  const float2* rope_cs;   // [rows, D/2] (cos, sin) already rounded to bf16 values
  int hq, hkv, d, t_alloc;
  // EPI_LOGITS
  float* logits_out;   // [rows, ld_logits] fp32 or null
  long long ld_logits;
  int logits_only_row; // >= 0: store only this row, at logits_out[0 .. n)  (prefill: last position)
  float* part_score;   // [rows, n_tiles] best (possibly Gumbel-perturbed) score in the tile
  int* part_idx;       // [rows, n_tiles] its global vocab id
  float* part_raw;     // [rows, n_tiles] its unperturbed logit
  float* part_max;     // [rows, n_tiles] max logit in the tile
  float* part_sum;     // [rows, n_tiles] sum exp(logit - max)
  int n_tiles;
  int vocab_offset;
  float scale, softcap, inv_temp;
  int round_bf16;
  int gumbel;          // 1 = add Gumbel noise (weighted sampling)
  const uint32_t* rng_state;  // [4]: step, seed_lo, seed_hi, unused (device memory: graph replays see updates)
  int row_offset;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ---- epilogues: thread owns weight row n (n_local in the tile) and 16 consecutive rows r ----

__device__ __forceinline__ void epi_store_bf16(const EpiArgs& e, const GemmParams& p, const float (&v)[16], int r0, int n) {
  if (n >= p.n) return;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int r = r0 + j;
    if (r < p.rows) e.out[(long long)r * e.ld_out + n] = __float2bfloat16_rn(v[j]);
  }
}

__device__ __forceinline__ void epi_residual(const EpiArgs& e, const GemmParams& p, const float (&v)[16], int r0, int n) {
  if (n >= p.n) return;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int r = r0 + j;
    if (r < p.rows) {
      const long long o = (long long)r * e.ld_out + n;
      e.out[o] = __float2bfloat16_rn(__bfloat162float(e.resid[o]) + bf16r(v[j]));
    }
  }
}

// linears.py:425-476 with mlp_activations [silu, linear]; every intermediate is a bf16 array there.
__device__ __forceinline__ void epi_swiglu(const EpiArgs& e, const GemmParams& p, const float (&v)[16], int r0, int n, int lane) {
  const int m = (n >> 5) * 16 + (lane & 15);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float other = __shfl_xor_sync(0xffffffffu, v[j], 16);
    const float a = bf16r(lane < 16 ? v[j] : other);
    const float b = bf16r(lane < 16 ? other : v[j]);
    const float sg = bf16r(1.0f / (1.0f + expf(-a)));
    const float act = bf16r(a * sg);
    const int r = r0 + j;
    if (lane < 16 && r < p.rows && n < p.n) e.out[(long long)r * e.ld_out + m] = __float2bfloat16_rn(act * b);
  }
}

// embeddings.py:304-315 (half-split rotation, bf16 arithmetic) + kvcache.py:626-718 (append).
// `exch` is a [128][17] fp32 exchange buffer: the rotation partner of feature d is d +- D/2,
// which lives in another epilogue warp.
__device__ __forceinline__ void epi_qkv_rope(const EpiArgs& e, const GemmParams& p, const float (&v)[16], int r0, int n,
                                              int n_local, float* exch) {
  const int D = e.d, half = D >> 1;
  const int head = n / D, d = n - head * D;
  const bool is_q = head < e.hq, is_k = !is_q && head < e.hq + e.hkv;
  const bool rot = (is_q || is_k) && n < p.n;
  float own[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    own[j] = bf16r(v[j]);
    exch[n_local * 17 + j] = own[j];
  }
  epi_bar_sync();
  const bool first_half = d < half;
  const int partner = first_half ? n_local + half : n_local - half;
  const int fi = first_half ? d : d - half;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int r = r0 + j;
    if (r >= p.rows || n >= p.n) continue;
    float val = own[j];
    if (rot) {
      const float other = exch[partner * 17 + j];
      const float2 cs = e.rope_cs[r * half + fi];
      // first:  a*cos - b*sin   second: b*cos + a*sin   (own = a resp. b)
      const float t1 = bf16r(own[j] * cs.x), t2 = bf16r(other * cs.y);
      val = first_half ? (t1 - t2) : (t1 + t2);
    }
    const bf16 o = __float2bfloat16_rn(val);
    if (is_q) {
      e.q_out[(long long)r * (e.hq * D) + n] = o;
    } else {
      const int wr = e.write_row[r];
      if (wr >= 0) {
        const int kvh = is_k ? head - e.hq : head - e.hq - e.hkv;
        const long long off = (((long long)e.plane[r] * e.hkv + kvh) * e.t_alloc + wr) * D + d;
        (is_k ? e.k_cache : e.v_cache)[off] = o;
      }
    }
  }
  epi_bar_sync();
}

// decoders.py:537-589 (logits) + inference_utils.py:66-84 (greedy / weighted as Gumbel arg-max)
// + inference_utils.py:55-63 (log-sum-exp partials).  `red` is [4][16][5] fp32 scratch.
__device__ __forceinline__ void epi_logits(const EpiArgs& e, const GemmParams& p, const float (&v)[16], int r0, int n, int lane,
                                            int quarter, int epi_tid, int tile, float* red) {
  const bool valid = n < p.n;
  const int gid = e.vocab_offset + n;
  const uint32_t step = e.gumbel ? e.rng_state[0] : 0u;
  const uint64_t seed = e.gumbel ? (uint64_t(e.rng_state[2]) << 32) | e.rng_state[1] : 0ull;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int r = r0 + j;
    if (r >= p.rows) continue;  // warp-uniform: every lane of the warp shares r
    float lg = v[j];
    if (e.round_bf16) lg = bf16r(lg);
    lg *= e.scale;
    if (e.softcap != 0.0f) lg = tanhf(lg / e.softcap) * e.softcap;
    if (e.logits_out != nullptr && valid) {
      if (e.logits_only_row < 0) e.logits_out[(long long)r * e.ld_logits + n] = lg;
      else if (r == e.logits_only_row) e.logits_out[n] = lg;
    }
    float raw = valid ? lg : -INFINITY;
    float score = raw;
    if (e.gumbel && valid) score = lg * e.inv_temp + gumbel_noise(seed, step, uint32_t(e.row_offset + r), uint32_t(gid));
    int idx = gid;
    float mx = raw;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float s2 = __shfl_xor_sync(0xffffffffu, score, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
      const float r2 = __shfl_xor_sync(0xffffffffu, raw, o);
      if (s2 > score || (s2 == score && i2 < idx)) { score = s2; idx = i2; raw = r2; }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    float ex = valid ? expf(lg - mx) : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ex += __shfl_xor_sync(0xffffffffu, ex, o);
    if (lane == 0) {
      float* t = red + (quarter * 16 + j) * 5;
      t[0] = score; t[1] = __int_as_float(idx); t[2] = raw; t[3] = mx; t[4] = ex;
    }
  }
  epi_bar_sync();
  if (epi_tid < 16 && r0 + epi_tid < p.rows) {
    float score = -INFINITY, raw = -INFINITY, mx = -INFINITY, sum = 0.0f;
    int idx = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float* t = red + (q * 16 + epi_tid) * 5;
      const int i2 = __float_as_int(t[1]);
      if (t[0] > score || (t[0] == score && i2 < idx)) { score = t[0]; idx = i2; raw = t[2]; }
      const float m2 = fmaxf(mx, t[3]);
      if (m2 > -INFINITY) sum = sum * expf(mx - m2) + t[4] * expf(t[3] - m2);
      mx = m2;
    }
    const long long o = (long long)(r0 + epi_tid) * e.n_tiles + tile;
    e.part_score[o] = score; e.part_idx[o] = idx; e.part_raw[o] = raw; e.part_max[o] = mx; e.part_sum[o] = sum;
  }
  epi_bar_sync();
}

// ---- the kernel -------------------------------------------------------------------------

struct GemmSmemTail {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
  uint32_t is_last;
};

__host__ __device__ inline size_t gemm_smem_bytes(int stages, int r_tile) {
  return 1024 + size_t(stages) * (kWTileBytes + r_tile * kBlockK * 2) + sizeof(GemmSmemTail) + 16;
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x, const GemmParams p, const EpiArgs e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int x_tile_bytes = p.r_tile * kBlockK * 2;
  const int stage_bytes = kWTileBytes + x_tile_bytes;
  GemmSmemTail* tail = reinterpret_cast<GemmSmemTail*>(smem + size_t(p.stages) * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int n0 = tile * kTileN;
  const int kb_total = p.k / kBlockK;
  const int kb0 = int((long long)split * kb_total / p.splits);
  const int kb1 = int((long long)(split + 1) * kb_total / p.splits);
  const int nkb = kb1 - kb0;
  const uint32_t tmem_cols = p.r_tile < 32 ? 32u : uint32_t(p.r_tile);

  griddep_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    mbar_init(&tail->tmem_full, 1);
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

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int pre = nkb < p.stages ? nkb : p.stages;
      for (int i = 0; i < pre; ++i) {  // weights do not depend on the previous kernel
        mbar_expect_tx(&tail->full[i], uint32_t(stage_bytes));
        tma_load_2d(smem + size_t(i) * stage_bytes, &tm_w, (kb0 + i) * kBlockK, n0, &tail->full[i], kEvictFirst);
      }
      griddep_wait();
      for (int i = 0; i < pre; ++i)
        tma_load_2d(smem + size_t(i) * stage_bytes + kWTileBytes, &tm_x, (kb0 + i) * kBlockK, 0, &tail->full[i], kEvictLast);
      for (int i = pre; i < nkb; ++i) {
        const int s = i % p.stages;
        mbar_wait(&tail->empty[s], ((i / p.stages) & 1) ^ 1);
        mbar_expect_tx(&tail->full[s], uint32_t(stage_bytes));
        tma_load_2d(smem + size_t(s) * stage_bytes, &tm_w, (kb0 + i) * kBlockK, n0, &tail->full[s], kEvictFirst);
        tma_load_2d(smem + size_t(s) * stage_bytes + kWTileBytes, &tm_x, (kb0 + i) * kBlockK, 0, &tail->full[s], kEvictLast);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kTileN, p.r_tile);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % p.stages;
        mbar_wait(&tail->full[s], (i / p.stages) & 1);
        tcgen05_fence_after();
        const uint64_t da = umma_desc_sw128(smem + size_t(s) * stage_bytes);
        const uint64_t db = umma_desc_sw128(smem + size_t(s) * stage_bytes + kWTileBytes);
#pragma unroll
        for (int kk = 0; kk < kBlockK / kUmmaK; ++kk)
          umma_bf16(tmem_base, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc, uint32_t((i | kk) != 0));
        umma_commit(&tail->empty[s]);
      }
      umma_commit(&tail->tmem_full);
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 2..5 own TMEM lane quarters 2,3,0,1 =====
    const int quarter = warp & 3;
    const int n_local = quarter * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    const int n = n0 + n_local;
    float* scratch = reinterpret_cast<float*>(smem);  // stage memory is free once tmem_full fires
    mbar_wait(&tail->tmem_full, 0);
    tcgen05_fence_after();
    griddep_wait();
    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16);
    const int chunks = p.r_tile / 16;
    bool run_epilogue = true;
    if (p.splits > 1) {
      float* mine = p.ws + ((size_t)(tile * p.splits + split) * p.r_tile) * kTileN;
      for (int c = 0; c < chunks; ++c) {
        float v[16];
        tmem_ld_x16(taddr + uint32_t(c * 16), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) __stcg(mine + (size_t)(c * 16 + j) * kTileN + n_local, v[j]);
      }
      __threadfence();
      epi_bar_sync();
      if (epi_tid == 0) {
        const int old = atomicAdd(p.tickets + tile, 1);
        const bool last = old == p.splits - 1;
        if (last) p.tickets[tile] = 0;
        tail->is_last = last ? 1u : 0u;
      }
      epi_bar_sync();
      run_epilogue = tail->is_last != 0;
      if (run_epilogue) __threadfence();
    }
    if (run_epilogue) {
      for (int c = 0; c < chunks; ++c) {
        float v[16];
        if (p.splits > 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.0f;
          for (int s = 0; s < p.splits; ++s) {
            const float* part = p.ws + ((size_t)(tile * p.splits + s) * p.r_tile + c * 16) * kTileN + n_local;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += ldcg_f32(part + (size_t)j * kTileN);
          }
        } else {
          tmem_ld_x16(taddr + uint32_t(c * 16), v);
        }
        const int r0 = c * 16;
        if (EPI == EPI_STORE_BF16) epi_store_bf16(e, p, v, r0, n);
        if (EPI == EPI_RESIDUAL) epi_residual(e, p, v, r0, n);
        if (EPI == EPI_SWIGLU) epi_swiglu(e, p, v, r0, n, lane);
        if (EPI == EPI_QKV_ROPE) epi_qkv_rope(e, p, v, r0, n, n_local, scratch);
        if (EPI == EPI_LOGITS) epi_logits(e, p, v, r0, n, lane, quarter, epi_tid, tile, scratch);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mtx
