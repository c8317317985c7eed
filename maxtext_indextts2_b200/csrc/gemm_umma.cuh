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
// Small matrices are split along K over the CTAs of one thread-block cluster.  Each CTA
// parks its partial accumulator tile in its own shared memory; after a cluster barrier CTA c
// sums rows [c*R/S, (c+1)*R/S) of all S partials over distributed shared memory (fixed
// order, so the result is deterministic) and runs the epilogue for that slice -- a
// reduce-scatter with no global-memory round trip and no straggler.
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
constexpr int kPartialPitch = 132;  // fp32 words per row of a parked partial tile (conflict-free, 16-byte aligned)
constexpr int kLogitPitch = 129;    // fp32 words per row of the transposed logits tile

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
  int splits;   // K splits = cluster size along y (1, 2, 4, 8 or 16)
  int stages;   // smem ring depth
  long long* trace;  // debug: clock64 timeline of CTA (0,0), 128 slots, or null
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
  float* part_max;     // [rows, n_tiles] max logit in the tile           (want_lse)
  float* part_sum;     // [rows, n_tiles] sum exp(logit - max)            (want_lse)
  int n_tiles;
  int vocab_offset;
  float scale, softcap, inv_temp;
  int round_bf16;
  int gumbel;          // 1 = add Gumbel noise (weighted sampling)
  int want_lse;        // 1 = also produce the log-sum-exp partials (return_log_prob)
  const uint32_t* rng_state;  // [4]: step, seed_lo, seed_hi, unused (device memory: graph replays see updates)
  int row_offset;
  // gemm_rows.cuh only: RMSNorm fused into the GEMMs around it (norm scales folded into the weights).  A residual epilogue
  // leaves the per-row sums of squares of its output, one partial per 128-feature tile, in ss_out[tile * ss_pitch + row]; a
  // GEMM whose input is that output multiplies its accumulator by rstd[row] = rsqrt(sum_t ss_in[t * ss_pitch + row] / ss_dim + eps).
  const float* ss_in;
  float* ss_out;
  int ss_tiles, ss_pitch, ss_dim;
  float ss_eps;
  int act_gelu;  // gemm_rows.cuh, EPI_SWIGLU: 1 = the gate activation is gelu (tanh form) instead of silu (gemma3)
  int kv_fp8;    // gemm_rows.cuh, quantising QKV epilogue: 1 = float8_e4m3fn bytes (kv_quant_dtype fp8) instead of int8
  // gemm_rows.cuh only: int8 KV cache (kvcache.py:36-90, kv_quant_axis dkv).  Non-null: key / value rows are appended as 64 bytes
  // u = clip(rint(x * 127.5 / scale), -128, 127) + 128 at the same row index of kq_cache / vq_cache, with scale = max|x| over the
  // head's dims in k_scale / v_scale [planes, Hkv, t_alloc]; k_cache / v_cache are then unused.
  uint8_t* kq_cache;
  uint8_t* vq_cache;
  float* k_scale;
  float* v_scale;
};


// ---- thread-block cluster helpers ----
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// ---- epilogues: thread owns weight row n (n_local in the tile) and rows r0 .. r0+15, valid while r < r_lim ----

__device__ __forceinline__ void epi_store_bf16(const EpiArgs& e, const GemmParams& p, const float (&v)[16], int r0, int r_lim, int n) {
  if (n >= p.n) return;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int r = r0 + j;
    if (r < r_lim) e.out[(long long)r * e.ld_out + n] = __float2bfloat16_rn(v[j]);
  }
}

__device__ __forceinline__ void epi_residual(const EpiArgs& e, const GemmParams& p, const float (&v)[16], int r0, int r_lim, int n) {
  if (n >= p.n) return;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int r = r0 + j;
    if (r < r_lim) {
      const long long o = (long long)r * e.ld_out + n;
      e.out[o] = __float2bfloat16_rn(__bfloat162float(e.resid[o]) + bf16r(v[j]));
    }
  }
}

// linears.py:425-476 with mlp_activations [silu, linear]; every intermediate is a bf16 array there.
__device__ __forceinline__ void epi_swiglu(const EpiArgs& e, const GemmParams& p, const float (&v)[16], int r0, int r_lim, int n, int lane) {
  const int m = (n >> 5) * 16 + (lane & 15);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float other = __shfl_xor_sync(0xffffffffu, v[j], 16);
    const float a = bf16r(lane < 16 ? v[j] : other);
    const float b = bf16r(lane < 16 ? other : v[j]);
    const float sg = bf16r(1.0f / (1.0f + expf(-a)));
    const float act = bf16r(a * sg);
    const int r = r0 + j;
    if (lane < 16 && r < r_lim && n < p.n) e.out[(long long)r * e.ld_out + m] = __float2bfloat16_rn(act * b);
  }
}

// embeddings.py:304-315 (half-split rotation, bf16 arithmetic) + kvcache.py:626-718 (append).
// `exch` is a [128][17] fp32 exchange buffer: the rotation partner of feature d is d +- D/2,
// which lives in another epilogue warp.
__device__ __forceinline__ void epi_qkv_rope(const EpiArgs& e, const GemmParams& p, const float (&v)[16], int r0, int r_lim, int n,
                                              int n_local, float* exch, int pitch = 17) {
  const int D = e.d, half = D >> 1;
  const int head = n / D, d = n - head * D;
  const bool is_q = head < e.hq, is_k = !is_q && head < e.hq + e.hkv;
  const bool rot = (is_q || is_k) && n < p.n;
  float own[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    own[j] = bf16r(v[j]);
    if (j < pitch - 1) exch[n_local * pitch + j] = own[j];
  }
  epi_bar_sync();
  const bool first_half = d < half;
  const int partner = first_half ? n_local + half : n_local - half;
  const int fi = first_half ? d : d - half;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int r = r0 + j;
    if (r >= r_lim || n >= p.n) continue;
    float val = own[j];
    if (rot) {
      const float other = exch[partner * pitch + j];
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

// decoders.py:537-589: logits = bf16(dot) (* 1/sqrt(E) when tied), optional tanh soft cap, fp32.
__device__ __forceinline__ float logit_transform(const EpiArgs& e, float acc) {
  float lg = e.round_bf16 ? bf16r(acc) : acc;
  lg *= e.scale;
  if (e.softcap != 0.0f) lg = tanhf(lg / e.softcap) * e.softcap;
  return lg;
}

// Logits epilogue for one group of up to 64 rows: the 128 x RG tile is transposed through shared
// memory so that each thread then scans a contiguous run of vocabulary entries of ONE row --
// arg-max with the lowest index winning ties (jnp.argmax, inference_utils.py:76), optionally on
// Gumbel-perturbed scores (jax.random.categorical, :78), and optionally the (max, sum exp)
// pair for log-softmax (:55-63).  `lt` / `st` are [RG][129] fp32 tiles, `comb` is [8][64][5].
__device__ __forceinline__ void epi_logits_group(const EpiArgs& e, const GemmParams& p, uint32_t taddr, int g0, int RG, int n,
                                                 int n_local, int epi_tid, int tile, float* lt, float* st, float* comb) {
  const bool valid = n < p.n;
  const int gid = e.vocab_offset + n;
  uint32_t step = 0;
  uint64_t seed = 0;
  if (e.gumbel) {
    step = e.rng_state[0];
    seed = (uint64_t(e.rng_state[2]) << 32) | e.rng_state[1];
  }
  for (int c = 0; c < RG / 16; ++c) {
    float v[16];
    tmem_ld_x16(taddr + uint32_t(g0 + c * 16), v);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int rl = c * 16 + j;  // row within the group
      const int r = g0 + rl;
      float lg = -INFINITY, sc = -INFINITY;
      if (valid && r < p.rows) {
        lg = logit_transform(e, v[j]);
        if (e.logits_out != nullptr) {
          if (e.logits_only_row < 0) e.logits_out[(long long)r * e.ld_logits + n] = lg;
          else if (r == e.logits_only_row) e.logits_out[n] = lg;
        }
        sc = e.gumbel ? lg * e.inv_temp + gumbel_noise(seed, step, uint32_t(e.row_offset + r), uint32_t(gid)) : lg;
      }
      lt[rl * kLogitPitch + n_local] = lg;
      if (e.gumbel) st[rl * kLogitPitch + n_local] = sc;
    }
  }
  epi_bar_sync();
  // thread -> (row, part): a warp covers 32 consecutive rows of one part, so the scans are conflict-free
  const int nparts = kEpiThreads / RG, span = kTileN / nparts;
  const int rl = epi_tid % RG, part = epi_tid / RG;
  const float* lrow = lt + rl * kLogitPitch + part * span;
  const float* srow = (e.gumbel ? st : lt) + rl * kLogitPitch + part * span;
  float best = -INFINITY, raw = -INFINITY, mx = -INFINITY, sum = 0.0f;
  int bi = 0;
  for (int i = 0; i < span; ++i) {
    const float s = srow[i];
    if (s > best) { best = s; bi = i; }  // strict: the first (lowest-index) maximum is kept
  }
  raw = lrow[bi];
  if (e.want_lse) {
    for (int i = 0; i < span; ++i) mx = fmaxf(mx, lrow[i]);
    if (mx > -INFINITY)
      for (int i = 0; i < span; ++i) sum += expf(lrow[i] - mx);
  }
  float* cb = comb + (part * 64 + rl) * 5;
  cb[0] = best; cb[1] = __int_as_float(e.vocab_offset + tile * kTileN + part * span + bi); cb[2] = raw; cb[3] = mx; cb[4] = sum;
  epi_bar_sync();
  if (epi_tid < RG && g0 + epi_tid < p.rows) {
    best = -INFINITY; raw = -INFINITY; mx = -INFINITY; sum = 0.0f;
    int idx = 0x7fffffff;
    for (int q = 0; q < nparts; ++q) {  // parts are in increasing vocabulary order: strict > keeps the lowest index
      const float* t = comb + (q * 64 + epi_tid) * 5;
      if (t[0] > best) { best = t[0]; idx = __float_as_int(t[1]); raw = t[2]; }
      if (e.want_lse) {
        const float m2 = fmaxf(mx, t[3]);
        if (m2 > -INFINITY) sum = sum * expf(mx - m2) + t[4] * expf(t[3] - m2);
        mx = m2;
      }
    }
    const long long o = (long long)(g0 + epi_tid) * e.n_tiles + tile;
    e.part_score[o] = best; e.part_idx[o] = idx; e.part_raw[o] = raw;
    if (e.want_lse) { e.part_max[o] = mx; e.part_sum[o] = sum; }
  }
  epi_bar_sync();
}

template <int EPI>
__device__ __forceinline__ void run_epilogue(const EpiArgs& e, const GemmParams& p, const float (&v)[16], int r0, int r_lim, int n,
                                             int n_local, int lane, float* exch, int pitch = 17) {
  if (EPI == EPI_STORE_BF16) epi_store_bf16(e, p, v, r0, r_lim, n);
  if (EPI == EPI_RESIDUAL) epi_residual(e, p, v, r0, r_lim, n);
  if (EPI == EPI_SWIGLU) epi_swiglu(e, p, v, r0, r_lim, n, lane);
  if (EPI == EPI_QKV_ROPE) epi_qkv_rope(e, p, v, r0, r_lim, n, n_local, exch, pitch);
}

// ---- the kernel -------------------------------------------------------------------------

struct GemmSmemTail {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
  uint32_t pad;
};

// Shared memory after the 1024-byte alignment slack: max(pipeline stages, epilogue scratch) + tail.
__host__ __device__ inline size_t gemm_main_region_bytes(int stages, int r_tile, int splits, int epi) {
  size_t pipe = size_t(stages) * (kWTileBytes + r_tile * kBlockK * 2);
  size_t scratch = 0;
  if (splits > 1) scratch += size_t(r_tile) * kPartialPitch * 4 + 128 * 17 * 4 + size_t(r_tile / splits) * kPartialPitch * 4;  // parked partial, exchange, reduced slice
  else if (epi == EPI_QKV_ROPE) scratch += 128 * 17 * 4;          // rotation exchange
  if (epi == EPI_LOGITS) scratch += 2 * 64 * kLogitPitch * 4 + 8 * 64 * 5 * 4;
  size_t m = pipe > scratch ? pipe : scratch;
  return (m + 1023) / 1024 * 1024;
}
__host__ __device__ inline size_t gemm_smem_bytes(int stages, int r_tile, int splits, int epi) {
  return 1024 + gemm_main_region_bytes(stages, r_tile, splits, epi) + sizeof(GemmSmemTail) + 16;
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x, const GemmParams p, const EpiArgs e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space
  const int x_tile_bytes = p.r_tile * kBlockK * 2;
  const int stage_bytes = kWTileBytes + x_tile_bytes;
  GemmSmemTail* tail = reinterpret_cast<GemmSmemTail*>(smem + gemm_main_region_bytes(p.stages, p.r_tile, p.splits, EPI));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;  // split == rank in the (1, splits, 1) cluster
  const int n0 = tile * kTileN;
  const int kb_total = p.k / kBlockK;
  const int kb0 = int((long long)split * kb_total / p.splits);
  const int kb1 = int((long long)(split + 1) * kb_total / p.splits);
  const int nkb = kb1 - kb0;
  const uint32_t tmem_cols = p.r_tile < 32 ? 32u : uint32_t(p.r_tile);

  const int tl = timeline_begin(10 + EPI);
  griddep_launch_dependents();
  long long* trace = (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) ? p.trace : nullptr;
  const long long t_start = clock64();

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
      if (trace) trace[0] = clock64() - t_start;
      for (int i = 0; i < pre; ++i) {  // weights do not depend on the previous kernel
        mbar_expect_tx(&tail->full[i], uint32_t(stage_bytes));
        tma_load_2d(smem + size_t(i) * stage_bytes, &tm_w, (kb0 + i) * kBlockK, n0, &tail->full[i], kEvictFirst);
      }
      griddep_wait();
      if (trace) trace[1] = clock64() - t_start;
      for (int i = 0; i < pre; ++i)
        tma_load_2d(smem + size_t(i) * stage_bytes + kWTileBytes, &tm_x, (kb0 + i) * kBlockK, 0, &tail->full[i], kEvictLast);
      for (int i = pre; i < nkb; ++i) {
        const int s = i % p.stages;
        mbar_wait(&tail->empty[s], ((i / p.stages) & 1) ^ 1);
        if (trace && i < 30) trace[2 + i] = clock64() - t_start;
        mbar_expect_tx(&tail->full[s], uint32_t(stage_bytes));
        tma_load_2d(smem + size_t(s) * stage_bytes, &tm_w, (kb0 + i) * kBlockK, n0, &tail->full[s], kEvictFirst);
        tma_load_2d(smem + size_t(s) * stage_bytes + kWTileBytes, &tm_x, (kb0 + i) * kBlockK, 0, &tail->full[s], kEvictLast);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kTileN, p.r_tile);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % p.stages;
        mbar_wait(&tail->full[s], (i / p.stages) & 1);
        tcgen05_fence_after();
        if (trace && i < 30) trace[40 + i] = clock64() - t_start;
        const uint64_t da = umma_desc_sw128(smem + size_t(s) * stage_bytes);
        const uint64_t db = umma_desc_sw128(smem + size_t(s) * stage_bytes + kWTileBytes);
#pragma unroll
        for (int kk = 0; kk < kBlockK / kUmmaK; ++kk)
          umma_bf16(tmem_base, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc, uint32_t((i | kk) != 0));
        umma_commit(&tail->empty[s]);
      }
      umma_commit(&tail->tmem_full);
      if (trace) trace[39] = clock64() - t_start;
    }
    __syncwarp();
  }

  // ===== epilogue: warps 2..5 own TMEM lane quarters 2,3,0,1 =====
  const int quarter = warp & 3;
  const int n_local = quarter * 32 + lane;
  const int epi_tid = threadIdx.x - 64;
  const int n = n0 + n_local;
  const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16);
  float* scratch = reinterpret_cast<float*>(smem);  // pipeline stages are free once tmem_full fires
  if (warp >= 2) {
    mbar_wait(&tail->tmem_full, 0);
    tcgen05_fence_after();
    griddep_wait();
    if (trace && epi_tid == 0) trace[80] = clock64() - t_start;
  }

  if (EPI == EPI_LOGITS) {
    if (warp >= 2) {
      float* lt = scratch;
      float* st = lt + 64 * kLogitPitch;
      float* comb = st + 64 * kLogitPitch;
      const int RG = p.r_tile < 64 ? p.r_tile : 64;
      for (int g0 = 0; g0 < p.r_tile && g0 < p.rows; g0 += RG)
        epi_logits_group(e, p, taddr, g0, RG, n, n_local, epi_tid, tile, lt, st, comb);
    }
  } else if (p.splits == 1) {
    if (warp >= 2) {
      for (int c = 0; c < p.r_tile / 16 && c * 16 < p.rows; ++c) {
        float v[16];
        tmem_ld_x16(taddr + uint32_t(c * 16), v);
        run_epilogue<EPI>(e, p, v, c * 16, p.rows, n, n_local, lane, scratch);
      }
    }
  } else {
    // ---- split-K: park the partial tile, cluster barrier, reduce-scatter over DSMEM ----
    float* partial = scratch;                                // [r_tile][kPartialPitch]
    float* exch = scratch + p.r_tile * kPartialPitch;        // not aliased: peers read `partial` during the epilogue
    if (warp >= 2) {
      for (int c = 0; c < p.r_tile / 16; ++c) {
        float v[16];
        tmem_ld_x16(taddr + uint32_t(c * 16), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) partial[(c * 16 + j) * kPartialPitch + n_local] = v[j];
      }
    }
    if (trace && epi_tid == 0) trace[82] = clock64() - t_start;
    cluster_sync_all();
    if (trace && epi_tid == 0) trace[83] = clock64() - t_start;
    if (warp >= 2) {
      const int rpc = p.r_tile / p.splits;  // rows reduced and finished by this CTA
      const int r_begin = split * rpc;
      const int r_lim = min(p.rows, r_begin + rpc);
      const uint32_t my = smem_u32(partial);
      // Pull phase: 16-byte remote loads, one float4 column group of one row per thread and
      // peer, all peers in flight at once; the sum runs in split order (deterministic).
      float* reduced = exch + 128 * 17;  // [rpc][kPartialPitch] local staging of the reduced slice
      for (int u = epi_tid; u < rpc * 32; u += kEpiThreads) {
        const int rr = u >> 5, c4 = u & 31;
        const uint32_t off = uint32_t(((r_begin + rr) * kPartialPitch + c4 * 4) * 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < p.splits; s0 += 8) {
          float4 t[8];
#pragma unroll
          for (int ss = 0; ss < 8; ++ss) {
            t[ss] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + ss < p.splits) t[ss] = ld_dsmem_f32x4(dsmem_addr(my, uint32_t(s0 + ss)) + off);
          }
#pragma unroll
          for (int ss = 0; ss < 8; ++ss) {
            acc.x += t[ss].x; acc.y += t[ss].y; acc.z += t[ss].z; acc.w += t[ss].w;
          }
        }
        *reinterpret_cast<float4*>(reduced + rr * kPartialPitch + c4 * 4) = acc;
      }
      epi_bar_sync();
      if (trace && epi_tid == 0) trace[84] = clock64() - t_start;
      for (int r0 = r_begin; r0 < r_begin + rpc && r0 < p.rows; r0 += 16) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (r0 + j < r_begin + rpc) ? reduced[(r0 - r_begin + j) * kPartialPitch + n_local] : 0.0f;
        run_epilogue<EPI>(e, p, v, r0, r_lim, n, n_local, lane, exch);
      }
    }
    if (trace && epi_tid == 0) trace[85] = clock64() - t_start;
    cluster_sync_all();  // keep every CTA's partial alive until all peers have read it
    if (trace && epi_tid == 0) trace[86] = clock64() - t_start;
  }

  tcgen05_fence_before();
  __syncthreads();
  if (trace && threadIdx.x == 64) trace[81] = clock64() - t_start;
  timeline_end(tl);
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mtx
