// The whole decode step as ONE persistent kernel.
//
// The step is a chain of small dependent phases (per layer: norm, QKV, attention, out-proj, norm,
// MLP up, MLP down; then norm, logits).  As separate kernels each phase boundary costs ~10 us of
// drain / dependency-release / ramp latency against ~3 us of bandwidth time (profiles/, round 1).
// Here 2 CTAs per SM stay resident for the whole step; a phase boundary is a grid barrier
// (one atomic per CTA), and the TMA producer warp of every CTA keeps running AHEAD of the barrier,
// pulling the next phase's weight tiles into the shared-memory ring while the current phase's
// epilogue and the barrier complete -- weights never depend on activations.
//
// Roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2-5 =
// workers: GEMM epilogues, RMSNorm rows, the attention loop, the grid barrier.
// Split-K GEMMs reduce over DSMEM inside groups of S consecutive CTAs of a cluster of 8, synchronised
// by remote mbarrier arrives (no cluster-wide barrier, so the producer/MMA warps never stall on it).
//
// Shared memory (<= 112 KB so that two CTAs fit on an SM):
//   [0, 72 KB)       3-stage ring of (128x64 weight tile | r_tile x 64 activation tile); the same bytes
//                    hold the attention K/V tiles (4 warps x 16 KB) during the attention phase
//   [72 KB, 104 KB)  worker scratch: parked split-K partial [64][128] fp32 | attention merge buffer |
//                    logits transpose tiles | RoPE exchange buffer of unsplit QKV tiles
//   [104 KB, ..)     RoPE exchange buffer of split QKV tiles, barriers, flags
#pragma once

#include "attention.cuh"
#include "gemm_umma.cuh"
#include "step_kernels.cuh"

namespace mtx {

constexpr int kMegaThreads = 192;
constexpr int kMegaStages = 3;
constexpr int kMegaCluster = 8;
constexpr int kMegaMaxRTile = 64;
constexpr int kMegaStageBytes = kWTileBytes + kMegaMaxRTile * kBlockK * 2;  // 24 KB
constexpr int kMegaRingBytes = kMegaStages * kMegaStageBytes;               // 72 KB
constexpr int kMegaPartialPitch = 128;                                      // fp32 words per row of a parked partial
constexpr int kMegaScratchBytes = kMegaMaxRTile * kMegaPartialPitch * 4;    // 32 KB
constexpr int kMegaExchPitch = 9;                                           // split QKV epilogues finish <= 8 rows per CTA
constexpr int kMegaExchBytes = 128 * kMegaExchPitch * 4;

struct MegaTail {
  uint64_t full[kMegaStages];
  uint64_t empty[kMegaStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t peer_ready[4];  // index log2(S): arrivals from the S CTAs of a split group
  uint64_t peer_done[4];
  uint64_t attn_bars[2 * kAttnWarps];
  uint32_t tmem_base;
  volatile uint32_t barriers_passed;  // grid barriers this CTA has completed (written by worker thread 0)
  volatile uint32_t attn_done;        // attention phases this CTA has completed
  uint32_t pad;
};

__host__ __device__ inline size_t mega_smem_bytes() {
  return 1024 + kMegaRingBytes + kMegaScratchBytes + kMegaExchBytes + sizeof(MegaTail) + 64;
}

struct MegaParams {
  // model
  int L, E, HD, M, qkv_n, V;
  int hq, hkv, d, t_alloc, num_slots;
  int rows, r_tile;
  float eps;
  // K splits (1, 2, 4 or 8) of the four per-layer GEMMs
  int s_qkv, s_oproj, s_up, s_down;
  // activations [r_tile, *] bf16
  bf16 *x, *h, *n, *q, *attn, *act;
  // weights that are not behind tensor maps
  const bf16 *embedding, *attn_norm, *mlp_norm, *final_norm;
  bf16 *k_cache, *v_cache;
  long long kv_layer_elems;
  // row descriptors (prepare_rows_kernel)
  const int *token, *plane, *write_row;
  const float2* rope_cs;
  AttnParams attn_args;  // plane_base is patched per layer
  EpiArgs logits;        // logits epilogue arguments
  unsigned int* grid_bar;  // zeroed before every launch
  long long* trace;        // debug: globaltimer at [barrier k][arrive|release][cta], or null
};

// ---- small helpers -----------------------------------------------------------------------

__device__ __forceinline__ void mbar_arrive_count(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t cta_rank) {
  const uint32_t remote = dsmem_addr(smem_u32(local_bar), cta_rank);
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("mtx: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned int* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until(volatile uint32_t* flag, uint32_t want) {
  const long long t0 = clock64();
  while (*flag < want) {
    if (clock64() - t0 > 4000000000LL) {
      printf("mtx: flag wait timed out (block %d thread %d want %u have %u)\n", blockIdx.x, threadIdx.x, want, *flag);
      __trap();
    }
  }
}

// Grid barrier over all CTAs, executed by the 128 worker threads.  `k` = ordinal of this barrier.
__device__ __forceinline__ void grid_barrier(const MegaParams& p, MegaTail* tail, uint32_t k, int wtid) {
  __threadfence();
  fence_proxy_async_all();  // this phase's generic global writes are read through TMA by other CTAs
  epi_bar_sync();
  if (wtid == 0) {
    if (p.trace) p.trace[(2 * k) * gridDim.x + blockIdx.x] = (long long)globaltimer_ns();
    atomicAdd(p.grid_bar, 1u);
    const uint32_t target = k * gridDim.x;
    const long long t0 = clock64();
    while (ld_acquire_gpu(p.grid_bar) < target) {
      if (clock64() - t0 > 4000000000LL) {
        printf("mtx: grid barrier %u timed out (block %d)\n", k, blockIdx.x);
        __trap();
      }
    }
    __threadfence();
    if (p.trace) p.trace[(2 * k + 1) * gridDim.x + blockIdx.x] = (long long)globaltimer_ns();
    tail->barriers_passed = k;
    __threadfence_block();
  }
  epi_bar_sync();
}

// RMSNorm of one row by the 128 worker threads (normalizations.py:57-69), optional embedding gather.
__device__ __forceinline__ void mega_rmsnorm_row(const bf16* src, const bf16* scale, bf16* x_out, bf16* n_out, int E, float eps,
                                                 float* s_part, int wtid) {
  const int nvec = E / 8;
  float ss = 0.0f;
  for (int i = wtid; i < nvec; i += 128) {
    const uint4 raw = *reinterpret_cast<const uint4*>(src + i * 8);
    if (x_out != nullptr) *reinterpret_cast<uint4*>(x_out + i * 8) = raw;
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(w[j]), b = bf16_hi(w[j]);
      ss += a * a + b * b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  epi_bar_sync();
  if ((wtid & 31) == 0) s_part[wtid >> 5] = ss;
  epi_bar_sync();
  const float total = s_part[0] + s_part[1] + s_part[2] + s_part[3];
  const float rstd = 1.0f / sqrtf(total / float(E) + eps);
  for (int i = wtid; i < nvec; i += 128) {
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
    *reinterpret_cast<uint4*>(n_out + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Logits epilogue for 16 rows of one 128-wide vocabulary tile (see epi_logits_group): small scratch
// version for the persistent kernel.  lt/st = [16][129] fp32, comb = [8][16][5].
__device__ __forceinline__ void mega_logits_chunk(const EpiArgs& e, int n_valid, int rows, const float (&v)[16], int r0, int n, int n_local,
                                                  int wtid, int tile, float* lt, float* st, float* comb) {
  const bool valid = n < n_valid;
  const int gid = e.vocab_offset + n;
  uint32_t step = 0;
  uint64_t seed = 0;
  if (e.gumbel) {
    step = e.rng_state[0];
    seed = (uint64_t(e.rng_state[2]) << 32) | e.rng_state[1];
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int r = r0 + j;
    float lg = -INFINITY, sc = -INFINITY;
    if (valid && r < rows) {
      lg = logit_transform(e, v[j]);
      if (e.logits_out != nullptr) {
        if (e.logits_only_row < 0) e.logits_out[(long long)r * e.ld_logits + n] = lg;
        else if (r == e.logits_only_row) e.logits_out[n] = lg;
      }
      sc = e.gumbel ? lg * e.inv_temp + gumbel_noise(seed, step, uint32_t(e.row_offset + r), uint32_t(gid)) : lg;
    }
    lt[j * kLogitPitch + n_local] = lg;
    if (e.gumbel) st[j * kLogitPitch + n_local] = sc;
  }
  epi_bar_sync();
  const int rl = wtid & 15, part = wtid >> 4;  // 8 parts of 16 vocabulary entries
  const float* lrow = lt + rl * kLogitPitch + part * 16;
  const float* srow = (e.gumbel ? st : lt) + rl * kLogitPitch + part * 16;
  float best = -INFINITY, mx = -INFINITY, sum = 0.0f;
  int bi = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float s = srow[i];
    if (s > best) { best = s; bi = i; }
  }
  const float raw = lrow[bi];
  if (e.want_lse) {
#pragma unroll
    for (int i = 0; i < 16; ++i) mx = fmaxf(mx, lrow[i]);
    if (mx > -INFINITY) {
#pragma unroll
      for (int i = 0; i < 16; ++i) sum += expf(lrow[i] - mx);
    }
  }
  float* cb = comb + (part * 16 + rl) * 5;
  cb[0] = best; cb[1] = __int_as_float(e.vocab_offset + tile * kTileN + part * 16 + bi); cb[2] = raw; cb[3] = mx; cb[4] = sum;
  epi_bar_sync();
  if (wtid < 16 && r0 + wtid < rows) {
    float b2 = -INFINITY, r2 = -INFINITY, m2 = -INFINITY, s2 = 0.0f;
    int idx = 0x7fffffff;
    for (int q = 0; q < 8; ++q) {
      const float* t = comb + (q * 16 + wtid) * 5;
      if (t[0] > b2) { b2 = t[0]; idx = __float_as_int(t[1]); r2 = t[2]; }
      if (e.want_lse) {
        const float mm = fmaxf(m2, t[3]);
        if (mm > -INFINITY) s2 = s2 * expf(m2 - mm) + t[4] * expf(t[3] - mm);
        m2 = mm;
      }
    }
    const long long o = (long long)(r0 + wtid) * e.n_tiles + tile;
    e.part_score[o] = b2; e.part_idx[o] = idx; e.part_raw[o] = r2;
    if (e.want_lse) { e.part_max[o] = m2; e.part_sum[o] = s2; }
  }
  epi_bar_sync();
}

// ---- schedule shared by the producer, the MMA issuer and the workers ----------------------

struct GemmPhase {
  int n_tiles;   // 128-row weight tiles
  int kb_total;  // 64-element k-blocks
  int S;         // K splits = CTAs per split group
};

// Tile of this CTA in round `round` of a phase, or -1.  A cluster of 8 hosts 8/S split groups.
__device__ __forceinline__ int mega_tile(const GemmPhase& g, int round, int n_clusters) {
  const int cluster = blockIdx.x / kMegaCluster, rank = blockIdx.x % kMegaCluster;
  const int gpc = kMegaCluster / g.S;
  const int tile = (round * gpc + rank / g.S) * n_clusters + cluster;  // consecutive tiles go to different clusters
  return tile < g.n_tiles ? tile : -1;
}
__device__ __forceinline__ int mega_rounds(const GemmPhase& g, int n_clusters) {
  const int per_round = n_clusters * (kMegaCluster / g.S);
  return (g.n_tiles + per_round - 1) / per_round;
}

// ---- the kernel --------------------------------------------------------------------------

__global__ void __launch_bounds__(kMegaThreads, 2)
step_megakernel(const __grid_constant__ CUtensorMap tm_wqkv, const __grid_constant__ CUtensorMap tm_wo,
                const __grid_constant__ CUtensorMap tm_w01, const __grid_constant__ CUtensorMap tm_wout,
                const __grid_constant__ CUtensorMap tm_wlogits, const __grid_constant__ CUtensorMap tm_xn,
                const __grid_constant__ CUtensorMap tm_xattn, const __grid_constant__ CUtensorMap tm_xact,
                const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const MegaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  float* scratch = reinterpret_cast<float*>(smem + kMegaRingBytes);
  float* exch = reinterpret_cast<float*>(smem + kMegaRingBytes + kMegaScratchBytes);
  MegaTail* tail = reinterpret_cast<MegaTail*>(smem + kMegaRingBytes + kMegaScratchBytes + kMegaExchBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_clusters = gridDim.x / kMegaCluster;
  const int rank = blockIdx.x % kMegaCluster;
  const int stage_bytes = kWTileBytes + p.r_tile * kBlockK * 2;
  const uint32_t tmem_cols = uint32_t(2 * p.r_tile < 32 ? 32 : 2 * p.r_tile);
  const int tl = timeline_begin(20);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMegaStages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->tmem_full[b], 1);
      mbar_init(&tail->tmem_empty[b], 1);
    }
    for (int i = 1; i < 4; ++i) {
      mbar_init(&tail->peer_ready[i], 1u << i);
      mbar_init(&tail->peer_done[i], 1u << i);
    }
    for (int i = 0; i < 2 * kAttnWarps; ++i) mbar_init(&tail->attn_bars[i], 1);
    tail->barriers_passed = 0;
    tail->attn_done = 0;
    fence_barrier_init();
    tma_prefetch_desc(&tm_wqkv);
    tma_prefetch_desc(&tm_wo);
    tma_prefetch_desc(&tm_w01);
    tma_prefetch_desc(&tm_wout);
    tma_prefetch_desc(&tm_wlogits);
    tma_prefetch_desc(&tm_xn);
    tma_prefetch_desc(&tm_xattn);
    tma_prefetch_desc(&tm_xact);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 1) {
    tmem_alloc(&tail->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's mbarriers exist before any remote arrive
  tcgen05_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  const GemmPhase ph_qkv{(p.qkv_n + kTileN - 1) / kTileN, p.E / kBlockK, p.s_qkv};
  const GemmPhase ph_o{(p.E + kTileN - 1) / kTileN, p.HD / kBlockK, p.s_oproj};
  const GemmPhase ph_up{(2 * p.M + kTileN - 1) / kTileN, p.E / kBlockK, p.s_up};
  const GemmPhase ph_down{(p.E + kTileN - 1) / kTileN, p.M / kBlockK, p.s_down};
  const int logits_tiles = (p.V + kTileN - 1) / kTileN;

  // Barrier ordinals: per layer 7 (after norm1, qkv, attention, out-proj, norm2, up, down), then norm, (logits).
  // A GEMM phase may read its activations once `need` barriers have completed.
  if (warp == 0) {
    // =================================== TMA producer ===================================
    if (lane == 0) {
      uint32_t kc = 0;
      auto run_phase = [&](const CUtensorMap* tmw, const CUtensorMap* tmx, const GemmPhase& g, int w_row0, uint32_t need,
                           uint32_t need_attn, bool strided_tiles) {
        bool ready = false;
        int pend_stage[kMegaStages], pend_c0[kMegaStages], npend = 0;
        auto flush = [&]() {
          if (!ready) {
            spin_until(&tail->barriers_passed, need);
            __threadfence_block();
            __threadfence();
            fence_proxy_async_all();
            ready = true;
          }
          for (int i = 0; i < npend; ++i)
            tma_load_2d(ring + size_t(pend_stage[i]) * stage_bytes + kWTileBytes, tmx, pend_c0[i], 0, &tail->full[pend_stage[i]], kEvictLast);
          npend = 0;
        };
        if (need_attn > 0) {  // the ring holds K/V tiles until then
          spin_until(&tail->attn_done, need_attn);
          __threadfence_block();
        }
        const int rounds = strided_tiles ? (g.n_tiles + int(gridDim.x) - 1) / int(gridDim.x) : mega_rounds(g, n_clusters);
        for (int round = 0; round < rounds; ++round) {
          const int tile = strided_tiles ? (round * int(gridDim.x) + int(blockIdx.x) < g.n_tiles ? round * int(gridDim.x) + int(blockIdx.x) : -1)
                                         : mega_tile(g, round, n_clusters);
          if (tile < 0) continue;
          const int split = rank % g.S;
          const int kb0 = int((long long)split * g.kb_total / g.S), kb1 = int((long long)(split + 1) * g.kb_total / g.S);
          for (int kb = kb0; kb < kb1; ++kb) {
            const int s = int(kc % kMegaStages);
            if (kc >= uint32_t(kMegaStages)) {
              // the stage we are about to refill may still be waiting for a pending X load
              for (int i = 0; i < npend; ++i)
                if (pend_stage[i] == s) { flush(); break; }
              mbar_wait(&tail->empty[s], ((kc / kMegaStages) & 1) ^ 1);
            }
            mbar_expect_tx(&tail->full[s], uint32_t(stage_bytes));
            tma_load_2d(ring + size_t(s) * stage_bytes, tmw, kb * kBlockK, w_row0 + tile * kTileN, &tail->full[s], kEvictFirst);
            pend_stage[npend] = s;
            pend_c0[npend] = kb * kBlockK;
            ++npend;
            ++kc;
            if (ready || npend == kMegaStages) flush();
          }
        }
        flush();
      };
      for (int l = 0; l < p.L; ++l) {
        const uint32_t b0 = uint32_t(7 * l);
        run_phase(&tm_wqkv, &tm_xn, ph_qkv, l * p.qkv_n, b0 + 1, uint32_t(l), false);
        run_phase(&tm_wo, &tm_xattn, ph_o, l * p.E, b0 + 3, uint32_t(l + 1), false);
        run_phase(&tm_w01, &tm_xn, ph_up, l * 2 * p.M, b0 + 5, 0, false);
        run_phase(&tm_wout, &tm_xact, ph_down, l * p.E, b0 + 6, 0, false);
      }
      const GemmPhase ph_logits{logits_tiles, p.E / kBlockK, 1};
      run_phase(&tm_wlogits, &tm_xn, ph_logits, 0, uint32_t(7 * p.L + 1), 0, true);
    }
    __syncwarp();
  } else if (warp == 1) {
    // =================================== MMA issuer =====================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kTileN, p.r_tile);
      uint32_t kc = 0, uc = 0;
      auto run_phase = [&](const GemmPhase& g, bool strided_tiles) {
        const int rounds = strided_tiles ? (g.n_tiles + int(gridDim.x) - 1) / int(gridDim.x) : mega_rounds(g, n_clusters);
        for (int round = 0; round < rounds; ++round) {
          const int tile = strided_tiles ? (round * int(gridDim.x) + int(blockIdx.x) < g.n_tiles ? 0 : -1) : mega_tile(g, round, n_clusters);
          if (tile < 0) continue;
          const int split = rank % g.S;
          const int kb0 = int((long long)split * g.kb_total / g.S), kb1 = int((long long)(split + 1) * g.kb_total / g.S);
          const uint32_t buf = uc & 1u;
          if (uc >= 2) mbar_wait(&tail->tmem_empty[buf], ((uc >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          const uint32_t acc = tmem_base + buf * uint32_t(p.r_tile);
          for (int kb = kb0; kb < kb1; ++kb) {
            const int s = int(kc % kMegaStages);
            mbar_wait(&tail->full[s], (kc / kMegaStages) & 1);
            tcgen05_fence_after();
            const uint64_t da = umma_desc_sw128(ring + size_t(s) * stage_bytes);
            const uint64_t db = umma_desc_sw128(ring + size_t(s) * stage_bytes + kWTileBytes);
#pragma unroll
            for (int kk = 0; kk < kBlockK / kUmmaK; ++kk)
              umma_bf16(acc, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc, uint32_t((kb > kb0) || kk > 0));
            umma_commit(&tail->empty[s]);
            ++kc;
          }
          umma_commit(&tail->tmem_full[buf]);
          ++uc;
        }
      };
      for (int l = 0; l < p.L; ++l) {
        run_phase(ph_qkv, false);
        run_phase(ph_o, false);
        run_phase(ph_up, false);
        run_phase(ph_down, false);
      }
      const GemmPhase ph_logits{logits_tiles, p.E / kBlockK, 1};
      run_phase(ph_logits, true);
    }
    __syncwarp();
  } else {
    // =================================== workers =======================================
    const int wtid = threadIdx.x - 64;
    const int quarter = warp & 3;
    const int n_local = quarter * 32 + lane;
    const uint32_t tlane = uint32_t(quarter * 32) << 16;
    float* s_part = exch;  // 4 floats during norm phases
    uint32_t uc = 0, nbar = 0, attn_phase = 0;
    uint32_t ready_par[4] = {0, 0, 0, 0}, done_par[4] = {0, 0, 0, 0};
    int pending_done = 0;  // log2(S) of a split unit whose partial may still be read by peers

    griddep_wait();
    if (p.trace && wtid == 0) p.trace[gridDim.x + blockIdx.x] = (long long)globaltimer_ns();  // "barrier 0 release" = start

    int ev_n = 0;
    bool ev_on = false;
    long long* ev_base = p.trace ? p.trace + 2 * 200 * (long long)gridDim.x + (long long)blockIdx.x * 64 : nullptr;
    auto EV = [&](int id) {
      if (ev_base && ev_on && wtid == 0 && ev_n < 32) {
        ev_base[2 * ev_n] = id;
        ev_base[2 * ev_n + 1] = (long long)globaltimer_ns();
        ++ev_n;
      }
    };
    auto gemm_phase = [&](int epi, const GemmPhase& g, const EpiArgs& e, int n_valid) {
      GemmParams gp;
      gp.n = n_valid;
      gp.k = 0;
      gp.rows = p.rows;
      gp.r_tile = p.r_tile;
      gp.splits = g.S;
      gp.stages = kMegaStages;
      gp.trace = nullptr;
      const int rounds = mega_rounds(g, n_clusters);
      for (int round = 0; round < rounds; ++round) {
        const int tile = mega_tile(g, round, n_clusters);
        if (tile < 0) continue;
        const int n = tile * kTileN + n_local;
        const uint32_t buf = uc & 1u;
        EV(epi * 10 + 0);
        mbar_wait(&tail->tmem_full[buf], (uc >> 1) & 1);
        tcgen05_fence_after();
        EV(epi * 10 + 1);
        const uint32_t taddr = tmem_base + tlane + buf * uint32_t(p.r_tile);
        if (g.S == 1) {
          for (int c = 0; c < p.r_tile / 16 && c * 16 < p.rows; ++c) {
            float v[16];
            tmem_ld_x16(taddr + uint32_t(c * 16), v);
            if (epi == EPI_QKV_ROPE) run_epilogue<EPI_QKV_ROPE>(e, gp, v, c * 16, p.rows, n, n_local, lane, scratch, 17);
            if (epi == EPI_RESIDUAL) run_epilogue<EPI_RESIDUAL>(e, gp, v, c * 16, p.rows, n, n_local, lane, scratch);
            if (epi == EPI_SWIGLU) run_epilogue<EPI_SWIGLU>(e, gp, v, c * 16, p.rows, n, n_local, lane, scratch);
          }
          tcgen05_fence_before();
          epi_bar_sync();
          EV(epi * 10 + 5);
          if (wtid == 0) mbar_arrive(&tail->tmem_empty[buf]);
        } else {
          const int lg = g.S == 2 ? 1 : g.S == 4 ? 2 : 3;
          float* partial = scratch;
          if (pending_done) {  // peers must have finished reading the previous partial
            mbar_wait_cluster(&tail->peer_done[pending_done], done_par[pending_done]);
            done_par[pending_done] ^= 1;
            pending_done = 0;
          }
          for (int c = 0; c < p.r_tile / 16; ++c) {
            float v[16];
            tmem_ld_x16(taddr + uint32_t(c * 16), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) partial[(c * 16 + j) * kMegaPartialPitch + n_local] = v[j];
          }
          tcgen05_fence_before();
          epi_bar_sync();
          const int base = (rank / g.S) * g.S, split = rank % g.S;
          if (wtid == 0) {
            mbar_arrive(&tail->tmem_empty[buf]);
            for (int m = 0; m < g.S; ++m) mbar_arrive_remote(&tail->peer_ready[lg], uint32_t(base + m));
          }
          EV(epi * 10 + 2);
          mbar_wait_cluster(&tail->peer_ready[lg], ready_par[lg]);
          ready_par[lg] ^= 1;
          EV(epi * 10 + 3);
          const int rpc = p.r_tile / g.S;
          const int r_begin = split * rpc;
          const int r_lim = min(p.rows, r_begin + rpc);
          const uint32_t my = smem_u32(partial);
          for (int u = wtid; u < rpc * 32; u += kEpiThreads) {
            const int rr = u >> 5, c4 = u & 31;
            const uint32_t off = uint32_t(((r_begin + rr) * kMegaPartialPitch + c4 * 4) * 4);
            float4 t[8];
#pragma unroll
            for (int ss = 0; ss < 8; ++ss) {
              t[ss] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ss < g.S) t[ss] = ld_dsmem_f32x4(dsmem_addr(my, uint32_t(base + ss)) + off);
            }
            float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int ss = 0; ss < 8; ++ss) {
              acc4.x += t[ss].x; acc4.y += t[ss].y; acc4.z += t[ss].z; acc4.w += t[ss].w;
            }
            // rows [r_begin, r_begin + rpc) of my own partial are read by nobody else: reuse them
            *reinterpret_cast<float4*>(partial + (r_begin + rr) * kMegaPartialPitch + c4 * 4) = acc4;
          }
          epi_bar_sync();
          EV(epi * 10 + 4);
          if (wtid == 0)
            for (int m = 0; m < g.S; ++m) mbar_arrive_remote(&tail->peer_done[lg], uint32_t(base + m));
          pending_done = lg;
          for (int r0 = r_begin; r0 < r_begin + rpc && r0 < p.rows; r0 += 16) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (r0 + j < r_begin + rpc) ? partial[(r0 + j) * kMegaPartialPitch + n_local] : 0.0f;
            if (epi == EPI_QKV_ROPE) run_epilogue<EPI_QKV_ROPE>(e, gp, v, r0, r_lim, n, n_local, lane, exch, kMegaExchPitch);
            if (epi == EPI_RESIDUAL) run_epilogue<EPI_RESIDUAL>(e, gp, v, r0, r_lim, n, n_local, lane, exch);
            if (epi == EPI_SWIGLU) run_epilogue<EPI_SWIGLU>(e, gp, v, r0, r_lim, n, n_local, lane, exch);
          }
          EV(epi * 10 + 5);
        }
        ++uc;
      }
      // before the grid barrier every split group has consumed its partials (attention and logits reuse the scratch)
      if (pending_done) {
        mbar_wait_cluster(&tail->peer_done[pending_done], done_par[pending_done]);
        done_par[pending_done] ^= 1;
        pending_done = 0;
      }
      EV(epi * 10 + 6);
    };

    for (int l = 0; l < p.L; ++l) {
      ev_on = l == 1;
      // ---- RMSNorm before attention (layer 0: embedding gather) ----
      for (int r = blockIdx.x; r < p.rows; r += gridDim.x) {
        const bf16* src = l == 0 ? p.embedding + (long long)p.token[r] * p.E : p.x + (long long)r * p.E;
        mega_rmsnorm_row(src, p.attn_norm + (long long)l * p.E, l == 0 ? p.x + (long long)r * p.E : nullptr, p.n + (long long)r * p.E,
                         p.E, p.eps, s_part, wtid);
      }
      grid_barrier(p, tail, ++nbar, wtid);
      // ---- QKV + RoPE + append ----
      {
        EpiArgs e;
        memset(&e, 0, sizeof(e));
        e.q_out = p.q;
        e.k_cache = p.k_cache + p.kv_layer_elems * l;
        e.v_cache = p.v_cache + p.kv_layer_elems * l;
        e.plane = p.plane;
        e.write_row = p.write_row;
        e.rope_cs = p.rope_cs;
        e.hq = p.hq;
        e.hkv = p.hkv;
        e.d = p.d;
        e.t_alloc = p.t_alloc;
        gemm_phase(EPI_QKV_ROPE, ph_qkv, e, p.qkv_n);
      }
      grid_barrier(p, tail, ++nbar, wtid);
      // ---- attention over the valid rows of both cache segments ----
      {
        AttnParams ap = p.attn_args;
        ap.plane_base = l * p.num_slots;
        float* sm_o_all = scratch;
        float* sm_stat = scratch + kAttnWarps * 16 * 64;
        attn_process_items<64>(tm_k, tm_v, ap, ring, sm_o_all, tail->attn_bars, sm_stat, attn_phase, wtid, blockIdx.x, gridDim.x);
        fence_proxy_async();
        epi_bar_sync();
        if (wtid == 0) {
          __threadfence_block();
          tail->attn_done = uint32_t(l + 1);
        }
      }
      grid_barrier(p, tail, ++nbar, wtid);
      // ---- out-projection + residual ----
      {
        EpiArgs e;
        memset(&e, 0, sizeof(e));
        e.out = p.h;
        e.resid = p.x;
        e.ld_out = p.E;
        gemm_phase(EPI_RESIDUAL, ph_o, e, p.E);
      }
      grid_barrier(p, tail, ++nbar, wtid);
      // ---- RMSNorm before the MLP ----
      for (int r = blockIdx.x; r < p.rows; r += gridDim.x)
        mega_rmsnorm_row(p.h + (long long)r * p.E, p.mlp_norm + (long long)l * p.E, nullptr, p.n + (long long)r * p.E, p.E, p.eps, s_part, wtid);
      grid_barrier(p, tail, ++nbar, wtid);
      // ---- MLP up + SwiGLU ----
      {
        EpiArgs e;
        memset(&e, 0, sizeof(e));
        e.out = p.act;
        e.ld_out = p.M;
        gemm_phase(EPI_SWIGLU, ph_up, e, 2 * p.M);
      }
      grid_barrier(p, tail, ++nbar, wtid);
      // ---- MLP down + residual ----
      {
        EpiArgs e;
        memset(&e, 0, sizeof(e));
        e.out = p.x;
        e.resid = p.h;
        e.ld_out = p.E;
        gemm_phase(EPI_RESIDUAL, ph_down, e, p.E);
      }
      grid_barrier(p, tail, ++nbar, wtid);
    }
    // ---- final RMSNorm ----
    for (int r = blockIdx.x; r < p.rows; r += gridDim.x)
      mega_rmsnorm_row(p.x + (long long)r * p.E, p.final_norm, nullptr, p.n + (long long)r * p.E, p.E, p.eps, s_part, wtid);
    grid_barrier(p, tail, ++nbar, wtid);
    // ---- logits + per-tile sampling partials ----
    {
      float* lt = scratch;
      float* st = lt + 16 * kLogitPitch;
      float* comb = st + 16 * kLogitPitch;
      for (int tile = blockIdx.x; tile < logits_tiles; tile += gridDim.x) {
        const uint32_t buf = uc & 1u;
        mbar_wait(&tail->tmem_full[buf], (uc >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + tlane + buf * uint32_t(p.r_tile);
        const int n = tile * kTileN + n_local;
        for (int c = 0; c < p.r_tile / 16 && c * 16 < p.rows; ++c) {
          float v[16];
          tmem_ld_x16(taddr + uint32_t(c * 16), v);
          mega_logits_chunk(p.logits, p.V, p.rows, v, c * 16, n, n_local, wtid, tile, lt, st, comb);
        }
        tcgen05_fence_before();
        epi_bar_sync();
        if (wtid == 0) mbar_arrive(&tail->tmem_empty[buf]);
        ++uc;
      }
    }
  }

  if (p.trace && threadIdx.x == 64) {
    p.trace[blockIdx.x] = (long long)globaltimer_ns();  // slot of "barrier 0 arrive" = end of the CTA's work
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer could still address its shared memory
  timeline_end(tl);
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mtx
