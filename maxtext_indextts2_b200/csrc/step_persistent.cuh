// The whole decode step as ONE persistent kernel: one CTA per SM, no clusters.
//
// Why: as separate kernels every phase of a layer (QKV, attention, out-proj, MLP up, MLP down, two
// norms) costs 8-16 us of launch / drain / dependency latency against 0.5-10 us of HBM time
// (profiles/, round 1).  Here a phase boundary is one grid barrier (~1.3 us) and the HBM stream never
// stops at it: the weight producer of every CTA runs AHEAD of the barrier, pulling the next phases'
// weight tiles into its shared-memory ring while the current phase drains (weights never depend on
// activations); only the small activation tiles wait for the barrier.
//
// Per layer five phases, five grid barriers:
//   QKV GEMM (+RMSNorm of x, RoPE, KV append) | attention | out-proj (+residual, row sum-of-squares) |
//   MLP up (+RMSNorm of h, SwiGLU) | MLP down (+residual, row sum-of-squares)
// then one small final-norm phase and the logits GEMM with the sampling partials.
//
// RMSNorm is never a phase of its own: the epilogue that produces x (or h) also emits per-row partial
// sums of squares; the consuming GEMM is computed as rstd[row] * dot(W, bf16(x * scale)): the per-feature
// scale is applied to the activation k-blocks in shared memory, the per-row factor to the fp32
// accumulator in the epilogue (normalizations.py:57-69 rounds x * rstd to bf16 first).
//
// GEMM work tables are built on the host (engine.cu, pk_fill_phase).  CTAs that share a weight tile
// exchange fp32 partial tiles through an L2-resident workspace WITHOUT flags: a word of a slot is either
// a sentinel NaN or data, the CTA that owns a row slice polls the fragments themselves, sums them in
// fixed order (deterministic, no atomics on data), runs the epilogue and puts the sentinel back.
//
// Epilogues work on ROW-MAJOR float4 fragments (a warp = 128 consecutive output features of one row),
// so RoPE / SwiGLU partners are a warp shuffle away and every global store is a piece of a 256-byte
// coalesced row segment.  A warp's epilogue is one dependent instruction chain (~2.6 ns per
// instruction): these paths are written for instruction count and independent rows in flight.
//
// Roles (352 threads): warp 0 weight producer (TMA) | warp 1 tcgen05.mma issuer | warps 2-5 epilogue |
// warps 6-10 attention; during the GEMM phases warps 6-9 scale activation tiles, warp 10 is the
// activation producer (TMA) and all five help to finish MLP-up rows.  Every role walks the same static
// fill sequence with its own lean loop.
// Attention is stream-K over 64-row KV tiles at WARP granularity: every attention warp of the grid
// gets the same number of tiles (the partition and the merge plan are computed once per step: lengths do
// not change between layers); a (row, kv head) pair that spans warps is merged by the CTA that holds its
// first tile, from shared memory and -- for warps of other CTAs -- from flag-in-data parts in L2.
#pragma once

#include <type_traits>

#include "attention.cuh"
#include "gemm_umma.cuh"
#include "step_kernels.cuh"

namespace mtx {

constexpr int kPkThreads = 352;
constexpr int kPkStages = 5;
constexpr int kPkMaxRTile = 64;
constexpr int kPkStageBytes = kWTileBytes + kPkMaxRTile * kBlockK * 2;  // 24 KB
constexpr int kPkRingBytes = kPkStages * kPkStageBytes;                 // 120 KB
constexpr int kPkAttnWarps = 5;
constexpr int kPkAttnBytes = kPkAttnWarps * 2 * (64 * 64 * 2);          // 80 KB: per warp one K and one V tile (D = 64)
constexpr int kPkParkFloats = kPkMaxRTile * 128;                        // 32 KB, aliases the attention tiles
constexpr int kPkAccBufs = 4;                                           // TMEM accumulators
constexpr int kPkMaxUnits = 4;                                          // work units of one CTA in one phase
constexpr int kPkMaxSplit = 16;                                         // CTAs sharing one weight tile
constexpr int kPkMaxParts = 48;                                         // attention warps sharing one (row, kv head)
constexpr int kPkSlotFloats = 64 * 128;                                 // one partial tile in the exchange workspace
constexpr uint32_t kPkSentinel = 0xFFFFDEADu;                           // "not written yet" word of the exchange workspace (a NaN no MMA produces)
constexpr int kPkParkPitch = 136;                                       // floats per row of the parked logits tile (conflict-free scans)
constexpr int kPkAttnListMax = 32;                                      // KV tiles one attention warp may own per layer
constexpr int kPkAttnSlotFloats = 8 * 64 + 2 * 8;                       // one attention partial (O, m, l), G <= 8, D = 64

enum PkPhase : int { PK_QKV = 0, PK_OPROJ = 1, PK_UP = 2, PK_DOWN = 3, PK_LOGITS = 4, PK_END = 5 };

struct PkUnit {
  int tile;     // 128-row weight tile
  int kb0, kb1; // k-block range
  int c_first;  // first CTA contributing to this tile
  int S;        // number of contributing CTAs (consecutive)
};

struct PkTable {
  int n_units[4];
  int kbs[4];  // k-blocks of this CTA per phase
  PkUnit u[4][kPkMaxUnits];
};

struct PkTail {
  uint64_t full_w[kPkStages];
  uint64_t full_x[kPkStages];
  uint64_t xready[kPkStages];
  uint64_t empty[kPkStages];
  uint64_t tmem_full[kPkAccBufs];
  uint64_t tmem_empty[kPkAccBufs];
  uint64_t attn_bars[2 * kPkAttnWarps];
  uint32_t tmem_base;
  volatile uint32_t bar_done;  // grid barriers this CTA has seen complete
  float rstd[kPkMaxRTile];
  float red[8];
  // row descriptors of the step (constant across layers), staged once for the attention warps
  int r_len0[kPkMaxRTile], r_rf[kPkMaxRTile], r_rl[kPkMaxRTile], r_plane[kPkMaxRTile], r_prefix[kPkMaxRTile + 1];
  // attention: this CTA's warps own contiguous runs of KV tiles, the same for every layer of the step
  int a_count[kPkAttnWarps];
  int a_tail[kPkAttnWarps];               // part slot (pair * kPkMaxParts + ordinal) of the warp's tail segment, -1: none
  // merge plan of the CTA (built once per step: the partition does not change between layers): one job per pair
  // that starts in this CTA and is cut between its warps or continues in later CTAs
  int a_njobs;
  struct Job {
    int pair, row, head;       // (row, kv head) and pair = row * hkv + head
    int n_src;                 // partials parked on chip: src[i] = w (warp w's K buffer) or 8 + w (warp w's slot)
    int src[kPkAttnWarps];
    int n_parts;               // partials of later CTAs' warps in the pair's L2 workspace
  } a_jobs[kPkAttnWarps + 1];
  int2 a_list[kPkAttnWarps][kPkAttnListMax];  // .x = cache row of the tile (layer 0), .y = PkAttnMeta bits
  int a_row2[kPkAttnWarps][kPkAttnListMax];   // paged, 32-token pages: cache row of the tile's second page
  float a_scales[kPkAttnWarps][128];          // quantised cache: the current tile's 64 K scales | 64 V scales
  alignas(16) float a_slot[kPkAttnWarps][kPkAttnSlotFloats];  // partial of a warp's first segment (its K/V buffers stay busy)
  PkTable tab;
};

__host__ __device__ inline size_t pk_smem_bytes() { return 1024 + kPkRingBytes + kPkAttnBytes + sizeof(PkTail) + 64; }

struct PkParams {
  int L, E, HD, M, qkv_n, V;
  int hq, hkv, d, t_alloc, num_slots;
  int rows, r_tile;
  int P, T;
  float eps, softcap;
  // activations [r_tile, *] bf16
  bf16 *x, *h, *n, *q, *attn, *act;
  const bf16 *embedding, *attn_norm, *mlp_norm, *final_norm;
  bf16 *k_cache, *v_cache;
  long long kv_layer_elems;
  int kv_layer_rows;  // rows of one layer in the K/V tensor maps
  // step_persistent_kernel<KVQ != 0>: quantised decode cache (quantize_kvcache with kv_quant_axis=dkv, inference/kvcache.py:36-90):
  // bytes [L, slots, Hkv, T, 64] (u = q + 128, or float8_e4m3fn) with one fp32 scale = max|x| per row in k_scale / v_scale
  // [L, slots, Hkv, T]; tm_k / tm_v then address the byte caches (64-byte rows, SWIZZLE_64B)
  uint8_t *kq_cache, *vq_cache;
  float *k_scale, *v_scale;
  // attention=paged (page_map != null): k_cache / v_cache are the page pools [L, Hkv, num_pages, page_tokens, 64], a layer being one
  // plane of t_alloc = num_pages * page_tokens rows per kv head; row r is page group r, its token i lives in page
  // page_map[r][i / page_tokens].  page_tokens >= 32: a 64-row tile is a slice of one page, or two whole pages (4 KB boxes).
  const int* page_map;
  int page_tokens, num_pages, max_pages;
  // row descriptors (prepare_rows_kernel)
  const int *token, *plane, *write_row, *len0, *ring_first, *ring_len;
  const float2* rope_cs;
  const int* tile_prefix;  // [rows + 1] exclusive prefix of the per-row tile counts
  const int* attn_info;    // [0] CTAs with attention work, [1] total tiles (all kv heads), [2] 1 = one pair per CTA
  // attention merge workspace
  float* attn_part_o;   // [pairs, kPkMaxParts, G, D]
  // split-K exchange
  float* part_ws;  // [n_ctas * 4] slots of kPkSlotFloats
  float* ss_x;     // [E/128 rounded up, kPkMaxRTile] partial sums of squares of the rows of x, one per 128-feature tile
  float* ss_h;
  const PkTable* tables;  // [n_ctas]
  EpiArgs logits;
  unsigned int* grid_bar;  // zeroed by prepare_rows_kernel
  long long* trace;        // debug: globaltimer at [barrier k][arrive|release][cta], or null
  int trace_bars;          // barrier slots of the trace buffer (>= 2 + 5 L + 1, see mtx_step_trace_words)
  int variant;             // experiment switches (MTX_PK_VARIANT): bit 0 = L2 prefetch of the next layer's K/V tiles at the end of a
                           // warp's tile loop, bit 1 = L2 prefetch of this layer's K/V tiles at the start of the layer
  float attn_skew;         // attention tile partition: see pk_cta_lo (0 = equal shares)
  int fold;                // 1: the RMSNorm scales are folded into wqkv / w01 (mtx_model_config.norm_scales_folded): no scale pass
};

// ---- small helpers -----------------------------------------------------------------------

// Debug event log: (id, globaltimer) pairs, 32 per role and CTA, after the barrier stamps.
struct PkEv {
  long long* base;
  int n;
  bool on;
  bool brief;  // log only the end of the attention tile loop (610) and of the merges (501): fits every other layer
};
__device__ __forceinline__ PkEv pk_ev_make(const PkParams& p, int role) {
  PkEv e;
  e.base = p.trace ? p.trace + 2 * (long long)p.trace_bars * gridDim.x + ((long long)blockIdx.x * 3 + role) * 64 : nullptr;
  e.n = 0;
  e.on = false;
  e.brief = false;
  return e;
}
__device__ __forceinline__ void pk_ev(PkEv& e, int id) {
#ifdef MTX_PK_EVENTS  // per-role event log (tools/mega_trace.py); compiled out by default to keep the kernel small
  if (e.base != nullptr && e.on && e.n < 32 && (!e.brief || id == 610 || id == 501)) {
    e.base[2 * e.n] = id;
    e.base[2 * e.n + 1] = (long long)globaltimer_ns();
    ++e.n;
  }
#else
  (void)e;
  (void)id;
#endif
}

// Pulls one box of a tensor map into L2 (no shared-memory destination, no completion to wait for).
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned int* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ void pk_wait_flag(volatile uint32_t* flag, uint32_t want) {
  if (*flag >= want) return;
  const long long t0 = clock64();
  while (*flag < want) {  // plain spin: __nanosleep wakes microseconds late, and every waiter is on a critical path
    if (clock64() - t0 > 4000000000LL) mtx_wait_timeout(1, int(want), int(*flag));
  }
}

// Grid barrier k, executed by ONE thread of the CTA after the CTA's writers have met at a CTA barrier (the
// fence below is cumulative over their writes, as in cooperative-groups grid.sync()).
__device__ __forceinline__ void pk_grid_barrier(const PkParams& p, PkTail* tail, uint32_t k) {
  // The arrival counter is cumulative: a CTA with nothing to do in a phase must not run ahead and have its
  // arrival counted towards a barrier that is still collecting.
  pk_wait_flag(&tail->bar_done, k - 1);
  __threadfence();
  if (p.trace) p.trace[(2 * k) * gridDim.x + blockIdx.x] = (long long)globaltimer_ns();
  atomicAdd(p.grid_bar, 1u);
  const uint32_t target = k * gridDim.x;
  const long long t0 = clock64();
  // (Neither polling harder nor polling less pays: four relaxed polls in flight per CTA, a quarter of a round trip apart, made
  // every phase slower, 1.341 vs 1.280 ms/step -- the counter's L2 slice serves the arrivals and the polls -- and a pause of
  // 128 / 384 clocks between polls gave 1.290 / 1.301 vs 1.284: profiles/r2ac_barrier_poll.txt, r2ad_barrier_pause.txt.)
  while (ld_acquire_gpu(p.grid_bar) < target) {
    if (clock64() - t0 > 4000000000LL) mtx_wait_timeout(2, int(k), 0);
  }
  if (p.trace) p.trace[(2 * k + 1) * gridDim.x + blockIdx.x] = (long long)globaltimer_ns();
  tail->bar_done = k;
  __threadfence_block();
}

// Barrier ordinals: 1 after the embedding gather; layer l: 2+5l after QKV, 3+5l after attention, 4+5l after
// out-proj, 5+5l after MLP up, 6+5l after MLP down; then 2+5L after the final norm.  A phase may read its
// activations once `need` barriers have completed.
__device__ __forceinline__ uint32_t pk_need(int layer, int ph, int L) {
  if (ph == PK_QKV) return uint32_t(1 + 5 * layer);
  if (ph == PK_LOGITS) return uint32_t(2 + 5 * L);
  return uint32_t(2 + 5 * layer + ph);  // out-proj 3+5l, up 4+5l, down 5+5l
}

// Ring cursor: stage index and use parity of consecutive fills.
struct PkCursor {
  int s;
  uint32_t par;   // parity of the current use of stage s
  uint32_t uses;  // 0 while the ring is being filled for the first time
};
__device__ __forceinline__ void pk_cursor_init(PkCursor& c) {
  c.s = 0;
  c.par = 0;
  c.uses = 0;
}
__device__ __forceinline__ void pk_cursor_next(PkCursor& c) {
  if (++c.s == kPkStages) {
    c.s = 0;
    c.par ^= 1;
    c.uses = 1;
  }
}
__device__ __forceinline__ void pk_cursor_skip(PkCursor& c, int n) {
  const int t = c.s + n;
  const int wraps = t / kPkStages;
  c.s = t - wraps * kPkStages;
  if (wraps) c.uses = 1;
  c.par ^= uint32_t(wraps & 1);
}

__device__ __forceinline__ float4 ldcg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
// Polling loads of the flag-in-data exchanges: volatile, so the compiler can neither hoist them out of the spin
// loop nor merge re-reads; .cg keeps them out of L1.
// Pause between polling attempts without touching the memory pipeline (__nanosleep wakes microseconds late).
__device__ __forceinline__ void pk_backoff() {
  const long long t = clock64();
  while (clock64() - t < 256) {
  }
}
__device__ __forceinline__ float4 ld_poll_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float2 ld_poll_f2(const float* p) {
  float2 v;
  asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Rounds two fp32 values to bf16 (round-to-nearest-even) with ONE packed conversion and returns them as fp32.
__device__ __forceinline__ void bf16r2(float a, float b, float& ra, float& rb) {
  const uint32_t pk = pack_bf16x2(a, b);
  ra = bf16_lo(pk);
  rb = bf16_hi(pk);
}
__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) { return make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d)); }

// ---- row-major epilogues: a warp holds features n0 = tile*128 + 4*lane .. +3 of row r ----------------

// out[r, n] = bf16(resid[r, n] + bf16(acc)); also the row's partial sum of squares over this tile.
__device__ __forceinline__ void pk_epi_residual(const float4 a, const uint2 rv, int r, int tile, int lane, int N, bf16* out, float* ss,
                                                int ss_tiles) {
  const int n0 = tile * 128 + lane * 4;
  float sq = 0.0f;
  if (n0 < N) {
    float a0, a1, a2, a3;
    bf16r2(a.x, a.y, a0, a1);
    bf16r2(a.z, a.w, a2, a3);
    const uint2 pk = pack_bf16x4(bf16_lo(rv.x) + a0, bf16_hi(rv.x) + a1, bf16_lo(rv.y) + a2, bf16_hi(rv.y) + a3);
    *reinterpret_cast<uint2*>(out + (long long)r * N + n0) = pk;
    const float o0 = bf16_lo(pk.x), o1 = bf16_hi(pk.x), o2 = bf16_lo(pk.y), o3 = bf16_hi(pk.y);
    sq = o0 * o0 + o1 * o1 + o2 * o2 + o3 * o3;
  }
  sq = warp_sum(sq);
  if (lane == 0) ss[tile * kPkMaxRTile + r] = sq;
}

// linears.py:425-476 with mlp_activations [silu, linear]: rows of w01 are interleaved 16 at a time (gate, value),
// so the value of a gate feature sits 16 features (4 lanes) further.  The two lanes of a (gate, value) pair split
// the four outputs between them (two shuffles, two outputs per lane): these epilogues run as one dependent chain
// per warp, their length in instructions is their cost.
__device__ __forceinline__ void pk_epi_swiglu(const float4 a, int r, int tile, int lane, int N2, int M, bf16* act) {
  const bool gate_lane = (lane & 4) == 0;
  const float r0 = __shfl_xor_sync(0xffffffffu, gate_lane ? a.z : a.x, 4), r1 = __shfl_xor_sync(0xffffffffu, gate_lane ? a.w : a.y, 4);
  if (tile * 128 + (lane & ~4) * 4 >= N2) return;
  float g0, g1, v0, v1, s0, s1, t0, t1;
  bf16r2(gate_lane ? a.x : r0, gate_lane ? a.y : r1, g0, g1);
  bf16r2(gate_lane ? r0 : a.z, gate_lane ? r1 : a.w, v0, v1);
  bf16r2(__fdividef(1.0f, 1.0f + __expf(-g0)), __fdividef(1.0f, 1.0f + __expf(-g1)), s0, s1);
  bf16r2(g0 * s0, g1 * s1, t0, t1);
  const int m0 = tile * 64 + (lane >> 3) * 16 + (lane & 3) * 4 + (gate_lane ? 0 : 2);
  *reinterpret_cast<uint32_t*>(act + (long long)r * M + m0) = pack_bf16x2(t0 * v0, t1 * v1);
}

// Side inputs of the QKV epilogue for one row fragment: they do not depend on the accumulator, so they are
// requested together with the partial tiles.
struct PkQkvSide {
  float4 c01, c23;  // (cos, sin) pairs of the four features
  int plane, wr;
};
__device__ __forceinline__ PkQkvSide pk_qkv_side(const PkParams& p, int r, int tile, int lane) {
  PkQkvSide s;
  const int n0 = tile * 128 + lane * 4;
  const float4* cs = reinterpret_cast<const float4*>(p.rope_cs + r * 32 + (n0 & 31));
  s.c01 = cs[0];
  s.c23 = cs[1];
  s.plane = p.plane[r];
  s.wr = p.write_row[r];
  return s;
}

// embeddings.py:304-315 (half-split rotation in bf16) + kvcache.py:626-718 (append).  D = 64: the rotation partner
// of feature d is d +- 32, eight lanes away.
template <int KVQ>
__device__ __forceinline__ void pk_epi_qkv(const float4 a, const PkQkvSide& sd, int r, int tile, int lane, const PkParams& p, bf16* k_layer,
                                           bf16* v_layer, int layer) {
  float own[4], oth[4];
  bf16r2(a.x, a.y, own[0], own[1]);
  bf16r2(a.z, a.w, own[2], own[3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) oth[i] = __shfl_xor_sync(0xffffffffu, own[i], 8);
  const int n0 = tile * 128 + lane * 4;
  if (n0 >= p.qkv_n) return;
  const int head = n0 >> 6, d0 = n0 & 63;
  const bool is_q = head < p.hq, is_k = !is_q && head < p.hq + p.hkv;
  float val[4] = {own[0], own[1], own[2], own[3]};
  if (is_q || is_k) {
    const float c[4] = {sd.c01.x, sd.c01.z, sd.c23.x, sd.c23.z}, s[4] = {sd.c01.y, sd.c01.w, sd.c23.y, sd.c23.w};
    float t1[4], t2[4];
    bf16r2(own[0] * c[0], own[1] * c[1], t1[0], t1[1]);
    bf16r2(own[2] * c[2], own[3] * c[3], t1[2], t1[3]);
    bf16r2(oth[0] * s[0], oth[1] * s[1], t2[0], t2[1]);
    bf16r2(oth[2] * s[2], oth[3] * s[3], t2[2], t2[3]);
    const bool first = d0 < 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) val[i] = first ? (t1[i] - t2[i]) : (t1[i] + t2[i]);
  }
  const uint2 packed = pack_bf16x4(val[0], val[1], val[2], val[3]);
  if (is_q) {
    *reinterpret_cast<uint2*>(p.q + (long long)r * p.HD + n0) = packed;
  } else if (KVQ != 0) {
    // KVQuant.quantize over the head's 64 dims (kvcache.py:76-90), on the bf16 values the reference would have cached: the 16
    // lanes of the head agree on max|x|, every lane quantises its four values
    const float x0 = bf16_lo(packed.x), x1 = bf16_hi(packed.x), x2 = bf16_lo(packed.y), x3 = bf16_hi(packed.y);
    float mx = fmaxf(fmaxf(fabsf(x0), fabsf(x1)), fmaxf(fabsf(x2), fabsf(x3)));
    const uint32_t group = 0xFFFFu << (lane & 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(group, mx, o));
    if (sd.wr >= 0) {
      constexpr bool fp8 = KVQ == 2;
      const float inv = mx > 0.0f ? kv_quant_max(fp8) / mx : 0.0f;
      const int kvh = is_k ? head - p.hq : head - p.hq - p.hkv;
      const long long row = (long long)layer * p.kv_layer_rows + ((long long)sd.plane * p.hkv + kvh) * p.t_alloc + sd.wr;
      uint8_t* dst = (is_k ? p.kq_cache : p.vq_cache) + row * 64 + d0;
      *reinterpret_cast<uint32_t*>(dst) = kv_quant_pair(x0, x1, inv, fp8) | (kv_quant_pair(x2, x3, inv, fp8) << 16);
      if (d0 == 0) (is_k ? p.k_scale : p.v_scale)[row] = mx;
    }
  } else if (sd.wr >= 0) {
    const int kvh = is_k ? head - p.hq : head - p.hq - p.hkv;
    const long long off = (((long long)sd.plane * p.hkv + kvh) * p.t_alloc + sd.wr) * 64 + d0;
    *reinterpret_cast<uint2*>((is_k ? k_layer : v_layer) + off) = packed;
  }
}

// decoders.py:537-589: logits = bf16(dot) (* 1/sqrt(E) when tied), optional tanh soft cap, fp32.
__device__ __forceinline__ float4 pk_logit_transform(const EpiArgs& e, float4 a) {
  if (e.round_bf16) {
    bf16r2(a.x, a.y, a.x, a.y);
    bf16r2(a.z, a.w, a.z, a.w);
  }
  a.x *= e.scale; a.y *= e.scale; a.z *= e.scale; a.w *= e.scale;
  if (e.softcap != 0.0f) {
    a.x = tanhf(a.x / e.softcap) * e.softcap; a.y = tanhf(a.y / e.softcap) * e.softcap;
    a.z = tanhf(a.z / e.softcap) * e.softcap; a.w = tanhf(a.w / e.softcap) * e.softcap;
  }
  return a;
}

// Sampling partials of one parked logits tile ([rows][kPkParkPitch] fp32 accumulators in shared memory).  Two
// threads scan one row (alternating 4-wide chunks, so the 8 lanes of a shared-memory phase hit distinct banks):
// the row's best candidate of this tile (lowest index wins ties, jnp.argmax), optionally Gumbel-perturbed
// (jax.random.categorical), and the (max, sum exp) pair for log-softmax.  No cross-lane traffic but one shuffle.
__device__ __forceinline__ void pk_logits_scan(const float* park, int wtid, int rows, int tile, const EpiArgs& e, int V, uint32_t step, uint64_t seed) {
  const int r = wtid >> 1, half = wtid & 1;
  const bool active = r < rows;
  float best = -INFINITY, raw = -INFINITY, mx = -INFINITY, sum = 0.0f;
  int bi = 0x7fffffff;
  if (active) {
    const float* row = park + r * kPkParkPitch;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int c = (2 * i + half) * 4;
      const int n0 = tile * 128 + c;
      const float4 lg4 = pk_logit_transform(e, *reinterpret_cast<const float4*>(row + c));
      const float lg[4] = {lg4.x, lg4.y, lg4.z, lg4.w};
      // the four ids of the chunk share one Philox block when the shard starts at a multiple of four
      float gn[4] = {0.f, 0.f, 0.f, 0.f};
      if (e.gumbel) {
        if ((e.vocab_offset & 3) == 0) {
          const float4 g4 = gumbel_noise4(seed, step, uint32_t(e.row_offset + r), uint32_t(e.vocab_offset + n0));
          gn[0] = g4.x; gn[1] = g4.y; gn[2] = g4.z; gn[3] = g4.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) gn[q] = gumbel_noise(seed, step, uint32_t(e.row_offset + r), uint32_t(e.vocab_offset + n0 + q));
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (n0 + q < V) {
          const float sc = e.gumbel ? lg[q] * e.inv_temp + gn[q] : lg[q];
          if (sc > best) { best = sc; raw = lg[q]; bi = n0 + q; }  // ascending scan: strict > keeps the lowest index
          if (e.want_lse) {
            const float mn = fmaxf(mx, lg[q]);
            sum = sum * __expf(mx - mn) + __expf(lg[q] - mn);
            mx = mn;
          }
        }
      }
    }
  }
  {  // combine the two halves of the row
    const float b2 = __shfl_xor_sync(0xffffffffu, best, 1);
    const float r2 = __shfl_xor_sync(0xffffffffu, raw, 1);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, 1);
    if (b2 > best || (b2 == best && i2 < bi)) { best = b2; raw = r2; bi = i2; }
    if (e.want_lse) {
      const float m2 = __shfl_xor_sync(0xffffffffu, mx, 1), s2 = __shfl_xor_sync(0xffffffffu, sum, 1);
      const float mn = fmaxf(mx, m2);
      if (mn > -INFINITY) sum = sum * __expf(mx - mn) + s2 * __expf(m2 - mn);
      mx = mn;
    }
  }
  if (active && half == 0) {
    const long long o = (long long)r * e.n_tiles + tile;
    e.part_score[o] = best;
    e.part_idx[o] = e.vocab_offset + bi;
    e.part_raw[o] = raw;
    if (e.want_lse) { e.part_max[o] = mx; e.part_sum[o] = sum; }
  }
}

// Optional materialisation of the logits of one parked tile: a warp per row, 512-byte coalesced stores.
__device__ __forceinline__ void pk_logits_store(const float* park, int ew, int lane, int rows, int tile, const EpiArgs& e, int V) {
  const int n0 = tile * 128 + lane * 4;
  for (int r = ew; r < rows; r += 4) {
    float* dst = e.logits_only_row < 0 ? e.logits_out + (long long)r * e.ld_logits + n0 : (r == e.logits_only_row ? e.logits_out + n0 : nullptr);
    if (dst == nullptr) continue;
    const float4 lg = pk_logit_transform(e, *reinterpret_cast<const float4*>(park + r * kPkParkPitch + lane * 4));
    if (n0 + 3 < V && (e.ld_logits & 3) == 0) {
      *reinterpret_cast<float4*>(dst) = lg;
    } else {
      const float v[4] = {lg.x, lg.y, lg.z, lg.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (n0 + q < V) dst[q] = v[q];
    }
  }
}

// Finishes rows r_begin + w, r_begin + w + nw, ... < r_end of a shared MLP-up tile: sums the partial tiles of `parts`
// CTAs starting at c_first (flag-in-data exchange, see the GEMM phase), adds the CTA's own accumulator parked in
// shared memory (`own`, row-major, or null when the CTA's part went through the exchange like everybody's),
// applies the row's rstd and runs the SwiGLU epilogue.  Out of line and shared by the epilogue warps and the
// attention warps: the finish is bound by L2 round trips per row, so every idle warp of the CTA takes rows.
// Four rows x four parts are in flight per warp.
// (Scalars by value: a pointer to the kernel parameters would make every use a generic load from the constant bank,
// reloaded after each store.)
__device__ __noinline__ void pk_finish_swiglu(float* part_ws, bf16* act, int M, const float* rstd, const float* own, int tile, int c_first, int parts,
                                              int r_begin, int r_end, int w, int nw, int lane) {
  if (parts == 0) {  // the CTA holds the whole reduction: nothing to fetch; four rows at a time (independent chains)
    const float* src = own + lane * 4;
    for (int r = r_begin + w; r < r_end; r += 4 * nw) {
      float4 a[4];
      float rs[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rk = r + k * nw < r_end ? r + k * nw : r;
        a[k] = *reinterpret_cast<const float4*>(src + rk * 128);
        rs[k] = rstd[rk];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (r + k * nw < r_end)  // (warp-uniform)
          pk_epi_swiglu(make_float4(a[k].x * rs[k], a[k].y * rs[k], a[k].z * rs[k], a[k].w * rs[k]), r + k * nw, tile, lane, 2 * M, M, act);
    }
    return;
  }
  float* src0 = part_ws + (long long)(c_first * 4 + (tile & 3)) * kPkSlotFloats + lane * 4;
  const float sent = __uint_as_float(kPkSentinel);
  if (own != nullptr && parts <= 2) {
    // An owned tile with one or two helper parts: four rows per round, eight fragments in flight, arrival tested
    // on the sums (the sentinel is a quiet NaN).  The general loop below costs ~2.5x the instructions per row.
    float* src1 = src0 + (parts == 2 ? 4 * kPkSlotFloats : 0);  // parts == 1: the same fragment twice, counted once
    for (int r = r_begin + w; r < r_end; r += 4 * nw) {
      int rk[4];
      float4 a[4], t0[4], t1[4];
      float rs[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        rk[k] = r + k * nw < r_end ? r + k * nw : r;
        a[k] = *reinterpret_cast<const float4*>(own + rk[k] * 128 + lane * 4);
        rs[k] = rstd[rk[k]];
      }
      const long long t_spin = clock64();
      for (;;) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          t0[k] = ld_poll_f4(src0 + rk[k] * 128);
          t1[k] = ld_poll_f4(src1 + rk[k] * 128);
        }
        float chk = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) chk += ((t0[k].x + t0[k].y) + (t0[k].z + t0[k].w)) + ((t1[k].x + t1[k].y) + (t1[k].z + t1[k].w));
        bool ok = chk == chk;
        if (!ok) {
          ok = true;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ok = ok && __float_as_uint(t0[k].x) != kPkSentinel && __float_as_uint(t0[k].y) != kPkSentinel &&
                 __float_as_uint(t0[k].z) != kPkSentinel && __float_as_uint(t0[k].w) != kPkSentinel &&
                 __float_as_uint(t1[k].x) != kPkSentinel && __float_as_uint(t1[k].y) != kPkSentinel &&
                 __float_as_uint(t1[k].z) != kPkSentinel && __float_as_uint(t1[k].w) != kPkSentinel;
        }
        if (__all_sync(0xffffffffu, ok)) break;
        pk_backoff();
        if (clock64() - t_spin > 4000000000LL) mtx_wait_timeout(5, tile, r);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (r + k * nw >= r_end) continue;  // (warp-uniform)
        *reinterpret_cast<float4*>(src0 + rk[k] * 128) = make_float4(sent, sent, sent, sent);
        if (parts == 2) *reinterpret_cast<float4*>(src1 + rk[k] * 128) = make_float4(sent, sent, sent, sent);
        float4 v = a[k];
        v.x += t0[k].x; v.y += t0[k].y; v.z += t0[k].z; v.w += t0[k].w;
        if (parts == 2) { v.x += t1[k].x; v.y += t1[k].y; v.z += t1[k].z; v.w += t1[k].w; }
        pk_epi_swiglu(make_float4(v.x * rs[k], v.y * rs[k], v.z * rs[k], v.w * rs[k]), rk[k], tile, lane, 2 * M, M, act);
      }
    }
    return;
  }
  for (int r = r_begin + w; r < r_end; r += 4 * nw) {
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (own != nullptr && r + k * nw < r_end) acc[k] = *reinterpret_cast<const float4*>(own + (r + k * nw) * 128 + lane * 4);
    }
    for (int s0 = 0; s0 < parts; s0 += 4) {
      float4 t[4][4];
      const long long t_spin = clock64();
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int ss = 0; ss < 4; ++ss) {
            t[k][ss] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r + k * nw < r_end && s0 + ss < parts) t[k][ss] = ld_poll_f4(src0 + (long long)(s0 + ss) * 4 * kPkSlotFloats + (r + k * nw) * 128);
          }
        // (the sentinel is a quiet NaN: test the sum first, compare word by word only when it is NaN)
        float chk = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int ss = 0; ss < 4; ++ss) chk += (t[k][ss].x + t[k][ss].y) + (t[k][ss].z + t[k][ss].w);
        if (chk != chk) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int ss = 0; ss < 4; ++ss)
              ok = ok && __float_as_uint(t[k][ss].x) != kPkSentinel && __float_as_uint(t[k][ss].y) != kPkSentinel &&
                   __float_as_uint(t[k][ss].z) != kPkSentinel && __float_as_uint(t[k][ss].w) != kPkSentinel;
        }
        if (__all_sync(0xffffffffu, ok)) break;
        pk_backoff();
        if (clock64() - t_spin > 4000000000LL) mtx_wait_timeout(5, tile, r);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) {
          if (r + k * nw < r_end && s0 + ss < parts)
            *reinterpret_cast<float4*>(src0 + (long long)(s0 + ss) * 4 * kPkSlotFloats + (r + k * nw) * 128) = make_float4(sent, sent, sent, sent);
          acc[k].x += t[k][ss].x; acc[k].y += t[k][ss].y; acc[k].z += t[k][ss].z; acc[k].w += t[k][ss].w;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (r + k * nw < r_end) {
        const float rs = rstd[r + k * nw];
        pk_epi_swiglu(make_float4(acc[k].x * rs, acc[k].y * rs, acc[k].z * rs, acc[k].w * rs), r + k * nw, tile, lane, 2 * M, M, act);
      }
  }
}
// One unit of the MLP-up phase as seen by the finishing warps: S > 1, the tile is shared evenly (every sharer dumps
// its partial and finishes a row slice); S < 0, this CTA owns the tile (its own accumulator is parked on chip,
// -S helper CTAs starting at c_first contribute partials, the owner finishes every row); S == 0, helper (dumps
// only); S == 1, the CTA holds the whole reduction.
__device__ __forceinline__ void pk_finish_up_unit(const PkParams* pp, const float* rstd, const float* park, const PkUnit& un, bool solo, int cta, int w,
                                                  int nw, int lane) {
  if (un.S > 1) {
    const int si = cta - un.c_first;
    pk_finish_swiglu(pp->part_ws, pp->act, pp->M, rstd, nullptr, un.tile, un.c_first, un.S, si * pp->rows / un.S, (si + 1) * pp->rows / un.S, w, nw, lane);
  } else if (un.S < 0 || (un.S == 1 && solo)) {  // (solo: a whole tile that is the CTA's only unit of the phase, see pk_up_all_warps)
    pk_finish_swiglu(pp->part_ws, pp->act, pp->M, rstd, park, un.tile, un.c_first, un.S < 0 ? -un.S : 0, 0, pp->rows, w, nw, lane);
  }
}
// MLP up: do all warps of the CTA finish its units?  Yes when a tile is shared, and also for a single whole tile:
// one row's SwiGLU epilogue is a dependent chain of ~60 instructions, 64 rows on four warps take ~3 us.
__device__ __forceinline__ bool pk_up_all_warps(const PkTable& tab) {
  bool any = tab.n_units[PK_UP] == 1 && tab.u[PK_UP][0].S == 1;
  for (int u = 0; u < tab.n_units[PK_UP]; ++u) any = any || tab.u[PK_UP][u].S > 1 || tab.u[PK_UP][u].S < 0;
  return any;
}

// ---- attention: stream-K over 64-row KV tiles -----------------------------------------------------
//
// The valid KV tiles of all (row, kv head) pairs form one list (row-major over row, head, tile).  The active
// CTAs take equal contiguous runs of it and the five attention warps of a CTA equal contiguous sub-runs, so
// every warp streams the same number of bytes whatever the context lengths are.  A warp's run is a sequence
// of segments (maximal pieces inside one pair).  A pair that lies inside one warp is finished there; a pair
// cut between warps of one CTA is merged in SHARED memory after the CTA's warps have met; only the pairs cut
// between CTAs go through an L2 workspace (one part per warp, written when the warp's segment ends; the CTA
// that holds the pair's first tile merges).
// The partition depends on the context lengths only, so it is built once per step and reused by every layer.

enum PkAttnMeta : int {
  PKA_SEG_START = 1 << 20,   // first tile of a segment of this warp
  PKA_SEG_END = 1 << 21,     // last tile of a segment of this warp
  PKA_PAIR_FIRST = 1 << 22,  // tile 0 of its pair
  PKA_PAIR_LAST = 1 << 23,   // last tile of its pair: holds the row appended this step
  PKA_TAIL = 1 << 24,        // end of a segment whose pair began in another CTA: the partial goes to that CTA through L2
};
// meta bits [0,6) = valid rows - 1, [6,14) = row, [14,20) = kv head
__device__ __forceinline__ int pka_cnt(int m) { return (m & 63) + 1; }
__device__ __forceinline__ int pka_row(int m) { return (m >> 6) & 255; }
__device__ __forceinline__ int pka_head(int m) { return (m >> 14) & 63; }

struct PkAttnPos {
  int r, h, t, nt;
  int len0, rf, rl;
  int plane_row;  // cache row of (layer 0, plane, kv head) in the K/V tensor maps
};

__device__ __forceinline__ void pk_attn_load_row(PkAttnPos& a, const PkParams& p, const PkTail* tail, int R) {
  a.len0 = tail->r_len0[a.r];
  a.rf = tail->r_rf[a.r];
  a.rl = tail->r_rl[a.r];
  a.nt = attn_num_tiles(a.len0, a.rf, a.rl, R);
  a.plane_row = (tail->r_plane[a.r] * p.hkv + a.h) * p.T;
}

// Position of flattened tile index g.
__device__ __forceinline__ void pk_attn_seek(PkAttnPos& a, long long g, const PkParams& p, const PkTail* tail, int R, int lane) {
  int cnt = 0;
  for (int r0 = 0; r0 < p.rows; r0 += 32) {
    const int r = r0 + lane;
    const bool le = r < p.rows && (long long)tail->r_prefix[r] * p.hkv <= g;
    cnt += __popc(__ballot_sync(0xffffffffu, le));
  }
  a.r = cnt - 1;
  a.h = 0;
  pk_attn_load_row(a, p, tail, R);
  const int rem = int(g - (long long)tail->r_prefix[a.r] * p.hkv);
  a.h = rem / a.nt;
  a.t = rem - a.h * a.nt;
  a.plane_row += a.h * p.T;
}

__device__ __forceinline__ void pk_attn_next(PkAttnPos& a, const PkParams& p, const PkTail* tail, int R) {
  if (++a.t < a.nt) return;
  a.t = 0;
  if (++a.h < p.hkv) {
    a.plane_row += p.T;
    return;
  }
  a.h = 0;
  if (++a.r < p.rows) pk_attn_load_row(a, p, tail, R);
}

// CTA that owns flattened tile g when `nc` CTAs share `total` tiles.
// First tile of CTA c when `nc` CTAs share `total` tiles.  skew = 0: equal shares.  skew > 0 (experiment, MTX_PK_ATTN_SKEW): CTA c's
// share is proportional to 1 + skew (1/2 - c / nc) -- the CTAs with high indices finish their tile loops ~1 us later than the first
// ones at equal shares (profiles/r2z_dataflow_trace.txt, per-decile loop ends).
__device__ __forceinline__ int pk_cta_lo(int c, int nc, int total, float skew) {
  if (c >= nc) return total;
  if (skew == 0.0f) return int((long long)c * total / nc);
  const double x = double(c) / double(nc);
  return int((x * (1.0 + 0.5 * double(skew)) - 0.5 * double(skew) * x * x) * double(total));
}
__device__ __forceinline__ int pk_cta_of(long long g, long long nc, long long total, float skew) {
  if (skew == 0.0f) return int(((g + 1) * nc - 1) / total);
  int c = int(g * nc / total);
  while (c > 0 && pk_cta_lo(c, int(nc), int(total), skew) > g) --c;
  while (c + 1 < nc && pk_cta_lo(c + 1, int(nc), int(total), skew) <= g) ++c;
  return c;
}

// Tiles [wlo, whi) of attention warp a of CTA c (equal contiguous runs per CTA, equal sub-runs per warp).
__device__ __forceinline__ void pk_warp_range(int c, int a, int nc, int total, float skew, int& wlo, int& whi) {
  const int clo = pk_cta_lo(c, nc, total, skew), chi = pk_cta_lo(c + 1, nc, total, skew);
  wlo = clo + a * (chi - clo) / kPkAttnWarps;
  whi = clo + (a + 1) * (chi - clo) / kPkAttnWarps;
}
// Number of warps that hold tiles, from warp 0 of CTA c0 up to (not including) the first warp starting at or after g_stop.
__device__ __noinline__ int pk_count_warps(int c0, int g_stop, int nc, int total, float skew) {
  int count = 0;
  for (int c = c0; c < nc; ++c)
    for (int a = 0; a < kPkAttnWarps; ++a) {
      int wlo, whi;
      pk_warp_range(c, a, nc, total, skew, wlo, whi);
      if (wlo >= g_stop) return count;
      if (whi > wlo) ++count;
    }
  return count;
}

// Once per step: the tile list of attention warp `aw` of this CTA, and the CTA's place in the exchange of the
// pairs cut between CTAs.  A pair is merged by the CTA that holds its first tile (the "head"); every warp of a
// later CTA that holds tiles of the pair (the "tail") writes its partial straight from registers into its own
// slot of the pair's L2 workspace when its segment ends, i.e. DURING the tile loop: by the time the head CTA
// has merged its own warps the tail parts are already there.  Slots are numbered by the warp's ordinal among the
// tail warps; both sides derive it from the partition alone.
__device__ __forceinline__ void pk_attn_build_list(const PkParams& p, PkTail* tail, int cta, int aw, int lane) {
  const int nc = p.attn_info[0], total = p.attn_info[1];
  const bool pair_mode = p.attn_info[2] != 0;  // CTA c owns the whole pair c (few short pairs: nothing crosses CTAs)
  const int R = p.T - p.P;
  int n = 0, tail_slot = -1;
  if (cta < nc && total > 0) {
    const float skew = p.attn_skew;
    int clo = pk_cta_lo(cta, nc, total, skew), chi = pk_cta_lo(cta + 1, nc, total, skew);
    if (pair_mode) {
      const int r = cta / p.hkv, h = cta - r * p.hkv;
      const int nt = attn_num_tiles(tail->r_len0[r], tail->r_rf[r], tail->r_rl[r], R);
      clo = tail->r_prefix[r] * p.hkv + h * nt;
      chi = clo + nt;
    }
    const int wlo = clo + aw * (chi - clo) / kPkAttnWarps, whi = clo + (aw + 1) * (chi - clo) / kPkAttnWarps;
    n = whi - wlo;
    if (n > kPkAttnListMax) n = kPkAttnListMax;  // excluded by the host-side check (pk_usable)
    if (n > 0) {
      PkAttnPos pos;
      pk_attn_seek(pos, wlo, p, tail, R, lane);
      // the warp's first segment is a tail if its pair began in an earlier CTA
      const int g_first = tail->r_prefix[pos.r] * p.hkv + pos.h * pos.nt;
      bool is_tail = !pair_mode && g_first < clo;
      if (is_tail) {
        const int c_first = pk_cta_of(g_first, nc, total, skew);
        const int ordinal = pk_count_warps(c_first + 1, wlo, nc, total, skew);
        if (ordinal >= kPkMaxParts) __trap();  // excluded by prepare_rows_kernel's choice of nc
        tail_slot = (pos.r * p.hkv + pos.h) * kPkMaxParts + ordinal;
      }
      for (int i = 0; i < n; ++i) {
        const TileLoc loc = attn_tile(pos.t, pos.len0, pos.rf, pos.rl, p.P, R);
        int meta = (loc.cnt - 1) | (pos.r << 6) | (pos.h << 14);
        if (i == 0 || pos.t == 0) meta |= PKA_SEG_START;
        if (pos.t == 0) meta |= PKA_PAIR_FIRST;
        if (i == n - 1 || pos.t == pos.nt - 1) {
          meta |= PKA_SEG_END;
          if (is_tail) meta |= PKA_TAIL;
          is_tail = false;
        }
        if (pos.t == pos.nt - 1) meta |= PKA_PAIR_LAST;
        if (lane == 0) tail->a_list[aw][i] = make_int2(pos.plane_row + loc.p0, meta);
        pk_attn_next(pos, p, tail, R);
      }
    }
    if (p.page_map != nullptr) {
      // Paged: .x so far is kv head * T + the tile's first token; turn it into the pool row of the page that holds the token,
      // one tile per lane.  Once per step: the pages of a sequence are the same in every layer.
      __syncwarp();
      if (lane < n) {
        const int2 e = tail->a_list[aw][lane];
        const int h = pka_head(e.y), tok = e.x - h * p.T;  // (group rows are plane 0: plane_row = h * T)
        const int* pm = p.page_map + (long long)pka_row(e.y) * p.max_pages;
        const int pt = p.page_tokens;
        const int first = pm[tok / pt];
        const int row = (h * p.num_pages + first) * pt + (tok & (pt - 1));
        // (a tile whose valid rows end inside its first 32-token page loads that page twice: entries of the page map past the
        //  sequence's pages are not valid)
        const int second = (pt == 32 && pka_cnt(e.y) > 32) ? pm[tok / pt + 1] : first;
        tail->a_list[aw][lane].x = row;
        tail->a_row2[aw][lane] = (h * p.num_pages + second) * pt;
      }
      __syncwarp();
    }
    if (aw == 0) {
      // Merge plan.  Walk the pairs that start inside [clo, chi); a pair that leaves the warp it starts in needs a
      // merge.  Its partial in warp w is w's last segment (parked in w's K buffer) when it runs to the end of w's
      // range, else w's first segment (parked in w's slot).
      int nj = 0, g = clo;
      while (g < chi) {
        PkAttnPos pos;
        pk_attn_seek(pos, g, p, tail, R, lane);
        const int g_first = g - pos.t, g_end = g_first + pos.nt;
        if (g_first >= clo) {
          int w_a = 0;
          while (w_a < kPkAttnWarps - 1 && clo + (w_a + 1) * (chi - clo) / kPkAttnWarps <= g_first) ++w_a;
          if (g_end > clo + (w_a + 1) * (chi - clo) / kPkAttnWarps && nj < kPkAttnWarps + 1) {
            PkTail::Job jb;
            jb.pair = pos.r * p.hkv + pos.h;
            jb.row = pos.r;
            jb.head = pos.h;
            jb.n_src = 0;
            for (int w = w_a; w < kPkAttnWarps; ++w) {
              const int wl = clo + w * (chi - clo) / kPkAttnWarps, wh = clo + (w + 1) * (chi - clo) / kPkAttnWarps;
              if (wl >= g_end) break;
              if (wh > wl) jb.src[jb.n_src++] = wh <= g_end ? w : 8 + w;
            }
            jb.n_parts = g_end > chi ? pk_count_warps(cta + 1, g_end, nc, total, skew) : 0;
            if (lane == 0) tail->a_jobs[nj] = jb;
            ++nj;
          }
        }
        g = g_end;
      }
      if (lane == 0) tail->a_njobs = nj;
    }
  } else if (aw == 0 && lane == 0) {
    tail->a_njobs = 0;
  }
  if (lane == 0) {
    tail->a_count[aw] = n;
    tail->a_tail[aw] = tail_slot;
  }
}

// Phase 2 of the CTA's attention: the merge jobs of the static plan (pk_attn_build_list).  Partials are [head][64] fp32 arrays
// followed by (m, l) per head, whatever arithmetic layout produced them.
__device__ __forceinline__ void pk_attention_merge(const PkParams& p, PkTail* tail, uint8_t* attn_tiles, int aw, int lane, PkEv& ev) {
  constexpr int D = 64;
  constexpr float kLog2e = 1.4426950408889634f;
  const int G = p.hq / p.hkv;
  // ---- phase 2: the CTA's merge jobs (static plan, see pk_attn_build_list) ----
  if (lane == 0) pk_ev(ev, 610);
  named_bar_sync(3, kPkAttnWarps * 32);
  if (lane == 0) pk_ev(ev, 611);
  // One thread per (head, 4 dims) unit of a job, so 160 / (16 G) jobs run side by side; a thread folds the job's
  // on-chip partials, then the parts of later CTAs' warps from L2 (flag-in-data: a word is the sentinel or data;
  // they were written during those warps' tile loops, so normally all are present), four parts per round trip.
  // Fixed order: deterministic.  A thread puts the words it has read back to the sentinel.
  const int n_units = G * (D / 4);
  const int t = aw * 32 + lane;
  const int lanes = (kPkAttnWarps * 32) / n_units;  // jobs in flight
  const int jl = t / n_units, unit = t - jl * n_units;
  const int gq = unit / (D / 4), d4 = unit - gq * (D / 4);
  const int stride = (G * D + 2 * G + 3) & ~3;
  const float sent = __uint_as_float(kPkSentinel);
  const uint32_t head_mask = 0xFFFFu << (lane & 16);  // the 16 threads of a head sit in one half warp
  if (jl < lanes) {
    for (int j = jl; j < tail->a_njobs; j += lanes) {
      const int n_src = tail->a_jobs[j].n_src, n_parts = tail->a_jobs[j].n_parts, pair = tail->a_jobs[j].pair;
      float M = -INFINITY, Ls = 0.0f;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      auto fold = [&](const float2 ml, const float4 o4) {
        const float Mn = fmaxf(M, ml.x);
        const float so = ex2_approx((M - Mn) * kLog2e), sn = ex2_approx((ml.x - Mn) * kLog2e);
        Ls = Ls * so + ml.y * sn;
        acc.x = acc.x * so + o4.x * sn; acc.y = acc.y * so + o4.y * sn;
        acc.z = acc.z * so + o4.z * sn; acc.w = acc.w * so + o4.w * sn;
        M = Mn;
      };
      float* gbase = p.attn_part_o + (long long)pair * kPkMaxParts * stride;
      // the first round of remote parts is requested before the on-chip partials are folded
      float2 ml[4];
      float4 o4[4];
      auto request = [&](int c0) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          if (c0 + cc < n_parts) {
            const float* part = gbase + (long long)(c0 + cc) * stride;
            ml[cc] = ld_poll_f2(part + G * D + gq * 2);
            o4[cc] = ld_poll_f4(part + gq * D + d4 * 4);
          }
      };
      request(0);
      for (int c = 0; c < n_src; ++c) {  // on-chip partials, folded like the remote ones (online rescaling)
        const int code = tail->a_jobs[j].src[c];
        const float* src = code < 8 ? reinterpret_cast<const float*>(attn_tiles + code * 2 * 8192) : tail->a_slot[code - 8];
        fold(*reinterpret_cast<const float2*>(src + G * D + gq * 2), *reinterpret_cast<const float4*>(src + gq * D + d4 * 4));
      }
      for (int c0 = 0; c0 < n_parts; c0 += 4) {
        const long long t_spin = clock64();
        for (;;) {
          bool ok = true;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            if (c0 + cc < n_parts)
              ok = ok && __float_as_uint(ml[cc].x) != kPkSentinel && __float_as_uint(ml[cc].y) != kPkSentinel &&
                   __float_as_uint(o4[cc].x) != kPkSentinel && __float_as_uint(o4[cc].y) != kPkSentinel &&
                   __float_as_uint(o4[cc].z) != kPkSentinel && __float_as_uint(o4[cc].w) != kPkSentinel;
          if (ok) break;
          pk_backoff();
          if (clock64() - t_spin > 4000000000LL) mtx_wait_timeout(4, pair, c0);
          request(c0);
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          if (c0 + cc < n_parts) {
            *reinterpret_cast<float4*>(gbase + (long long)(c0 + cc) * stride + gq * D + d4 * 4) = make_float4(sent, sent, sent, sent);
            fold(ml[cc], o4[cc]);
          }
        if (c0 + 4 < n_parts) request(c0 + 4);
      }
      const float inv = 1.0f / Ls;
      *reinterpret_cast<uint2*>(p.attn + (long long)tail->a_jobs[j].row * p.hq * D + (long long)tail->a_jobs[j].head * G * D + gq * D + d4 * 4) =
          pack_bf16x4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
      if (n_parts > 0) {
        __syncwarp(head_mask);  // all 16 threads of the head have read its (m, l) words
        if (d4 == 0)
          for (int c = 0; c < n_parts; ++c) *reinterpret_cast<float2*>(gbase + (long long)c * stride + G * D + gq * 2) = make_float2(sent, sent);
      }
    }
  }
  if (lane == 0) pk_ev(ev, 691);
}

// One thread requests a 64-row K (or V) tile: cache row `row` of the layer, or -- 32-token pages -- the two pages at `row`, `row2`.
__device__ __forceinline__ void pk_issue_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, const PkParams& p, int layer_row, int row,
                                              int row2, uint32_t bytes = 8192) {
  mbar_expect_tx(bar, bytes);
  if (p.page_map != nullptr && p.page_tokens == 32) {
    tma_load_2d(dst, tm, 0, layer_row + row, bar, kEvictFirst);
    tma_load_2d(dst + 4096, tm, 0, layer_row + row2, bar, kEvictFirst);
  } else {
    tma_load_2d(dst, tm, 0, layer_row + row, bar, kEvictFirst);
  }
}

// The whole CTA's attention for one layer, executed by the attention warps.
//   phase 1: every warp walks its tile list (TMA K/V tiles -> S = Q K^T -> online softmax -> O += P V)
//   phase 2: partial segments are merged in shared memory; pairs shared with other CTAs go through L2
// `primed`: the first K/V tiles were already requested (before the grid barrier).
__device__ __forceinline__ void pk_attention_cta(const CUtensorMap& tm_k, const CUtensorMap& tm_v, const PkParams& p, PkTail* tail, int layer,
                                                 uint8_t* attn_tiles, uint32_t& phase, int cta, int aw, int lane, bool primed, PkEv& ev) {
  constexpr int D = 64;
  constexpr float kLog2e = 1.4426950408889634f;
  const int G = p.hq / p.hkv;
  const int layer_row = layer * p.kv_layer_rows;
  const int* list2 = tail->a_row2[aw];
  const int gid = lane >> 2, tid4 = lane & 3;
  uint8_t* k_tile = attn_tiles + aw * 2 * 8192;
  uint8_t* v_tile = k_tile + 8192;
  uint64_t* bar_k = &tail->attn_bars[2 * aw];
  uint64_t* bar_v = bar_k + 1;
  const uint32_t kb = smem_u32(k_tile), vb = smem_u32(v_tile);
  const int n = tail->a_count[aw];
  const int2* list = tail->a_list[aw];

  if (n > 0 && !primed && lane == 0) {
    fence_proxy_async_all();  // K/V rows appended by the QKV epilogue (generic stores) are read through TMA
    pk_issue_tile(k_tile, &tm_k, bar_k, p, layer_row, list[0].x, list2[0]);
    pk_issue_tile(v_tile, &tm_v, bar_v, p, layer_row, list[0].x, list2[0]);
  }

  // Transposed formulation: S^T = K Q^T and O^T = V^T P^T, so the 16-row MMA dimension carries KV rows / head
  // dimensions and the 8-column dimension carries the (up to 8) query heads of the group: no padding of the
  // heads to 16, half the MMAs, exponentials and accumulator registers of the head-major form.
  //   thread (gid, tid4) owns heads h0 = 2 tid4, h1 = h0 + 1
  //   s[mb][0..3]  = S^T[kv = 16 mb + gid (+8 for 2,3)][h0, h1]
  //   o[db][0..3]  = O^T[d  = 16 db + gid (+8 for 2,3)][h0, h1]
  uint32_t qb[D / 16][2];
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;  // per head h0 / h1 (l: this thread's rows only)
  float o[D / 16][4];
  bool seg_first = false;   // the running segment starts at tile 0 of its pair
  const int a_row = (lane & 7) + 8 * ((lane >> 3) & 1), a_chunk = lane >> 4;  // ldmatrix addressing, K tile as A
  const int v_row = (lane & 7) + 8 * (lane >> 4), v_chunk = (lane >> 3) & 1;   // ldmatrix.trans addressing, V^T as A

  for (int i = 0; i < n; ++i) {
    const int2 e = list[i];
    const int meta = e.y;
    const int r = pka_row(meta), h = pka_head(meta), cnt = pka_cnt(meta);
    if (meta & PKA_SEG_START) {
      const bf16* qrow = p.q + (long long)r * p.hq * D + (long long)h * G * D + gid * D + tid4 * 2;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        qb[kk][0] = gid < G ? __ldcg(reinterpret_cast<const uint32_t*>(qrow + kk * 16)) : 0u;
        qb[kk][1] = gid < G ? __ldcg(reinterpret_cast<const uint32_t*>(qrow + kk * 16 + 8)) : 0u;
      }
      m0 = m1 = -INFINITY;
      l0 = l1 = 0.0f;
#pragma unroll
      for (int db = 0; db < D / 16; ++db) o[db][0] = o[db][1] = o[db][2] = o[db][3] = 0.0f;
      seg_first = (meta & PKA_PAIR_FIRST) != 0;
    }
    const bool more = i + 1 < n;

    // ---- S^T = K Q^T over the 64 rows of the tile ----
    float s[4][4];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) s[mb][0] = s[mb][1] = s[mb][2] = s[mb][3] = 0.0f;
    mbar_wait(bar_k, phase);
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const int row = 16 * mb + a_row;
        const uint32_t addr = kb + row * 128 + ((((2 * kk + a_chunk) & 7) ^ (row & 7)) << 4);
        uint32_t a[4];
        ldmatrix_x4(addr, a[0], a[1], a[2], a[3]);
        mma_m16n8k16_bf16(s[mb], a, qb[kk][0], qb[kk][1]);
      }
    }
    // K tile consumed: refill it with the next tile's keys while the softmax and P V run
    fence_proxy_async();
    __syncwarp();
    if (more && lane == 0) pk_issue_tile(k_tile, &tm_k, bar_k, p, layer_row, list[i + 1].x, list2[i + 1]);
    // ---- mask + online softmax: a head's scores live in the 8 threads of equal tid4 ----
    if (p.softcap != 0.0f) {
#pragma unroll
      for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int q = 0; q < 4; ++q) s[mb][q] = tanhf(s[mb][q] / p.softcap) * p.softcap;
    }
    if (cnt < 64) {
#pragma unroll
      for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (16 * mb + gid + (q >> 1) * 8 >= cnt) s[mb][q] = -INFINITY;
    }
    float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      tm0 = fmaxf(tm0, fmaxf(s[mb][0], s[mb][2]));
      tm1 = fmaxf(tm1, fmaxf(s[mb][1], s[mb][3]));
    }
#pragma unroll
    for (int sh = 4; sh < 32; sh <<= 1) {
      tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, sh));
      tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, sh));
    }
    const float nm0 = fmaxf(m0, tm0), nm1 = fmaxf(m1, tm1);
    const float a0 = ex2_approx((m0 - nm0) * kLog2e), a1 = ex2_approx((m1 - nm1) * kLog2e);  // ex2(-inf) = 0 on the first tile
    m0 = nm0;
    m1 = nm1;
    l0 *= a0;
    l1 *= a1;
    const float ms0 = m0 * kLog2e, ms1 = m1 * kLog2e;
    uint32_t pb[4][2];  // P^T as the B operand of O^T += V^T P^T, one k-step per 16 KV rows
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      const float p0 = ex2_approx(fmaf(s[mb][0], kLog2e, -ms0)), p1 = ex2_approx(fmaf(s[mb][1], kLog2e, -ms1));
      const float p2 = ex2_approx(fmaf(s[mb][2], kLog2e, -ms0)), p3 = ex2_approx(fmaf(s[mb][3], kLog2e, -ms1));
      l0 += p0 + p2;
      l1 += p1 + p3;
      // probabilities are cast to the value dtype before the PV product (kernels/ragged_attention.py:156);
      // the transpose turns (kv row, head pair) fragments into (head, kv row pair) fragments
      pb[mb][0] = movmatrix_trans(pack_bf16x2(p0, p1));
      pb[mb][1] = movmatrix_trans(pack_bf16x2(p2, p3));
    }
#pragma unroll
    for (int db = 0; db < D / 16; ++db) {
      o[db][0] *= a0;
      o[db][1] *= a1;
      o[db][2] *= a0;
      o[db][3] *= a1;
    }
    // ---- O^T += V^T P^T ----
    mbar_wait(bar_v, phase);
    if (cnt < 64) {  // rows past the valid count may hold anything: zero them (0 * NaN != 0)
      const int nvec = (64 - cnt) * 8;
      for (int q = lane; q < nvec; q += 32) *reinterpret_cast<uint4*>(v_tile + (cnt + q / 8) * 128 + (q & 7) * 16) = make_uint4(0, 0, 0, 0);
      __syncwarp();
    }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
#pragma unroll
      for (int db = 0; db < D / 16; ++db) {
        const int row = 16 * mb + v_row;
        const uint32_t addr = vb + row * 128 + ((((2 * db + v_chunk) & 7) ^ (row & 7)) << 4);
        uint32_t a[4];
        ldmatrix_x4_trans(addr, a[0], a[1], a[2], a[3]);
        mma_m16n8k16_bf16(o[db], a, pb[mb][0], pb[mb][1]);
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (more && lane == 0) pk_issue_tile(v_tile, &tm_v, bar_v, p, layer_row, list[i + 1].x, list2[i + 1]);
    phase ^= 1;

    if (meta & PKA_SEG_END) {
#pragma unroll
      for (int sh = 4; sh < 32; sh <<= 1) {
        l0 += __shfl_xor_sync(0xffffffffu, l0, sh);
        l1 += __shfl_xor_sync(0xffffffffu, l1, sh);
      }
      const int h0 = tid4 * 2;
      if (seg_first && (meta & PKA_PAIR_LAST)) {
        // the whole pair lived in this warp: normalise, transpose back to (head, d pair) fragments and store
        const long long out_base = (long long)r * p.hq * D + (long long)h * G * D;
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
        for (int db = 0; db < D / 16; ++db) {
          const uint32_t lo = movmatrix_trans(pack_bf16x2(o[db][0] * i0, o[db][1] * i1));  // (head gid, d = 16 db + 2 tid4 ..)
          const uint32_t hi = movmatrix_trans(pack_bf16x2(o[db][2] * i0, o[db][3] * i1));  // (head gid, d = 16 db + 8 + 2 tid4 ..)
          if (gid < G) {
            *reinterpret_cast<uint32_t*>(p.attn + out_base + gid * D + 16 * db + tid4 * 2) = lo;
            *reinterpret_cast<uint32_t*>(p.attn + out_base + gid * D + 16 * db + 8 + tid4 * 2) = hi;
          }
        }
      } else {
        // park the partial ([head][d] fp32, then (m, l) per head): the warp's last segment goes to its (now idle)
        // K buffer, an earlier one to its slot; a tail segment goes to its slot of the pair's L2 workspace
        const bool final_seg = !more, to_l2 = (meta & PKA_TAIL) != 0;
        const int stride = (G * D + 2 * G + 3) & ~3;  // 16-byte aligned parts
        float* dst = to_l2 ? p.attn_part_o + (long long)tail->a_tail[aw] * stride : final_seg ? reinterpret_cast<float*>(k_tile) : tail->a_slot[aw];
#pragma unroll
        for (int db = 0; db < D / 16; ++db) {
          const int d = 16 * db + gid;
          if (h0 < G) {
            dst[h0 * D + d] = o[db][0];
            dst[h0 * D + d + 8] = o[db][2];
          }
          if (h0 + 1 < G) {
            dst[(h0 + 1) * D + d] = o[db][1];
            dst[(h0 + 1) * D + d + 8] = o[db][3];
          }
        }
        if (gid == 0) {
          if (h0 < G) *reinterpret_cast<float2*>(dst + G * D + h0 * 2) = make_float2(m0, l0);
          if (h0 + 1 < G) *reinterpret_cast<float2*>(dst + G * D + (h0 + 1) * 2) = make_float2(m1, l1);
        }
      }
    }
  }

  // The HBM is idle for most of the GEMM phases that follow: ask for the next layer's K/V tiles of this warp now, so that its
  // next tile loop streams from L2 (the list is the same for every layer; a tile is 8 KB, a warp holds ~6).
  if ((p.variant & 1) && layer + 1 < p.L && p.page_map == nullptr) {
    const int next_layer_row = layer_row + p.kv_layer_rows;
    for (int i = lane; i < n; i += 32) {
      tma_prefetch_l2_2d(&tm_k, 0, next_layer_row + list[i].x);
      tma_prefetch_l2_2d(&tm_v, 0, next_layer_row + list[i].x);
    }
  }
  pk_attention_merge(p, tail, attn_tiles, aw, lane, ev);
}

// The same phase over a quantised cache (step_persistent_kernel<KVQ != 0>; F8: float8_e4m3fn bytes, else u = q + 128).  The tile
// walk, segment bookkeeping and merge plan are those of pk_attention_cta; the arithmetic of a tile is attn_process_items_q8's
// (attention.cuh) -- 4 KB byte tiles (SWIZZLE_64B), bytes converted exactly to fp16 in registers, fp16 MMAs, permuted head dims,
// k_scale / v_scale folded into the scores / probabilities -- in the transposed form of pk_attention_cta (half the MMAs).  Byte tiles are half
// the size, so a warp's 16 KB buffer is TWO stages of [K tile 4 KB | V tile 4 KB]: tile i + 2 is requested when tile i has been
// consumed and the whole buffer stays in flight (with one stage the loop was latency-bound: 30 us per layer against 18 for bf16).
// One mbarrier per stage (the K and the V tile of a stage complete together); bit s of `phase` is the parity of stage s.
// Partials are parked over stage 0.
template <bool F8>
__device__ __forceinline__ void pk_attention_cta_q8(const CUtensorMap& tm_k, const CUtensorMap& tm_v, const PkParams& p, PkTail* tail, int layer,
                                                    uint8_t* attn_tiles, uint32_t& phase, int cta, int aw, int lane, bool primed, PkEv& ev) {
  constexpr int D = 64;
  constexpr uint32_t kTileBytes = 64 * D;
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr float kInv = F8 ? kF8Inv : kQ8Inv;
  const int G = p.hq / p.hkv;
  const int layer_row = layer * p.kv_layer_rows;
  const int gid = lane >> 2, tid4 = lane & 3;
  uint8_t* stage0 = attn_tiles + aw * 2 * 8192;
  float* s_ks = tail->a_scales[aw];
  float* s_vs = s_ks + 64;
  uint64_t* bars = &tail->attn_bars[2 * aw];  // [stage]
  const int n = tail->a_count[aw];
  const int2* list = tail->a_list[aw];

  auto issue = [&](int i) {  // one thread: K and V tile i into stage i & 1
    uint8_t* st = stage0 + (i & 1) * 8192;
    mbar_expect_tx(&bars[i & 1], 2 * kTileBytes);
    tma_load_2d(st, &tm_k, 0, layer_row + list[i].x, &bars[i & 1], kEvictFirst);
    tma_load_2d(st + kTileBytes, &tm_v, 0, layer_row + list[i].x, &bars[i & 1], kEvictFirst);
  };
  if (lane == 0) {
    fence_proxy_async_all();  // rows appended by the QKV epilogue (generic stores) are read through TMA
    if (n > 0 && !primed) issue(0);
    if (n > 1) issue(1);
  }
  // The scales of a tile (two K and two V rows per lane) are requested TWO tiles ahead, when the tile's bytes are: an L2 round trip
  // under load is longer than one tile's arithmetic.  L2 loads (ld.cg): this step's row was written by another SM.
  struct Scales { float k0, k1, v0, v1; };
  auto request_scales = [&](int i) {
    Scales sc;
    const long long row0 = (long long)layer_row + list[i].x;
    const int c = pka_cnt(list[i].y);
    sc.k0 = lane < c ? __ldcg(p.k_scale + row0 + lane) : 0.0f;
    sc.k1 = lane + 32 < c ? __ldcg(p.k_scale + row0 + lane + 32) : 0.0f;
    sc.v0 = lane < c ? __ldcg(p.v_scale + row0 + lane) : 0.0f;
    sc.v1 = lane + 32 < c ? __ldcg(p.v_scale + row0 + lane + 32) : 0.0f;
    return sc;
  };
  Scales sc_cur = {0.f, 0.f, 0.f, 0.f}, sc_next = sc_cur;
  if (n > 0) sc_cur = request_scales(0);
  if (n > 1) sc_next = request_scales(1);

  // Transposed products, as in pk_attention_cta: S^T = K Q^T and O^T = V^T P^T, the 16-row MMA dimension carrying kv rows /
  // head dims and the 8-column dimension the (up to 8) heads of the group.
  //   thread (gid, tid4) owns heads h0 = 2 tid4, h1 = h0 + 1
  //   s[mb][0..3] = S^T[kv = 16 mb + gid (+8 for 2,3)][h0, h1]
  //   o[db][0..3] = O^T[d = 8 gid + 2 db (+1 for 2,3)][h0, h1]   (head dims permuted: a thread holds 8 consecutive dims)
  // K as the A operand: k-slots (2 tid4, 2 tid4 + 1, 2 tid4 + 8, 2 tid4 + 9) of k-step kk are head dims 16 tid4 + 4 kk + (0..3),
  // i.e. bytes 4 kk .. 4 kk + 3 of ONE 16-byte load per kv row; the query fragments are gathered with the same permutation.
  uint32_t qb[4][2];
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
  float o[4][4];
  bool seg_first = false;

  for (int i = 0; i < n; ++i) {
    const int meta = list[i].y;
    const int r = pka_row(meta), h = pka_head(meta), cnt = pka_cnt(meta);
    if (meta & PKA_SEG_START) {
      const bf16* qrow = p.q + (long long)r * p.hq * D + (long long)h * G * D + gid * D + 16 * tid4;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        qb[kk][0] = gid < G ? bf16x2_to_f16x2(__ldcg(reinterpret_cast<const uint32_t*>(qrow + 4 * kk))) : 0u;
        qb[kk][1] = gid < G ? bf16x2_to_f16x2(__ldcg(reinterpret_cast<const uint32_t*>(qrow + 4 * kk + 2))) : 0u;
      }
      m0 = m1 = -INFINITY;
      l0 = l1 = 0.0f;
#pragma unroll
      for (int db = 0; db < 4; ++db) o[db][0] = o[db][1] = o[db][2] = o[db][3] = 0.0f;
      seg_first = (meta & PKA_PAIR_FIRST) != 0;
    }
    const bool more = i + 1 < n;
    const uint32_t kb = smem_u32(stage0 + (i & 1) * 8192), vb = kb + kTileBytes;
    // this tile's scales to shared memory (per-row reads below), the next tile's requested
    __syncwarp();
    s_ks[lane] = sc_cur.k0 * kInv;
    s_ks[lane + 32] = sc_cur.k1 * kInv;
    s_vs[lane] = sc_cur.v0 * kInv;
    s_vs[lane + 32] = sc_cur.v1 * kInv;
    __syncwarp();
    sc_cur = sc_next;
    if (i + 2 < n) sc_next = request_scales(i + 2);

    mbar_wait(&bars[i & 1], (phase >> (i & 1)) & 1u);
    phase ^= 1u << (i & 1);
    uint32_t pb[4][2];
    kv8t_scores<F8>(kb, s_ks, s_vs, cnt, p.softcap, qb, m0, m1, l0, l1, o, pb, gid, tid4);
    kv8t_pv<F8>(vb, cnt, pb, o, gid, tid4);
    // the stage is consumed: refill it with tile i + 2
    fence_proxy_async();
    __syncwarp();
    if (i + 2 < n && lane == 0) issue(i + 2);

    if (meta & PKA_SEG_END) {
#pragma unroll
      for (int sh = 4; sh < 32; sh <<= 1) {
        l0 += __shfl_xor_sync(0xffffffffu, l0, sh);
        l1 += __shfl_xor_sync(0xffffffffu, l1, sh);
      }
      const int h0 = tid4 * 2;
      // head h0: dims 8 gid .. 8 gid + 7 = o[0][0], o[0][2], o[1][0], o[1][2], ...; head h0 + 1: o[.][1], o[.][3]
      if (seg_first && (meta & PKA_PAIR_LAST)) {
        bf16* dst = p.attn + (long long)r * p.hq * D + (long long)h * G * D + 8 * gid;
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        if (h0 < G)
          *reinterpret_cast<uint4*>(dst + h0 * D) = make_uint4(pack_bf16x2(o[0][0] * i0, o[0][2] * i0), pack_bf16x2(o[1][0] * i0, o[1][2] * i0),
                                                               pack_bf16x2(o[2][0] * i0, o[2][2] * i0), pack_bf16x2(o[3][0] * i0, o[3][2] * i0));
        if (h0 + 1 < G)
          *reinterpret_cast<uint4*>(dst + (h0 + 1) * D) = make_uint4(pack_bf16x2(o[0][1] * i1, o[0][3] * i1), pack_bf16x2(o[1][1] * i1, o[1][3] * i1),
                                                                     pack_bf16x2(o[2][1] * i1, o[2][3] * i1), pack_bf16x2(o[3][1] * i1, o[3][3] * i1));
      } else {
        const bool final_seg = !more, to_l2 = (meta & PKA_TAIL) != 0;
        const int stride = (G * D + 2 * G + 3) & ~3;
        float* dst = to_l2 ? p.attn_part_o + (long long)tail->a_tail[aw] * stride : final_seg ? reinterpret_cast<float*>(stage0) : tail->a_slot[aw];
        if (h0 < G) {
          float* row = dst + h0 * D + 8 * gid;
          *reinterpret_cast<float4*>(row) = make_float4(o[0][0], o[0][2], o[1][0], o[1][2]);
          *reinterpret_cast<float4*>(row + 4) = make_float4(o[2][0], o[2][2], o[3][0], o[3][2]);
          if (gid == 0) *reinterpret_cast<float2*>(dst + G * D + h0 * 2) = make_float2(m0, l0);
        }
        if (h0 + 1 < G) {
          float* row = dst + (h0 + 1) * D + 8 * gid;
          *reinterpret_cast<float4*>(row) = make_float4(o[0][1], o[0][3], o[1][1], o[1][3]);
          *reinterpret_cast<float4*>(row + 4) = make_float4(o[2][1], o[2][3], o[3][1], o[3][3]);
          if (gid == 0) *reinterpret_cast<float2*>(dst + G * D + (h0 + 1) * 2) = make_float2(m1, l1);
        }
      }
    }
  }
  pk_attention_merge(p, tail, attn_tiles, aw, lane, ev);
}

// ---- the kernel --------------------------------------------------------------------------

// KVQ: 0 = bf16 KV cache, 1 = int8, 2 = float8_e4m3fn (kv_quant_axis dkv) -- separate instantiations, so the bf16 kernel's
// code and register allocation do not change with the quantised attention path.
template <int KVQ>
__global__ void __launch_bounds__(kPkThreads, 1)
step_persistent_kernel(const __grid_constant__ CUtensorMap tm_wqkv, const __grid_constant__ CUtensorMap tm_wo,
                       const __grid_constant__ CUtensorMap tm_w01, const __grid_constant__ CUtensorMap tm_wout,
                       const __grid_constant__ CUtensorMap tm_wlogits, const __grid_constant__ CUtensorMap tm_x,
                       const __grid_constant__ CUtensorMap tm_attn, const __grid_constant__ CUtensorMap tm_h,
                       const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_n,
                       const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ PkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space
  uint8_t* ring = smem;
  uint8_t* attn_tiles = smem + kPkRingBytes;
  float* park = reinterpret_cast<float*>(attn_tiles);
  PkTail* tail = reinterpret_cast<PkTail*>(smem + kPkRingBytes + kPkAttnBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ctas = int(gridDim.x), cta = int(blockIdx.x);
  const int x_bytes = p.r_tile * kBlockK * 2;
  const uint32_t tmem_cols = uint32_t(kPkAccBufs * p.r_tile < 32 ? 32 : kPkAccBufs * p.r_tile);
  const int kb_e = p.E / kBlockK;
  const int logits_tiles_all = (p.V + kTileN - 1) / kTileN;
  const int logits_tiles = cta < logits_tiles_all ? (logits_tiles_all - cta + n_ctas - 1) / n_ctas : 0;
  const int ss_tiles = (p.E + 127) / 128;
  const int tl = timeline_begin(20);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPkStages; ++s) {
      mbar_init(&tail->full_w[s], 1);
      mbar_init(&tail->full_x[s], 1);
      mbar_init(&tail->xready[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int b = 0; b < kPkAccBufs; ++b) {
      mbar_init(&tail->tmem_full[b], 1);
      mbar_init(&tail->tmem_empty[b], 1);
    }
    for (int i = 0; i < 2 * kPkAttnWarps; ++i) mbar_init(&tail->attn_bars[i], 1);
    tail->bar_done = 0;
    fence_barrier_init();
    tma_prefetch_desc(&tm_wqkv);
    tma_prefetch_desc(&tm_wo);
    tma_prefetch_desc(&tm_w01);
    tma_prefetch_desc(&tm_wout);
    tma_prefetch_desc(&tm_wlogits);
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_attn);
    tma_prefetch_desc(&tm_h);
    tma_prefetch_desc(&tm_act);
    tma_prefetch_desc(&tm_n);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  {  // this CTA's work table
    const int* src = reinterpret_cast<const int*>(p.tables + cta);
    int* dst = reinterpret_cast<int*>(&tail->tab);
    for (int i = threadIdx.x; i < int(sizeof(PkTable) / 4); i += kPkThreads) dst[i] = src[i];
  }
  if (warp == 1) {
    tmem_alloc(&tail->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tail->tmem_base;
  const PkTable& tab = tail->tab;

  // Calls f(layer, phase, weight tile, first k-block, last k-block + 1) for every work unit of this CTA in issue
  // order.  All GEMM roles walk the same sequence; logits tiles come as units of phase PK_LOGITS.
  auto for_each_unit = [&](auto&& f) {
    for (int l = 0; l < p.L; ++l)
      for (int ph = 0; ph < 4; ++ph) {
        const int nu = tab.n_units[ph];
        for (int u = 0; u < nu; ++u) {
          const PkUnit un = tab.u[ph][u];
          f(l, ph, un.tile, un.kb0, un.kb1);
        }
      }
    for (int j = 0; j < logits_tiles; ++j) f(p.L, int(PK_LOGITS), cta + j * n_ctas, 0, kb_e);
  };

  if (warp == 0) {
    // =================================== weight producer =================================
    // Runs ahead of the grid barriers, bounded only by the ring: a stage is refilled as soon as the MMAs that
    // read it have completed.
    if (lane == 0) {
      PkCursor c;
      pk_cursor_init(c);
      for_each_unit([&](int l, int ph, int tile, int kb0, int kb1) {
        const CUtensorMap* tmw = ph == PK_QKV ? &tm_wqkv : ph == PK_OPROJ ? &tm_wo : ph == PK_UP ? &tm_w01 : ph == PK_DOWN ? &tm_wout : &tm_wlogits;
        const int n_ph = ph == PK_QKV ? p.qkv_n : ph == PK_UP ? 2 * p.M : ph == PK_LOGITS ? 0 : p.E;
        const int row = l * n_ph + tile * kTileN;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (c.uses) mbar_wait(&tail->empty[c.s], c.par ^ 1);
          mbar_expect_tx(&tail->full_w[c.s], uint32_t(kWTileBytes));
          tma_load_2d(ring + size_t(c.s) * kPkStageBytes, tmw, kb * kBlockK, row, &tail->full_w[c.s], kEvictFirst);
          pk_cursor_next(c);
        }
      });
    }
    __syncwarp();
  } else if (warp == 1) {
    // =================================== MMA issuer =====================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kTileN, p.r_tile);
      PkCursor c;
      pk_cursor_init(c);
      uint32_t uc = 0;
      for_each_unit([&](int, int, int, int kb0, int kb1) {
        const uint32_t buf = uc % kPkAccBufs;
        const uint32_t acc = tmem_base + buf * uint32_t(p.r_tile);
        if (uc >= uint32_t(kPkAccBufs)) mbar_wait(&tail->tmem_empty[buf], ((uc / kPkAccBufs) & 1) ^ 1);
        tcgen05_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&tail->full_w[c.s], c.par);
          mbar_wait(&tail->xready[c.s], c.par);
          tcgen05_fence_after();
          const uint64_t da = umma_desc_sw128(ring + size_t(c.s) * kPkStageBytes);
          const uint64_t db = umma_desc_sw128(ring + size_t(c.s) * kPkStageBytes + kWTileBytes);
#pragma unroll
          for (int kk = 0; kk < kBlockK / kUmmaK; ++kk)
            umma_bf16(acc, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc, uint32_t((kb > kb0) || kk > 0));
          umma_commit(&tail->empty[c.s]);
          pk_cursor_next(c);
        }
        umma_commit(&tail->tmem_full[buf]);
        ++uc;
      });
    }
    __syncwarp();
  } else if (warp < 6) {
    // =================================== epilogue warps ==================================
    const int wtid = threadIdx.x - 64;
    const int ew = warp - 2;  // 0..3: row-major work is dealt by this index
    const int quarter = warp & 3;
    const int n_local = quarter * 32 + lane;
    const uint32_t tlane = uint32_t(quarter * 32) << 16;
    uint32_t uc = 0, nbar = 0;
    PkEv ev = pk_ev_make(p, 0);
    bool up_shared = false;  // this CTA shares MLP-up tiles with other CTAs: all its warps help to finish them
    up_shared = pk_up_all_warps(tab);
    const bool up_solo = tab.n_units[PK_UP] == 1 && tab.u[PK_UP][0].S == 1;

    griddep_wait();
    if (p.trace && wtid == 0) p.trace[gridDim.x + blockIdx.x] = (long long)globaltimer_ns();  // "barrier 0 release" = start

    // Closes a phase: generic writes that other CTAs read through TMA get their proxy fence, the CTA's epilogue
    // threads meet, one thread runs the grid barrier (its gpu-scope fence is cumulative over the CTA's writes).
    auto phase_end = [&](bool wait_all) {
      if (wtid == 0) pk_ev(ev, 90);
      fence_proxy_async_all();
      named_bar_sync(1, 128);
      if (wtid == 0) pk_ev(ev, 91);
      ++nbar;
      if (wtid == 0) pk_grid_barrier(p, tail, nbar);
      if (wait_all) named_bar_sync(1, 128);
    };

    // ---- embedding gather (embeddings.py:154) + row sums of squares of x ----
    {
      const int nvec = p.E / 8;
      for (int r = cta; r < p.rows; r += n_ctas) {
        const bf16* src = p.embedding + (long long)p.token[r] * p.E;
        float ss = 0.0f;
        for (int i = wtid; i < nvec; i += 128) {
          const uint4 raw = *reinterpret_cast<const uint4*>(src + i * 8);
          *reinterpret_cast<uint4*>(p.x + (long long)r * p.E + i * 8) = raw;
          const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = bf16_lo(w[j]), b = bf16_hi(w[j]);
            ss += a * a + b * b;
          }
        }
        ss = warp_sum(ss);
        named_bar_sync(1, 128);
        if (lane == 0) tail->red[ew] = ss;
        named_bar_sync(1, 128);
        if (wtid < ss_tiles) p.ss_x[wtid * kPkMaxRTile + r] = wtid == 0 ? tail->red[0] + tail->red[1] + tail->red[2] + tail->red[3] : 0.0f;
      }
      phase_end(false);
    }

    // One GEMM phase of this CTA: dump every accumulator (partial tile to the exchange workspace, or straight to
    // the epilogue when the CTA owns the whole reduction), then finish the row slices it owns.
    // `ph_tag` selects the epilogue kind at compile time (QKV / SwiGLU / residual); out-proj and MLP down share the
    // residual body and differ only in `tph` (work table, buffers): per layer the instruction footprint of the CTA
    // has to stay inside the instruction cache, every copy of this body costs ~20 KB of it.
    auto gemm_phase = [&](auto ph_tag, int tph, int layer) {
      constexpr int ph = decltype(ph_tag)::value;
      const int nu = tab.n_units[tph];
      const int N = ph == PK_QKV ? p.qkv_n : ph == PK_UP ? 2 * p.M : p.E;
      bf16* k_layer = p.k_cache + p.kv_layer_elems * layer;
      bf16* v_layer = p.v_cache + p.kv_layer_elems * layer;
      if ((ph == PK_QKV || ph == PK_UP) && nu > 0) {
        // rstd of every row from the per-tile sums of squares the previous epilogue left: two threads per row,
        // eight independent loads each; off the critical path (the MMAs of the phase run meanwhile)
        pk_wait_flag(&tail->bar_done, pk_need(layer, ph, p.L));
        const float* ss = ph == PK_QKV ? p.ss_x : p.ss_h;
        const int row = wtid >> 1, half = wtid & 1;
        float tot = 0.0f;
        if (row < p.rows) {
          for (int t0 = half * 8; t0 < ss_tiles; t0 += 16) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = t0 + i < ss_tiles ? __ldcg(ss + (t0 + i) * kPkMaxRTile + row) : 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) tot += v[i];
          }
        }
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        if (half == 0 && row < p.r_tile) tail->rstd[row] = row < p.rows ? 1.0f / sqrtf(tot / float(p.E) + p.eps) : 0.0f;
        named_bar_sync(1, 128);
      }
      const bf16* resid = tph == PK_OPROJ ? p.x : p.h;
      bf16* res_out = tph == PK_OPROJ ? p.h : p.x;
      float* ss_out = tph == PK_OPROJ ? p.ss_h : p.ss_x;
      // Side inputs of a row fragment do not depend on the accumulator: they are requested before the fragment's
      // partial tiles are polled, so the epilogue itself never waits on memory.
      struct Side {
        PkQkvSide q;
        uint2 rv;
      };
      auto side = [&](int r, int tile) {
        Side sd;
        sd.q.c01 = sd.q.c23 = make_float4(0.f, 0.f, 0.f, 0.f);
        sd.q.plane = sd.q.wr = 0;
        sd.rv = make_uint2(0u, 0u);
        const int n0 = tile * 128 + lane * 4;
        if (ph == PK_QKV) sd.q = pk_qkv_side(p, r, tile, lane);
        else if (ph != PK_UP && n0 < N) sd.rv = __ldcg(reinterpret_cast<const uint2*>(resid + (long long)r * N + n0));
        return sd;
      };
      auto epilogue = [&](float4 a, const Side& sd, int r, int tile) {
        if (ph == PK_QKV || ph == PK_UP) {
          const float rs = tail->rstd[r];
          a.x *= rs; a.y *= rs; a.z *= rs; a.w *= rs;
        }
        if (ph == PK_QKV) pk_epi_qkv<KVQ>(a, sd.q, r, tile, lane, p, k_layer, v_layer, layer);
        else if (ph == PK_UP) pk_epi_swiglu(a, r, tile, lane, N, p.M, p.act);
        else pk_epi_residual(a, sd.rv, r, tile, lane, N, res_out, ss_out, ss_tiles);
      };
      for (int u = 0; u < nu; ++u) {
        const PkUnit un = tab.u[tph][u];
        const uint32_t buf = uc % kPkAccBufs;
        mbar_wait(&tail->tmem_full[buf], (uc / kPkAccBufs) & 1);
        tcgen05_fence_after();
        if (wtid == 0) pk_ev(ev, 100 * tph + 12);
        const uint32_t taddr = tmem_base + tlane + buf * uint32_t(p.r_tile);
        // exchange slot in global memory, or the park buffer in shared memory for a tile this CTA owns alone
        if (un.S == 1 || un.S < 0) {
          for (int c = 0; c * 16 < p.rows; ++c) {
            float v[16];
            tmem_ld_x16(taddr + uint32_t(c * 16), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) park[(c * 16 + j) * 128 + n_local] = v[j];
          }
        } else {
          float* dst = p.part_ws + (long long)(cta * 4 + (un.tile & 3)) * kPkSlotFloats;
          for (int c = 0; c * 16 < p.rows; ++c) {
            float v[16];
            tmem_ld_x16(taddr + uint32_t(c * 16), v);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c * 16 + j < p.rows) dst[(c * 16 + j) * 128 + n_local] = v[j];
          }
        }
        tcgen05_fence_before();
        if (wtid == 0) pk_ev(ev, 100 * tph + 13);
        // Polling loads issued back to back can starve the SM's own pending stores (then every CTA waits for
        // everybody): the poll loop below pauses between attempts so that the store queue always drains.
        named_bar_sync(1, 128);
        if (wtid == 0) mbar_arrive(&tail->tmem_empty[buf]);
        if (un.S == 1 && !(ph == PK_UP && up_solo)) {
          // four rows at a time: one row's epilogue is a single dependent chain (shuffles, exponentials, roundings)
          for (int r0 = ew; r0 < p.rows; r0 += 16) {
            float4 a4[4];
            Side sd4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = r0 + 4 * j < p.rows ? r0 + 4 * j : r0;
              a4[j] = *reinterpret_cast<const float4*>(park + r * 128 + lane * 4);
              sd4[j] = side(r, un.tile);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (r0 + 4 * j < p.rows) epilogue(a4[j], sd4[j], r0 + 4 * j, un.tile);
          }
          named_bar_sync(1, 128);  // the park buffer is rewritten by the next unit
        }
        ++uc;
      }
      // Finish the row slices this CTA owns.  The exchange has no flags: a word of the workspace is either the
      // sentinel or data, so the reader simply re-requests a fragment until all of its words have arrived, then
      // puts the sentinel back for the next layer (the same thread reads the same words every layer).
      // MLP up: the attention warps help to finish; they need rstd and this CTA's parked accumulator
      if (ph == PK_UP && up_shared) asm volatile("bar.arrive 2, %0;" ::"r"(128 + kPkAttnWarps * 32) : "memory");
      for (int u = 0; u < nu; ++u) {
        const PkUnit un = tab.u[tph][u];
        if ((un.S == 1 && !(ph == PK_UP && up_solo)) || un.S == 0) continue;
        if (ph == PK_UP) {  // every warp of the CTA takes rows (the attention warps call the same routine)
          if (wtid == 0) pk_ev(ev, 100 * tph + 14);
          pk_finish_up_unit(&p, tail->rstd, park, un, up_solo, cta, ew, 4 + kPkAttnWarps, lane);
          if (wtid == 0) pk_ev(ev, 100 * tph + 15);
#ifdef MTX_PK_ICACHE_PROBE  // the same (idempotent) work again: how much of the first pass was instruction fetch?
          if (up_solo) pk_finish_up_unit(&p, tail->rstd, park, un, up_solo, cta, ew, 4 + kPkAttnWarps, lane);
          if (wtid == 0) pk_ev(ev, 100 * tph + 16);
#endif
          continue;
        }
        const int si = cta - un.c_first;
        const int r_begin = si * p.rows / un.S, r_end = (si + 1) * p.rows / un.S;
        float* src0 = p.part_ws + (long long)(un.c_first * 4 + (un.tile & 3)) * kPkSlotFloats + lane * 4;
        const float sent = __uint_as_float(kPkSentinel);
        // two rows x eight parts = 16 fragments in flight per warp: the loop is bound by L2 round trips
        for (int r = r_begin + ew; r < r_end; r += 8) {
          const int r2 = r + 4;
          const bool two = r2 < r_end;
          const Side sd_a = side(r, un.tile), sd_b = side(two ? r2 : r, un.tile);
          float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int s0 = 0; s0 < un.S; s0 += 8) {
            float4 ta[8], tb[8], sa, sb;
            const long long t_spin = clock64();
            for (;;) {
#pragma unroll
              for (int ss = 0; ss < 8; ++ss) {
                ta[ss] = tb[ss] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s0 + ss < un.S) {
                  const float* ps = src0 + (long long)(s0 + ss) * 4 * kPkSlotFloats;
                  ta[ss] = ld_poll_f4(ps + r * 128);
                  if (two) tb[ss] = ld_poll_f4(ps + r2 * 128);
                }
              }
              // The sentinel is a quiet NaN, so the sums the epilogue needs anyway tell whether every word has
              // arrived: one test instead of 64 comparisons.  Genuine NaN / Inf - Inf data takes the exact test.
              sa = ta[0];
              sb = tb[0];
#pragma unroll
              for (int ss = 1; ss < 8; ++ss) {
                sa.x += ta[ss].x; sa.y += ta[ss].y; sa.z += ta[ss].z; sa.w += ta[ss].w;
                sb.x += tb[ss].x; sb.y += tb[ss].y; sb.z += tb[ss].z; sb.w += tb[ss].w;
              }
              const float chk = ((sa.x + sa.y) + (sa.z + sa.w)) + ((sb.x + sb.y) + (sb.z + sb.w));
              bool ok = chk == chk;
              if (!ok) {
                ok = true;
#pragma unroll
                for (int ss = 0; ss < 8; ++ss)
                  ok = ok && __float_as_uint(ta[ss].x) != kPkSentinel && __float_as_uint(ta[ss].y) != kPkSentinel &&
                       __float_as_uint(ta[ss].z) != kPkSentinel && __float_as_uint(ta[ss].w) != kPkSentinel &&
                       __float_as_uint(tb[ss].x) != kPkSentinel && __float_as_uint(tb[ss].y) != kPkSentinel &&
                       __float_as_uint(tb[ss].z) != kPkSentinel && __float_as_uint(tb[ss].w) != kPkSentinel;
              }
              if (__all_sync(0xffffffffu, ok)) break;
              pk_backoff();  // the missing fragments are still in flight somewhere: let stores (ours too) drain
              if (clock64() - t_spin > 4000000000LL) mtx_wait_timeout(3, tph, un.tile);
            }
#pragma unroll
            for (int ss = 0; ss < 8; ++ss) {
              if (s0 + ss < un.S) {
                float* ps = src0 + (long long)(s0 + ss) * 4 * kPkSlotFloats;
                *reinterpret_cast<float4*>(ps + r * 128) = make_float4(sent, sent, sent, sent);
                if (two) *reinterpret_cast<float4*>(ps + r2 * 128) = make_float4(sent, sent, sent, sent);
              }
            }
            acc_a.x += sa.x; acc_a.y += sa.y; acc_a.z += sa.z; acc_a.w += sa.w;
            acc_b.x += sb.x; acc_b.y += sb.y; acc_b.z += sb.z; acc_b.w += sb.w;
          }
          epilogue(acc_a, sd_a, r, un.tile);
          if (two) epilogue(acc_b, sd_b, r2, un.tile);
        }
      }
    };

    for (int l = 0; l < p.L; ++l) {
      ev.on = l == 1;
#pragma unroll 1  // one copy of the phase body
      for (int ph = 0; ph < 4; ++ph) {
        if (ph == PK_QKV) gemm_phase(std::integral_constant<int, PK_QKV>{}, ph, l);
        else if (ph == PK_UP) gemm_phase(std::integral_constant<int, PK_UP>{}, ph, l);
        else gemm_phase(std::integral_constant<int, PK_DOWN>{}, ph, l);
        if (ph == PK_UP && up_shared) {  // the attention warps' share of the MLP-up rows is stored too
          asm volatile("bar.sync 4, %0;" ::"r"(128 + kPkAttnWarps * 32) : "memory");
        }
        phase_end(l == p.L - 1 && ph == PK_DOWN);
        if (ph == PK_QKV) ++nbar;  // the attention phase's barrier is run by the attention warps
      }
    }

    // ---- final RMSNorm (decoders.py:537-589) into the activation buffer the logits GEMM streams ----
    {
      const int nvec = p.E / 8;
      for (int r = cta; r < p.rows; r += n_ctas) {
        float tot = 0.0f;
        for (int t = 0; t < ss_tiles; ++t) tot += __ldcg(p.ss_x + t * kPkMaxRTile + r);
        const float rstd = 1.0f / sqrtf(tot / float(p.E) + p.eps);
        for (int i = wtid; i < nvec; i += 128) {
          const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(p.x + (long long)r * p.E + i * 8));
          const uint4 sc = *reinterpret_cast<const uint4*>(p.final_norm + i * 8);
          const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
          const uint32_t s[4] = {sc.x, sc.y, sc.z, sc.w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float y0, y1;
            bf16r2(bf16_lo(w[j]) * rstd, bf16_hi(w[j]) * rstd, y0, y1);
            o[j] = pack_bf16x2(y0 * bf16_lo(s[j]), y1 * bf16_hi(s[j]));
          }
          *reinterpret_cast<uint4*>(p.n + (long long)r * p.E + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      phase_end(false);
    }

    // ---- logits + per-tile sampling partials ----
    {
      uint32_t step = 0;
      uint64_t seed = 0;
      if (p.logits.gumbel) {
        step = p.logits.rng_state[0];
        seed = (uint64_t(p.logits.rng_state[2]) << 32) | p.logits.rng_state[1];
      }
      for (int j = 0; j < logits_tiles; ++j) {
        const int tile = cta + j * n_ctas;
        const uint32_t buf = uc % kPkAccBufs;
        ev.on = false;
        if (wtid == 0) pk_ev(ev, 410);
        mbar_wait(&tail->tmem_full[buf], (uc / kPkAccBufs) & 1);
        tcgen05_fence_after();
        if (wtid == 0) pk_ev(ev, 411);
        const uint32_t taddr = tmem_base + tlane + buf * uint32_t(p.r_tile);
        for (int c = 0; c * 16 < p.rows; ++c) {
          float v[16];
          tmem_ld_x16(taddr + uint32_t(c * 16), v);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) park[(c * 16 + jj) * kPkParkPitch + n_local] = v[jj];
        }
        tcgen05_fence_before();
        named_bar_sync(1, 128);
        if (wtid == 0) mbar_arrive(&tail->tmem_empty[buf]);
        if (wtid == 0) pk_ev(ev, 412);
        if (p.logits.logits_out != nullptr) pk_logits_store(park, ew, lane, p.rows, tile, p.logits, p.V);
        pk_logits_scan(park, wtid, p.rows, tile, p.logits, p.V, step, seed);
        named_bar_sync(1, 128);
        if (wtid == 0) pk_ev(ev, 413);
        ++uc;
      }
    }
    if (p.trace && wtid == 0) p.trace[blockIdx.x] = (long long)globaltimer_ns();  // slot of "barrier 0 arrive" = end of work
  } else {
    // ============== attention warps; between attention phases: activation producer and normaliser ==============
    const int aw = warp - 6;             // 0..5
    const int atid = threadIdx.x - 192;  // 0..159
    const bool xformer = aw < 4;         // warps 6-9 normalise activation tiles during the QKV and MLP-up phases
    const bool xproducer = aw == 4;      // warp 10 requests the activation tiles of every GEMM phase
    uint32_t kv_phase = 0;
    PkEv ev = pk_ev_make(p, aw == kPkAttnWarps - 1 ? 2 : 1);  // debug: first and last attention warp
    if (aw != 0 && aw != kPkAttnWarps - 1) ev.base = nullptr;
    PkEv evx = pk_ev_make(p, 2);
    evx.base = nullptr;

    // ---- once per step: stage the row descriptors, cut this warp's share of the attention tiles ----
    griddep_wait();
    for (int r = atid; r < p.rows; r += kPkAttnWarps * 32) {
      tail->r_len0[r] = p.len0[r];
      tail->r_rf[r] = p.ring_first[r];
      tail->r_rl[r] = p.ring_len[r];
      tail->r_plane[r] = p.plane[r];
      tail->r_prefix[r] = p.tile_prefix[r];
    }
    if (atid == 0) tail->r_prefix[p.rows] = p.tile_prefix[p.rows];
    named_bar_sync(3, kPkAttnWarps * 32);
    pk_attn_build_list(p, tail, cta, aw, lane);
    __syncwarp();
    // The first K/V tiles of a layer may be requested before the QKV phase has finished unless the tile holds
    // the row that phase appends, or the QKV epilogue of this CTA parks a whole tile in the same shared memory
    // (single-CTA tiles; the other phases' parked tiles are long consumed when the next layer's QKV phase starts).
    bool may_prime = tail->a_count[aw] > 0 && (tail->a_list[aw][0].y & PKA_PAIR_LAST) == 0;
    for (int u = 0; u < tab.n_units[PK_QKV]; ++u)
      if (tab.u[PK_QKV][u].S == 1) may_prime = false;

    bool up_shared = false;  // see the epilogue warps
    up_shared = pk_up_all_warps(tab);
    PkCursor cur;  // shared fill sequence (activation producer and normaliser walk it in step with the weights)
    pk_cursor_init(cur);

    // Activation tiles of one GEMM phase: requested as soon as the phase's grid barrier has completed and the
    // ring stage is free.  Normalised phases deliver to full_x (the normaliser signals xready), the others
    // deliver to xready directly.
    auto x_phase = [&](int ph, int layer) {
      const int nu = ph == PK_LOGITS ? logits_tiles : tab.n_units[ph];
      if (nu == 0) return;
      const bool xform = (ph == PK_QKV || ph == PK_UP) && !p.fold;
      const CUtensorMap* tmx = ph == PK_QKV ? &tm_x : ph == PK_OPROJ ? &tm_attn : ph == PK_UP ? &tm_h : ph == PK_DOWN ? &tm_act : &tm_n;
      pk_wait_flag(&tail->bar_done, pk_need(layer, ph, p.L));
      fence_proxy_async_all();  // activations written with generic stores by other CTAs, read through TMA
      evx.on = layer == 1;
      for (int u = 0; u < nu; ++u) {
        const int kb0 = ph == PK_LOGITS ? 0 : tab.u[ph][u].kb0, kb1 = ph == PK_LOGITS ? kb_e : tab.u[ph][u].kb1;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (cur.uses) mbar_wait(&tail->empty[cur.s], cur.par ^ 1);
          pk_ev(evx, 100 * ph + 1);
          uint64_t* bar = xform ? &tail->full_x[cur.s] : &tail->xready[cur.s];
          if (!xform) mbar_arrive(&tail->full_x[cur.s]);  // keep every barrier of the stage in step
          mbar_expect_tx(bar, uint32_t(x_bytes));
          tma_load_2d(ring + size_t(cur.s) * kPkStageBytes + kWTileBytes, tmx, kb * kBlockK, 0, bar, kEvictLast);
          pk_cursor_next(cur);
        }
      }
    };

    // normalizations.py:57-69, reassociated: the per-feature scale is applied in place to the activation k-blocks TMA
    // just delivered (n' = bf16(x * scale)); the per-row factor rstd multiplies the fp32 accumulator in the epilogue
    // (dot(W, x * scale) * rstd).  The reference rounds x * rstd to bf16 before the scale; here that rounding is
    // skipped (one rounding less, |rel. difference| <= 2^-9 per term before the dot product averages it out): the
    // per-row factor then no longer sits between a tile's arrival and its MMA, nor does the phase start wait for
    // the row statistics.
    auto transform_phase = [&](int ph, int layer) {
      const int nu = tab.n_units[ph];
      if (nu == 0) return;
      if (atid == 0) pk_ev(ev, 100 * ph + 20);
      // lane -> 16-byte chunks (row, physical chunk pc): pc = lane & 7, rows (lane >> 3) + 4 j; the logical chunk
      // pc ^ (row & 7) takes two values (j even / odd).
      const int pc = lane & 7, rq = lane >> 3;
      const int lc_a = pc ^ (rq & 7), lc_b = pc ^ ((rq + 4) & 7);
      const bf16* scale = (ph == PK_QKV ? p.attn_norm : p.mlp_norm) + (long long)layer * p.E;
      // k-block of this warp's fill number f of the phase (-1 past the end); the norm scales of the next owned
      // k-block are requested one fill ahead, the first ones before the phase's barrier completes
      auto kb_of_fill = [&](int f) {
        for (int u = 0; u < nu; ++u) {
          const int n = tab.u[ph][u].kb1 - tab.u[ph][u].kb0;
          if (f < n) return tab.u[ph][u].kb0 + f;
          f -= n;
        }
        return -1;
      };
      uint4 sc_a = make_uint4(0, 0, 0, 0), sc_b = sc_a;
      {
        const int kbn = kb_of_fill(aw);
        if (kbn >= 0) {
          sc_a = *reinterpret_cast<const uint4*>(scale + kbn * kBlockK + lc_a * 8);
          sc_b = *reinterpret_cast<const uint4*>(scale + kbn * kBlockK + lc_b * 8);
        }
      }
      pk_wait_flag(&tail->bar_done, pk_need(layer, ph, p.L));
      if (atid == 0) pk_ev(ev, 100 * ph + 21);
      // Each warp normalises whole k-blocks on its own (fill i of the phase belongs to warp i % 4), so the k-blocks
      // of a phase are processed in parallel and no CTA-level barrier sits between a tile's arrival and its MMA.
      int fi = 0;
      for (int u = 0; u < nu; ++u) {
        const PkUnit un = tab.u[ph][u];
        for (int kb = un.kb0; kb < un.kb1; ++kb, ++fi) {
          if ((fi & 3) == aw) {
            if (atid == 0) pk_ev(ev, 100 * ph + 22);
            mbar_wait(&tail->full_x[cur.s], cur.par);
            if (atid == 0) pk_ev(ev, 100 * ph + 23);
            uint8_t* xt = ring + size_t(cur.s) * kPkStageBytes + kWTileBytes;
            // four chunks at a time (loads, math, stores): the in-place update would otherwise serialise one
            // dependent chain per chunk.  Rows past r_tile lie in the stage's unused activation space: harmless.
            uint8_t* lane_base = xt + rq * 128 + pc * 16;
#pragma unroll 1  // keep the body small: it is entered once per k-block and must stay resident in the instruction cache
            for (int j0 = 0; j0 < 16; j0 += 4) {
              uint4 raw[4];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) raw[jj] = *reinterpret_cast<const uint4*>(lane_base + (j0 + jj) * 512);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const uint4 sc = (jj & 1) ? sc_b : sc_a;
                // one packed bf16 multiply per feature pair: bf16(x * scale), rounded once
                const __nv_bfloat162 p0 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&raw[jj].x), *reinterpret_cast<const __nv_bfloat162*>(&sc.x));
                const __nv_bfloat162 p1 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&raw[jj].y), *reinterpret_cast<const __nv_bfloat162*>(&sc.y));
                const __nv_bfloat162 p2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&raw[jj].z), *reinterpret_cast<const __nv_bfloat162*>(&sc.z));
                const __nv_bfloat162 p3 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&raw[jj].w), *reinterpret_cast<const __nv_bfloat162*>(&sc.w));
                raw[jj] = make_uint4(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1),
                                     *reinterpret_cast<const uint32_t*>(&p2), *reinterpret_cast<const uint32_t*>(&p3));
              }
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) *reinterpret_cast<uint4*>(lane_base + (j0 + jj) * 512) = raw[jj];
            }
            {
              const int kbn = kb_of_fill(fi + 4);
              if (kbn >= 0) {
                sc_a = *reinterpret_cast<const uint4*>(scale + kbn * kBlockK + lc_a * 8);
                sc_b = *reinterpret_cast<const uint4*>(scale + kbn * kBlockK + lc_b * 8);
              }
            }
            if (atid == 0) pk_ev(ev, 100 * ph + 28);
            fence_proxy_async();  // the MMA reads the tile through the async proxy
            if (atid == 0) pk_ev(ev, 100 * ph + 29);
            __syncwarp();
            if (lane == 0) mbar_arrive(&tail->xready[cur.s]);
            if (atid == 0) pk_ev(ev, 100 * ph + 24);
          }
          pk_cursor_next(cur);
        }
      }
    };

    // Between attention phases: warps 6-9 normalise (QKV, MLP up), lane 0 of warp 10 requests activation tiles,
    // everybody else just keeps the ring cursor in step.
    auto gemm_duty = [&](int ph, int layer) {
      const bool xform = (ph == PK_QKV || ph == PK_UP) && !p.fold;
      if (xformer && xform) {
        transform_phase(ph, layer);
      } else if (xproducer && lane == 0) {
        x_phase(ph, layer);
      } else {
        pk_cursor_skip(cur, tab.kbs[ph]);
      }
    };

    ev.brief = aw == kPkAttnWarps - 1;
    for (int l = 0; l < p.L; ++l) {
      ev.on = ev.brief ? (l & 1) == 0 && l < 32 : l == 1;
#pragma unroll 1  // one copy of the duty code
      for (int ph = 0; ph < 4; ++ph) {
        if (ph == PK_QKV && ((p.variant & 2) || ((p.variant & 1) && l == 0)) && p.page_map == nullptr) {
          // this layer's K/V tiles towards L2 while the QKV phase runs (the row appended by that phase is written later: L2 stays coherent)
          const int lrow = l * p.kv_layer_rows;
          const int n_own = tail->a_count[aw];
          for (int i = lane; i < n_own; i += 32) {
            tma_prefetch_l2_2d(&tm_k, 0, lrow + tail->a_list[aw][i].x);
            tma_prefetch_l2_2d(&tm_v, 0, lrow + tail->a_list[aw][i].x);
          }
        }
        gemm_duty(ph, l);
        if (ph == PK_UP && up_shared) {
          __syncwarp();
          asm volatile("bar.sync 2, %0;" ::"r"(128 + kPkAttnWarps * 32) : "memory");  // rstd and the parked accumulator are in place
          for (int u = 0; u < tab.n_units[PK_UP]; ++u)
            pk_finish_up_unit(&p, tail->rstd, park, tab.u[PK_UP][u], tab.n_units[PK_UP] == 1, cta, 4 + aw, 4 + kPkAttnWarps, lane);
          fence_proxy_async_all();
          asm volatile("bar.arrive 4, %0;" ::"r"(128 + kPkAttnWarps * 32) : "memory");
        }
        if (ph != PK_QKV) continue;
        __syncwarp();
        // ---- attention over the valid rows of both cache segments ----
        if (may_prime && lane == 0) {
          uint8_t* kt = attn_tiles + aw * 2 * 8192;
          if constexpr (KVQ != 0) {  // byte tiles: K and V of the first tile into stage 0, one barrier (pk_attention_cta_q8)
            const int row0 = l * p.kv_layer_rows + tail->a_list[aw][0].x;
            mbar_expect_tx(&tail->attn_bars[2 * aw], 8192);
            tma_load_2d(kt, &tm_k, 0, row0, &tail->attn_bars[2 * aw], kEvictFirst);
            tma_load_2d(kt + 4096, &tm_v, 0, row0, &tail->attn_bars[2 * aw], kEvictFirst);
          } else {
            pk_issue_tile(kt, &tm_k, &tail->attn_bars[2 * aw], p, l * p.kv_layer_rows, tail->a_list[aw][0].x, tail->a_row2[aw][0]);
            pk_issue_tile(kt + 8192, &tm_v, &tail->attn_bars[2 * aw + 1], p, l * p.kv_layer_rows, tail->a_list[aw][0].x, tail->a_row2[aw][0]);
          }
        }
        pk_wait_flag(&tail->bar_done, uint32_t(2 + 5 * l));
        if (lane == 0) pk_ev(ev, 500);
        if constexpr (KVQ != 0) pk_attention_cta_q8<KVQ == 2>(tm_k, tm_v, p, tail, l, attn_tiles, kv_phase, cta, aw, lane, may_prime, ev);
        else pk_attention_cta(tm_k, tm_v, p, tail, l, attn_tiles, kv_phase, cta, aw, lane, may_prime, ev);
        if (lane == 0) pk_ev(ev, 501);
        fence_proxy_async_all();
        named_bar_sync(3, kPkAttnWarps * 32);
        if (atid == 0) pk_grid_barrier(p, tail, uint32_t(3 + 5 * l));
      }
      __syncwarp();
    }
    if (xproducer && lane == 0) x_phase(PK_LOGITS, p.L);
    __syncwarp();
  }

  tcgen05_fence_before();
  __syncthreads();
  timeline_end(tl);
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mtx
