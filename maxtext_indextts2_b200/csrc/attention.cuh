// GQA decode attention over the valid rows of the two KV-cache segments.
//
// Replaces AttentionOp.__call__ in autoregressive mode (MaxText/layers/attentions.py:1399-1466):
// two masked dot-product attentions (prefill segment, AR ring) merged by
// normalize_attention (:1376-1397).  Mathematically that is one softmax over the rows whose
// segment id is active; this kernel reads ONLY those rows (the reference reads every
// allocated row and masks).
//
// Layout: one layer of the cache is [num_slots, Hkv, T, D] bf16 -- the sequence of a
// (slot, kv-head) is contiguous, rows [0,P) prefill, rows [P,T) the AR ring.
//
// Work decomposition (split-KV): the valid rows of a (row, kv-head) are cut into 64-row
// tiles that never straddle a segment boundary; 4 tiles = one work item = one CTA iteration
// (one tile per warp).  A persistent grid walks a device-built work list, so ragged contexts
// cost no empty CTAs and the launch is graph-capturable.  Each warp TMA-loads its K and V
// tile (SWIZZLE_128B) into its own smem, computes S = Q K^T for the G query heads of the
// group at once (so K/V are read once per group), an fp32 softmax with quad shuffles, and
// O = P V.  Warps merge in smem; items merge through an L2 workspace, the last CTA to
// arrive for a (row, kv-head) writing the bf16 result.
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"

namespace mtx {

constexpr int kAttnThreads = 128;
constexpr int kAttnTileRows = 64;   // kv rows per warp tile
constexpr int kAttnWarps = 4;       // tiles per work item

struct AttnParams {
  const bf16* q;          // [rows, Hq*D]
  bf16* out;              // [rows, Hq*D]
  const int* plane;       // [rows]
  const int* len0;        // [rows] valid rows of the prefill segment
  const int* ring_first;  // [rows] ring offset of the first valid AR row
  const int* ring_len;    // [rows] number of valid AR rows
  const int* work_items;  // [work_count] (row << 16) | chunk
  const int* work_count;  // [1]
  float* part_o;          // [rows, Hkv, max_chunks, G, D]
  float* part_ml;         // [rows, Hkv, max_chunks, G, 2]
  int* tickets;           // [rows, Hkv], zero between launches
  int rows, hq, hkv, P, T, max_chunks;
  int tiles_per_item;     // 64-row tiles one CTA iteration covers (multiple of 4); >= all tiles: no cross-CTA merge
  int plane_base;         // layer * num_slots (row coordinate base in the cache tensor map)
  float softcap;
  long long* trace;  // debug: clock64 timeline of warp 0 of CTA 0 (8 slots per item), or null
  // mtx_ragged_attention (the gpu_ragged_attention contract, attentions.py:761-815):
  int seq_major;     // 1: K/V are [plane, T, Hkv, D] (the reference's logical cache layout): tile = rows of one head at stride Hkv*D
  float* out_max;    // [rows, Hq] or null.  Non-null: `out` is left UNNORMALISED (sum_t exp(s_t - max) v_t) and the row's
  float* out_sum;    // [rows, Hq]   (max, sum) are returned for the caller's cross-segment merge (attentions.py:1376-1397)
  // Sliding-window layers in AUTOREGRESSIVE mode (attentions.py:600-602,624-631: the window is taken over the CACHE INDICES of
  // each segment, next_pos = kv_seq_len - 1): only prefill rows [skip0, P) and ring indices [ring_off, R) can be attended.
  // len0 / ring_first / ring_len then describe the valid rows INSIDE those two sub-ranges (prepare_rows_kernel, *_w arrays).
  int skip0;      // first prefill row of the window (0: no window)
  int ring_off;   // first ring index of the window (0: no window)
  int ring_size;  // ring indices in the window (0: the whole ring, T - P)
  const int* skip0_rows;  // [rows] or null: per-row addition to skip0 (prefill positions run as decode rows: position window)
  // Paged KV cache (attention=paged: inference/paged_attention.py:302-346, page_manager.py:49-91).  page_map != null: K / V are page
  // pools [Hkv, num_pages, tokens_per_page, D] (one layer at page_row_base rows into the tensor map), token i of row r (= page
  // group r) lives in page page_map[r][i / tokens_per_page]; len0 is the group's sequence length, the ring is empty.
  // The tensor maps' box is min(tokens_per_page, 64) rows: a 64-row tile is one slice of a page or 64 / tokens_per_page pages.
  const int* page_map;   // [page groups, max_pages]
  int tokens_per_page;   // power of two >= 8
  int num_pages, max_pages;
  long long page_row_base;
};

struct TileLoc { int p0, cnt; };

__host__ __device__ inline int attn_num_tiles(int len0, int ring_first, int ring_len, int R) {
  const int a = ring_len < R - ring_first ? ring_len : R - ring_first;
  const int b = ring_len - a;
  return (len0 + 63) / 64 + (a + 63) / 64 + (b + 63) / 64;
}

__host__ __device__ inline int attn_max_tiles(int P, int T) { return (P + 63) / 64 + (T - P + 63) / 64 + 1; }

// Chunks of a row's valid sequence when one work item covers `tpi` tiles.
__host__ __device__ inline int attn_max_chunks(int P, int T, int tpi = kAttnWarps) {
  return (attn_max_tiles(P, T) + tpi - 1) / tpi;
}

// Tiles per work item: the whole sequence of a (row, kv head) when there are at least as many such
// pairs as CTA slots worth filling (no cross-CTA merge at all); otherwise cut so that roughly
// `target_items` items exist for the longest possible context.
__host__ inline int attn_tiles_per_item(int rows, int hkv, int P, int T, int target_items) {
  const int pairs = rows * hkv;
  const int mt = attn_max_tiles(P, T);
  if (pairs >= target_items) return (mt + kAttnWarps - 1) / kAttnWarps * kAttnWarps;
  int tpi = (mt * pairs + target_items - 1) / target_items;
  tpi = (tpi + kAttnWarps - 1) / kAttnWarps * kAttnWarps;
  return tpi < kAttnWarps ? kAttnWarps : tpi;
}

// Physical start row and valid count of 64-row tile t of a row's valid sequence.
__device__ __forceinline__ TileLoc attn_tile(int t, int len0, int ring_first, int ring_len, int P, int R, int skip0 = 0, int ring_off = 0) {
  const int a = min(ring_len, R - ring_first);
  const int b = ring_len - a;
  const int n0 = (len0 + 63) >> 6, na = (a + 63) >> 6, nb = (b + 63) >> 6;
  TileLoc loc;
  loc.cnt = 0;
  loc.p0 = 0;
  if (t < n0) {
    loc.p0 = skip0 + t * 64;
    loc.cnt = min(64, len0 - t * 64);
  } else if (t < n0 + na) {
    const int u = t - n0;
    loc.p0 = P + ring_off + ring_first + u * 64;
    loc.cnt = min(64, a - u * 64);
  } else if (t < n0 + na + nb) {
    const int u = t - n0 - na;
    loc.p0 = P + ring_off + u * 64;
    loc.cnt = min(64, b - u * 64);
  }
  return loc;
}

// One block: per-row chunk counts -> prefix sum -> (row, chunk) list.
__global__ void attn_build_worklist_kernel(const int* len0, const int* ring_first, const int* ring_len, int rows, int P, int T,
                                           int tpi, int* work_items, int* work_count) {
  __shared__ int s_off[257];
  griddep_launch_dependents();
  griddep_wait();
  const int R = T - P;
  const int tid = threadIdx.x;
  int chunks = 0;
  if (tid < rows) chunks = (attn_num_tiles(len0[tid], ring_first[tid], ring_len[tid], R) + tpi - 1) / tpi;
  s_off[tid + 1] = chunks;
  if (tid == 0) s_off[0] = 0;
  __syncthreads();
  if (tid == 0)
    for (int i = 1; i <= 256; ++i) s_off[i] += s_off[i - 1];
  __syncthreads();
  if (tid < rows) {
    const int base = s_off[tid];
    for (int c = 0; c < chunks; ++c) work_items[base + c] = (tid << 16) | c;
  }
  if (tid == 0) *work_count = s_off[rows];
}

// One thread requests 64-row tile `t` (physical row `row` when the cache is contiguous) of kv head `h` for row `r`: the whole tile,
// or -- paged with pages shorter than the tile -- the pages that hold its `cnt` valid rows (the rest of the tile keeps stale
// bytes: the scores of those rows are masked by selection and the value rows are zeroed before the P V product).
template <int D>
__device__ __forceinline__ void attn_issue_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, const AttnParams& p, int r, int h, int t,
                                                int cnt, int col0, int row) {
  constexpr int kSub = D / 64;
  if (p.page_map == nullptr) {
    mbar_expect_tx(bar, 64 * D * 2);
#pragma unroll
    for (int s = 0; s < kSub; ++s) tma_load_2d(dst + s * 8192, tm, col0 + s * 64, row, bar, kEvictFirst);
    return;
  }
  const int tpp = p.tokens_per_page;
  const int* pm = p.page_map + (long long)r * p.max_pages;
  const int tok = t * 64;
  if (tpp >= 64) {
    const long long prow = p.page_row_base + ((long long)h * p.num_pages + pm[tok / tpp]) * tpp + (tok & (tpp - 1));
    mbar_expect_tx(bar, 64 * D * 2);
#pragma unroll
    for (int s = 0; s < kSub; ++s) tma_load_2d(dst + s * 8192, tm, s * 64, int(prow), bar, kEvictFirst);
  } else {
    const int nb = (cnt + tpp - 1) / tpp;
    mbar_expect_tx(bar, nb * tpp * D * 2);
    for (int j = 0; j < nb; ++j) {
      const long long prow = p.page_row_base + ((long long)h * p.num_pages + pm[tok / tpp + j]) * tpp;
#pragma unroll
      for (int s = 0; s < kSub; ++s) tma_load_2d(dst + s * 8192 + j * tpp * 128, tm, s * 64, int(prow), bar, kEvictFirst);
    }
  }
}

// The attention work loop of 128 threads (4 warps).  `tiles` = [warp][K tile | V tile] (1024-byte
// aligned), `sm_o_all` = merge buffer [warp][o_rows][D] fp32, `bars` = 2 initialised mbarriers per warp
// (their current parity in `phase`, updated on return), `sm_stat` = 128 floats + 1 int.  Items are
// taken from the work list starting at `first_item` with stride `item_stride`.  `tid` is 0..127; the
// 128 threads synchronise with the named barrier of epi_bar_sync().
template <int D>
__device__ __forceinline__ void attn_process_items(const CUtensorMap& tm_k, const CUtensorMap& tm_v, const AttnParams& p, uint8_t* tiles,
                                                   float* sm_o_all, uint64_t* bars, float* sm_stat, uint32_t& phase, int tid,
                                                   int first_item, int item_stride) {
  constexpr int kSub = D / 64;                 // 64-wide (128-byte) sub-tiles per row
  constexpr int kTileBytes = 64 * D * 2;       // one warp's K (or V) tile
  constexpr float kLog2e = 1.4426950408889634f;
  uint8_t* smem = tiles;
  const int G = p.hq / p.hkv;
  const int o_rows = G <= 8 ? 8 : 16;
  float* sm_m = sm_stat;                 // [4][16]
  float* sm_l = sm_m + kAttnWarps * 16;  // [4][16]
  volatile int* s_last_p = reinterpret_cast<volatile int*>(sm_l + kAttnWarps * 16);

  const int warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tid4 = lane & 3;
  const int mtx_i = lane >> 3, lrow = lane & 7;
  const int R = p.ring_size > 0 ? p.ring_size : p.T - p.P;
  const int TPI = p.tiles_per_item;
  uint8_t* k_tile = smem + warp * 2 * kTileBytes;
  uint8_t* v_tile = k_tile + kTileBytes;
  float* sm_o = sm_o_all + warp * o_rows * D;
  uint64_t* bar_k = bars + 2 * warp;
  uint64_t* bar_v = bar_k + 1;
  const uint32_t kb = smem_u32(k_tile), vb = smem_u32(v_tile);

  const int n_items = *p.work_count * p.hkv;
  long long* trace = (p.trace != nullptr && blockIdx.x == 0 && tid == 0) ? p.trace : nullptr;
  const long long t_start = clock64();
  int iter = 0;

  for (int item = first_item; item < n_items; item += item_stride) {
    const int packed = p.work_items[item / p.hkv];
    const int h = item % p.hkv;
    const int r = packed >> 16, chunk = packed & 0xffff;
    const int len0 = p.len0[r], rf = p.ring_first[r], rl = p.ring_len[r];
    const int skip0 = p.skip0 + (p.skip0_rows != nullptr ? p.skip0_rows[r] : 0);
    const int nt = attn_num_tiles(len0, rf, rl, R);
    const int n_chunks = (nt + TPI - 1) / TPI;
    const int t_begin = chunk * TPI, t_end = min(nt, t_begin + TPI);
    const int plane_row = p.seq_major ? (p.plane_base + p.plane[r]) * p.T : ((p.plane_base + p.plane[r]) * p.hkv + h) * p.T;
    const int col0 = p.seq_major ? h * D : 0;  // element column of the head inside a cache row
    if (trace && iter < 15) trace[iter * 8 + 0] = clock64() - t_start;

    // first tile of this warp: start both loads before touching Q
    int t = t_begin + warp;
    TileLoc loc = attn_tile(t < t_end ? t : nt, len0, rf, rl, p.P, R, skip0, p.ring_off);  // t >= nt gives an empty tile
    if (t < t_end && lane == 0) {
      attn_issue_tile<D>(k_tile, &tm_k, bar_k, p, r, h, t, loc.cnt, col0, plane_row + loc.p0);
      attn_issue_tile<D>(v_tile, &tm_v, bar_v, p, r, h, t, loc.cnt, col0, plane_row + loc.p0);
    }

    // Q fragments (A operand, rows = query heads of the group, zero-padded to 16)
    uint32_t qf[D / 16][4];
    {
      const bf16* qrow = p.q + (long long)r * p.hq * D + (long long)h * G * D;
#pragma unroll
      for (int tt = 0; tt < D / 16; ++tt) {
        const int d = tt * 16 + tid4 * 2;
        qf[tt][0] = gid < G ? *reinterpret_cast<const uint32_t*>(qrow + gid * D + d) : 0u;
        qf[tt][1] = gid + 8 < G ? *reinterpret_cast<const uint32_t*>(qrow + (gid + 8) * D + d) : 0u;
        qf[tt][2] = gid < G ? *reinterpret_cast<const uint32_t*>(qrow + gid * D + d + 8) : 0u;
        qf[tt][3] = gid + 8 < G ? *reinterpret_cast<const uint32_t*>(qrow + (gid + 8) * D + d + 8) : 0u;
      }
    }

    // running softmax state of this warp over its tiles (rows gid and gid+8 of the fragment)
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    float o[D / 8][4];
#pragma unroll
    for (int j = 0; j < D / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f;

    for (; t < t_end; t += kAttnWarps) {
      const int cnt = loc.cnt;
      const int tn = t + kAttnWarps;
      const TileLoc nloc = attn_tile(tn < t_end ? tn : nt, len0, rf, rl, p.P, R, skip0, p.ring_off);
      // ---- S = Q K^T over the 64 rows of the tile ----
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
      mbar_wait(bar_k, phase);
      if (trace && iter < 15 && t == t_begin) trace[iter * 8 + 1] = clock64() - t_start;
#pragma unroll
      for (int tt = 0; tt < D / 16; ++tt) {
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          const int row = 8 * (2 * jp + (mtx_i >> 1)) + lrow;
          const int c = 2 * tt + (mtx_i & 1);
          const uint32_t addr = kb + (c >> 3) * 8192 + row * 128 + (((c & 7) ^ (row & 7)) << 4);
          uint32_t b00, b01, b10, b11;
          ldmatrix_x4(addr, b00, b01, b10, b11);
          mma_m16n8k16_bf16(s[2 * jp], qf[tt], b00, b01);
          mma_m16n8k16_bf16(s[2 * jp + 1], qf[tt], b10, b11);
        }
      }
      // K tile consumed: refill it with the next tile's keys while the softmax and P V run
      fence_proxy_async();
      __syncwarp();
      if (tn < t_end && lane == 0) attn_issue_tile<D>(k_tile, &tm_k, bar_k, p, r, h, tn, nloc.cnt, col0, plane_row + nloc.p0);
      // ---- mask + online softmax (quad shuffles) ----
      float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = 8 * j + tid4 * 2 + (e & 1);
          float x = s[j][e];
          if (p.softcap != 0.0f) x = tanhf(x / p.softcap) * p.softcap;
          s[j][e] = col < cnt ? x : -INFINITY;
        }
        tm0 = fmaxf(tm0, fmaxf(s[j][0], s[j][1]));
        tm1 = fmaxf(tm1, fmaxf(s[j][2], s[j][3]));
      }
      tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1));
      tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
      tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1));
      tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
      const float nm0 = fmaxf(m0, tm0), nm1 = fmaxf(m1, tm1);
      const float a0 = exp2f((m0 - nm0) * kLog2e), a1 = exp2f((m1 - nm1) * kLog2e);  // exp2(-inf) = 0 on the first tile
      m0 = nm0;
      m1 = nm1;
      l0 *= a0;
      l1 *= a1;
      uint32_t pa[4][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p0 = exp2f((s[j][0] - m0) * kLog2e), p1 = exp2f((s[j][1] - m0) * kLog2e);
        const float p2 = exp2f((s[j][2] - m1) * kLog2e), p3 = exp2f((s[j][3] - m1) * kLog2e);
        l0 += p0 + p1;
        l1 += p2 + p3;
        // probabilities are cast to the value dtype before the PV product
        // (kernels/ragged_attention.py:156 `unnormalized.astype(v.dtype)`)
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
      // ---- O += P V ----
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
      // V tile consumed: refill
      fence_proxy_async();
      __syncwarp();
      if (tn < t_end && lane == 0) attn_issue_tile<D>(v_tile, &tm_v, bar_v, p, r, h, tn, nloc.cnt, col0, plane_row + nloc.p0);
      phase ^= 1;
      loc = nloc;
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    if (trace && iter < 15) trace[iter * 8 + 4] = clock64() - t_start;

    // ---- merge the four warps of the item in shared memory ----
    if (m0 > -INFINITY) {
#pragma unroll
      for (int j = 0; j < D / 8; ++j) {
        const int d = 8 * j + tid4 * 2;
        if (gid < G) *reinterpret_cast<float2*>(sm_o + gid * D + d) = make_float2(o[j][0], o[j][1]);
        if (gid + 8 < G) *reinterpret_cast<float2*>(sm_o + (gid + 8) * D + d) = make_float2(o[j][2], o[j][3]);
      }
    }
    if (tid4 == 0) {
      sm_m[warp * 16 + gid] = m0;
      sm_m[warp * 16 + gid + 8] = m1;
      sm_l[warp * 16 + gid] = l0;
      sm_l[warp * 16 + gid + 8] = l1;
    }
    epi_bar_sync();
    if (trace && iter < 15) trace[iter * 8 + 5] = clock64() - t_start;

    const long long out_base = (long long)r * p.hq * D + (long long)h * G * D;
    const long long part_base = ((long long)(r * p.hkv + h) * p.max_chunks + chunk) * G;
    for (int e = tid; e < G * D; e += kAttnThreads) {
      const int g = e / D, d = e - g * D;
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, sm_m[w * 16 + g]);
      float L = 0.0f, O = 0.0f;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) {
        const float mw = sm_m[w * 16 + g];
        if (mw > -INFINITY) {
          const float sc = exp2f((mw - M) * kLog2e);
          L += sm_l[w * 16 + g] * sc;
          O += sm_o_all[(w * o_rows + g) * D + d] * sc;
        }
      }
      if (n_chunks == 1) {
        p.out[out_base + e] = __float2bfloat16_rn(p.out_max != nullptr ? O : O / L);
        if (p.out_max != nullptr && d == 0) {
          p.out_max[(long long)r * p.hq + h * G + g] = M;
          p.out_sum[(long long)r * p.hq + h * G + g] = L;
        }
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
      epi_bar_sync();
      if (tid == 0) {
        const int old = atomicAdd(p.tickets + r * p.hkv + h, 1);
        const int last = old == n_chunks - 1;
        if (last) p.tickets[r * p.hkv + h] = 0;
        *s_last_p = last;
      }
      epi_bar_sync();
      if (*s_last_p) {
        __threadfence();
        const long long pb = (long long)(r * p.hkv + h) * p.max_chunks * G;
        for (int e = tid; e < G * D; e += kAttnThreads) {
          const int g = e / D, d = e - g * D;
          float M = -INFINITY;
          for (int c = 0; c < n_chunks; ++c) M = fmaxf(M, __ldcg(p.part_ml + (pb + c * G + g) * 2));
          float L = 0.0f, O = 0.0f;
          for (int c = 0; c < n_chunks; ++c) {
            const float sc = exp2f((__ldcg(p.part_ml + (pb + c * G + g) * 2) - M) * kLog2e);
            L += __ldcg(p.part_ml + (pb + c * G + g) * 2 + 1) * sc;
            O += __ldcg(p.part_o + (pb + c * G + g) * D + d) * sc;
          }
          p.out[out_base + e] = __float2bfloat16_rn(p.out_max != nullptr ? O : O / L);
          if (p.out_max != nullptr && d == 0) {
            p.out_max[(long long)r * p.hq + h * G + g] = M;
            p.out_sum[(long long)r * p.hq + h * G + g] = L;
          }
        }
      }
    }
    if (trace && iter < 15) trace[iter * 8 + 6] = clock64() - t_start;
    epi_bar_sync();  // the merge buffer is rewritten by the next item
    if (trace && iter < 15) trace[iter * 8 + 7] = clock64() - t_start;
    ++iter;
  }
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads)
decode_attn_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const AttnParams p) {
  constexpr int kTileBytes = 64 * D * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space
  const int G = p.hq / p.hkv;
  const int o_rows = G <= 8 ? 8 : 16;
  // [warp][K tile | V tile] | merge buffer [warp][o_rows][D] fp32 | barriers | merge statistics
  float* sm_o_all = reinterpret_cast<float*>(smem + kAttnWarps * 2 * kTileBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_o_all + kAttnWarps * o_rows * D);
  float* sm_stat = reinterpret_cast<float*>(bars + 2 * kAttnWarps);
  const int tl = timeline_begin(3);
  griddep_launch_dependents();
  if ((threadIdx.x & 31) == 0) {
    mbar_init(bars + 2 * (threadIdx.x >> 5), 1);
    mbar_init(bars + 2 * (threadIdx.x >> 5) + 1, 1);
    fence_barrier_init();
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  __syncthreads();
  griddep_wait();
  uint32_t phase = 0;
  attn_process_items<D>(tm_k, tm_v, p, smem, sm_o_all, bars, sm_stat, phase, threadIdx.x, blockIdx.x, gridDim.x);
  timeline_end(tl);
}


__host__ inline size_t attn_smem_bytes(int D, int G) {
  const size_t o_rows = G <= 8 ? 8 : 16;
  return 1024 + size_t(kAttnWarps) * 2 * (64 * D * 2) + kAttnWarps * o_rows * D * 4 + 2 * kAttnWarps * 8 + 2 * kAttnWarps * 16 * 4 + 32;
}

// ---- int8 KV cache (quantize_kvcache, kv_quant_dtype int8, kv_quant_axis dkv: MaxText/inference/kvcache.py:36-90) ---------
//
// A cache row of one kv head is 64 unsigned bytes u = q + 128 with q = clip(rint(x * 127.5 / scale), -128, 127) and one fp32
// scale = max|x| over the head's 64 dims (KVQuant.quantize with the "dkv" axis); x ~ q * scale / 127.5.  The tile loop below is
// the bf16 one with three changes:
//   * K / V tiles are 4 KB (TMA, SWIZZLE_64B); bytes become EXACT fp16 integers in registers with one byte permute per pair
//     (0x6400 | u = 1024 + u, minus 1152) and the MMAs run in fp16 (the queries, bf16 values, convert exactly);
//   * the per-row scales are folded in where they are per-MMA-row constants: k_scale[kv] / 127.5 multiplies the scores,
//     v_scale[kv] / 127.5 multiplies the probabilities before the P V product;
//   * a dot product does not care about the order of its terms, so the head dims are PERMUTED to make every operand fragment
//     a vector load: a thread takes 16 consecutive bytes of a K row for all four k-steps of S = Q K^T (the query fragments are
//     gathered with the same permutation), and 8 consecutive bytes of a V row for all eight output blocks of O = P V (the
//     output columns come out permuted: element (n-block j, column c) is head dim 8 c + j).
constexpr int kQ8WarpBytes = 9216;  // per warp: K tile 4 KB | V tile 4 KB | 64 K scales + 64 V scales (+ pad)
constexpr float kQ8Inv = 1.0f / 127.5f;
constexpr float kF8Inv = 1.0f / 448.0f;  // kv_quant_dtype fp8: value = e4m3(x * 448 / scale) (kvcache.py:38,86-88), x ~ value * scale / 448

__device__ __forceinline__ void mma_m16n8k16_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Two bytes of `w` (selected by `sel`: 0x4140 = bytes 0,1; 0x4342 = bytes 2,3; 0x5140 = byte 0 of w and byte 0 of the second word ...)
// as the exact fp16 pair (u_lo - 128, u_hi - 128).
__device__ __forceinline__ uint32_t q8_pair_f16(uint32_t w, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x64646464u), "r"(sel));
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(0x64806480u));
  return r;
}
// The same two bytes read as float8_e4m3fn values (kv_quant_dtype fp8): one permute to pair them, one packed conversion (exact).
__device__ __forceinline__ uint32_t f8_pair_f16(uint32_t w, uint32_t sel) {
  uint32_t t, r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(w), "r"(0u), "r"((sel & 0xFu) | ((sel >> 4) & 0xF0u)));
  asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f16x2.e4m3x2 %0, lo; }" : "=r"(r) : "r"(t));
  return r;
}
template <bool F8>
__device__ __forceinline__ uint32_t kv8_pair_f16(uint32_t w, uint32_t sel) {
  return F8 ? f8_pair_f16(w, sel) : q8_pair_f16(w, sel);
}
__device__ __forceinline__ uint32_t bf16x2_to_f16x2(uint32_t v) {
  const __half2 h = __floats2half2_rn(bf16_lo(v), bf16_hi(v));
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Transposes an 8x8 matrix of 16-bit elements held one row pair per thread (thread (gid, tid4) holds M[gid][2 tid4 .. +1]).
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_shared_v2(uint32_t addr) {
  uint2 v;
  asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}

// ---- the same tile in TRANSPOSED form, for groups of up to 8 heads: S^T = K Q^T and O^T = V^T P^T, the 16-row MMA dimension
// carrying kv rows / head dims and the 8-column dimension the heads (no padding of the group to 16 rows: half the MMAs).
//   thread (gid, tid4) owns heads h0 = 2 tid4, h1 = h0 + 1;  m / l: running max and (this thread's rows of the) sum per head
//   qb[kk][0..1] = fp16 Q[head gid][16 tid4 + 4 kk + (0,1)], [.. + (2,3)]   (zero for gid >= G)
//   o[db][0..3]  = O^T[d = 8 gid + 2 db (+1 for 2,3)][h0, h1]: a thread holds 8 consecutive dims of its two heads
// K as the A operand: k-slots (2 tid4, 2 tid4 + 1, 2 tid4 + 8, 2 tid4 + 9) of k-step kk are head dims 16 tid4 + 4 kk + (0..3), i.e.
// bytes 4 kk .. 4 kk + 3 of ONE 16-byte load per kv row (the query fragments carry the same permutation).
// kv8t_scores: scores of the tile at shared address kb (64 rows x 64 bytes, SWIZZLE_64B), dequantised with s_ks[row], masked past
// `cnt`, online-softmax update of (m, l, o); returns P^T (times s_vs[row]) as the fp16 B operand pb of the P V product.
template <bool F8>
__device__ __forceinline__ void kv8t_scores(uint32_t kb, const float* s_ks, const float* s_vs, int cnt, float softcap, const uint32_t (&qb)[4][2],
                                            float& m0, float& m1, float& l0, float& l1, float (&o)[4][4], uint32_t (&pb)[4][2], int gid, int tid4) {
  constexpr float kLog2e = 1.4426950408889634f;
  float s[4][4];
#pragma unroll
  for (int mb = 0; mb < 4; ++mb) s[mb][0] = s[mb][1] = s[mb][2] = s[mb][3] = 0.0f;
  {
    uint4 wa[4], wb[4];  // the thread's 16 bytes of kv rows 16 mb + gid and 16 mb + gid + 8
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      const int ra = 16 * mb + gid, rb = ra + 8;
      wa[mb] = ld_shared_v4(kb + ra * 64 + ((tid4 ^ ((ra >> 1) & 3)) << 4));
      wb[mb] = ld_shared_v4(kb + rb * 64 + ((tid4 ^ ((rb >> 1) & 3)) << 4));
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {  // (independent accumulators back to back)
        const uint32_t xa = kk == 0 ? wa[mb].x : kk == 1 ? wa[mb].y : kk == 2 ? wa[mb].z : wa[mb].w;
        const uint32_t xb = kk == 0 ? wb[mb].x : kk == 1 ? wb[mb].y : kk == 2 ? wb[mb].z : wb[mb].w;
        const uint32_t a[4] = {kv8_pair_f16<F8>(xa, 0x4140u), kv8_pair_f16<F8>(xb, 0x4140u), kv8_pair_f16<F8>(xa, 0x4342u),
                               kv8_pair_f16<F8>(xb, 0x4342u)};
        mma_m16n8k16_f16(s[mb], a, qb[kk][0], qb[kk][1]);
      }
    }
  }
  // dequantise (per kv row), mask, online softmax: a head's scores live in the 8 threads of equal tid4
  float vsr[4][2];
  float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
  for (int mb = 0; mb < 4; ++mb) {
    const float ka = s_ks[16 * mb + gid], kc = s_ks[16 * mb + gid + 8];
    vsr[mb][0] = s_vs[16 * mb + gid];
    vsr[mb][1] = s_vs[16 * mb + gid + 8];
    s[mb][0] *= ka;
    s[mb][1] *= ka;
    s[mb][2] *= kc;
    s[mb][3] *= kc;
    if (softcap != 0.0f) {
#pragma unroll
      for (int q = 0; q < 4; ++q) s[mb][q] = tanhf(s[mb][q] / softcap) * softcap;
    }
    if (16 * mb + gid >= cnt) s[mb][0] = s[mb][1] = -INFINITY;
    if (16 * mb + gid + 8 >= cnt) s[mb][2] = s[mb][3] = -INFINITY;
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
#pragma unroll
  for (int mb = 0; mb < 4; ++mb) {
    const float p0 = ex2_approx(fmaf(s[mb][0], kLog2e, -ms0)), p1 = ex2_approx(fmaf(s[mb][1], kLog2e, -ms1));
    const float p2 = ex2_approx(fmaf(s[mb][2], kLog2e, -ms0)), p3 = ex2_approx(fmaf(s[mb][3], kLog2e, -ms1));
    l0 += p0 + p2;
    l1 += p1 + p3;
    pb[mb][0] = movmatrix_trans(pack_f16x2(p0 * vsr[mb][0], p1 * vsr[mb][0]));
    pb[mb][1] = movmatrix_trans(pack_f16x2(p2 * vsr[mb][1], p3 * vsr[mb][1]));
  }
#pragma unroll
  for (int db = 0; db < 4; ++db) {
    o[db][0] *= a0;
    o[db][1] *= a1;
    o[db][2] *= a0;
    o[db][3] *= a1;
  }
}

// O^T += V^T P^T over the tile at shared address vb: a thread reads the 8 bytes V[row][8 gid .. 8 gid + 7] of kv rows
// 16 mb + 2 tid4 (+1, +8, +9); m-block db takes bytes 2 db (row gid of the fragment) and 2 db + 1 (row gid + 8).
template <bool F8>
__device__ __forceinline__ void kv8t_pv(uint32_t vb, int cnt, const uint32_t (&pb)[4][2], float (&o)[4][4], int gid, int tid4) {
#pragma unroll
  for (int mb = 0; mb < 4; ++mb) {
    uint2 w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int row = 16 * mb + 2 * tid4 + (q & 1) + 8 * (q >> 1);
      w[q] = ld_shared_v2(vb + row * 64 + ((((gid >> 1) ^ ((row >> 1) & 3)) << 4) | ((gid & 1) << 3)));
      if (cnt < 64 && row >= cnt) w[q].x = w[q].y = F8 ? 0u : 0x80808080u;  // rows past the valid count may hold anything: make them zero
    }
#pragma unroll
    for (int db = 0; db < 4; ++db) {
      const uint32_t w0 = db < 2 ? w[0].x : w[0].y, w1 = db < 2 ? w[1].x : w[1].y, w2 = db < 2 ? w[2].x : w[2].y, w3 = db < 2 ? w[3].x : w[3].y;
      // bytes (2 db, 2 db + 1) mod 4 of the two rows of a pair -> (row0[b], row1[b], row0[b + 1], row1[b + 1])
      const uint32_t selg = (db & 1) ? 0x7362u : 0x5140u;
      uint32_t g01, g23;
      asm("prmt.b32 %0, %1, %2, %3;" : "=r"(g01) : "r"(w0), "r"(w1), "r"(selg));
      asm("prmt.b32 %0, %1, %2, %3;" : "=r"(g23) : "r"(w2), "r"(w3), "r"(selg));
      const uint32_t a[4] = {kv8_pair_f16<F8>(g01, 0x4140u), kv8_pair_f16<F8>(g01, 0x4342u), kv8_pair_f16<F8>(g23, 0x4140u),
                             kv8_pair_f16<F8>(g23, 0x4342u)};
      mma_m16n8k16_f16(o[db], a, pb[mb][0], pb[mb][1]);
    }
  }
}

// The attention work loop of 128 threads (4 warps) over an int8 cache, head_dim 64.  Same contract as attn_process_items;
// `tiles` = [warp][kQ8WarpBytes], k_scale / v_scale = fp32 per cache row, indexed like the rows of the K/V tensor maps.
// TR: groups of up to 8 heads run the tile in the transposed form (kv8t_scores / kv8t_pv): half the MMAs and half the
// accumulator registers of the heads-in-rows form below, which stays for wider groups.
template <bool F8, bool TR>
__device__ __forceinline__ void attn_process_items_q8(const CUtensorMap& tm_k, const CUtensorMap& tm_v, const AttnParams& p, const float* k_scale,
                                                      const float* v_scale, uint8_t* tiles, float* sm_o_all, uint64_t* bars, float* sm_stat,
                                                      uint32_t& phase, int tid, int first_item, int item_stride) {
  constexpr int D = 64;
  constexpr int kTileBytes = 64 * D;  // one warp's K (or V) tile: 64 rows of 64 bytes
  constexpr float kLog2e = 1.4426950408889634f;
  const int G = p.hq / p.hkv;
  const int o_rows = G <= 8 ? 8 : 16;
  float* sm_m = sm_stat;                 // [4][16]
  float* sm_l = sm_m + kAttnWarps * 16;  // [4][16]
  volatile int* s_last_p = reinterpret_cast<volatile int*>(sm_l + kAttnWarps * 16);

  const int warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tid4 = lane & 3;
  const int R = p.T - p.P;
  const int TPI = p.tiles_per_item;
  uint8_t* k_tile = tiles + warp * kQ8WarpBytes;
  uint8_t* v_tile = k_tile + kTileBytes;
  float* s_ks = reinterpret_cast<float*>(k_tile + 2 * kTileBytes);  // [64] K scales of the tile (already / 127.5)
  float* s_vs = s_ks + 64;                                          // [64] V scales
  float* sm_o = sm_o_all + warp * o_rows * D;
  uint64_t* bar_k = bars + 2 * warp;
  uint64_t* bar_v = bar_k + 1;
  const uint32_t kb = smem_u32(k_tile), vb = smem_u32(v_tile);

  const int n_items = *p.work_count * p.hkv;
  for (int item = first_item; item < n_items; item += item_stride) {
    const int packed = p.work_items[item / p.hkv];
    const int h = item % p.hkv;
    const int r = packed >> 16, chunk = packed & 0xffff;
    const int len0 = p.len0[r], rf = p.ring_first[r], rl = p.ring_len[r];
    const int nt = attn_num_tiles(len0, rf, rl, R);
    const int n_chunks = (nt + TPI - 1) / TPI;
    const int t_begin = chunk * TPI, t_end = min(nt, t_begin + TPI);
    const int plane_row = ((p.plane_base + p.plane[r]) * p.hkv + h) * p.T;

    int t = t_begin + warp;
    TileLoc loc = attn_tile(t < t_end ? t : nt, len0, rf, rl, p.P, R);  // t >= nt gives an empty tile
    if (t < t_end && lane == 0) {
      mbar_expect_tx(bar_k, kTileBytes);
      tma_load_2d(k_tile, &tm_k, 0, plane_row + loc.p0, bar_k, kEvictFirst);
      mbar_expect_tx(bar_v, kTileBytes);
      tma_load_2d(v_tile, &tm_v, 0, plane_row + loc.p0, bar_v, kEvictFirst);
    }

    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    if constexpr (TR) {
      uint32_t qb[4][2];
      {
        const bf16* qrow = p.q + (long long)r * p.hq * D + (long long)h * G * D + gid * D + 16 * tid4;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          qb[kk][0] = gid < G ? bf16x2_to_f16x2(*reinterpret_cast<const uint32_t*>(qrow + 4 * kk)) : 0u;
          qb[kk][1] = gid < G ? bf16x2_to_f16x2(*reinterpret_cast<const uint32_t*>(qrow + 4 * kk + 2)) : 0u;
        }
      }
      float o[4][4];
#pragma unroll
      for (int db = 0; db < 4; ++db) o[db][0] = o[db][1] = o[db][2] = o[db][3] = 0.0f;
      // a tile's scales (two K and two V rows per lane) are requested one tile ahead, with the tile's bytes
      float sk0, sk1, sv0, sv1;
      auto request_scales = [&](const TileLoc& tl) {
        const long long row0 = (long long)plane_row + tl.p0;
        sk0 = lane < tl.cnt ? __ldg(k_scale + row0 + lane) : 0.0f;
        sk1 = lane + 32 < tl.cnt ? __ldg(k_scale + row0 + lane + 32) : 0.0f;
        sv0 = lane < tl.cnt ? __ldg(v_scale + row0 + lane) : 0.0f;
        sv1 = lane + 32 < tl.cnt ? __ldg(v_scale + row0 + lane + 32) : 0.0f;
      };
      request_scales(loc);
      for (; t < t_end; t += kAttnWarps) {
        const int cnt = loc.cnt;
        const int tn = t + kAttnWarps;
        const TileLoc nloc = attn_tile(tn < t_end ? tn : nt, len0, rf, rl, p.P, R);
        __syncwarp();
        s_ks[lane] = sk0 * (F8 ? kF8Inv : kQ8Inv);
        s_ks[lane + 32] = sk1 * (F8 ? kF8Inv : kQ8Inv);
        s_vs[lane] = sv0 * (F8 ? kF8Inv : kQ8Inv);
        s_vs[lane + 32] = sv1 * (F8 ? kF8Inv : kQ8Inv);
        __syncwarp();
        if (tn < t_end) request_scales(nloc);
        uint32_t pb[4][2];
        mbar_wait(bar_k, phase);
        kv8t_scores<F8>(kb, s_ks, s_vs, cnt, p.softcap, qb, m0, m1, l0, l1, o, pb, gid, tid4);
        fence_proxy_async();
        __syncwarp();
        if (tn < t_end && lane == 0) {
          mbar_expect_tx(bar_k, kTileBytes);
          tma_load_2d(k_tile, &tm_k, 0, plane_row + nloc.p0, bar_k, kEvictFirst);
        }
        mbar_wait(bar_v, phase);
        kv8t_pv<F8>(vb, cnt, pb, o, gid, tid4);
        fence_proxy_async();
        __syncwarp();
        if (tn < t_end && lane == 0) {
          mbar_expect_tx(bar_v, kTileBytes);
          tma_load_2d(v_tile, &tm_v, 0, plane_row + nloc.p0, bar_v, kEvictFirst);
        }
        phase ^= 1;
        loc = nloc;
      }
#pragma unroll
      for (int sh = 4; sh < 32; sh <<= 1) {
        l0 += __shfl_xor_sync(0xffffffffu, l0, sh);
        l1 += __shfl_xor_sync(0xffffffffu, l1, sh);
      }
      // the warp's partial to the merge buffer: head h0 = 2 tid4 holds dims 8 gid .. 8 gid + 7 in o[.][0], o[.][2]; h0 + 1 in o[.][1], o[.][3]
      const int h0 = 2 * tid4;
      if (h0 < G) {
        float* row = sm_o + h0 * D + 8 * gid;
        *reinterpret_cast<float4*>(row) = make_float4(o[0][0], o[0][2], o[1][0], o[1][2]);
        *reinterpret_cast<float4*>(row + 4) = make_float4(o[2][0], o[2][2], o[3][0], o[3][2]);
      }
      if (h0 + 1 < G) {
        float* row = sm_o + (h0 + 1) * D + 8 * gid;
        *reinterpret_cast<float4*>(row) = make_float4(o[0][1], o[0][3], o[1][1], o[1][3]);
        *reinterpret_cast<float4*>(row + 4) = make_float4(o[2][1], o[2][3], o[3][1], o[3][3]);
      }
      if (gid == 0) {
        sm_m[warp * 16 + h0] = m0;
        sm_m[warp * 16 + h0 + 1] = m1;
        sm_l[warp * 16 + h0] = l0;
        sm_l[warp * 16 + h0 + 1] = l1;
      }
    } else {
    // Q fragments (A operand: rows = query heads of the group, zero-padded to 16) as fp16, head dims gathered with the K
    // permutation: k-slot (2 tid4 + e) of k-step tt is head dim 16 tid4 + 4 tt + e, k-slot (2 tid4 + 8 + e) is dim 16 tid4 + 4 tt + 2 + e
    uint32_t qf[4][4];
    {
      const bf16* qrow = p.q + (long long)r * p.hq * D + (long long)h * G * D;
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) {
        const int d = 16 * tid4 + 4 * tt;
        qf[tt][0] = gid < G ? bf16x2_to_f16x2(*reinterpret_cast<const uint32_t*>(qrow + gid * D + d)) : 0u;
        qf[tt][1] = gid + 8 < G ? bf16x2_to_f16x2(*reinterpret_cast<const uint32_t*>(qrow + (gid + 8) * D + d)) : 0u;
        qf[tt][2] = gid < G ? bf16x2_to_f16x2(*reinterpret_cast<const uint32_t*>(qrow + gid * D + d + 2)) : 0u;
        qf[tt][3] = gid + 8 < G ? bf16x2_to_f16x2(*reinterpret_cast<const uint32_t*>(qrow + (gid + 8) * D + d + 2)) : 0u;
      }
    }

    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f;

    for (; t < t_end; t += kAttnWarps) {
      const int cnt = loc.cnt;
      const int tn = t + kAttnWarps;
      const TileLoc nloc = attn_tile(tn < t_end ? tn : nt, len0, rf, rl, p.P, R);
      // the tile's scales: two rows per lane, staged for the per-column reads below
      {
        const long long row0 = (long long)plane_row + loc.p0;
        __syncwarp();
        s_ks[lane] = lane < cnt ? __ldg(k_scale + row0 + lane) * (F8 ? kF8Inv : kQ8Inv) : 0.0f;
        s_ks[lane + 32] = lane + 32 < cnt ? __ldg(k_scale + row0 + lane + 32) * (F8 ? kF8Inv : kQ8Inv) : 0.0f;
        s_vs[lane] = lane < cnt ? __ldg(v_scale + row0 + lane) * (F8 ? kF8Inv : kQ8Inv) : 0.0f;
        s_vs[lane + 32] = lane + 32 < cnt ? __ldg(v_scale + row0 + lane + 32) * (F8 ? kF8Inv : kQ8Inv) : 0.0f;
        __syncwarp();
      }
      // ---- S = Q K^T over the 64 rows of the tile: n-block j = kv rows 8 j .. 8 j + 7, one 16-byte load per row ----
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
      mbar_wait(bar_k, phase);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = 8 * j + gid;
        const uint32_t addr = kb + row * 64 + ((tid4 ^ ((row >> 1) & 3)) << 4);
        uint32_t w[4];
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) mma_m16n8k16_f16(s[j], qf[tt], kv8_pair_f16<F8>(w[tt], 0x4140u), kv8_pair_f16<F8>(w[tt], 0x4342u));
      }
      // K tile consumed: refill it with the next tile's keys while the softmax and P V run
      fence_proxy_async();
      __syncwarp();
      if (tn < t_end && lane == 0) {
        mbar_expect_tx(bar_k, kTileBytes);
        tma_load_2d(k_tile, &tm_k, 0, plane_row + nloc.p0, bar_k, kEvictFirst);
      }
      // ---- dequantise the scores (per kv column), mask, online softmax ----
      float vs[8][2];
      float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 ks2 = *reinterpret_cast<const float2*>(s_ks + 8 * j + tid4 * 2);
        const float2 vs2 = *reinterpret_cast<const float2*>(s_vs + 8 * j + tid4 * 2);
        vs[j][0] = vs2.x;
        vs[j][1] = vs2.y;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = 8 * j + tid4 * 2 + (e & 1);
          float x = s[j][e] * ((e & 1) ? ks2.y : ks2.x);
          if (p.softcap != 0.0f) x = tanhf(x / p.softcap) * p.softcap;
          s[j][e] = col < cnt ? x : -INFINITY;
        }
        tm0 = fmaxf(tm0, fmaxf(s[j][0], s[j][1]));
        tm1 = fmaxf(tm1, fmaxf(s[j][2], s[j][3]));
      }
      tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1));
      tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
      tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1));
      tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
      const float nm0 = fmaxf(m0, tm0), nm1 = fmaxf(m1, tm1);
      const float a0 = exp2f((m0 - nm0) * kLog2e), a1 = exp2f((m1 - nm1) * kLog2e);  // exp2(-inf) = 0 on the first tile
      m0 = nm0;
      m1 = nm1;
      l0 *= a0;
      l1 *= a1;
      uint32_t pa[4][4];  // P (times the V scale of its column) as the fp16 A operand, one k-step per 16 kv rows
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p0 = exp2f((s[j][0] - m0) * kLog2e), p1 = exp2f((s[j][1] - m0) * kLog2e);
        const float p2 = exp2f((s[j][2] - m1) * kLog2e), p3 = exp2f((s[j][3] - m1) * kLog2e);
        l0 += p0 + p1;
        l1 += p2 + p3;
        pa[j >> 1][(j & 1) * 2 + 0] = pack_f16x2(p0 * vs[j][0], p1 * vs[j][1]);
        pa[j >> 1][(j & 1) * 2 + 1] = pack_f16x2(p2 * vs[j][0], p3 * vs[j][1]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j][0] *= a0;
        o[j][1] *= a0;
        o[j][2] *= a1;
        o[j][3] *= a1;
      }
      // ---- O += P V: k-step u = kv rows 16 u .. 16 u + 15; B column gid of n-block j is head dim 8 gid + j, so a thread reads the
      // 8 bytes V[row][8 gid .. 8 gid + 7] of its four kv rows and builds the B fragments of all eight n-blocks from them ----
      mbar_wait(bar_v, phase);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t lo[4], hi[4];  // rows 16 u + 2 tid4, + 1, + 8, + 9: bytes 0..3 and 4..7 of the 8-byte run
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int row = 16 * u + 2 * tid4 + (q & 1) + 8 * (q >> 1);
          const uint32_t addr = vb + row * 64 + ((((gid >> 1) ^ ((row >> 1) & 3)) << 4) | ((gid & 1) << 3));
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(lo[q]), "=r"(hi[q]) : "r"(addr));
          if (row >= cnt) lo[q] = hi[q] = F8 ? 0u : 0x80808080u;  // rows past the valid count may hold anything: make them zero (u = 128; e4m3 +0)
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // byte j of the run: rows (2 tid4, 2 tid4 + 1) -> b0, rows (2 tid4 + 8, 2 tid4 + 9) -> b1
          const uint32_t w0 = j < 4 ? lo[0] : hi[0], w1 = j < 4 ? lo[1] : hi[1], w2 = j < 4 ? lo[2] : hi[2], w3 = j < 4 ? lo[3] : hi[3];
          uint32_t g0, g1;
          const uint32_t selg = 0x0040u + 0x0011u * uint32_t(j & 3);  // (byte j of the first word, byte j of the second word)
          asm("prmt.b32 %0, %1, %2, %3;" : "=r"(g0) : "r"(w0), "r"(w1), "r"(selg));
          asm("prmt.b32 %0, %1, %2, %3;" : "=r"(g1) : "r"(w2), "r"(w3), "r"(selg));
          mma_m16n8k16_f16(o[j], pa[u], kv8_pair_f16<F8>(g0, 0x4140u), kv8_pair_f16<F8>(g1, 0x4140u));
        }
      }
      // V tile consumed: refill
      fence_proxy_async();
      __syncwarp();
      if (tn < t_end && lane == 0) {
        mbar_expect_tx(bar_v, kTileBytes);
        tma_load_2d(v_tile, &tm_v, 0, plane_row + nloc.p0, bar_v, kEvictFirst);
      }
      phase ^= 1;
      loc = nloc;
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // ---- merge the four warps of the item in shared memory (accumulator column 2 tid4 + e of n-block j is head dim 16 tid4 + 8 e + j) ----
    if (m0 > -INFINITY) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int d0 = 16 * tid4 + j, d1 = d0 + 8;
        if (gid < G) { sm_o[gid * D + d0] = o[j][0]; sm_o[gid * D + d1] = o[j][1]; }
        if (gid + 8 < G) { sm_o[(gid + 8) * D + d0] = o[j][2]; sm_o[(gid + 8) * D + d1] = o[j][3]; }
      }
    }
    if (tid4 == 0) {
      sm_m[warp * 16 + gid] = m0;
      sm_m[warp * 16 + gid + 8] = m1;
      sm_l[warp * 16 + gid] = l0;
      sm_l[warp * 16 + gid + 8] = l1;
    }
    }  // (heads-in-rows form)
    epi_bar_sync();

    const long long out_base = (long long)r * p.hq * D + (long long)h * G * D;
    const long long part_base = ((long long)(r * p.hkv + h) * p.max_chunks + chunk) * G;
    for (int e = tid; e < G * D; e += kAttnThreads) {
      const int g = e / D, d = e - g * D;
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, sm_m[w * 16 + g]);
      float L = 0.0f, O = 0.0f;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) {
        const float mw = sm_m[w * 16 + g];
        if (mw > -INFINITY) {
          const float sc = exp2f((mw - M) * kLog2e);
          L += sm_l[w * 16 + g] * sc;
          O += sm_o_all[(w * o_rows + g) * D + d] * sc;
        }
      }
      if (n_chunks == 1) {
        p.out[out_base + e] = __float2bfloat16_rn(p.out_max != nullptr ? O : O / L);
        if (p.out_max != nullptr && d == 0) {
          p.out_max[(long long)r * p.hq + h * G + g] = M;
          p.out_sum[(long long)r * p.hq + h * G + g] = L;
        }
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
      epi_bar_sync();
      if (tid == 0) {
        const int old = atomicAdd(p.tickets + r * p.hkv + h, 1);
        const int last = old == n_chunks - 1;
        if (last) p.tickets[r * p.hkv + h] = 0;
        *s_last_p = last;
      }
      epi_bar_sync();
      if (*s_last_p) {
        __threadfence();
        const long long pb = (long long)(r * p.hkv + h) * p.max_chunks * G;
        for (int e = tid; e < G * D; e += kAttnThreads) {
          const int g = e / D, d = e - g * D;
          float M = -INFINITY;
          for (int c = 0; c < n_chunks; ++c) M = fmaxf(M, __ldcg(p.part_ml + (pb + c * G + g) * 2));
          float L = 0.0f, O = 0.0f;
          for (int c = 0; c < n_chunks; ++c) {
            const float sc = exp2f((__ldcg(p.part_ml + (pb + c * G + g) * 2) - M) * kLog2e);
            L += __ldcg(p.part_ml + (pb + c * G + g) * 2 + 1) * sc;
            O += __ldcg(p.part_o + (pb + c * G + g) * D + d) * sc;
          }
          p.out[out_base + e] = __float2bfloat16_rn(p.out_max != nullptr ? O : O / L);
          if (p.out_max != nullptr && d == 0) {
            p.out_max[(long long)r * p.hq + h * G + g] = M;
            p.out_sum[(long long)r * p.hq + h * G + g] = L;
          }
        }
      }
    }
    epi_bar_sync();  // the merge buffer is rewritten by the next item
  }
}


template <bool F8, bool TR>
__global__ void __launch_bounds__(kAttnThreads, TR ? 4 : 3)  // (transposed form: 128 registers, four CTAs = 16 warps per SM; else 168, three)
decode_attn_q8_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const AttnParams p, const float* k_scale,
                      const float* v_scale) {
  constexpr int D = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int G = p.hq / p.hkv;
  const int o_rows = G <= 8 ? 8 : 16;
  float* sm_o_all = reinterpret_cast<float*>(smem + kAttnWarps * kQ8WarpBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_o_all + kAttnWarps * o_rows * D);
  float* sm_stat = reinterpret_cast<float*>(bars + 2 * kAttnWarps);
  const int tl = timeline_begin(4);
  griddep_launch_dependents();
  if ((threadIdx.x & 31) == 0) {
    mbar_init(bars + 2 * (threadIdx.x >> 5), 1);
    mbar_init(bars + 2 * (threadIdx.x >> 5) + 1, 1);
    fence_barrier_init();
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  __syncthreads();
  griddep_wait();
  uint32_t phase = 0;
  attn_process_items_q8<F8, TR>(tm_k, tm_v, p, k_scale, v_scale, smem, sm_o_all, bars, sm_stat, phase, threadIdx.x, blockIdx.x, gridDim.x);
  timeline_end(tl);
}

__host__ inline size_t attn_q8_smem_bytes(int G) {
  const size_t o_rows = G <= 8 ? 8 : 16;
  return 1024 + size_t(kAttnWarps) * kQ8WarpBytes + kAttnWarps * o_rows * 64 * 4 + 2 * kAttnWarps * 8 + 2 * kAttnWarps * 16 * 4 + 32;
}

}  // namespace mtx
