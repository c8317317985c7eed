// Paged KV cache (attention=paged): the writes into the page pools.
//
// Reference: MaxText/inference/paged_attention.py (PagedAttentionOp: pools `key_pages` / `value_pages` of shape
// [num_kv_heads, num_pages, tokens_per_page, head_dim], :152-160), MaxText/inference/page_manager.py (PageState :49-91),
// MaxText/maxengine.py:1104-1131 (insert of a prefix's pages).  The reads are in attention.cuh (AttnParams.page_map).
//
// Here one pool holds all layers: [L, Hkv, num_pages, tokens_per_page, D] bf16.
#pragma once

#include "common.cuh"

namespace mtx {

// update_decode_step_pages (paged_attention.py:446-471): the step's key / value of EVERY row goes to
// pages[h, active_page[g], active_page_position[g]] of the row's page group g -- also for groups without an active page, whose
// (0, 0) entry is the never-allocated page 0 (page_manager.py:113-115).  k_new / v_new: [rows, Hkv, D] (the rotated keys and the
// values the QKV epilogue left).  One CTA per row, 16 bytes per thread.
struct PagedAppendArgs {
  const bf16 *k_new, *v_new;
  bf16 *k_pages, *v_pages;  // this layer's pools [Hkv, num_pages, tokens_per_page, D]
  const int *group;         // [rows] page group of the row, or null: row r is group r
  const int *active_page, *active_pos;
  int hkv, d, num_pages, tokens_per_page;
};

__global__ void paged_append_kernel(const PagedAppendArgs a) {
  griddep_launch_dependents();
  griddep_wait();
  const int r = blockIdx.x;
  const int g = a.group != nullptr ? a.group[r] : r;
  const int page = a.active_page[g], pos = a.active_pos[g];
  const int vec_per_head = a.d / 8;
  for (int i = threadIdx.x; i < a.hkv * vec_per_head; i += blockDim.x) {
    const int h = i / vec_per_head, v = i - h * vec_per_head;
    const long long src = ((long long)r * a.hkv + h) * a.d + v * 8;
    const long long dst = (((long long)h * a.num_pages + page) * a.tokens_per_page + pos) * a.d + v * 8;
    *reinterpret_cast<uint4*>(a.k_pages + dst) = *reinterpret_cast<const uint4*>(a.k_new + src);
    *reinterpret_cast<uint4*>(a.v_pages + dst) = *reinterpret_cast<const uint4*>(a.v_new + src);
  }
}

// The prefix of one sequence into the pages of its group (maxengine.py:1104-1131 `_copy_paged`: prefix page i of every kv head
// goes to pool page page_map[slot][i], i < num_pages_used[slot]).  The prefix is [L, Hkv, src_rows, D] -- prefill's cache rows,
// i.e. the reference's prefix pages [Hkv, P / tokens_per_page, tokens_per_page, D] read as [Hkv, P, D] -- of which the first
// n_tokens rows are copied (whole 16-byte vectors; the tail of the last page keeps what it held: positions past the sequence
// length are never read).  Grid-stride over (layer, head, token, vector).
struct PagedInsertArgs {
  const bf16 *k_src, *v_src;
  bf16 *k_pages, *v_pages;  // [L, Hkv, num_pages, tokens_per_page, D]
  const int* page_map_row;  // [max_pages] of the destination group
  int layers, hkv, d, src_rows, n_tokens, num_pages, tokens_per_page;
};

__global__ void paged_insert_kernel(const PagedInsertArgs a) {
  const int vec = a.d / 8;
  const long long total = (long long)a.layers * a.hkv * a.n_tokens * vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = int(i % vec);
    long long q = i / vec;
    const int tok = int(q % a.n_tokens);
    q /= a.n_tokens;
    const int h = int(q % a.hkv), l = int(q / a.hkv);
    const int page = a.page_map_row[tok / a.tokens_per_page], pos = tok % a.tokens_per_page;
    const long long src = (((long long)l * a.hkv + h) * a.src_rows + tok) * a.d + v * 8;
    const long long dst = ((((long long)l * a.hkv + h) * a.num_pages + page) * a.tokens_per_page + pos) * a.d + v * 8;
    *reinterpret_cast<uint4*>(a.k_pages + dst) = __ldg(reinterpret_cast<const uint4*>(a.k_src + src));
    *reinterpret_cast<uint4*>(a.v_pages + dst) = __ldg(reinterpret_cast<const uint4*>(a.v_src + src));
  }
}

// PageManager.update_decode_pages on the device (page_manager.py:332-412 `_update_decode_pages_global`; the reference runs it as a
// jitted function on device arrays too): every active group gets one more token; the groups that crossed a page boundary get, in
// group order, the lowest free pages (index >= 1) while free pages last.  ONE CTA of 256 threads, all of them calling; at most
// 256 groups.  The arrays are PageState's (has_active_page as int32 0 / 1) and are updated in place.
struct PageStateDev {
  int* page_status;           // [num_pages]
  int* page_map;              // [groups, max_pages_per_group]
  int* num_pages_used;        // [groups]
  int* sequence_lengths;      // [groups]
  int* active_page;           // [groups]
  const int* has_active_page; // [groups]
  int* active_page_position;  // [groups]
  int num_pages, groups, max_pages_per_group, tokens_per_page;
};

__device__ __forceinline__ int page_block_exclusive_scan(int v, int* s_warp, int& total) {
  // exclusive prefix sum over the 256 threads of the CTA (8 warps); s_warp: 9 ints of shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // (s_warp may still be read from a previous scan)
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += s_warp[w];
  total = 0;
  for (int w = 0; w < 8; ++w) total += s_warp[w];
  return base + inc - v;
}

__device__ __forceinline__ void page_update_decode(const PageStateDev& a, int* s_warp, int* s_free) {
  const int g = threadIdx.x;
  const int tpp = a.tokens_per_page;
  bool need = false;
  int used = 0;
  if (g < a.groups) {
    const int active = a.has_active_page[g] != 0 ? 1 : 0;
    const int len = a.sequence_lengths[g] + active;
    used = a.num_pages_used[g];
    const int required = (len + tpp - 1) / tpp;
    need = active && required > used && required <= a.max_pages_per_group;
    a.sequence_lengths[g] = len;
    if (active) a.active_page_position[g] = (len - 1) % tpp;
  }
  int n_need;
  const int rank = page_block_exclusive_scan(need ? 1 : 0, s_warp, n_need);
  if (n_need == 0) return;  // (uniform: every thread holds the same total)
  // the first n_need free pages, ascending: every thread owns a contiguous run of pages [lo, hi) of [1, num_pages)
  const int per = (a.num_pages - 1 + 255) / 256;
  const int lo = 1 + threadIdx.x * per, hi = min(a.num_pages, lo + per);
  int n_free = 0;
  for (int pg = lo; pg < hi; ++pg) n_free += a.page_status[pg] == 0 ? 1 : 0;
  int total_free;
  int k = page_block_exclusive_scan(n_free, s_warp, total_free);
  for (int pg = lo; pg < hi && k < n_need; ++pg)
    if (a.page_status[pg] == 0) s_free[k++] = pg;
  __syncthreads();
  if (need && rank < total_free) {  // later groups find no free page and keep their state
    const int pg = s_free[rank];
    a.page_status[pg] = 1;
    a.page_map[(long long)g * a.max_pages_per_group + used] = pg;
    a.num_pages_used[g] = used + 1;
    a.active_page[g] = pg;
  }
}

__global__ void __launch_bounds__(256) page_update_decode_kernel(const PageStateDev a) {
  __shared__ int s_warp[9];
  __shared__ int s_free[256];
  page_update_decode(a, s_warp, s_free);
}

}  // namespace mtx
