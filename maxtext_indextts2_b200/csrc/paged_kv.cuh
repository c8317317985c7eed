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

}  // namespace mtx
