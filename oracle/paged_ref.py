"""ORACLE -- CPU restatement of the reference's paged KV cache (``attention=paged``).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/`` may import it; the product package never does.

What it restates (file:line relative to the reference tree HyperBlaze456/maxtext-indextts2):

* page bookkeeping   MaxText/inference/page_manager.py: ``initialize_page_state`` :93-127,
                     ``_find_next_free_page_index`` :130-157, ``_release_pages_for_group`` :160-206,
                     ``_reserve_pages_for_group`` :209-312, ``_update_decode_pages_global`` :332-412 -- written as the
                     same loops, one page / one group at a time (the product's page_manager.py is vectorised numpy);
* page writes        MaxText/inference/paged_attention.py ``update_decode_step_pages`` :446-471,
                     ``update_prefill_step_pages`` :403-444, MaxText/maxengine.py ``_copy_paged`` :1104-1131;
* paged decode attention  ``paged_attention_v1_decode`` :302-346.  The kernel it calls is third-party
                     (jax.experimental.pallas.ops.tpu.paged_attention, jax==0.6.2 per constraints_gpu.txt:88): its
                     published contract is, per sequence, softmax(q . K[:length]^T) V[:length] over the tokens of the
                     sequence's pages in page_indices order, with q used as given (MaxText folds 1/sqrt(D) into the query
                     kernel, attentions.py:1900-1904).  Restated here in fp32 on the bf16 inputs.

PARITY STATUS.  The bookkeeping is pinned by the known-answer cases of the reference's own test
(MaxText/tests/inference/page_manager_test.py, restated in tests/test_page_manager.py).  The attention has no
reference-held vectors and the reference cannot run here (no jax): parity unpinned, like oracle/decode_ref.py.
"""

from __future__ import annotations

import copy

import torch


# ---- page bookkeeping (plain Python lists / ints) ------------------------------------------------------------------------

def initialize_page_state(num_pages: int, max_page_groups: int, max_pages_per_group: int) -> dict:
  status = [0] * num_pages
  status[0] = 1  # page_manager.py:113-115
  return {
      "page_status": status,
      "page_map": [[0] * max_pages_per_group for _ in range(max_page_groups)],
      "num_pages_used": [0] * max_page_groups,
      "sequence_lengths": [0] * max_page_groups,
      "active_page": [0] * max_page_groups,
      "has_active_page": [False] * max_page_groups,
      "active_page_position": [0] * max_page_groups,
  }


def find_next_free_page_index(status: list) -> int:
  """:130-157: lowest index >= 1 whose status is 0, else -1."""
  for i in range(1, len(status)):
    if status[i] == 0:
      return i
  return -1


def release_pages_for_group(state: dict, group: int, max_pages_per_group: int) -> dict:
  s = copy.deepcopy(state)
  valid = state["num_pages_used"][group]
  for i in range(max_pages_per_group):
    page = state["page_map"][group][i]
    if i < valid and page > 0:
      s["page_status"][page] = 0
  s["num_pages_used"][group] = 0
  s["sequence_lengths"][group] = 0
  s["active_page"][group] = 0
  s["has_active_page"][group] = False
  s["active_page_position"][group] = 0
  return s


def reserve_pages_for_group(state: dict, group: int, true_length: int, tokens_per_page: int, max_pages_per_group: int) -> dict:
  needed = (true_length + tokens_per_page - 1) // tokens_per_page
  last_pos = (true_length - 1) % tokens_per_page
  next_write = (last_pos + 1) % tokens_per_page
  free = sum(1 for x in state["page_status"] if x == 0)
  if not (free >= needed and needed <= max_pages_per_group):
    return state
  s = copy.deepcopy(state)
  for i in range(needed):
    page = find_next_free_page_index(s["page_status"])
    if page >= 0:
      s["page_status"][page] = 1
      s["page_map"][group][i] = page
      s["num_pages_used"][group] += 1
  s["sequence_lengths"][group] = true_length
  s["active_page"][group] = s["page_map"][group][needed - 1]
  s["has_active_page"][group] = True
  s["active_page_position"][group] = next_write
  return s


def update_prefill_pages(state: dict, group: int, true_length: int, tokens_per_page: int, max_pages_per_group: int) -> dict:
  """:315-329."""
  return reserve_pages_for_group(release_pages_for_group(state, group, max_pages_per_group), group, true_length, tokens_per_page,
                                 max_pages_per_group)


def update_decode_pages(state: dict, tokens_per_page: int, max_pages_per_group: int) -> dict:
  """:332-412."""
  groups = len(state["sequence_lengths"])
  s = copy.deepcopy(state)
  needs = [False] * groups
  for g in range(groups):
    active = state["has_active_page"][g]
    s["sequence_lengths"][g] = state["sequence_lengths"][g] + (1 if active else 0)
    if active:
      s["active_page_position"][g] = (s["sequence_lengths"][g] - 1) % tokens_per_page
    required = (s["sequence_lengths"][g] + tokens_per_page - 1) // tokens_per_page
    needs[g] = active and required > state["num_pages_used"][g] and required <= max_pages_per_group
  for g in range(groups):
    page = find_next_free_page_index(s["page_status"])
    if needs[g] and page >= 0:
      s["page_status"][page] = 1
      s["page_map"][g][s["num_pages_used"][g]] = page
      s["num_pages_used"][g] += 1
      s["active_page"][g] = page
  return s


# ---- page pools -----------------------------------------------------------------------------------------------------------

def update_decode_step_pages(k_pages: torch.Tensor, v_pages: torch.Tensor, key: torch.Tensor, value: torch.Tensor, state: dict):
  """paged_attention.py:446-471.  pools [Hkv, num_pages, tokens_per_page, D]; key / value [B, Hkv, D]; in place."""
  for b in range(key.shape[0]):
    page, pos = state["active_page"][b], state["active_page_position"][b]
    k_pages[:, page, pos] = key[b]
    v_pages[:, page, pos] = value[b]


def copy_prefix_pages(k_pages: torch.Tensor, v_pages: torch.Tensor, k_prefix: torch.Tensor, v_prefix: torch.Tensor, state: dict, slot: int):
  """maxengine.py:1104-1131: prefix pages [Hkv, n, tokens_per_page, D] -> pool pages page_map[slot][i], i < num_pages_used."""
  for i in range(state["num_pages_used"][slot]):
    dst = state["page_map"][slot][i]
    k_pages[:, dst] = k_prefix[:, i]
    v_pages[:, dst] = v_prefix[:, i]


def paged_attention(q: torch.Tensor, k_pages: torch.Tensor, v_pages: torch.Tensor, lengths, page_map, softcap: float = 0.0) -> torch.Tensor:
  """q [B, Hq, D]; pools [Hkv, num_pages, tokens_per_page, D]; lengths [B]; page_map [B, pages_per_sequence] -> [B, Hq, D] fp32
  (rows of length 0: zeros)."""
  B, Hq, D = q.shape
  Hkv, _, tpp, _ = k_pages.shape
  G = Hq // Hkv
  out = torch.zeros(B, Hq, D, dtype=torch.float32)
  for b in range(B):
    n = int(lengths[b])
    if n == 0:
      continue
    pages = [int(page_map[b][i]) for i in range((n + tpp - 1) // tpp)]
    k = k_pages[:, pages].reshape(Hkv, -1, D)[:, :n].float()
    v = v_pages[:, pages].reshape(Hkv, -1, D)[:, :n].float()
    qb = q[b].float().reshape(Hkv, G, D)
    s = torch.einsum("kgd,ksd->kgs", qb, k)
    if softcap:
      s = torch.tanh(s / softcap) * softcap
    p = torch.softmax(s, dim=-1)
    out[b] = torch.einsum("kgs,ksd->kgd", p, v).reshape(Hq, D)
  return out
