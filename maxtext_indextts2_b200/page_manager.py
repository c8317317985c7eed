"""Page manager of the paged KV cache (``attention=paged``), the host side.

Mirrors ``MaxText/inference/page_manager.py``: ``PageState`` (:49-91), ``initialize_page_state`` (:93-127) and
``PageManager`` (:415-621) with the same method names, arguments, error messages and allocation order (the lowest free
page index >= 1 first; page 0 is never handed out, :113-115).  The reference keeps the state in jnp arrays and updates
it with jitted loops; the state is a few hundred integers, so here it is numpy on the host, updated in place-free style
(every method returns a new ``PageState``) and uploaded to the device arrays a decode step reads
(``mtx_decode_state.page_map`` etc.) by ``MaxEngine``.  Integer logic: results are bit-exact with the reference's.
"""

from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass(frozen=True)
class PageState:
  """page_manager.py:49-91.  All arrays int32 (``has_active_page`` bool)."""

  page_status: np.ndarray           # [num_pages] 0 free / 1 allocated
  page_map: np.ndarray              # [max_page_groups, max_pages_per_group] global page index of the group's i-th page
  num_pages_used: np.ndarray        # [max_page_groups]
  sequence_lengths: np.ndarray      # [max_page_groups]
  active_page: np.ndarray           # [max_page_groups]
  has_active_page: np.ndarray       # [max_page_groups] bool
  active_page_position: np.ndarray  # [max_page_groups]

  def replace(self, **changes) -> "PageState":
    return dataclasses.replace(self, **changes)


def initialize_page_state(num_pages: int, max_page_groups: int, max_pages_per_group: int) -> PageState:
  """page_manager.py:93-127: everything zero, page 0 marked used."""
  status = np.zeros((num_pages,), dtype=np.int32)
  status[0] = 1
  z = lambda: np.zeros((max_page_groups,), dtype=np.int32)
  return PageState(
      page_status=status,
      page_map=np.zeros((max_page_groups, max_pages_per_group), dtype=np.int32),
      num_pages_used=z(),
      sequence_lengths=z(),
      active_page=z(),
      has_active_page=np.zeros((max_page_groups,), dtype=bool),
      active_page_position=z(),
  )


def _release_pages_for_group(state: PageState, group: int) -> PageState:
  """page_manager.py:160-206: free the group's pages (entries with index > 0 only) and clear its fields."""
  status = state.page_status.copy()
  pages = state.page_map[group, : int(state.num_pages_used[group])]
  status[pages[pages > 0]] = 0

  def cleared(a, value=0):
    a = a.copy()
    a[group] = value
    return a

  return state.replace(
      page_status=status,
      num_pages_used=cleared(state.num_pages_used),
      sequence_lengths=cleared(state.sequence_lengths),
      active_page=cleared(state.active_page),
      has_active_page=cleared(state.has_active_page, False),
      active_page_position=cleared(state.active_page_position),
  )


def _reserve_pages_for_group(state: PageState, group: int, true_length: int, tokens_per_page: int, max_pages_per_group: int) -> PageState:
  """page_manager.py:209-312: the ceil(true_length / tokens_per_page) lowest free pages, or nothing at all."""
  needed = (true_length + tokens_per_page - 1) // tokens_per_page
  free = np.flatnonzero(state.page_status[1:] == 0) + 1  # ascending: the order _find_next_free_page_index hands them out
  # (the reference counts free pages over the whole array; page 0 is always marked used, so the counts agree)
  if int(np.sum(state.page_status == 0)) < needed or needed > max_pages_per_group:
    return state
  take = free[:needed].astype(np.int32)
  status = state.page_status.copy()
  status[take] = 1
  page_map = state.page_map.copy()
  page_map[group, :needed] = take
  used = state.num_pages_used.copy()
  used[group] += needed

  def with_value(a, value):
    a = a.copy()
    a[group] = value
    return a

  return state.replace(
      page_status=status,
      page_map=page_map,
      num_pages_used=used,
      sequence_lengths=with_value(state.sequence_lengths, true_length),
      active_page=with_value(state.active_page, page_map[group, needed - 1]),
      has_active_page=with_value(state.has_active_page, True),
      active_page_position=with_value(state.active_page_position, true_length % tokens_per_page),
  )


def _update_decode_pages_global(state: PageState, tokens_per_page: int, max_pages_per_group: int) -> PageState:
  """page_manager.py:332-412: one more token for every active group; groups that crossed a page boundary get the lowest
  free page, in group order, while free pages last."""
  active = state.has_active_page
  lengths = state.sequence_lengths + active  # (int32 + bool stays int32)
  position = np.where(active, (lengths - 1) % tokens_per_page, state.active_page_position)
  required = (lengths + (tokens_per_page - 1)) // tokens_per_page
  needs = active & (required > state.num_pages_used) & (required <= max_pages_per_group)
  status, page_map = state.page_status, state.page_map
  used, active_page = state.num_pages_used, state.active_page
  if needs.any():
    groups = np.flatnonzero(needs)
    free = np.flatnonzero(status[1:] == 0) + 1
    n = min(groups.size, free.size)  # later groups find no free page and keep their state (can_allocate false)
    groups, take = groups[:n], free[:n].astype(np.int32)
    status, page_map, used, active_page = status.copy(), page_map.copy(), used.copy(), active_page.copy()
    status[take] = 1
    page_map[groups, used[groups]] = take
    used[groups] += 1
    active_page[groups] = take
  # (the common step hands out no page: page_status / page_map / num_pages_used / active_page are then the SAME arrays)
  return PageState(status, page_map, used, lengths, active_page, active, position)


class PageManager:
  """page_manager.py:415-621."""

  def __init__(self, config):
    self.num_pages: int = int(config.pagedattn_num_pages)
    self.tokens_per_page: int = int(config.pagedattn_tokens_per_page)
    self.max_target_length: int = int(config.max_target_length)
    # global_batch_size_to_load in the reference (:471); one process per GPU holds per_device_batch_size page groups
    self.max_page_groups: int = int(config.per_device_batch_size)
    self.max_pages_per_group: int = int(config.pagedattn_max_pages_per_group)
    self._validate_init_params()

  def _validate_init_params(self) -> None:
    if self.max_pages_per_group <= 0:
      raise ValueError("`pagedattn_max_pages_per_group` must be positive.")
    min_required = (self.max_target_length + self.tokens_per_page - 1) // self.tokens_per_page
    if self.max_pages_per_group < min_required:
      raise ValueError(
          f"`pagedattn_max_pages_per_group` ({self.max_pages_per_group}) is insufficient for `max_target_length` "
          f"({self.max_target_length}). Needs {min_required}."
      )
    if self.num_pages <= 1:
      raise ValueError("`pagedattn_num_pages` must be greater than 1.")
    if self.tokens_per_page <= 0:
      raise ValueError("`pagedattn_tokens_per_page` must be positive.")
    if self.max_page_groups <= 0:
      raise ValueError("`pagedattn_max_page_groups` must be positive.")

  def update_prefill_pages(self, page_state: PageState, page_group_id: int, true_length: int) -> PageState:
    """:493-536: release the group's pages, then reserve the pages of a `true_length`-token sequence (or none)."""
    if page_group_id < 0 or page_group_id >= self.max_page_groups:
      raise ValueError(f"PageManager: page_group_id ({page_group_id}) out of range [0, {self.max_page_groups})")
    if true_length <= 0 or true_length > self.max_target_length:
      raise ValueError(f"PageManager: true_length ({true_length}) out of range (0, {self.max_target_length}]")
    released = _release_pages_for_group(page_state, page_group_id)
    return _reserve_pages_for_group(released, page_group_id, int(true_length), self.tokens_per_page, self.max_pages_per_group)

  def update_decode_pages(self, page_state: PageState) -> PageState:
    """:538-563."""
    return _update_decode_pages_global(page_state, self.tokens_per_page, self.max_pages_per_group)

  def release_pages(self, page_state: PageState, page_group_id: int) -> PageState:
    """:565-598."""
    if page_group_id < 0 or page_group_id >= self.max_page_groups:
      raise ValueError(f"PageManager: page_group_id ({page_group_id}) out of range [0, {self.max_page_groups})")
    return _release_pages_for_group(page_state, page_group_id)

  def get_initial_page_state(self) -> PageState:
    """:600-621."""
    return initialize_page_state(self.num_pages, self.max_page_groups, self.max_pages_per_group)
