"""Page manager (attention=paged): the product's numpy implementation against the known answers of the reference's own test
(MaxText/tests/inference/page_manager_test.py -- the cases are cited per test) and, on random operation sequences, against the
loop-for-loop restatement in oracle/paged_ref.py.  Integer logic: every comparison is exact.  CPU only."""

import numpy as np
import pytest

from maxtext_indextts2_b200 import page_manager as pm_lib
from maxtext_indextts2_b200 import pyconfig
from oracle import paged_ref

NUM_PAGES, TPP, P, T, GROUPS = 128, 8, 128, 256, 4  # page_manager_test.py:33-38
MAX_PAGES = (T + TPP - 1) // TPP


def make_pm(**kw):
  keys = dict(per_device_batch_size=GROUPS, max_prefill_predict_length=P, max_target_length=T, pagedattn_num_pages=NUM_PAGES,
              pagedattn_tokens_per_page=TPP, pagedattn_max_pages_per_group=MAX_PAGES)
  keys.update(kw)
  return pm_lib.PageManager(pyconfig.initialize(None, **keys))


def as_dict(state: pm_lib.PageState) -> dict:
  return {f: getattr(state, f).tolist() for f in ("page_status", "page_map", "num_pages_used", "sequence_lengths", "active_page",
                                                   "has_active_page", "active_page_position")}


def consistent(state: pm_lib.PageState):
  # page_manager_test.py:283-325: allocated pages == mapped pages + the reserved page 0; a group's pages are distinct
  assert int(state.page_status.sum()) == int(state.num_pages_used.sum()) + 1
  for g in range(state.page_map.shape[0]):
    used = state.page_map[g, : state.num_pages_used[g]]
    assert len(np.unique(used)) == len(used) and np.all(state.page_status[used] == 1) and np.all(used > 0)


def test_initialization():  # :94-104
  state = make_pm().get_initial_page_state()
  assert state.page_status[0] == 1 and np.all(state.page_status[1:] == 0)
  assert state.page_map.shape == (GROUPS, MAX_PAGES) and state.page_status.shape == (NUM_PAGES,)
  for f in ("num_pages_used", "sequence_lengths", "active_page", "active_page_position"):
    assert getattr(state, f).shape == (GROUPS,) and not getattr(state, f).any()
  assert not state.has_active_page.any()
  assert as_dict(state) == paged_ref.initialize_page_state(NUM_PAGES, GROUPS, MAX_PAGES)


def test_reserve_prefill_group():  # :106-149
  pm = make_pm()
  state = pm.update_prefill_pages(pm.get_initial_page_state(), 0, 12)
  assert state.sequence_lengths[0] == 12 and state.num_pages_used[0] == 2 and state.has_active_page[0]
  assert state.page_map[0, :2].tolist() == [1, 2]  # the lowest free indices >= 1 (:130-157)
  assert state.active_page[0] == 2 and state.active_page_position[0] == 12 % TPP
  consistent(state)


def test_reserve_prefill_no_space():  # :151-170
  pm = make_pm()
  full = pm.get_initial_page_state().replace(page_status=np.ones((NUM_PAGES,), dtype=np.int32))
  state = pm.update_prefill_pages(full, 0, 12)
  assert np.all(state.page_status == 1) and state.sequence_lengths[0] == 0 and state.num_pages_used[0] == 0
  assert not state.has_active_page[0]


def test_reserve_prefill_edge_cases():  # :172-191
  pm = make_pm()
  init = pm.get_initial_page_state()
  s = pm.update_prefill_pages(init, 1, 2 * TPP)
  assert (s.sequence_lengths[1], s.num_pages_used[1], bool(s.has_active_page[1]), s.active_page_position[1]) == (2 * TPP, 2, True, 0)
  s = pm.update_prefill_pages(init, 2, 5)
  assert (s.sequence_lengths[2], s.num_pages_used[2], bool(s.has_active_page[2]), s.active_page_position[2]) == (5, 1, True, 5)


def test_release_pages():  # :193-229
  pm = make_pm()
  s = pm.update_prefill_pages(pm.get_initial_page_state(), 1, 20)
  pages = s.page_map[1, : s.num_pages_used[1]].copy()
  assert len(pages) == 3
  r = pm.release_pages(s, 1)
  assert r.sequence_lengths[1] == 0 and r.num_pages_used[1] == 0 and not r.has_active_page[1]
  assert np.all(r.page_status[pages] == 0) and int(r.page_status.sum()) == 1


def test_update_decode_pages():  # :231-281
  pm = make_pm()
  init = pm.get_initial_page_state()
  assert as_dict(pm.update_decode_pages(init)) == as_dict(init)  # no active group: nothing changes
  # a sequence that exactly fills its page: the next token opens a second page
  s = pm.update_prefill_pages(init, 0, TPP)
  assert (s.num_pages_used[0], s.sequence_lengths[0], s.active_page_position[0]) == (1, TPP, 0)
  d = pm.update_decode_pages(s)
  assert d.sequence_lengths[0] == TPP + 1 and d.num_pages_used[0] == 2
  first, second = d.page_map[0, 0], d.page_map[0, 1]
  assert first != second and d.active_page[0] == second and d.active_page_position[0] == 0
  # a partial page: the token stays inside it
  s = pm.update_prefill_pages(init, 1, 5)
  d = pm.update_decode_pages(s)
  assert (d.sequence_lengths[1], d.num_pages_used[1], d.active_page[1], d.active_page_position[1]) == (6, 1, s.active_page[1], 5)


def test_group_and_length_boundaries():  # :327-355, :412-453
  pm = make_pm()
  init = pm.get_initial_page_state()
  s = pm.update_prefill_pages(init, GROUPS - 1, 1)
  assert (s.sequence_lengths[GROUPS - 1], s.num_pages_used[GROUPS - 1], bool(s.has_active_page[GROUPS - 1])) == (1, 1, True)
  s = pm.update_prefill_pages(init, 0, T)
  assert s.sequence_lengths[0] == T and s.num_pages_used[0] == MAX_PAGES
  for bad in (-1, GROUPS):
    with pytest.raises(ValueError, match="page_group_id"):
      pm.update_prefill_pages(init, bad, 1)
    with pytest.raises(ValueError, match="page_group_id"):
      pm.release_pages(init, bad)
  for bad in (-1, 0, T + 1):
    with pytest.raises(ValueError, match="true_length"):
      pm.update_prefill_pages(init, 0, bad)


def test_init_validation():  # page_manager.py:475-491
  with pytest.raises(ValueError, match="insufficient"):
    make_pm(pagedattn_max_pages_per_group=MAX_PAGES - 1)
  with pytest.raises(ValueError, match="greater than 1"):
    make_pm(pagedattn_num_pages=1)
  assert make_pm(pagedattn_max_pages_per_group=-1).max_pages_per_group == MAX_PAGES  # pyconfig.py:611-614


def test_repeated_allocation_deallocation():  # :490-541
  pm = make_pm()
  state = pm.get_initial_page_state()
  for rep in range(3):
    for g in range(GROUPS):
      state = pm.update_prefill_pages(state, g, (g + 1) * 7 + rep)
      assert state.sequence_lengths[g] == (g + 1) * 7 + rep and state.has_active_page[g]
    consistent(state)
    for g in range(GROUPS):
      state = pm.release_pages(state, g)
    assert int(state.num_pages_used.sum()) == 0 and int(state.page_status.sum()) == 1


@pytest.mark.parametrize("seed,num_pages", [(0, 128), (1, 24), (2, 12)])
def test_random_sequences_against_oracle(seed, num_pages):
  """Prefill / decode / release in random order, with pools small enough to run dry: every field equal after every operation."""
  tpp, t = 8, 64
  max_pages = t // tpp
  pm = make_pm(pagedattn_num_pages=num_pages, max_target_length=t, max_prefill_predict_length=32, pagedattn_max_pages_per_group=max_pages)
  rng = np.random.default_rng(seed)
  state, ref = pm.get_initial_page_state(), paged_ref.initialize_page_state(num_pages, GROUPS, max_pages)
  for _ in range(400):
    op = rng.integers(0, 10)
    if op < 2:
      g, n = int(rng.integers(0, GROUPS)), int(rng.integers(1, 33))
      state, ref = pm.update_prefill_pages(state, g, n), paged_ref.update_prefill_pages(ref, g, n, tpp, max_pages)
    elif op < 3:
      g = int(rng.integers(0, GROUPS))
      state, ref = pm.release_pages(state, g), paged_ref.release_pages_for_group(ref, g, max_pages)
    else:
      state, ref = pm.update_decode_pages(state), paged_ref.update_decode_pages(ref, tpp, max_pages)
    assert as_dict(state) == ref
