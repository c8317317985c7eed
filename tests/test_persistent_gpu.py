"""The persistent step kernel (csrc/step_persistent.cuh) against the per-kernel path and the CPU oracle.

Both GPU paths implement the same arithmetic (same bf16 rounding points); they differ in the order of the
fp32 partial sums of the split-K GEMMs and of the attention partials, so the bf16-rounded logits agree to
about two bf16 ulps (stated: atol 5e-2 + rtol 2e-2 on logits of magnitude <= ~6, where one ulp is 0.016-0.031)
and greedy tokens agree except at near-ties.
"""

import os

import numpy as np
import pytest
import torch

from maxtext_indextts2_b200 import maxengine, pyconfig
from oracle import decode_ref as ref
from tests.helpers import make_params, random_tokens

pytestmark = pytest.mark.gpu


def _mid_config(batch, **kw):
  base = dict(
      base_num_decoder_layers=3, base_emb_dim=384, base_num_query_heads=6, base_num_kv_heads=2, head_dim=64, base_mlp_dim=768,
      vocab_size=5000, per_device_batch_size=batch, max_prefill_predict_length=64, max_target_length=192,
      weight_dtype="bfloat16", attention="dot_product", scan_layers=False, materialize_logits=True)
  base.update(kw)
  return pyconfig.initialize(None, **base)


def _engine(cfg, persistent, graph=False):
  """MTX_PERSISTENT is read when the engine binds its buffers."""
  old = os.environ.get("MTX_PERSISTENT")
  os.environ["MTX_PERSISTENT"] = "1" if persistent else "0"
  try:
    engine = maxengine.MaxEngine(cfg, use_cuda_graph=graph)
    dparams = engine.load_params(make_params(cfg))
  finally:
    if old is None:
      del os.environ["MTX_PERSISTENT"]
    else:
      os.environ["MTX_PERSISTENT"] = old
  return engine, dparams


def _launches_per_step(engine, dparams, state):
  n0 = engine.lib.mtx_launch_count()
  state, _ = engine.generate(dparams, state)
  torch.cuda.synchronize()
  return int(engine.lib.mtx_launch_count() - n0), state


def _ragged(batch, P, R, seed):
  rng = np.random.Generator(np.random.PCG64(seed))
  pl = rng.integers(1, P + 1, size=batch)
  al = rng.integers(0, R - 8, size=batch)
  al[pl < P] = 0  # a slot decodes only after its whole prompt
  return pl, al


@pytest.mark.parametrize("batch", [1, 7, 33, 64])
def test_persistent_matches_per_kernel_path(batch):
  cfg = _mid_config(batch)
  P, R = cfg.max_prefill_predict_length, cfg.max_target_length - cfg.max_prefill_predict_length
  pl, al = _ragged(batch, P, R, seed=batch)
  def run(persistent, forced=None):
    engine, dparams = _engine(cfg, persistent)
    state = engine.fill_synthetic_context(pl, al, seed=11)
    logits, tokens = [], []
    for step in range(6):
      if step == 0:
        n, state = _launches_per_step(engine, dparams, state)
        assert (n == 3) == persistent, f"{n} launches per step with persistent={persistent}"
      else:
        state, _ = engine.generate(dparams, state)
      logits.append(state["logits"].float().cpu().clone())
      tokens.append(state["tokens"].cpu().clone())
      if forced is not None:  # teacher-force the other run's tokens so both see the same history
        state["tokens"].copy_(forced[step])
    return logits, tokens

  first = run(True)
  results = [first, run(False, forced=first[1])]
  for step, (a, b) in enumerate(zip(results[0][0], results[1][0])):
    torch.testing.assert_close(a, b, rtol=2e-2, atol=5e-2, msg=lambda m: f"step {step}: {m}")
  # greedy tokens: identical except where the top-2 margin is within the logit tolerance
  for step, (ta, tb) in enumerate(zip(results[0][1], results[1][1])):
    for r in np.nonzero((ta != tb).numpy().reshape(-1))[0]:
      row = results[1][0][step][r, 0]
      assert abs(row[int(ta[r])] - row[int(tb[r])]) <= 1e-1, f"step {step} row {r}: tokens differ beyond a near-tie"


def test_persistent_is_deterministic():
  cfg = _mid_config(64)
  P, R = cfg.max_prefill_predict_length, cfg.max_target_length - cfg.max_prefill_predict_length
  pl, al = _ragged(64, P, R, seed=3)
  runs = []
  for _ in range(2):
    engine, dparams = _engine(cfg, True, graph=True)
    state = engine.fill_synthetic_context(pl, al, seed=5)
    out = []
    for _ in range(6):
      state, result = engine.generate(dparams, state)
      out.append((state["logits"].cpu().clone(), result.data.cpu().clone()))
    runs.append(out)
    del engine
  for (la, ta), (lb, tb) in zip(*runs):
    assert torch.equal(la, lb) and torch.equal(ta, tb)  # fixed summation orders: bit-identical


def test_persistent_full_batch_against_the_oracle():
  """64 slots with ragged prompts, decoded past the ring wrap of the slots that joined first."""
  cfg = _mid_config(64, base_num_decoder_layers=2, max_prefill_predict_length=32, max_target_length=48, vocab_size=2000)
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=False)  # eager, so that every launch is counted
  dparams = engine.load_params(params)
  prompts = random_tokens((64, 32), cfg.vocab_size, seed=21)
  rng = np.random.Generator(np.random.PCG64(2))
  lengths = rng.integers(1, 33, size=64)
  ostate, state = oracle.init_decode_state(), engine.init_decode_state()
  for slot in range(64):
    n = int(lengths[slot])
    padded = torch.zeros(oracle.P, dtype=torch.int64)
    padded[:n] = prompts[slot, :n]
    oprefix, ofirst = oracle.prefill(padded, n)
    ostate = oracle.insert(oprefix, ostate, slot)
    prefix, _ = engine.prefill(params=dparams, padded_tokens=padded, true_length=n)
    prefix["tokens"].fill_(int(ofirst))
    state = engine.insert(prefix, state, slot)
  n0 = engine.lib.mtx_launch_count()
  for step in range(20):  # ring of 16 rows: wraps
    ostate, odata = oracle.generate(ostate)
    state, result = engine.generate(dparams, state)
    torch.testing.assert_close(state["logits"].cpu(), ostate["logits"], rtol=1e-1, atol=1e-1)
    state["tokens"].copy_(odata[:, :1])
  torch.cuda.synchronize()
  assert engine.lib.mtx_launch_count() - n0 == 3 * 20  # prepare + persistent step + finalize


def test_wide_head_groups_fall_back_to_the_per_kernel_path():
  """More than 8 query heads per kv head do not fit the persistent kernel's attention MMA: the engine must
  switch to the per-kernel path on its own, with the same results contract."""
  cfg = _mid_config(2, base_num_query_heads=12, base_num_kv_heads=1, base_emb_dim=256, base_mlp_dim=512, base_num_decoder_layers=2)
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=False)
  dparams = engine.load_params(params)
  prompts = random_tokens((2, 64), cfg.vocab_size, seed=4)
  ostate, state = oracle.init_decode_state(), engine.init_decode_state()
  for slot, n in enumerate((9, 40)):
    padded = torch.zeros(oracle.P, dtype=torch.int64)
    padded[:n] = prompts[slot, :n]
    oprefix, ofirst = oracle.prefill(padded, n)
    ostate = oracle.insert(oprefix, ostate, slot)
    prefix, _ = engine.prefill(params=dparams, padded_tokens=padded, true_length=n)
    prefix["tokens"].fill_(int(ofirst))
    state = engine.insert(prefix, state, slot)
  n, state = _launches_per_step(engine, dparams, state)
  assert n > 3
  ostate, odata = oracle.generate(ostate)
  torch.testing.assert_close(state["logits"].cpu(), ostate["logits"], rtol=1e-1, atol=1e-1)


@pytest.mark.parametrize("batch,heads", [(2, (6, 2)), (24, (6, 2)), (3, (8, 1)), (5, (10, 2))])
def test_persistent_long_contexts_cross_cta_merges(batch, heads):
  """Long contexts with few slots: a (row, kv head) pair spans many CTAs and warps, so the attention phase runs
  its cross-CTA merge plan with tens of remote parts per pair (several rounds of four), pairs that start and end
  inside one CTA, and -- with 8 query heads per kv head -- one merge job at a time.  Checked against the
  per-kernel path on the same synthetic cache."""
  hq, hkv = heads
  cfg = _mid_config(batch, base_num_query_heads=hq, base_num_kv_heads=hkv, base_emb_dim=hq * 64, max_prefill_predict_length=1024,
                    max_target_length=2560, base_num_decoder_layers=2, vocab_size=3000)
  rng = np.random.Generator(np.random.PCG64(100 + batch))
  pl = rng.integers(700, 1025, size=batch)
  al = rng.integers(0, 1500, size=batch)
  al[pl < 1024] = 0
  pl[0], al[0] = 1024, 1500  # the longest possible pair: 40 tiles

  def run(persistent, forced=None):
    engine, dparams = _engine(cfg, persistent)
    state = engine.fill_synthetic_context(pl, al, seed=13)
    logits, tokens = [], []
    for step in range(3):
      n, state = _launches_per_step(engine, dparams, state)
      assert (n == 3) == persistent
      logits.append(state["logits"].float().cpu().clone())
      tokens.append(state["tokens"].cpu().clone())
      if forced is not None:
        state["tokens"].copy_(forced[step])
    return logits, tokens

  a = run(True)
  b = run(False, forced=a[1])
  for step, (la, lb) in enumerate(zip(a[0], b[0])):
    torch.testing.assert_close(la, lb, rtol=2e-2, atol=5e-2, msg=lambda m: f"step {step}: {m}")


@pytest.mark.parametrize("batch", [4, 64])
def test_persistent_at_indextts2_scale_against_the_per_kernel_path(batch):
  """The work tables of the real model (24 layers, emb 1280, mlp 5120: MLP up as 80 owned tiles + 68 helper CTAs,
  8- and 14-way split-K elsewhere) against the per-kernel path on a synthetic cache.  24 layers deep, two correct
  summation orders differ by ~0.02 on average and up to ~0.15 on single logits of magnitude ~4 (bf16 ulp 0.03):
  stated tolerance atol = 0.25 on every logit, mean |d| <= 0.04."""
  cfg = pyconfig.initialize(None, model_name="indextts2-t2s", per_device_batch_size=batch, max_prefill_predict_length=256,
                            max_target_length=512, materialize_logits=True, vocab_size=20480)
  pl, al = _ragged(batch, 256, 256, seed=batch)

  def run(persistent, forced=None):
    engine, dparams = _engine(cfg, persistent)
    state = engine.fill_synthetic_context(pl, al, seed=11)
    logits, tokens = [], []
    for step in range(2):
      n, state = _launches_per_step(engine, dparams, state)
      assert (n == 3) == persistent
      logits.append(state["logits"].float().cpu().clone())
      tokens.append(state["tokens"].cpu().clone())
      if forced is not None:
        state["tokens"].copy_(forced[step])
    return logits, tokens

  a = run(True)
  b = run(False, forced=a[1])
  for la, lb in zip(a[0], b[0]):
    d = (la - lb).abs()
    assert d.max() <= 0.25 and d.mean() <= 0.04, f"max {d.max():.3f} mean {d.mean():.4f}"
