"""The Gemma-3 decoder block (MaxText/layers/gemma3.py:36-197) on the B200 against the CPU oracle: q/k RMSNorm + query scalar,
post-attention / post-MLP norms, GELU gate, tied logits, five sliding-window layers per global one with their own RoPE base.
Prefill (position-space window), insert, and decode steps (cache-index-space window of AUTOREGRESSIVE mode,
attentions.py:600-602,624-631), including steps after the AR ring has wrapped.  Tolerances as tests/test_engine_gpu.py."""

import numpy as np
import pytest
import torch

from maxtext_indextts2_b200 import maxengine
from oracle import decode_ref as ref
from tests.helpers import make_params, random_tokens, small_config
from tests.test_engine_gpu import _lockstep, _prefill_both

pytestmark = pytest.mark.gpu


def gemma_config(**kw):
  base = dict(
      model_name="gemma3-27b",  # head_dim 128 geometry; query scalar = (base_emb_dim // base_num_query_heads) ** -0.5
      base_num_decoder_layers=7,  # local x5, global, local
      base_emb_dim=256,
      base_num_query_heads=4,
      base_num_kv_heads=2,
      head_dim=64,
      base_mlp_dim=512,
      vocab_size=1024,
      sliding_window_size=16,
      max_prefill_predict_length=32,  # > window: the prefill segment is cut by the window
      max_target_length=96,           # ring of 64 > window: so is the ring
      per_device_batch_size=3,
      materialize_logits=True,
  )
  base.update(kw)
  return small_config(**base)


@pytest.mark.parametrize("head_dim,model", [(64, "gemma3-27b"), (128, "gemma3-27b"), (256, "gemma3-4b")])
def test_gemma3_prefill_and_decode_match_oracle(head_dim, model):
  """head_dim 256 (gemma3-1b/4b/12b; query scalar head_dim ** -0.5) runs the transposed decode attention of
  csrc/attention_wide.cuh, and its prompt positions go through the same kernel as decode rows."""
  cfg = gemma_config(head_dim=head_dim, model_name=model)
  assert cfg.decoder_block == "gemma3" and cfg.logits_via_embedding
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  engine = maxengine.MaxEngine(cfg)
  dparams = engine.load_params(params)
  prompts = random_tokens((3, 32), cfg.vocab_size, seed=5)
  ostate, state = _prefill_both(engine, dparams, oracle, oracle.init_decode_state(), engine.init_decode_state(), prompts, [32, 25, 20])  # (longer than max_prefill - window: every row's local window holds keys)
  # 80 steps: the 64-row ring wraps, and the window [48, 64) of ring indices is crossed twice
  # (atol 0.15: with the two extra RMSNorms per layer single logits of the 240 x 1024 compared leave the 1e-1 band by ~0.01)
  near = _lockstep(engine, dparams, oracle, ostate, state, steps=80, atol=0.15)
  assert near <= 8, f"{near} near-ties in 240 tokens"


def test_gemma3_window_wider_than_the_cache_is_full_attention():
  """sliding_window_size >= both segments: the local layers differ from the global one only by their RoPE base."""
  cfg = gemma_config(sliding_window_size=4096, base_num_decoder_layers=6, per_device_batch_size=2)
  params = make_params(cfg)
  f32 = ref.DecodeOracle(cfg, params, faithful=False)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=False)
  dparams = engine.load_params(params)
  prompts = random_tokens((2, 32), cfg.vocab_size, seed=6)
  ostate, state = _prefill_both(engine, dparams, f32, f32.init_decode_state(), engine.init_decode_state(), prompts, [17, 32])
  for _ in range(6):
    ostate, odata = f32.generate(ostate)
    state, result = engine.generate(dparams, state)
    want = ostate["logits"]
    got = state["logits"].cpu()
    assert (got - want).abs().max() <= 2**-5 * max(1.0, float(want.abs().max()))
    state["tokens"].copy_(odata[:, :1].to(state["tokens"].device))


@pytest.mark.parametrize("head_dim,model", [(64, "gemma3-27b"), (256, "gemma3-4b")])
def test_gemma3_chunked_prefill_with_a_window_inside_the_chunks(head_dim, model):
  """A 550-token prompt in three chunks of 256 positions with sliding_window_size 100: the position window
  (attentions.py:624-631) cuts inside and across chunks, in prefill_attn_kernel (head_dim 64) and in the decode kernel
  that runs prompt positions as rows (head_dim 256); then decode steps on the inserted prefix."""
  cfg = gemma_config(head_dim=head_dim, model_name=model, sliding_window_size=100, max_prefill_predict_length=600,
                     max_target_length=640, per_device_batch_size=2, base_num_decoder_layers=6, prefill_chunk_size=256)
  params = make_params(cfg)
  f32 = ref.DecodeOracle(cfg, params, faithful=False)
  engine = maxengine.MaxEngine(cfg)
  dparams = engine.load_params(params)
  prompts = random_tokens((2, 600), cfg.vocab_size, seed=13)
  ostate, state = f32.init_decode_state(), engine.init_decode_state()
  for slot, n in enumerate((550, 300)):
    padded = torch.zeros(600, dtype=torch.int64)
    padded[:n] = prompts[slot, :n]
    oprefix, ofirst = f32.prefill(padded, n)
    ostate = f32.insert(oprefix, ostate, slot)
    prefix, _ = engine.prefill(params=dparams, padded_tokens=padded, true_length=n)
    want, got = oprefix["logits"][0], prefix["logits"].cpu()[0]
    assert (got - want).abs().max() <= 2**-5 * max(1.0, float(want.abs().max()))
    prefix["tokens"].fill_(int(ofirst))
    state = engine.insert(prefix, state, slot)
  for _ in range(4):
    ostate, odata = f32.generate(ostate)
    state, result = engine.generate(dparams, state)
    want, got = ostate["logits"], state["logits"].cpu()
    assert (got - want).abs().max() <= 2**-5 * max(1.0, float(want.abs().max()))
    state["tokens"].copy_(odata[:, :1].to(state["tokens"].device))
