"""Pins the CPU oracle with the reference's own invariants for the decode path.

The reference stores no golden vectors for this path; its tests compare
the model with itself.  Each test here restates one of them against ``oracle/``:

* AR steps == full-sequence forward    MaxText/tests/attention_test.py:361-406 (1e-2),
                                       MaxText/tests/model_test.py:119-191 (1e-1)
* RoPE vs complex-number rotation      MaxText/tests/llama_test.py:79-106 (rtol 1e-1, atol 1e-4)
* RoPE / scale commutation             MaxText/tests/llama_test.py:108-127
* GQA decode vs reference_gqa          MaxText/tests/kernels_test.py:90-110
* prefill / insert / generate shapes   MaxText/tests/maxengine_test.py:111-164
* Philox-4x32-10 known-answer vectors  (Random123 kat_vectors; pins the sampler's bit stream)
"""

import copy
import math

import numpy as np
import pytest
import torch

from oracle import decode_ref as ref
from tests.helpers import make_params, random_tokens, small_config


def _teacher_forced_ar_logits(oracle, tokens, prefill_len):
  """prefill(prefill_len) + AR steps with the true next tokens; returns [B, T-prefill_len, V]."""
  B, T = tokens.shape
  state = oracle.init_decode_state()
  for b in range(B):
    padded = torch.zeros(oracle.P, dtype=torch.int64)
    padded[:prefill_len] = tokens[b, :prefill_len]
    prefix, _ = oracle.prefill(padded, prefill_len)
    state = oracle.insert(prefix, state, b)
  outs = []
  for t in range(prefill_len, T):
    state["tokens"] = tokens[:, t : t + 1].to(torch.int32)
    state, _ = oracle.generate(state)
    outs.append(state["logits"][:, 0])
  return torch.stack(outs, dim=1)


@pytest.mark.parametrize("faithful,tol", [(False, 2e-4), (True, 1e-1)])
def test_autoregression_matches_full_forward(faithful, tol):
  cfg = small_config()
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=faithful)
  T = 20  # 4 prefill + 16 AR steps = exactly the ring length (no overwrite)
  tokens = random_tokens((2, T), cfg.vocab_size)
  full = oracle.forward_full(tokens)
  ar = _teacher_forced_ar_logits(oracle, tokens, prefill_len=4)
  torch.testing.assert_close(ar, full[:, 4:], rtol=tol, atol=tol)


def test_autoregression_with_ring_wrap_and_late_insert():
  """A slot inserted later has its AR rows at a ring offset (maxengine.py:1060-1067)."""
  cfg = small_config(per_device_batch_size=2, max_prefill_predict_length=8, max_target_length=20)
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=False)
  tokens = random_tokens((2, 18), cfg.vocab_size, seed=5)
  full = oracle.forward_full(tokens)
  state = oracle.init_decode_state()
  padded = torch.zeros(oracle.P, dtype=torch.int64)
  padded[:3] = tokens[0, :3]
  prefix, _ = oracle.prefill(padded, 3)
  state = oracle.insert(prefix, state, 0)
  # slot 0 decodes alone for 5 steps (slot 1 runs on garbage, as in the reference)
  for t in range(3, 8):
    state["tokens"][0] = tokens[0, t]
    state, _ = oracle.generate(state)
    torch.testing.assert_close(state["logits"][0, 0], full[0, t], rtol=2e-4, atol=2e-4)
  # now slot 1 arrives; its AR rows start at ring index 5
  padded = torch.zeros(oracle.P, dtype=torch.int64)
  padded[:6] = tokens[1, :6]
  prefix, _ = oracle.prefill(padded, 6)
  state = oracle.insert(prefix, state, 1)
  assert int(state["cache"]["ar_index"]) == 5
  assert int(state["cache"]["ar_lengths"][1]) == 0
  for i in range(7):
    state["tokens"][0] = tokens[0, 8 + i]
    state["tokens"][1] = tokens[1, 6 + i]
    state, _ = oracle.generate(state)
    torch.testing.assert_close(state["logits"][0, 0], full[0, 8 + i], rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(state["logits"][1, 0], full[1, 6 + i], rtol=2e-4, atol=2e-4)
  # 12 steps on a ring of 12: the index wrapped to 0
  assert int(state["cache"]["ar_index"]) == 0


def _complex_rope(x, positions, theta=10000.0):
  """Independent RoPE on interleaved pairs (llama reference), then permuted to half-split."""
  B, T, H, D = x.shape
  freqs = 1.0 / (theta ** (np.arange(0, D, 2)[: D // 2].astype(np.float64) / D))
  ang = positions[:, :, None].astype(np.float64) * freqs[None, None, :]
  cis = np.exp(1j * ang)[:, :, None, :]
  xc = x[..., : D // 2].astype(np.float64) + 1j * x[..., D // 2 :].astype(np.float64)
  out = xc * cis
  return np.concatenate([out.real, out.imag], axis=-1)


def test_rope_matches_complex_rotation():
  cfg = small_config(head_dim=128)
  oracle = ref.DecodeOracle(cfg, make_params(cfg), faithful=False)
  rng = np.random.default_rng(0)
  x = rng.normal(1, 0.5, (1, 8, 4, 128)).astype(np.float32)
  pos = np.arange(8)[None, :]
  got = oracle.rope(torch.from_numpy(x), torch.from_numpy(pos)).numpy()
  want = _complex_rope(x, pos)
  np.testing.assert_allclose(got, want, rtol=1e-1, atol=1e-4)
  np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


def test_rope_commutes_with_scaling():
  cfg = small_config(head_dim=128)
  oracle = ref.DecodeOracle(cfg, make_params(cfg), faithful=False)
  rng = np.random.default_rng(1)
  x = torch.from_numpy(rng.normal(1, 0.5, (1, 8, 4, 128)).astype(np.float32))
  pos = torch.arange(8)[None, :]
  a = oracle.rope(x, pos) * (128**-0.5)
  b = oracle.rope(x * (128**-0.5), pos)
  torch.testing.assert_close(a, b, rtol=1e-1, atol=1e-4)


def test_two_segment_attention_matches_reference_gqa():
  """prefill segment + AR segment merged (attentions.py:1376-1397) == one softmax over valid rows."""
  cfg = small_config(base_num_query_heads=8, base_num_kv_heads=2, max_prefill_predict_length=32, max_target_length=96)
  oracle = ref.DecodeOracle(cfg, make_params(cfg), faithful=False)
  g = torch.Generator().manual_seed(0)
  B, Hq, Hkv, D, P, R = 4, 8, 2, 64, 32, 64
  bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
  q = bf(torch.randn(B, 1, Hq, D, generator=g))
  Kp, Vp = bf(torch.randn(B, P, Hkv, D, generator=g)), bf(torch.randn(B, P, Hkv, D, generator=g))
  Ka, Va = bf(torch.randn(B, R, Hkv, D, generator=g)), bf(torch.randn(B, R, Hkv, D, generator=g))
  plen = torch.tensor([32, 5, 0, 17])
  alen = torch.tensor([1, 64, 9, 30])
  pmask = torch.arange(P)[None, :] < plen[:, None]
  amask = torch.arange(R)[None, :] < alen[:, None]
  o_p, m_p, l_p = oracle._local_attention(q, Kp, Vp, pmask[:, None, :])
  o_a, m_a, l_a = oracle._local_attention(q, Ka, Va, amask[:, None, :])
  merged = oracle._normalize_attention([o_p, o_a], [m_p, m_a], [l_p, l_a])[:, 0]
  want, _, _ = ref.gqa_decode_ref(
      q[:, 0], torch.cat([Kp, Ka], 1), torch.cat([Vp, Va], 1), torch.cat([pmask, amask], 1), p_bf16=False
  )
  # kernels_test.py:108-110 bounds: max 1.5e-1, mean 1e-2; fp32 mode is far inside them
  assert (merged - want).abs().max() < 1e-5


def test_engine_shapes_and_dtypes():
  cfg = small_config()
  oracle = ref.DecodeOracle(cfg, make_params(cfg), faithful=True)
  state = oracle.init_decode_state()
  assert state["logits"].shape == (2, 1, cfg.vocab_size) and state["logits"].dtype == torch.float32
  padded = torch.zeros(oracle.P, dtype=torch.int64)
  padded[:4] = torch.tensor([3, 1, 4, 1])
  prefix, first = oracle.prefill(padded, 4)
  assert prefix["logits"].shape == (1, 1, cfg.vocab_size)
  assert int(prefix["next_pos"]) == 4 and int(prefix["generated_tokens"]) == 0
  state = oracle.insert(prefix, state, 1)
  assert int(state["tokens"][1]) == int(first)
  state, data = oracle.generate(state)
  assert data.shape == (2, 3) and data.dtype == torch.int32  # token, valid, length
  assert data[:, 1].tolist() == [1, 1] and data[:, 2].tolist() == [1, 1]
  assert state["next_pos"][1].item() == 5


def test_faithful_logits_are_bf16_values_and_greedy_takes_first_max():
  cfg = small_config()
  oracle = ref.DecodeOracle(cfg, make_params(cfg), faithful=True)
  logits = oracle.forward_full(random_tokens((1, 6), cfg.vocab_size))
  assert torch.equal(logits, logits.to(torch.bfloat16).to(torch.float32))
  lg = torch.tensor([[0.0, 2.0, 2.0, -1.0]])
  assert ref.sampling(lg, "greedy").item() == 1


def test_philox_known_answers():
  # Random123 kat_vectors, philox4x32 with 10 rounds
  cases = [
      ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
      ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
      (
          (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344),
          (0xA4093822, 0x299F31D0),
          (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1),
      ),
  ]
  for ctr, key, want in cases:
    got = ref.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], key[0], key[1])
    assert tuple(int(g[0]) for g in got) == want


def test_sampling_strategies():
  g = torch.Generator().manual_seed(3)
  logits = torch.randn(3, 1, 500, generator=g) * 3
  # topk: the sample is always one of the k largest
  for step in range(5):
    tok = ref.sampling(logits, "topk", topk=7, temperature=0.8, seed=11, step=step)
    top = torch.topk(logits, 7, dim=-1).indices
    assert all(tok[b, 0] in top[b, 0] for b in range(3))
  # nucleus: cutoff logit is the first sorted logit whose cumulative mass reaches p
  cut = ref.nucleus_cutoff(logits.reshape(3, 500), 0.6)
  for b in range(3):
    row = logits[b, 0]
    probs = torch.softmax(row, -1)
    mass_ge = probs[row >= cut[b]].sum()
    mass_gt = probs[row > cut[b]].sum()
    assert mass_ge >= 0.6 - 1e-6 and mass_gt < 0.6 + 1e-6
  tok = ref.sampling(logits, "nucleus", nucleus_topp=0.6, temperature=1.0, seed=5, step=0)
  assert all(logits[b, 0, tok[b, 0]] >= cut[b] for b in range(3))
  # weighted: empirical distribution follows softmax(logits / T)
  small = torch.tensor([[1.0, 0.0, -1.0, 2.0]]).repeat(4000, 1)
  tok = ref.sampling(small, "weighted", temperature=1.0, seed=1, step=0)
  freq = torch.bincount(tok, minlength=4).float() / 4000
  torch.testing.assert_close(freq, torch.softmax(small[0], -1), atol=0.03, rtol=0)
  with pytest.raises(ValueError):
    ref.sampling(logits, "topk", topk=0)
  with pytest.raises(ValueError):
    ref.sampling(logits, "beam")


def test_log_prob_of_chosen_token():
  logits = torch.tensor([[[0.0, math.log(3.0)]]])
  lp = ref.log_prob_of_chosen_token(logits, torch.tensor([[1]]))
  assert abs(lp.item() - math.log(0.75)) < 1e-6


# ---- gemma3 block (layers/gemma3.py:36-197) ----------------------------------------------------------------------


def _gemma_cfg(**kw):
  base = dict(model_name="gemma3-27b", base_num_decoder_layers=7, base_emb_dim=128, base_num_query_heads=4, base_num_kv_heads=2,
              head_dim=64, base_mlp_dim=256, vocab_size=512, sliding_window_size=8, max_prefill_predict_length=16,
              max_target_length=48, per_device_batch_size=2)
  base.update(kw)
  return small_config(**base)


def test_gemma3_config_and_layer_pattern():
  from maxtext_indextts2_b200 import gemma3, pyconfig

  cfg = _gemma_cfg()
  assert [gemma3.get_attention_type(i) for i in range(7)] == ["local_sliding"] * 5 + ["global", "local_sliding"]  # gemma3.py:36-48
  assert gemma3.get_query_pre_attn_scalar(cfg) == (cfg.base_emb_dim // cfg.base_num_query_heads) ** -0.5         # gemma3.py:55-56
  with pytest.raises(ValueError):
    _gemma_cfg(model_name="default", decoder_block="gemma3", mlp_activations=["gelu", "linear"], use_post_attn_norm=True,
               use_post_ffw_norm=True, logits_via_embedding=True)  # gemma3.py:57-58: the scalar is chosen by model name
  cfg4 = pyconfig.initialize(None, model_name="gemma3-4b")  # head_dim 256
  assert cfg4.head_dim == 256 and gemma3.get_query_pre_attn_scalar(cfg4) == 256**-0.5
  with pytest.raises(ValueError):
    pyconfig.initialize(None, model_name="gemma3-4b", head_dim=512)


def test_gemma3_ar_steps_equal_full_forward_when_the_window_covers_the_cache():
  """tests/attention_test.py:361-406 for the gemma3 block.  With sliding_window_size >= the segment lengths the cache-index
  window of AUTOREGRESSIVE mode (attentions.py:600-602) admits every valid row and equals the position window of the
  full-sequence graph as long as the sequence is shorter than the window."""
  cfg = _gemma_cfg(sliding_window_size=64)
  params = make_params(cfg)
  o = ref.DecodeOracle(cfg, params, faithful=False)
  toks = random_tokens((1, 24), cfg.vocab_size, seed=9)
  full = o.forward_full(toks)
  state = o.init_decode_state()
  prefix, _ = o.prefill(toks[0, :16], 16)
  state = o.insert(prefix, state, 0)
  torch.testing.assert_close(prefix["logits"][0, 0], full[0, 15], rtol=1e-4, atol=1e-4)
  for t in range(16, 24):
    state["tokens"][0] = int(toks[0, t])
    logits = o.step_logits(state)
    state["next_pos"] = state["next_pos"] + 1
    torch.testing.assert_close(logits[0, 0], full[0, t], rtol=2e-4, atol=2e-4)


def test_gemma3_local_layers_mask_by_cache_index_in_ar_mode():
  """attentions.py:600-602,624-631: in AUTOREGRESSIVE mode next_pos = kv_seq_len - 1 of the segment, so a local layer sees the
  last `sliding_window_size` INDICES of the prefill segment and of the ring."""
  cfg = _gemma_cfg()
  o = ref.DecodeOracle(cfg, make_params(cfg), faithful=False)
  assert o._window_ar(16).nonzero()[:, 0].tolist() == list(range(8, 16))
  assert o._window_ar(32).nonzero()[:, 0].tolist() == list(range(24, 32))
  w = o._window_full(12)
  assert w[11].nonzero()[:, 0].tolist() == list(range(4, 12)) and w[3].nonzero()[:, 0].tolist() == [0, 1, 2, 3]
