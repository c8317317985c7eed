"""CPU checks of the test infrastructure that ties the CUDA engine's state to the oracle (oracle/mirror.py) and of
the checkpoint-layout helpers (params.unscan_params): no GPU needed."""

import types

import pytest
import torch

from maxtext_indextts2_b200 import maxengine, params as params_lib
from oracle import decode_ref as ref
from oracle import mirror
from tests.helpers import make_params, random_tokens, small_config


def test_device_layout_unpacks_to_the_oracle_weights():
  cfg = small_config()
  params = make_params(cfg)
  want = ref.prepare_weights(params, cfg)
  got = mirror.oracle_weights_from_device(maxengine.pack_params(params, cfg, torch.device("cpu")), cfg)
  assert torch.equal(got.embedding, want.embedding) and torch.equal(got.final_scale, want.final_scale)
  assert torch.equal(got.logits, want.logits)
  for a, b in zip(got.layers, want.layers):
    assert set(a) == set(b)
    for k in a:
      assert torch.equal(a[k], b[k]), k


def test_scanned_checkpoint_layout_is_unstacked():
  """configs/base.yml keeps the reference default scan_layers=True: load_params must accept decoder/layers/... with
  the layer index on axis param_scan_axis=1 (decoders.py:427)."""
  cfg = small_config(scan_layers=True)
  params = make_params(cfg)
  scanned = params_lib.scan_params(params, cfg)
  assert "layers" in scanned["params"]["decoder"] and "layers_0" not in scanned["params"]["decoder"]
  q = scanned["params"]["decoder"]["layers"]["self_attention"]["query"]["kernel"]
  assert q.shape == (cfg.emb_dim, cfg.num_decoder_layers, cfg.num_query_heads, cfg.head_dim)
  a = maxengine.pack_params(scanned, cfg, torch.device("cpu")).tensors
  b = maxengine.pack_params(params, cfg, torch.device("cpu")).tensors
  for k in a:
    assert torch.equal(a[k], b[k]), k
  wa, wb = ref.prepare_weights(scanned, cfg), ref.prepare_weights(params, cfg)
  assert all(torch.equal(x[k], y[k]) for x, y in zip(wa.layers, wb.layers) for k in x)
  bad = {"params": {"token_embedder": params["params"]["token_embedder"], "decoder": {"decoder_norm": params["params"]["decoder"]["decoder_norm"]}}}
  with pytest.raises(ValueError, match="layers"):
    maxengine.pack_params(bad, cfg, torch.device("cpu"))


def _engine_view_of(ostate, oracle, planes):
  """The engine's device layout ([L, planes, Hkv, T, D], lengths) holding the oracle state `ostate`."""
  cfg, c = oracle.cfg, ostate["cache"]
  L, Hkv, D = cfg.num_decoder_layers, cfg.num_kv_heads, cfg.head_dim
  k = torch.randn(L, planes, Hkv, oracle.T, D).to(torch.bfloat16)  # rows outside the valid ranges hold anything
  v = torch.randn(L, planes, Hkv, oracle.T, D).to(torch.bfloat16)
  for l in range(L):
    k[l, : oracle.B, :, : oracle.P] = c["prefill_key"][l].permute(0, 2, 1, 3).to(torch.bfloat16)
    v[l, : oracle.B, :, : oracle.P] = c["prefill_value"][l].permute(0, 2, 1, 3).to(torch.bfloat16)
    k[l, : oracle.B, :, oracle.P :] = c["ar_key"][l].permute(0, 2, 1, 3).to(torch.bfloat16)
    v[l, : oracle.B, :, oracle.P :] = c["ar_value"][l].permute(0, 2, 1, 3).to(torch.bfloat16)
  plen = torch.zeros(planes, dtype=torch.int32)
  alen = torch.zeros(planes, dtype=torch.int32)
  plen[: oracle.B] = c["prefill_segment_id"].sum(-1)
  alen[: oracle.B] = c["ar_lengths"]
  return types.SimpleNamespace(_k=k, _v=v, _prefill_len=plen, _ar_lengths=alen, _ar_index=torch.tensor([c["ar_index"]], dtype=torch.int32),
                               _next_pos=ostate["next_pos"].clone(), _generated=ostate["generated_tokens"].clone(), _tokens=ostate["tokens"].clone())


def test_mirrored_state_continues_like_the_original():
  """An oracle run (ragged prompts, a late insert, past the ring wrap) is exported in the engine's layout and mirrored
  back into a fresh oracle holding a subset of the slots: both must produce the same logits from then on."""
  cfg = small_config(per_device_batch_size=3, max_prefill_predict_length=8, max_target_length=14)
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  prompts = random_tokens((3, 8), cfg.vocab_size, seed=3)
  state = oracle.init_decode_state()
  for slot, n in ((0, 8), (1, 3)):
    prefix, _ = oracle.prefill(prompts[slot], n)
    state = oracle.insert(prefix, state, slot)
  for _ in range(4):
    state, _ = oracle.generate(state)
  prefix, _ = oracle.prefill(prompts[2], 5)
  state = oracle.insert(prefix, state, 2)  # joins late: its ring rows start at index 4
  for _ in range(4):  # 8 steps on a ring of 6 rows: slots 0 and 1 have wrapped
    state, _ = oracle.generate(state)
  fake = _engine_view_of(state, oracle, planes=4)
  sub = mirror.make_oracle(cfg, oracle.w, 2, faithful=True)
  sstate = mirror.mirror_state(fake, sub, [2, 0])
  assert torch.equal(sstate["cache"]["ar_segment_id"], state["cache"]["ar_segment_id"][[2, 0]])
  assert torch.equal(sstate["cache"]["prefill_segment_id"], state["cache"]["prefill_segment_id"][[2, 0]])
  for _ in range(3):
    state, data = oracle.generate(state)
    sstate, sdata = sub.generate(sstate)
    assert torch.equal(sstate["logits"], state["logits"][[2, 0]])
    assert torch.equal(sdata, data[[2, 0]])


def test_mismatch_classification_uses_bf16_ulps_of_the_top_logit():
  row = torch.tensor([4.0, 4.02, 3.0])
  c = mirror.classify_mismatch(row, 0, 1)
  assert mirror.bf16_ulp(4.02) == 2.0**-5 and c["strict"] and abs(c["ulps"] - 0.02 / 2.0**-5) < 1e-3
  assert not mirror.classify_mismatch(row, 2, 1)["strict"]
