"""The callers either side of the decode step: the audio-id wire format (CPU) and the continuous-batching loop
(GPU) that mirrors MaxText/inference/offline_engine.py:473-715."""

import json

import numpy as np
import pytest
import torch

from maxtext_indextts2_b200 import audio_tokens


def test_audio_token_map_follows_the_reference_construction(tmp_path):
  m = audio_tokens.AudioTokenMap.gemma3_indextts2()
  assert m.vocab_size == 264192  # stats.adjusted_vocab_size, a multiple of 256
  codes = torch.arange(audio_tokens.NUM_CODES)
  ids = m.to_token_ids(codes)
  assert torch.equal(m.to_codes(ids), codes.to(torch.int32)) and ids.unique().numel() == audio_tokens.NUM_CODES
  # extend_tokenizer.py: unused tokens 6.. first, the added tokens after the soft token (shifted down by one), markers, padding
  assert int(m.to_token_ids(torch.tensor([0]))) == 6 and int(m.to_token_ids(torch.tensor([98]))) == 104
  assert int(m.to_token_ids(torch.tensor([99]))) == 256001 and int(m.to_token_ids(torch.tensor([6241]))) == 262143
  assert int(m.to_token_ids(torch.tensor([6242]))) == 262144 and int(m.to_token_ids(torch.tensor([8191]))) == 264093
  assert int(m.to_codes(torch.tensor([264094]))) == audio_tokens.BEGIN_TEXT_AUDIO_ID
  assert int(m.to_codes(torch.tensor([264095]))) == audio_tokens.BEGIN_AUDIO_AUDIO_ID
  assert int(m.to_codes(torch.tensor([264191]))) == audio_tokens.PADDING_AUDIO_ID and int(m.to_codes(torch.tensor([5]))) == -2
  assert bool(m.is_audio(torch.tensor([6]))) and not bool(m.is_audio(torch.tensor([264094])))
  # both JSON forms the reference writes
  adjusted = {"embedding_to_audio": {str(i): int(a) for i, a in enumerate(m.embedding_to_audio.tolist()) if a != -2},
              "stats": {"adjusted_embedding_size": 264192}}
  p = tmp_path / "audio_token_mapping_adjusted.json"
  p.write_text(json.dumps(adjusted))
  again = audio_tokens.AudioTokenMap.from_json(str(p))
  assert torch.equal(again.embedding_to_audio, m.embedding_to_audio) and torch.equal(again.audio_to_embedding, m.audio_to_embedding)
  raw = {"audio_mappings": {str(i if i < 262144 else i + 1): int(a) for i, a in enumerate(m.embedding_to_audio.tolist()) if a != -2}}
  p2 = tmp_path / "audio_token_mapping.json"
  p2.write_text(json.dumps(raw))
  assert torch.equal(audio_tokens.AudioTokenMap.from_json(str(p2), vocab_size=264192).embedding_to_audio, m.embedding_to_audio)


@pytest.mark.gpu
def test_continuous_batching_completes_every_prompt_like_a_dedicated_run():
  """Seven prompts through three slots: slots are refilled as sequences end (EOS or length), every completion equals what the
  oracle says for that prompt alone (teacher-forced, near-ties allowed as elsewhere)."""
  from maxtext_indextts2_b200 import maxengine, offline_engine
  from oracle import decode_ref as ref
  from tests.helpers import make_params, random_tokens, small_config

  cfg = small_config(per_device_batch_size=3, max_prefill_predict_length=16, max_target_length=48, return_log_prob=True)
  params = make_params(cfg)
  rng = np.random.Generator(np.random.PCG64(5))
  prompts = [random_tokens((int(n),), cfg.vocab_size, seed=20 + i).numpy() for i, n in enumerate(rng.integers(2, 17, size=7))]
  probe = offline_engine.OfflineEngine(cfg, params=params)
  base = probe.batch_inference(prompts, max_decode_length=12)
  eos = int(base[2].token_ids[4])  # a token that really occurs: prompt 2 must stop there
  eng = offline_engine.OfflineEngine(cfg, params=params, eos_ids=[eos])
  outs = eng.batch_inference(prompts, max_decode_length=12)
  assert [o.index for o in outs] == list(range(7)) and eng.prefills_run == 7
  assert eng.steps_run < 7 * 11  # slots ran side by side
  oracle_cfg = small_config(per_device_batch_size=1, max_prefill_predict_length=16, max_target_length=48)
  oracle = ref.DecodeOracle(oracle_cfg, params, faithful=True)
  for o, prompt in zip(outs, prompts):
    toks = o.token_ids.tolist()
    assert 1 <= len(toks) <= 12 and o.logprobs.shape == (len(toks),)
    if eos in toks:
      assert toks.index(eos) == len(toks) - 1  # nothing after EOS
    else:
      assert len(toks) == 12
    padded = torch.zeros(16, dtype=torch.int64)
    padded[: len(prompt)] = torch.from_numpy(prompt)
    prefix, first = oracle.prefill(padded, len(prompt))
    state = oracle.insert(prefix, oracle.init_decode_state(), 0)
    logits = prefix["logits"][0, 0]
    for t in toks:
      want = int(torch.argmax(logits))
      assert t == want or abs(float(logits[t]) - float(logits[want])) <= 2**-6 * max(1.0, float(logits.abs().max()))
      state["tokens"] = torch.tensor([[t]], dtype=torch.int32)
      state, _ = oracle.generate(state)
      logits = state["logits"][0, 0]
