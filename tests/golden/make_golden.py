"""Generates tests/golden/decode_small.npz and gemma3_small.npz from the CPU oracle (run from the repo root).

The reference (Python/JAX) cannot run in this image, so these vectors are outputs of
oracle/decode_ref.py -- they pin the oracle against drift and give the GPU tests a fixture
that does not need the oracle at run time.  Config: tests/helpers.small_config() defaults,
2 slots, prompts from PCG64(1234), greedy, 12 decode steps, dtype-faithful mode.  gemma3_small.npz: the gemma3 block
(gemma3_config below: 7 layers = five local, one global, one local; window 8 < prefill segment 16 < ring 32), 20 decode steps.
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import decode_ref as ref  # noqa: E402
from tests.helpers import make_params, random_tokens, small_config  # noqa: E402


def gemma3_config(**kw):
  base = dict(model_name="gemma3-27b", base_num_decoder_layers=7, base_emb_dim=128, base_num_query_heads=4, base_num_kv_heads=2,
              head_dim=64, base_mlp_dim=256, vocab_size=512, sliding_window_size=8, max_prefill_predict_length=16,
              max_target_length=48, per_device_batch_size=2)
  base.update(kw)
  return small_config(**base)


def generate(cfg=None, steps=12, lengths=(9, 16)):
  cfg = cfg or small_config()
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  prompts = random_tokens((2, 16), cfg.vocab_size)
  lengths = list(lengths)
  state = oracle.init_decode_state()
  first = []
  for slot in range(2):
    padded = torch.zeros(oracle.P, dtype=torch.int64)
    padded[: lengths[slot]] = prompts[slot, : lengths[slot]]
    prefix, tok = oracle.prefill(padded, lengths[slot])
    first.append(int(tok))
    state = oracle.insert(prefix, state, slot)
  tokens, logits = [], []
  for _ in range(steps):
    state, data = oracle.generate(state)
    tokens.append(data[:, 0].numpy().copy())
    logits.append(state["logits"][:, 0].numpy().copy())
  return dict(prompts=prompts.numpy(), lengths=np.array(lengths), first_tokens=np.array(first), tokens=np.stack(tokens),
              logits=np.stack(logits).astype(np.float32))


def generate_gemma3():
  return generate(gemma3_config(), steps=20, lengths=(12, 16))


if __name__ == "__main__":
  here = os.path.dirname(os.path.abspath(__file__))
  for name, data in (("decode_small.npz", generate()), ("gemma3_small.npz", generate_gemma3())):
    out = os.path.join(here, name)
    np.savez_compressed(out, **data)
    print("wrote", out, os.path.getsize(out), "bytes")
