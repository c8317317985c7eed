"""Generates tests/golden/decode_small.npz from the CPU oracle (run from the repo root).

The reference (Python/JAX) cannot run in this image, so these vectors are outputs of
oracle/decode_ref.py -- they pin the oracle against drift and give the GPU tests a fixture
that does not need the oracle at run time.  Config: tests/helpers.small_config() defaults,
2 slots, prompts from PCG64(1234), greedy, 12 decode steps, dtype-faithful mode.
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import decode_ref as ref  # noqa: E402
from tests.helpers import make_params, random_tokens, small_config  # noqa: E402


def generate():
  cfg = small_config()
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  prompts = random_tokens((2, 16), cfg.vocab_size)
  lengths = [9, 16]
  state = oracle.init_decode_state()
  first = []
  for slot in range(2):
    padded = torch.zeros(oracle.P, dtype=torch.int64)
    padded[: lengths[slot]] = prompts[slot, : lengths[slot]]
    prefix, tok = oracle.prefill(padded, lengths[slot])
    first.append(int(tok))
    state = oracle.insert(prefix, state, slot)
  tokens, logits = [], []
  for _ in range(12):
    state, data = oracle.generate(state)
    tokens.append(data[:, 0].numpy().copy())
    logits.append(state["logits"][:, 0].numpy().copy())
  return dict(prompts=prompts.numpy(), lengths=np.array(lengths), first_tokens=np.array(first), tokens=np.stack(tokens),
              logits=np.stack(logits).astype(np.float32))


if __name__ == "__main__":
  out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decode_small.npz")
  np.savez_compressed(out, **generate())
  print("wrote", out, os.path.getsize(out), "bytes")
