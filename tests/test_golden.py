"""Golden vectors (tests/golden/decode_small.npz and gemma3_small.npz, made by tests/golden/make_golden.py).

CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA engine matches them without
the oracle in the loop (logits rtol = atol = 1e-1, greedy ids exact except documented near-ties)."""

import os

import numpy as np
import pytest
import torch

from tests.golden import make_golden
from tests.helpers import make_params, small_config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode_small.npz")
GOLDEN_GEMMA3 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gemma3_small.npz")


@pytest.mark.parametrize("which", ["llama2", "gemma3"])
def test_oracle_reproduces_golden_vectors(which):
  want = np.load(GOLDEN if which == "llama2" else GOLDEN_GEMMA3)
  got = make_golden.generate() if which == "llama2" else make_golden.generate_gemma3()
  for k in want.files:
    np.testing.assert_array_equal(got[k], want[k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["llama2", "gemma3"])
def test_engine_matches_golden_vectors(which):
  from maxtext_indextts2_b200 import maxengine

  g = np.load(GOLDEN if which == "llama2" else GOLDEN_GEMMA3)
  cfg = small_config(materialize_logits=True) if which == "llama2" else make_golden.gemma3_config(materialize_logits=True)
  engine = maxengine.MaxEngine(cfg)
  dparams = engine.load_params(make_params(cfg))
  state = engine.init_decode_state()
  for slot in range(2):
    n = int(g["lengths"][slot])
    padded = torch.zeros(cfg.max_prefill_predict_length, dtype=torch.int64)
    padded[:n] = torch.from_numpy(g["prompts"][slot, :n])
    prefix, _ = engine.prefill(params=dparams, padded_tokens=padded, true_length=n)
    prefix["tokens"].fill_(int(g["first_tokens"][slot]))
    state = engine.insert(prefix, state, slot)
  mismatches = 0
  for step in range(g["tokens"].shape[0]):
    state, result = engine.generate(dparams, state)
    np.testing.assert_allclose(state["logits"].cpu().numpy()[:, 0], g["logits"][step], rtol=1e-1, atol=1e-1 if which == "llama2" else 0.15)
    got = result.data.cpu().numpy()[:, 0]
    for b in range(2):
      if got[b] != g["tokens"][step, b]:
        row = g["logits"][step, b]
        assert abs(row[got[b]] - row[g["tokens"][step, b]]) <= 2**-6 * max(1.0, abs(row.max()))
        mismatches += 1
    state["tokens"].copy_(torch.from_numpy(g["tokens"][step]).reshape(2, 1))
  assert mismatches <= (1 if which == "llama2" else 2)
