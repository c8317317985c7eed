"""Shared fixtures for the parity tests: small configs, prompts, params."""

import numpy as np
import torch

from maxtext_indextts2_b200 import params as params_lib
from maxtext_indextts2_b200 import pyconfig


def small_config(**kw):
  base = dict(
      base_num_decoder_layers=2,
      base_emb_dim=128,
      base_num_query_heads=4,
      base_num_kv_heads=2,
      head_dim=64,
      base_mlp_dim=256,
      vocab_size=1024,
      per_device_batch_size=2,
      max_prefill_predict_length=16,
      max_target_length=32,
      weight_dtype="bfloat16",
      attention="dot_product",
      scan_layers=False,
  )
  base.update(kw)
  return pyconfig.initialize(None, **base)


def make_params(config, seed=0, perturb=True):
  p = params_lib.init_params(config, seed)
  if perturb:
    params_lib.perturb_norm_scales(p["params"], seed + 1)
  return p


def random_tokens(shape, vocab, seed=1234):
  rng = np.random.Generator(np.random.PCG64(seed))
  return torch.from_numpy(rng.integers(0, vocab, size=shape, dtype=np.int64))
