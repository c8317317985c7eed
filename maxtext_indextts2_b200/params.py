"""Parameter tree of the llama2-style decoder (and of the gemma3 block), with the reference's names and shapes.

Tree (unscanned layout; SURVEY 5 "Checkpoint / resume", names from
MaxText/layers/models.py:69, llama2.py:79,102,139, attentions.py:1865-1869,
linears.py:347,373,388, decoders.py:544,573)::

  params/token_embedder/embedding                       [V, E]
  params/decoder/layers_{i}/pre_self_attention_layer_norm/scale   [E]
  params/decoder/layers_{i}/self_attention/query/kernel  [E, Hq, D]
  params/decoder/layers_{i}/self_attention/key/kernel    [E, Hkv, D]
  params/decoder/layers_{i}/self_attention/value/kernel  [E, Hkv, D]
  params/decoder/layers_{i}/self_attention/out/kernel    [Hq, D, E]
  params/decoder/layers_{i}/mlp/mlp_layer_norm/scale     [E]
  params/decoder/layers_{i}/mlp/wi_0/kernel              [E, M]
  params/decoder/layers_{i}/mlp/wi_1/kernel              [E, M]
  params/decoder/layers_{i}/mlp/wo/kernel                [M, E]
  params/decoder/decoder_norm/scale                      [E]
  params/decoder/logits_dense/kernel                     [E, V]   (absent if logits_via_embedding)

The gemma3 block (layers/gemma3.py:86-171, attentions.py:1872-1888) names its norms
``pre_self_attention_norm`` / ``post_self_attention_norm`` / ``post_ffw_norm`` ([E] each; the MLP pre-norm stays
``mlp/mlp_layer_norm``) and adds ``self_attention/query_norm`` and ``key_norm`` ([D]).

Random init follows the reference's distributions (layers/initializers.py:31-43,
models.py:68, linears.py:106, attentions.py:1510,1900-1904) but draws from
numpy's PCG64 because JAX's bit streams are not reproducible outside JAX.
"""

from __future__ import annotations

import math

import numpy as np
import torch

_TRUNC_STD = 0.87962566103423978  # std of a unit normal truncated to [-2, 2]


def _normal(rng: np.random.Generator, shape, std: float, dtype: torch.dtype) -> torch.Tensor:
  n = int(np.prod(shape))
  out = torch.empty(n, dtype=dtype)
  step = 1 << 24
  for lo in range(0, n, step):
    hi = min(n, lo + step)
    chunk = rng.standard_normal(hi - lo, dtype=np.float32) * np.float32(std)
    out[lo:hi] = torch.from_numpy(chunk).to(dtype)
  return out.reshape(shape)


def _truncated_normal(rng: np.random.Generator, shape, std: float, dtype: torch.dtype) -> torch.Tensor:
  """variance_scaling(..., "truncated_normal"): unit normal cut at +-2, rescaled to `std`."""
  n = int(np.prod(shape))
  out = torch.empty(n, dtype=dtype)
  step = 1 << 24
  scale = np.float32(std / _TRUNC_STD)
  for lo in range(0, n, step):
    hi = min(n, lo + step)
    chunk = rng.standard_normal(hi - lo, dtype=np.float32)
    bad = np.abs(chunk) > 2.0
    while bad.any():
      chunk[bad] = rng.standard_normal(int(bad.sum()), dtype=np.float32)
      bad = np.abs(chunk) > 2.0
    out[lo:hi] = torch.from_numpy(chunk * scale).to(dtype)
  return out.reshape(shape)


def weight_torch_dtype(config) -> torch.dtype:
  return torch.bfloat16 if config.weight_dtype == "bfloat16" else torch.float32


def init_params(config, seed: int | None = None) -> dict:
  """Random-init parameter tree on the CPU (reference: maxtext_utils.py:916-921)."""
  seed = config.init_weights_seed if seed is None else seed
  rng = np.random.Generator(np.random.PCG64(seed))
  dt = weight_torch_dtype(config)
  E, Hq, Hkv, D = config.emb_dim, config.num_query_heads, config.num_kv_heads, config.head_dim
  M, V, L = config.mlp_dim, config.vocab_size, config.num_decoder_layers

  params: dict = {"token_embedder": {"embedding": _normal(rng, (V, E), 1.0, dt)}}
  dec: dict = {}
  gemma3 = config.decoder_block == "gemma3"
  for i in range(L):
    # attention kernels: variance_scaling(1.0, fan_in, normal); query additionally / sqrt(D)
    q = _normal(rng, (E, Hq, D), 1.0 / math.sqrt(E), torch.float32) / math.sqrt(D)
    k = _normal(rng, (E, Hkv, D), 1.0 / math.sqrt(E), dt)
    v = _normal(rng, (E, Hkv, D), 1.0 / math.sqrt(E), dt)
    o = _normal(rng, (Hq, D, E), 1.0 / math.sqrt(Hq * D), dt)
    attn = {
        "query": {"kernel": q.to(dt)},
        "key": {"kernel": k},
        "value": {"kernel": v},
        "out": {"kernel": o},
    }
    mlp = {
        "mlp_layer_norm": {"scale": torch.ones(E, dtype=dt)},
        "wi_0": {"kernel": _truncated_normal(rng, (E, M), 1.0 / math.sqrt(E), dt)},
        "wi_1": {"kernel": _truncated_normal(rng, (E, M), 1.0 / math.sqrt(E), dt)},
        "wo": {"kernel": _truncated_normal(rng, (M, E), 1.0 / math.sqrt(M), dt)},
    }
    if gemma3:
      # layers/gemma3.py:86-171: names of the four block norms; attentions.py:1872-1888: the q / k norms over head_dim
      attn["query_norm"] = {"scale": torch.ones(D, dtype=dt)}
      attn["key_norm"] = {"scale": torch.ones(D, dtype=dt)}
      dec[f"layers_{i}"] = {
          "pre_self_attention_norm": {"scale": torch.ones(E, dtype=dt)},
          "self_attention": attn,
          "post_self_attention_norm": {"scale": torch.ones(E, dtype=dt)},
          "mlp": mlp,
          "post_ffw_norm": {"scale": torch.ones(E, dtype=dt)},
      }
    else:
      dec[f"layers_{i}"] = {
          "pre_self_attention_layer_norm": {"scale": torch.ones(E, dtype=dt)},
          "self_attention": attn,
          "mlp": mlp,
      }
  dec["decoder_norm"] = {"scale": torch.ones(E, dtype=dt)}
  if not config.logits_via_embedding:
    dec["logits_dense"] = {"kernel": _truncated_normal(rng, (E, V), 1.0 / math.sqrt(E), dt)}
  params["decoder"] = dec
  return {"params": params}


def unscan_params(params: dict, config) -> dict:
  """Reference checkpoint layout with ``scan_layers=True`` (base.yml:261-262) -> the per-layer tree above.

  Under scan the decoder stores ONE subtree ``params/decoder/layers/...`` whose leaves carry the layer index on
  axis ``param_scan_axis`` (default 1; decoders.py:427): scale [E, L], query kernel [E, L, Hq, D], out kernel
  [Hq, L, D, E], wi_0 [E, L, M], wo [M, L, E].  A tree that already has ``layers_{i}`` is returned unchanged.
  """
  dec = params["params"]["decoder"]
  if "layers_0" in dec:
    return params
  if "layers" not in dec:
    raise ValueError(
        "parameter tree has neither params/decoder/layers_{i} (scan_layers=False) nor params/decoder/layers "
        "(scan_layers=True, layer index on axis param_scan_axis)")
  axis = int(config.param_scan_axis)
  L = config.num_decoder_layers

  def take(node, i):
    if isinstance(node, dict):
      return {k: take(v, i) for k, v in node.items()}
    if node.shape[axis] != L:
      raise ValueError(f"scanned leaf of shape {tuple(node.shape)} does not hold {L} layers on axis {axis}")
    return node.select(axis, i)

  new_dec = {k: v for k, v in dec.items() if k != "layers"}
  for i in range(L):
    new_dec[f"layers_{i}"] = take(dec["layers"], i)
  out = dict(params["params"])
  out["decoder"] = new_dec
  return {"params": out}


def scan_params(params: dict, config) -> dict:
  """Inverse of :func:`unscan_params` (tests: builds the layout a scanned reference checkpoint has)."""
  dec = params["params"]["decoder"]
  axis = int(config.param_scan_axis)
  L = config.num_decoder_layers

  def stack(nodes):
    if isinstance(nodes[0], dict):
      return {k: stack([n[k] for n in nodes]) for k in nodes[0]}
    return torch.stack(nodes, dim=axis)

  new_dec = {k: v for k, v in dec.items() if not k.startswith("layers_")}
  new_dec["layers"] = stack([dec[f"layers_{i}"] for i in range(L)])
  out = dict(params["params"])
  out["decoder"] = new_dec
  return {"params": out}


def perturb_norm_scales(params: dict, seed: int = 1) -> dict:
  """Give the RMSNorm scales non-trivial values (tests only need them != 1)."""
  rng = np.random.Generator(np.random.PCG64(seed))

  def walk(node):
    for k, v in node.items():
      if isinstance(v, dict):
        walk(v)
      elif k == "scale":
        node[k] = (1.0 + 0.1 * torch.from_numpy(rng.standard_normal(v.shape[0], dtype=np.float32))).to(v.dtype)

  walk(params)
  return params


def param_count(config) -> dict:
  E, Hq, Hkv, D = config.emb_dim, config.num_query_heads, config.num_kv_heads, config.head_dim
  M, V, L = config.mlp_dim, config.vocab_size, config.num_decoder_layers
  layer = E * Hq * D + 2 * E * Hkv * D + Hq * D * E + 3 * E * M + 2 * E
  if config.decoder_block == "gemma3":
    layer += 2 * E + 2 * D
  return {
      "layers": L * layer,
      "final_norm": E,
      "logits": 0 if config.logits_via_embedding else E * V,
      "embedding": V * E,
  }
