"""Configuration for the decode path, with the reference's key names.

Mirrors the behaviour of the reference's ``pyconfig.initialize``
(MaxText/pyconfig.py:1170) for the keys the decode step reads:

* ``argv[1]`` is a YAML file, everything after it is ``key=value``
  (pyconfig.py:426);
* ``M_<KEY>`` environment variables override YAML values (pyconfig.py:414-421)
  and may not be combined with a command-line value for the same key;
* an unknown key is an error (pyconfig.py:435-437);
* ``model_name`` merges ``configs/models/<name>.yml`` (pyconfig.py:682-703);
* derived keys ``emb_dim``, ``num_query_heads``, ``num_kv_heads``, ``mlp_dim``,
  ``num_decoder_layers`` (pyconfig.py:576-582);
* ``attn_logits_soft_cap`` / ``final_logits_soft_cap`` of 0.0 become ``None``
  (pyconfig.py:571-574);
* the result is read-only (pyconfig.py:1150-1168).

Only the parser is ours; the reference's depends on omegaconf, which this
decode path does not need.
"""

from __future__ import annotations

import math
import os
from typing import Any

import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
BASE_YML = os.path.join(_HERE, "configs", "base.yml")
_MAX_PREFIX = "M_"

_VALID_ATTENTION = ("autoselected", "dot_product", "flash", "cudnn_flash_te", "cudnn_flash_jax", "paged")
_VALID_SAMPLING = ("greedy", "weighted", "nucleus", "topk")
_VALID_AXIS_ORDER = ("0,1,2,3", "0,2,1,3")


def _string_to_bool(s: str) -> bool:
  if s.lower() == "true":
    return True
  if s.lower() == "false":
    return False
  raise ValueError(f"Can't convert {s} to bool")


def _parse_like(default: Any, text: str, key: str) -> Any:
  """Parse a ``key=value`` string with the type of the YAML default."""
  if isinstance(default, bool):
    return _string_to_bool(text)
  if isinstance(default, int):
    try:
      return int(text)
    except ValueError:
      # base.yml writes some float-valued keys as ints (e.g. nucleus_p: -1)
      return float(text)
  if isinstance(default, float):
    return float(text)
  if isinstance(default, str):
    return text
  if isinstance(default, list):
    return yaml.safe_load(text)
  raise ValueError(f"Couldn't parse value {text!r} for key {key}")


def _load_yaml(path: str) -> dict:
  with open(path, "r", encoding="utf-8") as f:
    return yaml.safe_load(f) or {}


def get_individual_scales(scale: int):
  """pyconfig.py:1133-1147: spread a power-of-two scale over emb/heads/mlp/layers."""
  log_2_scale = math.floor(math.log2(scale))
  if 2**log_2_scale != scale:
    raise ValueError("Global parameter scale should be a power of 2.")
  base_scale, rem = divmod(log_2_scale, 3)
  num_head_scale = base_scale + int(rem > 0)
  mlp_dim_scale = num_head_scale
  emb_scale = base_scale + int(rem > 1)
  layer_scale = base_scale
  return emb_scale, num_head_scale, mlp_dim_scale, layer_scale


class HyperParameters:
  """Read-only view of the merged keys (pyconfig.py:1150-1168)."""

  def __init__(self, keys: dict):
    object.__setattr__(self, "_keys", dict(keys))

  def __getattr__(self, attr):
    keys = object.__getattribute__(self, "_keys")
    if attr not in keys:
      raise ValueError(f"Requested key {attr}, not in config")
    return keys[attr]

  def __setattr__(self, attr, value):
    raise ValueError("Reinitialization of config is not allowed")

  def get_keys(self) -> dict:
    return dict(object.__getattribute__(self, "_keys"))


def _validate(keys: dict) -> None:
  if keys["attention"] not in _VALID_ATTENTION:
    raise ValueError("Invalid attention kernel was passed. Valid options ", _VALID_ATTENTION)
  if keys["compute_axis_order"] not in _VALID_AXIS_ORDER:
    raise ValueError("Invalid compute_axis_order was passed. Valid options are ", _VALID_AXIS_ORDER)
  if keys["decode_sampling_strategy"] not in _VALID_SAMPLING:
    raise ValueError(f"Sampling algorithm={keys['decode_sampling_strategy']!r} not supported!")
  if keys["quantize_kvcache"]:
    # inference/kvcache.py:36-90.  Implemented: int8 with kv_quant_axis "dkv" (one scale per token and kv head, fused into the QKV
    # epilogue) and "heads_and_dkv" (the reference's default: one scale per token over all kv heads; a small kernel after the QKV GEMM).
    if keys["kv_quant_dtype"] not in ("int8", "fp8"):
      raise ValueError(f"Invalid kv_quant_dtype: {keys['kv_quant_dtype']} (this decode path implements int8 and fp8)")
    if keys["kv_quant_axis"] not in ("dkv", "heads_and_dkv"):
      raise ValueError(f"Invalid KV quant axis cfg: {keys['kv_quant_axis']}")  # kvcache.py:73
    if keys["head_dim"] != 64:
      raise ValueError("quantize_kvcache is implemented for head_dim=64")
  if keys["quantization"] not in ("", None):
    raise ValueError("quantization must be '' on this decode path (bf16 weights only)")
  if keys["decoder_block"] not in ("llama2", "gemma3"):
    raise ValueError(f"decoder_block={keys['decoder_block']!r}: the llama2 and gemma3 blocks are on this path")
  if keys["dtype"] != "bfloat16":
    raise ValueError("dtype must be bfloat16 on this path")
  if keys["rope_type"] != "default":
    raise ValueError("only rope_type=default is on this path")
  if keys["decoder_block"] == "gemma3":
    # layers/gemma3.py:45-197 + configs/models/gemma3-*.yml
    if keys["mlp_activations"] != ["gelu", "linear"]:
      raise ValueError("the gemma3 block is implemented with mlp_activations=[gelu, linear]")
    if not (keys["use_post_attn_norm"] and keys["use_post_ffw_norm"]):
      raise ValueError("the gemma3 block is implemented with use_post_attn_norm and use_post_ffw_norm (as every gemma3 model yml sets them)")
    if keys["head_dim"] not in (64, 128, 256):
      raise ValueError(f"head_dim={keys['head_dim']}: the attention kernels of this path take head_dim 64, 128 or 256")
    if keys["head_dim"] == 256 and keys["base_num_query_heads"] // max(1, keys["base_num_kv_heads"]) > 8:
      raise ValueError("head_dim 256 needs at most 8 query heads per kv head")
    if keys["sliding_window_size"] <= 0:
      raise ValueError("Sliding_window_size must be set if Local Sliding attention type")  # attentions.py:625-626
    if not str(keys["model_name"]).startswith("gemma3"):
      raise ValueError(f"Unsupported model name: {keys['model_name']}")  # gemma3.py:57 (query_pre_attn_scalar is chosen by model name)
    if keys["quantize_kvcache"] or keys["scan_layers"] or int(keys["vocab_parallelism"]) != 1 or not keys["fold_norm_scales"]:
      raise ValueError("the gemma3 block is implemented with a bf16 KV cache, scan_layers=False, vocab_parallelism=1 and fold_norm_scales=True")
  else:
    if keys["mlp_activations"] != ["silu", "linear"]:
      raise ValueError("only mlp_activations=[silu, linear] is on this path")
    if keys["use_post_attn_norm"] or keys["use_post_ffw_norm"] or keys["sliding_window_size"]:
      raise ValueError("use_post_attn_norm / use_post_ffw_norm / sliding_window_size belong to the gemma3 block")
  if keys["max_target_length"] <= keys["max_prefill_predict_length"]:
    # inference/kvcache.py:408-412
    raise ValueError(
        f"max_target_length: {keys['max_target_length']} should be greater than max_prefill_length:"
        f" {keys['max_prefill_predict_length']}!"
    )
  if keys["attention"] == "paged":
    # inference/paged_attention.py + inference/page_manager.py: bf16 pages, llama2 block, one page group per decode slot
    tpp = int(keys["pagedattn_tokens_per_page"])
    if tpp < 8 or tpp & (tpp - 1):
      raise ValueError("pagedattn_tokens_per_page must be a power of two >= 8 on this path")
    if keys["quantize_kvcache"] or keys["decoder_block"] != "llama2" or int(keys["vocab_parallelism"]) != 1:
      raise ValueError("attention=paged is implemented for the llama2 block with a bf16 cache and vocab_parallelism=1")
    if keys["head_dim"] not in (64, 128):
      raise ValueError("attention=paged takes head_dim 64 or 128")
  for k in ("use_qk_norm", "use_iota_embed", "use_untrainable_positional_embedding", "fused_qkv", "fused_mlp"):
    if keys[k]:
      raise ValueError(f"{k}=True is outside this decode path")
  if keys["trainable_position_size"] > 0:
    raise ValueError("trainable_position_size > 0 is outside this decode path")


def initialize(argv=None, **kwargs) -> HyperParameters:
  """Build the config from ``[prog, yaml, key=value...]`` plus keyword overrides."""
  argv = list(argv) if argv else ["", BASE_YML]
  yml_path = argv[1] if len(argv) > 1 and "=" not in argv[1] else BASE_YML
  cli = [a for a in argv[1:] if "=" in a]

  keys = _load_yaml(yml_path)
  if "base_config" in keys:  # pyconfig.py:475-497
    parent = keys.pop("base_config")
    if not os.path.isabs(parent):
      parent = os.path.join(os.path.dirname(yml_path), parent)
    merged = _load_yaml(parent)
    merged.update(keys)
    keys = merged
  elif os.path.abspath(yml_path) != os.path.abspath(BASE_YML):
    merged = _load_yaml(BASE_YML)
    merged.update(keys)
    keys = merged

  overrides = {}
  for item in cli:
    k, v = item.split("=", 1)
    overrides[k] = v
  for k, v in kwargs.items():
    overrides[k] = v

  for k in overrides:
    if k not in keys:
      raise ValueError(f"Key {k} was passed at the command line but isn't in config.")

  # model yml first (so explicit overrides win), pyconfig.py:682-703
  model_name = overrides.get("model_name", os.environ.get(_MAX_PREFIX + "MODEL_NAME", keys["model_name"]))
  if model_name != "default":
    model_yml = os.path.join(_HERE, "configs", "models", f"{model_name}.yml")
    if not os.path.isfile(model_yml):
      raise ValueError(f"Model {model_name!r}: no such file {model_yml}")
    for k, v in _load_yaml(model_yml).items():
      if k not in keys:
        raise ValueError(f"Key {k} in {model_yml} isn't in config.")
      keys[k] = v

  for k in list(keys):
    env_key = _MAX_PREFIX + k.upper()
    if k in overrides and env_key in os.environ:
      raise ValueError(f"You are passing overrides by both CLI and ENV for `{k}`. This isn't allowed.")
    if k in overrides:
      v = overrides[k]
      keys[k] = _parse_like(keys[k], v, k) if isinstance(v, str) and not isinstance(keys[k], str) else v
    elif env_key in os.environ:
      keys[k] = _parse_like(keys[k], os.environ[env_key], k)

  if keys["attn_logits_soft_cap"] == 0.0:
    keys["attn_logits_soft_cap"] = None
  if keys["final_logits_soft_cap"] == 0.0:
    keys["final_logits_soft_cap"] = None

  if keys["pagedattn_max_pages_per_group"] <= 0:  # pyconfig.py:611-614
    keys["pagedattn_max_pages_per_group"] = (keys["max_target_length"] + keys["pagedattn_tokens_per_page"] - 1) // keys["pagedattn_tokens_per_page"]

  emb_scale, num_head_scale, mlp_dim_scale, layer_scale = get_individual_scales(keys["global_parameter_scale"])
  keys["emb_dim"] = 2**emb_scale * keys["base_emb_dim"]
  keys["num_query_heads"] = 2**num_head_scale * keys["base_num_query_heads"]
  keys["num_kv_heads"] = 2**num_head_scale * keys["base_num_kv_heads"]
  keys["mlp_dim"] = 2**mlp_dim_scale * keys["base_mlp_dim"]
  keys["num_decoder_layers"] = 2**layer_scale * keys["base_num_decoder_layers"]

  _validate(keys)
  return HyperParameters(keys)
