"""Constants of the Gemma-3 decoder block (MaxText/layers/gemma3.py:36-62): the 5:1 local/global attention
pattern and the scalar the query is multiplied by before attention."""

from __future__ import annotations

LOCAL_SLIDING = "local_sliding"
GLOBAL = "global"

# gemma3.py:36-43
GEMMA3_ATTENTION_PATTERN = (LOCAL_SLIDING, LOCAL_SLIDING, LOCAL_SLIDING, LOCAL_SLIDING, LOCAL_SLIDING, GLOBAL)


def get_attention_type(layer_id: int) -> str:
  """gemma3.py:46-48."""
  return GEMMA3_ATTENTION_PATTERN[layer_id % len(GEMMA3_ATTENTION_PATTERN)]


def get_query_pre_attn_scalar(config) -> float:
  """gemma3.py:51-58."""
  if config.model_name in ("gemma3-4b", "gemma3-12b"):
    return config.head_dim**-0.5
  if config.model_name == "gemma3-27b":
    return (config.base_emb_dim // config.base_num_query_heads) ** -0.5
  raise ValueError(f"Unsupported model name: {config.model_name}")


def local_layer_mask(num_layers: int) -> list:
  """1 for the layers that use sliding-window attention and the local RoPE base."""
  return [1 if get_attention_type(i) == LOCAL_SLIDING else 0 for i in range(num_layers)]
