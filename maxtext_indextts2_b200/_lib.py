"""ctypes binding of ``libmtx_b200.so`` (the C ABI in ``include/mtx_b200.h``).

The library is built in-tree by :func:`build` (``nvcc`` for sm_100a).  There is no
Python or CPU implementation behind these calls: if the library is missing, or the
process has no CUDA device, the functions raise.
"""

from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libmtx_b200.so")
HEADER = os.path.join(os.path.dirname(_PKG), "include", "mtx_b200.h")

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "--shared",
    "-Xcompiler",
    "-fPIC",
]

MTX_OK, MTX_ERR_ARG, MTX_ERR_CUDA, MTX_ERR_UNSUPPORTED = 0, 1, 2, 3
SAMPLING = {"greedy": 0, "weighted": 1, "nucleus": 2, "topk": 3}


class ModelConfig(ctypes.Structure):
  _fields_ = [
      ("num_layers", ctypes.c_int32),
      ("emb_dim", ctypes.c_int32),
      ("num_q_heads", ctypes.c_int32),
      ("num_kv_heads", ctypes.c_int32),
      ("head_dim", ctypes.c_int32),
      ("mlp_dim", ctypes.c_int32),
      ("vocab_size", ctypes.c_int32),
      ("vocab_offset", ctypes.c_int32),
      ("max_prefill_len", ctypes.c_int32),
      ("max_target_len", ctypes.c_int32),
      ("num_slots", ctypes.c_int32),
      ("max_rows", ctypes.c_int32),
      ("rms_eps", ctypes.c_float),
      ("rope_min_timescale", ctypes.c_float),
      ("rope_max_timescale", ctypes.c_float),
      ("attn_softcap", ctypes.c_float),
      ("final_softcap", ctypes.c_float),
      ("logits_scale", ctypes.c_float),
      ("logits_round_bf16", ctypes.c_int32),
      ("embedding_rows", ctypes.c_int32),
      ("kv_quant", ctypes.c_int32),
      ("norm_scales_folded", ctypes.c_int32),
      ("decoder_block", ctypes.c_int32),
      ("sliding_window", ctypes.c_int32),
      ("local_rope_max_timescale", ctypes.c_float),
      ("query_scalar", ctypes.c_float),
      ("paged_num_pages", ctypes.c_int32),
      ("paged_tokens_per_page", ctypes.c_int32),
      ("paged_max_pages_per_group", ctypes.c_int32),
      ("paged_device_state", ctypes.c_int32),
  ]


class Weights(ctypes.Structure):
  _fields_ = [
      (name, ctypes.c_void_p)
      for name in ("embedding", "attn_norm", "wqkv", "wo", "mlp_norm", "w01", "wout", "final_norm", "logits",
                   "q_norm", "k_norm", "post_attn_norm", "post_ffw_norm")
  ]


class DecodeState(ctypes.Structure):
  _fields_ = [
      (name, ctypes.c_void_p)
      for name in (
          "k_cache",
          "v_cache",
          "tokens",
          "next_pos",
          "generated",
          "prefill_len",
          "ar_lengths",
          "ar_index",
          "result",
          "log_prob",
          "logits",
          "rng_state",
          "kq_cache",
          "vq_cache",
          "k_scale",
          "v_scale",
          "k_pages",
          "v_pages",
          "page_map",
          "page_lengths",
          "active_page",
          "active_page_pos",
          "page_status",
          "num_pages_used",
          "has_active_page",
      )
  ]


def sources() -> list[str]:
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cc"))) + [HEADER]


def _jaxlib_include() -> list[str]:
  """-I for jaxlib's headers when jaxlib is importable: csrc/mtx_jax_ffi.cc then builds the XLA FFI handlers."""
  try:
    import importlib.util

    spec = importlib.util.find_spec("jaxlib")
    if spec is None or not spec.submodule_search_locations:
      return []
    inc = os.path.join(list(spec.submodule_search_locations)[0], "include")
    return ["-I", inc] if os.path.isfile(os.path.join(inc, "xla", "ffi", "api", "ffi.h")) else []
  except Exception:
    return []


def needs_build() -> bool:
  if not os.path.isfile(LIB_PATH):
    return True
  built = os.path.getmtime(LIB_PATH)
  return any(os.path.getmtime(s) > built for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
  """Compile every CUDA source of the package into ``libmtx_b200.so`` (sm_100a only)."""
  if not force and not needs_build():
    return LIB_PATH
  nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
  defines = ["-D" + d for d in os.environ.get("MTX_NVCC_DEFINES", "").split() if d]  # e.g. MTX_PK_EVENTS for tools/mega_trace.py
  cmd = ([nvcc] + NVCC_FLAGS + defines + _jaxlib_include() + (["-Xptxas", "-v"] if verbose else []) +
         ["-o", LIB_PATH, os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "mtx_jax_ffi.cc")])
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
  if verbose:
    print(proc.stderr)
  return LIB_PATH


_lib = None


def _declare(lib) -> None:
  c = ctypes
  vp, i32, f32, sz = c.c_void_p, c.c_int, c.c_float, c.c_size_t
  lib.mtx_last_error.restype = c.c_char_p
  lib.mtx_last_error.argtypes = []
  lib.mtx_build_info.restype = c.c_char_p
  lib.mtx_build_info.argtypes = []
  lib.mtx_launch_count.restype = c.c_uint64
  lib.mtx_launch_count.argtypes = []
  lib.mtx_engine_create.restype = i32
  lib.mtx_engine_create.argtypes = [c.POINTER(ModelConfig), c.POINTER(vp)]
  lib.mtx_engine_destroy.restype = i32
  lib.mtx_engine_destroy.argtypes = [vp]
  lib.mtx_engine_workspace_bytes.restype = sz
  lib.mtx_engine_workspace_bytes.argtypes = [vp]
  lib.mtx_engine_bind.restype = i32
  lib.mtx_engine_bind.argtypes = [vp, c.POINTER(Weights), c.POINTER(DecodeState), vp, sz]
  lib.mtx_engine_set_sampling.restype = i32
  lib.mtx_engine_set_sampling.argtypes = [vp, i32, i32, f32, f32]
  lib.mtx_decode_step.restype = i32
  lib.mtx_decode_step.argtypes = [vp, i32, vp]
  lib.mtx_decode_step_graph.restype = i32
  lib.mtx_decode_step_graph.argtypes = [vp, i32, vp]
  lib.mtx_decode_step_host.restype = i32
  lib.mtx_decode_step_host.argtypes = [vp, i32, vp, vp, vp, vp]
  lib.mtx_decode_step_host_sync.restype = i32
  lib.mtx_decode_step_host_sync.argtypes = [vp, i32, vp, vp, vp, vp]
  lib.mtx_decode_step_candidates.restype = i32
  lib.mtx_decode_step_candidates.argtypes = [vp, i32, vp, vp]
  lib.mtx_sample_logits.restype = i32
  lib.mtx_sample_logits.argtypes = [vp, vp, i32, c.c_longlong, i32, i32, vp, vp, vp]
  lib.mtx_candidate_floats.restype = sz
  lib.mtx_candidate_floats.argtypes = [vp]
  lib.mtx_engine_counter.restype = i32
  lib.mtx_engine_counter.argtypes = [vp, i32, c.POINTER(c.c_longlong)]
  lib.mtx_commit_candidates.restype = i32
  lib.mtx_commit_candidates.argtypes = [vp, i32, vp, i32, vp]
  lib.mtx_debug_set_trace.restype = None
  lib.mtx_debug_set_trace.argtypes = [vp]
  lib.mtx_step_trace_words.restype = sz
  lib.mtx_step_trace_words.argtypes = [vp]
  lib.mtx_debug_set_timeline.restype = i32
  lib.mtx_debug_set_timeline.argtypes = [vp]
  lib.mtx_profile_decode_step.restype = i32
  lib.mtx_profile_decode_step.argtypes = [vp, i32, vp, c.POINTER(c.c_float), c.POINTER(c.c_int32)]
  lib.mtx_prefill_chunk.restype = i32
  lib.mtx_prefill_chunk.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
  lib.mtx_insert_prefix.restype = i32
  lib.mtx_insert_prefix.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
  lib.mtx_rmsnorm.restype = i32
  lib.mtx_rmsnorm.argtypes = [vp, vp, vp, i32, i32, f32, vp]
  lib.mtx_linear.restype = i32
  lib.mtx_linear.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
  lib.mtx_outproj_residual.restype = i32
  lib.mtx_outproj_residual.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
  lib.mtx_mlp_scratch_bytes.restype = sz
  lib.mtx_mlp_scratch_bytes.argtypes = [i32, i32, i32]
  lib.mtx_mlp.restype = i32
  lib.mtx_mlp.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, f32, vp, vp]
  lib.mtx_attention_scratch_bytes.restype = sz
  lib.mtx_attention_scratch_bytes.argtypes = [i32] * 6
  lib.mtx_decode_attention.restype = i32
  lib.mtx_decode_attention.argtypes = [vp] * 8 + [i32] * 7 + [f32, vp, vp]
  lib.mtx_ragged_attention_scratch_bytes.restype = sz
  lib.mtx_ragged_attention_scratch_bytes.argtypes = [i32] * 5
  lib.mtx_ragged_attention.restype = i32
  lib.mtx_ragged_attention.argtypes = [vp] * 7 + [i32] * 6 + [f32, vp, vp]
  lib.mtx_qkv_rope_append_scratch_bytes.restype = sz
  lib.mtx_qkv_rope_append_scratch_bytes.argtypes = [i32, i32]
  lib.mtx_qkv_rope_append.restype = i32
  lib.mtx_qkv_rope_append.argtypes = [vp] * 8 + [i32] * 6 + [f32, f32, vp, vp]
  lib.mtx_paged_append.restype = i32
  lib.mtx_paged_append.argtypes = [vp] * 6 + [i32] * 5 + [vp]
  lib.mtx_paged_attention_scratch_bytes.restype = sz
  lib.mtx_paged_attention_scratch_bytes.argtypes = [i32] * 5
  lib.mtx_paged_attention.restype = i32
  lib.mtx_paged_attention.argtypes = [vp] * 6 + [i32] * 7 + [f32, vp, vp]
  lib.mtx_page_update_decode.restype = i32
  lib.mtx_page_update_decode.argtypes = [vp] * 7 + [i32] * 4 + [vp]
  lib.mtx_paged_insert.restype = i32
  lib.mtx_paged_insert.argtypes = [vp] * 5 + [i32] * 7 + [vp]
  lib.mtx_jax_ffi_available.restype = i32
  lib.mtx_jax_ffi_available.argtypes = []
  lib.mtx_engine_rebind_state.restype = i32
  lib.mtx_engine_rebind_state.argtypes = [vp, c.POINTER(DecodeState)]


def load():
  """Load the library; raise if it has not been built (no fallback exists)."""
  global _lib
  if _lib is None:
    if not os.path.isfile(LIB_PATH):
      raise RuntimeError(
          f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
          "This package has no CPU or PyTorch implementation of the decode step."
      )
    lib = ctypes.CDLL(LIB_PATH)
    _declare(lib)
    _lib = lib
  return _lib


def last_error() -> str:
  return load().mtx_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
  """Map an MTX_ERR_* code to the exception the reference raises in the same situation."""
  if rc == MTX_OK:
    return
  msg = last_error()
  if rc in (MTX_ERR_ARG, MTX_ERR_UNSUPPORTED):
    raise ValueError(msg)  # the reference raises ValueError for bad config / sampling (SURVEY 8b "Errors")
  raise RuntimeError(msg)


def exported_symbols() -> list[str]:
  """Function names declared in include/mtx_b200.h."""
  import re

  text = open(HEADER, "r", encoding="utf-8").read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(mtx_[a-z0-9_]+)\s*\(", text)))


def require_cuda():
  import torch

  if not torch.cuda.is_available():
    raise RuntimeError("no CUDA device: the decode step only exists as sm_100a kernels")
  return torch.device("cuda", torch.cuda.current_device())
