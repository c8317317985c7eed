"""Registration of the library's XLA FFI handlers with JAX (``csrc/mtx_jax_ffi.cc``).

The reference's host code is Python/JAX; with jax + jaxlib installed and the library built against jaxlib's
headers, :func:`register` makes every handler an XLA custom-call target and the helpers below wrap them as
jit-compatible functions with the reference's own call shapes (``AttentionOp.gpu_ragged_attention``,
MaxText/layers/attentions.py:761-815).  jax is NOT installed in this image (and cannot be: no network), so
everything here raises a clear error instead of importing it at module load; the ctypes path (``_lib.py``)
is what the tests exercise.
"""

from __future__ import annotations

import ctypes

from . import _lib

TARGETS = {
    "mtx_ragged_attention": "MtxRaggedAttention",
    "mtx_decode_attention": "MtxDecodeAttention",
    "mtx_qkv_rope_append": "MtxQkvRopeAppend",
    "mtx_decode_step": "MtxDecodeStep",
    "mtx_outproj_residual": "MtxOutprojResidual",
    "mtx_mlp": "MtxMlp",
    "mtx_paged_append": "MtxPagedAppend",
    "mtx_paged_attention": "MtxPagedAttention",
}


def available() -> bool:
  """True when the loaded library was built with the XLA FFI handlers and jax can be imported."""
  try:
    import jax  # noqa: F401
  except Exception:
    return False
  return bool(_lib.load().mtx_jax_ffi_available())


def register() -> None:
  """jax.ffi.register_ffi_target(name, capsule, platform="CUDA") for every handler."""
  if not available():
    raise RuntimeError(
        "XLA FFI handlers unavailable: needs jax/jaxlib importable and libmtx_b200.so built with jaxlib's headers "
        "(csrc/mtx_jax_ffi.cc compiles to a stub without xla/ffi/api/ffi.h)")
  import jax

  lib = _lib.load()
  for target, symbol in TARGETS.items():
    jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")


def ragged_attention(q, k, v, lengths, *, seq_major: bool = True, softcap: float = 0.0):
  """Drop-in body for ``AttentionOp.gpu_ragged_attention``'s wrapped kernel call: returns (out, max, sum)."""
  import jax
  import jax.numpy as jnp

  b, _, hq, d = q.shape
  seq, hkv = (k.shape[1], k.shape[2]) if seq_major else (k.shape[2], k.shape[1])
  nbytes = int(_lib.load().mtx_ragged_attention_scratch_bytes(b, hkv, hq, d, seq))
  scratch = jnp.zeros((nbytes,), jnp.uint8)
  out_t = jax.ShapeDtypeStruct(q.shape, q.dtype)
  stat_t = jax.ShapeDtypeStruct((b, hq), jnp.float32)
  out, m, l = jax.ffi.ffi_call("mtx_ragged_attention", (out_t, stat_t, stat_t))(
      q, k, v, lengths, scratch, seq_major=ctypes.c_int32(int(seq_major)).value, softcap=float(softcap))
  return out, m.reshape(b, 1, hq, 1), l.reshape(b, 1, hq, 1)
