"""CPU-side checks of the boundary: the library builds for sm_100a, loads, and exports every
symbol include/mtx_b200.h declares; argument validation that needs no GPU; config behaviour."""

import ctypes
import os
import subprocess

import pytest

from maxtext_indextts2_b200 import _lib, pyconfig


@pytest.fixture(scope="module")
def lib():
  _lib.build()
  return _lib.load()


def test_header_symbols_are_exported(lib):
  names = _lib.exported_symbols()
  assert len(names) >= 14 and "mtx_decode_step" in names and "mtx_decode_attention" in names
  for n in names:
    assert hasattr(lib, n), f"{n} declared in include/mtx_b200.h but not exported"


def _check_sass(lib):
  sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
  assert "sm_100a" in sass
  for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):  # tcgen05.mma, TMA load, tcgen05.ld
    assert mnemonic in sass, mnemonic
  assert b"sm_100a" in lib.mtx_build_info()


def test_built_for_sm_100a_with_tcgen05_and_tma(lib):
  """(per-kernel counts: tools/sass_summary.py -> profiles/sass_summary.txt)"""
  _check_sass(lib)


@pytest.mark.gpu
def test_library_loaded_on_the_gpu_box_is_the_sm_100a_build(lib):
  """The same check in the `-m gpu` run, on the library the GPU tests actually load."""
  _check_sass(lib)


def test_engine_create_validates_without_a_gpu(lib):
  cfg = _lib.ModelConfig(num_layers=2, emb_dim=128, num_q_heads=4, num_kv_heads=2, head_dim=80, mlp_dim=256, vocab_size=512,
                         max_prefill_len=8, max_target_len=16, num_slots=2, max_rows=2, rms_eps=1e-5, rope_min_timescale=1,
                         rope_max_timescale=10000, logits_scale=1.0, logits_round_bf16=1)
  h = ctypes.c_void_p()
  assert lib.mtx_engine_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.MTX_ERR_UNSUPPORTED
  assert "head_dim" in _lib.last_error()
  cfg.head_dim = 64
  cfg.max_rows = 1000
  assert lib.mtx_engine_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.MTX_ERR_ARG
  cfg.max_rows = 2
  assert lib.mtx_engine_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.MTX_OK
  assert lib.mtx_engine_workspace_bytes(h) > 0
  # not bound -> every compute call refuses
  assert lib.mtx_decode_step(h, 2, None) == _lib.MTX_ERR_ARG
  assert lib.mtx_engine_set_sampling(h, 3, 0, 0.0, 1.0) == _lib.MTX_ERR_ARG  # topk with k = 0 (inference_utils.py:106-107)
  assert lib.mtx_engine_set_sampling(h, 9, 0, 0.0, 1.0) == _lib.MTX_ERR_ARG
  assert lib.mtx_engine_destroy(h) == _lib.MTX_OK


def test_no_cpu_fallback():
  import torch

  from maxtext_indextts2_b200 import maxengine

  if torch.cuda.is_available():
    pytest.skip("GPU present")
  with pytest.raises(RuntimeError, match="no CUDA device"):
    maxengine.MaxEngine(pyconfig.initialize(None))


def test_config_keys_and_derived_values(monkeypatch):
  cfg = pyconfig.initialize(None, model_name="indextts2-t2s", per_device_batch_size=64)
  assert (cfg.emb_dim, cfg.num_query_heads, cfg.num_kv_heads, cfg.head_dim, cfg.mlp_dim) == (1280, 20, 4, 64, 5120)
  assert cfg.num_decoder_layers == 24 and cfg.vocab_size == 264192 and cfg.attn_logits_soft_cap is None
  cfg = pyconfig.initialize(["prog", pyconfig.BASE_YML, "head_dim=64", "decode_sampling_strategy=topk", "decode_sampling_top_k=5"])
  assert cfg.head_dim == 64 and cfg.decode_sampling_top_k == 5
  monkeypatch.setenv("M_VOCAB_SIZE", "4096")
  assert pyconfig.initialize(None).vocab_size == 4096
  with pytest.raises(ValueError):
    pyconfig.initialize(None, vocab_size=1)  # CLI and ENV for the same key (pyconfig.py:414-421)
  monkeypatch.delenv("M_VOCAB_SIZE")
  with pytest.raises(ValueError):
    pyconfig.initialize(None, quantize_kvcache=True)
  with pytest.raises(ValueError):
    pyconfig.initialize(None, max_target_length=64, max_prefill_predict_length=64)
  with pytest.raises(ValueError):
    cfg.head_dim = 1


def test_xla_ffi_handlers_type_check_against_the_stub_header():
  """csrc/mtx_jax_ffi.cc is compile-guarded on jaxlib's xla/ffi/api/ffi.h, which this image does not have: the handlers are
  type-checked against tests/stubs/xla/ffi/api/ffi.h instead (same names; XLA_FFI_DEFINE_HANDLER_SYMBOL static_asserts that a
  handler is invocable with its bound context, arguments, results and attributes, in order).  Not a substitute for building with
  jaxlib, but it keeps the never-compiled branch from rotting."""
  import shutil
  import subprocess

  gxx = shutil.which("g++")
  if gxx is None:
    pytest.skip("no g++")
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  src = os.path.join(root, "maxtext_indextts2_b200", "csrc", "mtx_jax_ffi.cc")
  cuda_inc = "/usr/local/cuda/include"
  proc = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(root, "tests", "stubs"), "-I", cuda_inc, src],
                        capture_output=True, text=True)
  assert proc.returncode == 0, proc.stderr[-3000:]
  text = open(src, encoding="utf-8").read()
  from maxtext_indextts2_b200 import jax_ffi

  for symbol in jax_ffi.TARGETS.values():  # every target jax_ffi.register() looks up is defined by the source
    assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({symbol}," in text
