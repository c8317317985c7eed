"""Op-level parity on the B200: each fused op, called through the C ABI, against the oracle.

Tolerances: attention |d| <= 2^-7 * max|out| against the fp32-score reference
(one bf16 output rounding + bf16 probabilities), and the reference's own ceiling
rtol = atol = 1e-2 (MaxText/tests/attention_test.py:406); dense outputs within one bf16 ulp
of the fp32-accumulated product.
"""

import ctypes

import numpy as np
import pytest
import torch

from maxtext_indextts2_b200 import _lib
from oracle import decode_ref as ref

pytestmark = pytest.mark.gpu


def _stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
  return ctypes.c_void_p(t.data_ptr())


def _round_rows(rows):
  t = 16
  while t < rows:
    t *= 2
  return t


@pytest.mark.parametrize("rows,E", [(1, 128), (5, 256), (64, 1280)])
def test_rmsnorm(rows, E):
  lib = _lib.load()
  g = torch.Generator().manual_seed(0)
  x = (torch.randn(rows, E, generator=g) * 3).to(torch.bfloat16)
  scale = (1 + 0.1 * torch.randn(E, generator=g)).to(torch.bfloat16)
  out = torch.zeros(rows, E, dtype=torch.bfloat16, device="cuda")
  xd, sd = x.cuda(), scale.cuda()  # keep the device copies alive across the asynchronous launch
  _lib.check(lib.mtx_rmsnorm(_ptr(xd), _ptr(sd), _ptr(out), rows, E, 1e-5, _stream()))
  torch.cuda.synchronize()

  class C:
    normalization_layer_epsilon = 1e-5

  o = ref.DecodeOracle.__new__(ref.DecodeOracle)
  o.cfg, o.faithful = C, True
  want = o.rms_norm(x.float(), scale.float())
  got = out.cpu().float()
  # identical up to the last bf16 bit of the two roundings
  assert (got - want).abs().max() <= 2**-7 * want.abs().max()
  assert (got != want).float().mean() < 0.02


@pytest.mark.parametrize(
    "rows,n,k,splits",
    [(1, 256, 128, 1), (16, 512, 256, 1), (64, 1792, 1280, 1), (64, 1792, 1280, 8), (64, 1280, 5120, 16),
     (37, 384, 1280, 4), (200, 1280, 1280, 1), (256, 640, 320, 4), (64, 1000, 192, 2), (3, 256, 1280, 16),
     # 65..256 rows: gemm_rows.cuh -- persistent with several tiles per CTA (double-buffered accumulators, both row-block
     # counts), split-K over clusters of 2 / 4 / 8, ragged N
     (256, 24000, 256, 1), (130, 40000, 128, 1), (100, 1792, 1280, 8), (256, 1280, 5120, 8), (256, 1000, 192, 2), (65, 10240, 1280, 1),
     (256, 1280, 1280, 4)],
)
def test_linear_tcgen05(rows, n, k, splits):
  lib = _lib.load()
  g = torch.Generator().manual_seed(rows * 1000 + n)
  rt = _round_rows(rows)
  x = torch.zeros(rt, k, dtype=torch.bfloat16)
  x[:rows] = torch.randn(rows, k, generator=g).to(torch.bfloat16)
  w = (torch.randn(n, k, generator=g) / np.sqrt(k)).to(torch.bfloat16)
  out = torch.zeros(rows, n, dtype=torch.bfloat16, device="cuda")
  xd, wd = x.cuda(), w.cuda()
  for _ in range(2):
    out.zero_()
    _lib.check(lib.mtx_linear(_ptr(xd), _ptr(wd), _ptr(out), rows, n, k, splits, _stream()))
    torch.cuda.synchronize()
    want = x[:rows].float() @ w.float().t()
    got = out.cpu().float()
    err = (got - want).abs()
    assert err.max() <= 2**-7 * want.abs().clamp(min=1.0).max(), f"max err {err.max()}"
    # at most a last-bit difference from the bf16 rounding of an fp32 sum taken in another order
    assert torch.all(err <= 2**-7 * want.abs() + 1e-3)


def _attention_case(B, Hq, Hkv, D, P, T, len0, ring_first, ring_len, softcap=0.0, slots=None, seed=0):
  lib = _lib.load()
  slots = slots or B
  g = torch.Generator().manual_seed(seed)
  q = torch.randn(B, Hq * D, generator=g).to(torch.bfloat16)
  K = torch.randn(slots, Hkv, T, D, generator=g).to(torch.bfloat16)
  V = torch.randn(slots, Hkv, T, D, generator=g).to(torch.bfloat16)
  plane = torch.randperm(slots, generator=g)[:B].to(torch.int32)
  R = T - P
  valid = torch.zeros(B, T, dtype=torch.bool)
  for b in range(B):
    valid[b, : len0[b]] = True
    for i in range(ring_len[b]):
      valid[b, P + (ring_first[b] + i) % R] = True
  Kb = K[plane.long()].permute(0, 2, 1, 3).float()  # [B, T, Hkv, D]
  Vb = V[plane.long()].permute(0, 2, 1, 3).float()
  want, _, _ = ref.gqa_decode_ref(q.float().reshape(B, Hq, D), Kb, Vb, valid, softcap=softcap, p_bf16=True)
  want = want.reshape(B, Hq * D)
  out = torch.zeros(B, Hq * D, dtype=torch.bfloat16, device="cuda")
  nbytes = lib.mtx_attention_scratch_bytes(B, Hkv, Hq, D, P, T)
  scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  i32 = lambda v: torch.tensor(v, dtype=torch.int32, device="cuda")
  args = (q.cuda(), K.cuda(), V.cuda(), plane.cuda(), i32(len0), i32(ring_first), i32(ring_len))
  _lib.check(
      lib.mtx_decode_attention(
          _ptr(args[0]), _ptr(args[1]), _ptr(args[2]), _ptr(args[3]), _ptr(args[4]), _ptr(args[5]), _ptr(args[6]), _ptr(out),
          B, slots, Hq, Hkv, D, P, T, softcap, _ptr(scratch), _stream(),
      )
  )
  torch.cuda.synchronize()
  got = out.cpu().float()
  err = (got - want).abs().max().item()
  bound = 2**-7 * want.abs().max().item()
  assert err <= bound, f"max |d| {err} > {bound}"
  torch.testing.assert_close(got, want, rtol=1e-2, atol=1e-2)


def test_attention_gqa5_ragged():
  # IndexTTS2-scale head geometry: 20 query heads over 4 kv heads, D = 64
  _attention_case(
      B=6, Hq=20, Hkv=4, D=64, P=256, T=768,
      len0=[256, 1, 0, 100, 64, 200], ring_first=[0, 500, 3, 0, 0, 448], ring_len=[1, 40, 300, 0, 512, 130],
      slots=9,
  )


def test_attention_single_row_and_full_ring():
  _attention_case(B=2, Hq=4, Hkv=2, D=64, P=64, T=128, len0=[1, 64], ring_first=[0, 17], ring_len=[0, 64])


def test_attention_head_dim_128_mha_and_softcap():
  _attention_case(B=3, Hq=8, Hkv=8, D=128, P=128, T=320, len0=[128, 30, 0], ring_first=[0, 100, 150], ring_len=[5, 192, 77], softcap=30.0)


def test_attention_kernels_test_shape():
  # MaxText/tests/kernels_test.py:31-110 geometry: B=4, Hq=32, Hkv=8, D=128, S=512, random lengths
  rng = np.random.default_rng(3)
  lens = rng.integers(1, 512, size=4).tolist()
  _attention_case(B=4, Hq=32, Hkv=8, D=128, P=512, T=576, len0=lens, ring_first=[0] * 4, ring_len=[0] * 4)


def test_attention_long_context_many_chunks():
  # 4k prompt + 1.5k decoded (BASELINE config C4 lengths) for one kv head group
  _attention_case(B=2, Hq=5, Hkv=1, D=64, P=4096, T=5632, len0=[4000, 4096], ring_first=[0, 1000], ring_len=[1536, 1536])


def test_unsupported_shapes_fail_loudly():
  lib = _lib.load()
  t = torch.zeros(16, dtype=torch.int32, device="cuda")
  rc = lib.mtx_decode_attention(_ptr(t), _ptr(t), _ptr(t), _ptr(t), _ptr(t), _ptr(t), _ptr(t), _ptr(t), 1, 1, 4, 2, 80, 8, 16, 0.0,
                                _ptr(t), _stream())
  assert rc == _lib.MTX_ERR_UNSUPPORTED and "head_dim" in _lib.last_error()
  rc = lib.mtx_linear(_ptr(t), _ptr(t), _ptr(t), 4, 128, 100, 1, _stream())
  assert rc == _lib.MTX_ERR_ARG
