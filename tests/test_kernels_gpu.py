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


@pytest.mark.parametrize("seq_major", [1, 0])
@pytest.mark.parametrize("B,Hq,Hkv,D,S", [(4, 20, 4, 64, 300), (3, 8, 8, 128, 130), (2, 5, 1, 64, 2048)])
def test_ragged_attention_returns_unnormalised_out_max_sum(B, Hq, Hkv, D, S, seq_major):
  """The contract of AttentionOp.gpu_ragged_attention (attentions.py:761-815): (q, k, v, lengths) -> (unnormalised out, max,
  sum) for ONE segment, in the reference's logical cache layout [B,S,Hkv,D] and in this library's [B,Hkv,S,D]; checked against
  reference_gqa's arithmetic (kernels/ragged_attention.py:122-161) and by merging two segments the way the caller does
  (normalize_attention, attentions.py:1376-1397)."""
  lib = _lib.load()
  g = torch.Generator().manual_seed(B * 100 + S)
  q = torch.randn(B, Hq * D, generator=g).to(torch.bfloat16)
  K = torch.randn(B, S, Hkv, D, generator=g).to(torch.bfloat16)
  V = torch.randn(B, S, Hkv, D, generator=g).to(torch.bfloat16)
  lengths = torch.randint(1, S + 1, (B,), generator=g).to(torch.int32)
  lengths[0] = S
  valid = torch.arange(S)[None, :] < lengths[:, None]
  want_o, want_m, want_l = ref.gqa_decode_ref(q.float().reshape(B, Hq, D), K.float(), V.float(), valid, p_bf16=True)
  kd = (K if seq_major else K.permute(0, 2, 1, 3).contiguous()).cuda()
  vd = (V if seq_major else V.permute(0, 2, 1, 3).contiguous()).cuda()
  qd, ld = q.cuda(), lengths.cuda()
  out = torch.zeros(B, Hq * D, dtype=torch.bfloat16, device="cuda")
  om = torch.zeros(B, Hq, dtype=torch.float32, device="cuda")
  ol = torch.zeros(B, Hq, dtype=torch.float32, device="cuda")
  scratch = torch.empty(lib.mtx_ragged_attention_scratch_bytes(B, Hkv, Hq, D, S), dtype=torch.uint8, device="cuda")
  _lib.check(lib.mtx_ragged_attention(_ptr(qd), _ptr(kd), _ptr(vd), _ptr(ld), _ptr(out), _ptr(om), _ptr(ol), B, S, Hq, Hkv, D, seq_major, 0.0,
                                      _ptr(scratch), _stream()))
  torch.cuda.synchronize()
  got_o = out.cpu().float().reshape(B, Hq, D)
  got_m, got_l = om.cpu(), ol.cpu()
  torch.testing.assert_close(got_m, want_m, rtol=1e-5, atol=1e-5)
  torch.testing.assert_close(got_l, want_l, rtol=2e-3, atol=1e-3)  # exp2-based exponentials, fp32 sums in tile order
  norm = got_o / got_l[..., None]
  assert (norm - want_o).abs().max() <= 2**-6 * want_o.abs().max()  # one more bf16 rounding than the normalised kernel (of the unnormalised sums)
  # the caller's merge (normalize_attention, attentions.py:1376-1397) of two calls over the two halves of the sequence equals the
  # single call, for the rows that reach into the second half
  S1 = S // 2
  la = torch.clamp(lengths, max=S1)
  lb = torch.clamp(lengths - S1, min=1)
  both = lengths > S1

  def call(kk, vv, ll, seq):
    o_ = torch.zeros(B, Hq * D, dtype=torch.bfloat16, device="cuda")
    m_ = torch.zeros(B, Hq, dtype=torch.float32, device="cuda")
    l_ = torch.zeros(B, Hq, dtype=torch.float32, device="cuda")
    kk, vv, ll = kk.contiguous().cuda(), vv.contiguous().cuda(), ll.cuda()
    _lib.check(lib.mtx_ragged_attention(_ptr(qd), _ptr(kk), _ptr(vv), _ptr(ll), _ptr(o_), _ptr(m_), _ptr(l_), B, seq, Hq, Hkv, D, 1, 0.0,
                                        _ptr(scratch), _stream()))
    torch.cuda.synchronize()
    return o_.cpu().float().reshape(B, Hq, D), m_.cpu(), l_.cpu()

  oa, ma, sa = call(K[:, :S1], V[:, :S1], la, S1)
  ob, mb, sb = call(K[:, S1:], V[:, S1:], lb, S - S1)
  gmax = torch.maximum(ma, mb)
  wa, wb = torch.exp(ma - gmax), torch.exp(mb - gmax)
  merged = (wa[..., None] * oa + wb[..., None] * ob) / (wa * sa + wb * sb)[..., None]
  assert (merged[both] - want_o[both]).abs().max() <= 2**-6 * want_o.abs().max()


@pytest.mark.parametrize("rows,E,Hq,Hkv,D", [(5, 256, 4, 2, 64), (64, 1280, 20, 4, 64), (200, 512, 8, 8, 128), (256, 1280, 20, 4, 64)])
def test_qkv_rope_append_op(rows, E, Hq, Hkv, D):
  """Attention.query/key/value + RotaryEmbedding + KVCache append as one fused op through the C ABI, against the oracle's
  dense + rope (embeddings.py:277-315) and the cache rows it must have written (kvcache.py:626-718)."""
  lib = _lib.load()
  g = torch.Generator().manual_seed(rows + E)
  rt = _round_rows(rows)
  n = torch.zeros(rt, E, dtype=torch.bfloat16)
  n[:rows] = torch.randn(rows, E, generator=g).to(torch.bfloat16)
  qkv_n = (Hq + 2 * Hkv) * D
  w = (torch.randn(qkv_n, E, generator=g) / np.sqrt(E)).to(torch.bfloat16)
  planes, T = rows + 3, 40
  pos = torch.randint(0, 3000, (rows,), generator=g).to(torch.int32)
  plane = torch.randperm(planes, generator=g)[:rows].to(torch.int32)
  write_row = torch.randint(0, T, (rows,), generator=g).to(torch.int32)
  write_row[rows // 2] = -1  # skipped
  kc = torch.zeros(planes, Hkv, T, D, dtype=torch.bfloat16, device="cuda")
  vc = torch.zeros_like(kc)
  qo = torch.zeros(rows, Hq * D, dtype=torch.bfloat16, device="cuda")
  scratch = torch.empty(lib.mtx_qkv_rope_append_scratch_bytes(rows, D), dtype=torch.uint8, device="cuda")
  nd, wd, pd, pld, wrd = n.cuda(), w.cuda(), pos.cuda(), plane.cuda(), write_row.cuda()
  _lib.check(lib.mtx_qkv_rope_append(_ptr(nd), _ptr(wd), _ptr(pd), _ptr(pld), _ptr(wrd), _ptr(qo), _ptr(kc), _ptr(vc), rows, E, Hq, Hkv, D, T,
                                     1.0, 10000.0, _ptr(scratch), _stream()))
  torch.cuda.synchronize()

  class C:
    rope_min_timescale, rope_max_timescale = 1, 10000

  o = ref.DecodeOracle.__new__(ref.DecodeOracle)
  o.cfg, o.faithful = C, True
  y = o.r(n[:rows].float() @ w.float().t())
  q = o.rope(y[:, : Hq * D].reshape(rows, 1, Hq, D), pos.reshape(rows, 1))[:, 0]
  k = o.rope(y[:, Hq * D : (Hq + Hkv) * D].reshape(rows, 1, Hkv, D), pos.reshape(rows, 1))[:, 0]
  v = y[:, (Hq + Hkv) * D :].reshape(rows, Hkv, D)
  tol = dict(rtol=2e-2, atol=2e-2)  # bf16 products of the rotation on top of the last-bit difference of the dense output
  torch.testing.assert_close(qo.cpu().float().reshape(rows, Hq, D), q, **tol)
  kcc, vcc = kc.cpu().float(), vc.cpu().float()
  written = torch.zeros(planes, T, dtype=torch.bool)
  for r in range(rows):
    if write_row[r] < 0:
      continue
    written[plane[r], write_row[r]] = True
    torch.testing.assert_close(kcc[plane[r], :, write_row[r]], k[r], **tol)
    torch.testing.assert_close(vcc[plane[r], :, write_row[r]], v[r], **tol)
  assert kcc.permute(0, 2, 1, 3)[~written].abs().max() == 0 and vcc.permute(0, 2, 1, 3)[~written].abs().max() == 0  # nothing else touched


class _OracleOps:
  """DecodeOracle's rounding rules (faithful mode) without a model around them."""

  def __init__(self, eps=1e-5):
    class C:
      normalization_layer_epsilon = eps
      decoder_block = "llama2"

    self.o = ref.DecodeOracle.__new__(ref.DecodeOracle)
    self.o.cfg, self.o.faithful, self.o.scores_f32, self.o.softmax_f32 = C, True, True, True


@pytest.mark.parametrize("rows,E,HD", [(1, 256, 256), (37, 1280, 1280), (64, 1280, 1280), (200, 1280, 1280), (256, 640, 512)])
def test_outproj_residual_matches_oracle(rows, E, HD):
  """mtx_outproj_residual: out = x + dense(attn, wo) (attentions.py out projection + llama2.py:139-140), bf16 roundings as the oracle's."""
  lib = _lib.load()
  g = torch.Generator().manual_seed(rows + E)
  rt = _round_rows(rows)
  attn = torch.zeros(rt, HD, dtype=torch.bfloat16)
  attn[:rows] = torch.randn(rows, HD, generator=g).to(torch.bfloat16)
  wo = (torch.randn(E, HD, generator=g) / np.sqrt(HD)).to(torch.bfloat16)
  x = torch.randn(rows, E, generator=g).to(torch.bfloat16)
  ad, wd, xd = attn.cuda(), wo.cuda(), x.cuda()
  out = torch.zeros(rows, E, dtype=torch.bfloat16, device="cuda")
  _lib.check(lib.mtx_outproj_residual(_ptr(ad), _ptr(wd), _ptr(xd), _ptr(out), rows, E, HD, _stream()))
  torch.cuda.synchronize()
  o = _OracleOps().o
  want = o.r(x.float() + o.dense(attn[:rows].float(), wo.float().t()))
  got = out.cpu().float()
  err = (got - want).abs()
  assert err.max() <= 2**-6 * want.abs().max()  # two bf16 roundings (projection, sum), fp32 sums in another order
  assert (err > 2**-8 * want.abs().clamp(min=1.0)).float().mean() < 0.02


@pytest.mark.parametrize("rows,E,M", [(3, 256, 512), (64, 1280, 5120), (130, 1280, 5120), (256, 384, 768)])
def test_mlp_block_matches_oracle(rows, E, M):
  """mtx_mlp: out = h + MlpBlock(RMSNorm(h)) (linears.py:425-476, llama2.py:150-163) against DecodeOracle._mlp on the same weights."""
  lib = _lib.load()
  g = torch.Generator().manual_seed(rows * 7 + M)
  rt = _round_rows(rows)
  h = torch.zeros(rt, E, dtype=torch.bfloat16)
  h[:rows] = torch.randn(rows, E, generator=g).to(torch.bfloat16)
  scale = (1 + 0.1 * torch.randn(E, generator=g)).to(torch.bfloat16)
  w0 = (torch.randn(E, M, generator=g) / np.sqrt(E)).to(torch.bfloat16)  # [E, M] as the reference's wi_0 kernel
  w1 = (torch.randn(E, M, generator=g) / np.sqrt(E)).to(torch.bfloat16)
  wout = (torch.randn(M, E, generator=g) / np.sqrt(M)).to(torch.bfloat16)
  # mtx_weights.w01: rows of wi_0^T / wi_1^T interleaved in groups of 16; wout^T
  w01 = torch.stack([w0.t().reshape(M // 16, 16, E), w1.t().reshape(M // 16, 16, E)], dim=1).reshape(2 * M, E).contiguous()
  hd, sd, w01d, woutd = h.cuda(), scale.cuda(), w01.cuda(), wout.t().contiguous().cuda()
  out = torch.zeros(rows, E, dtype=torch.bfloat16, device="cuda")
  scratch = torch.empty(lib.mtx_mlp_scratch_bytes(rows, E, M), dtype=torch.uint8, device="cuda")
  _lib.check(lib.mtx_mlp(_ptr(hd), _ptr(sd), _ptr(w01d), _ptr(woutd), _ptr(out), rows, E, M, 1e-5, _ptr(scratch), _stream()))
  torch.cuda.synchronize()
  o = _OracleOps().o
  lw = dict(mlp_scale=scale.float(), w0=w0.float(), w1=w1.float(), wout=wout.float())
  hf = h[:rows].float()
  want = o.r(hf + o._mlp(lw, hf))
  got = out.cpu().float()
  err = (got - want).abs()
  assert err.max() <= 2**-5 * want.abs().max()  # four bf16 roundings deep, hardware exp in the sigmoid
  assert err.mean() <= 2**-9 * want.abs().mean() + 1e-3
