"""int8 KV cache (MaxText/inference/kvcache.py:36-90, 658-736; configs/base.yml:104-112) on the B200 against the
oracle with the same quantiser: quantize_kvcache=True, kv_quant_dtype=int8, kv_quant_axis=dkv (one scale per token and kv head)
or heads_and_dkv (the reference's default: one scale per token over all kv heads).

Tolerances: logits vs the quantised dtype-faithful oracle rtol = atol = 1e-1 (the reference's ceiling, as for bf16); cache bytes
equal to the oracle's quantiser up to one step of the int8 grid on isolated elements (fp32 product order)."""

import numpy as np
import pytest
import torch

from maxtext_indextts2_b200 import maxengine, pyconfig
from oracle import decode_ref as ref
from oracle import mirror
from tests.helpers import make_params, random_tokens, small_config

pytestmark = pytest.mark.gpu

QUANT = dict(quantize_kvcache=True, kv_quant_dtype="int8", kv_quant_axis="dkv")
AXES = ["dkv", "heads_and_dkv", "fp8-dkv", "fp8-heads_and_dkv"]  # kv_quant_dtype int8 unless prefixed


def _quant(axis):
  if axis.startswith("fp8-"):
    return dict(QUANT, kv_quant_axis=axis[4:], kv_quant_dtype="fp8")
  return dict(QUANT, kv_quant_axis=axis)


def test_config_accepts_only_the_implemented_quantiser():
  assert pyconfig.initialize(None, head_dim=64, **QUANT).quantize_kvcache
  assert pyconfig.initialize(None, head_dim=64, quantize_kvcache=True).kv_quant_axis == "heads_and_dkv"  # the reference default
  with pytest.raises(ValueError, match="axis"):
    pyconfig.initialize(None, head_dim=64, quantize_kvcache=True, kv_quant_axis="heads")
  assert pyconfig.initialize(None, head_dim=64, quantize_kvcache=True, kv_quant_axis="dkv", kv_quant_dtype="fp8").kv_quant_dtype == "fp8"
  with pytest.raises(ValueError, match="kv_quant_dtype"):
    pyconfig.initialize(None, head_dim=64, quantize_kvcache=True, kv_quant_axis="dkv", kv_quant_dtype="int4")


@pytest.mark.parametrize("axis", AXES)
def test_prefill_insert_and_decode_with_int8_cache_match_the_quantised_oracle(axis):
  cfg = small_config(per_device_batch_size=3, max_prefill_predict_length=16, max_target_length=28, materialize_logits=True, **_quant(axis))
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=False)
  dparams = engine.load_params(params)
  prompts = random_tokens((3, 16), cfg.vocab_size, seed=12)
  ostate, state = oracle.init_decode_state(), engine.init_decode_state()
  assert state["cache"]["key"].dtype == torch.uint8 and state["cache"]["key_scale"].dtype == torch.float32
  for slot, n in enumerate((16, 5, 9)):
    oprefix, ofirst = oracle.prefill(prompts[slot], n)
    ostate = oracle.insert(oprefix, ostate, slot)
    prefix, _ = engine.prefill(params=dparams, padded_tokens=prompts[slot], true_length=n)
    torch.testing.assert_close(prefix["logits"].cpu()[0], oprefix["logits"][0], rtol=1e-1, atol=1e-1)
    prefix["tokens"].fill_(int(ofirst))
    state = engine.insert(prefix, state, slot)
    # the inserted rows: bytes and scales against the oracle's quantiser applied to the ENGINE's bf16 prefix
    for name, src in (("key", prefix["cache"]["key"]), ("value", prefix["cache"]["value"])):
      x = src.float().cpu()  # [L, Hkv, n, D]
      scale = x.abs().amax(-1)
      if axis.endswith("heads_and_dkv"):  # one scale per token: the maximum over the kv heads too, stored once per head
        scale = scale.amax(1, keepdim=True).expand_as(scale)
      got_s = state["cache"][name + "_scale"][:, slot, :, :n].cpu()
      torch.testing.assert_close(got_s, scale, rtol=0, atol=0)
      raw = state["cache"][name][:, slot, :, :n].cpu()
      if axis.startswith("fp8-"):
        want = (x * (448.0 / scale.clamp(min=1e-30))[..., None]).to(torch.float8_e4m3fn).to(torch.float32)
        got = raw.view(torch.float8_e4m3fn).to(torch.float32)
        d = (got - want).abs()
        # (x * (448 / scale) lands within an fp32 ulp of a rounding boundary for a few elements: one e4m3 step, <= 1/8 relative)
        assert (d <= 0.126 * want.abs().clamp(min=2.0**-6)).all() and (d > 0).float().mean() < 5e-3
        continue
      q = torch.clamp(torch.round(x * (127.5 / scale.clamp(min=1e-30))[..., None]), -128, 127)
      got_q = raw.to(torch.int32) - 128
      d = (got_q - q.to(torch.int32)).abs()
      assert d.max() <= 1 and (d > 0).float().mean() < 5e-3  # x * (127.5 / scale) lands within an fp32 ulp of a .5 boundary for ~0.1 % of the elements
  for step in range(16):  # ring of 12 rows: wraps
    ostate, odata = oracle.generate(ostate)
    state, result = engine.generate(dparams, state)
    torch.testing.assert_close(state["logits"].cpu(), ostate["logits"], rtol=1e-1, atol=1e-1)
    state["tokens"].copy_(odata[:, :1])
  assert int(state["cache"]["cache_ar_index"].item()) == 4


@pytest.mark.parametrize("axis", AXES)
@pytest.mark.parametrize("batch", [5, 64, 200])
def test_int8_decode_on_a_synthetic_cache_against_the_oracle(batch, axis):
  """Ragged contexts on a random int8 cache (the persistent kernel, or the 128-row-block GEMMs padded to 128 rows, 128, 256), the oracle
  following a few slots through oracle/mirror.py on the dequantised cache; also the bf16 engine on the dequantised cache for
  reference: the two CUDA paths must agree far more tightly than either does with the CPU oracle."""
  cfg = pyconfig.initialize(
      None, base_num_decoder_layers=3, base_emb_dim=384, base_num_query_heads=10, base_num_kv_heads=2, head_dim=64, base_mlp_dim=768,
      vocab_size=5000, per_device_batch_size=batch, max_prefill_predict_length=192, max_target_length=448, weight_dtype="bfloat16",
      attention="dot_product", scan_layers=False, materialize_logits=True, **_quant(axis))
  rng = np.random.Generator(np.random.PCG64(batch))
  pl = rng.integers(1, 193, size=batch)
  al = rng.integers(0, 240, size=batch)
  al[pl < 192] = 0
  pl[0], al[0] = 192, 239
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=False)
  dparams = engine.load_params(make_params(cfg))
  state = engine.fill_synthetic_context(pl, al, seed=3)
  slots = sorted({0, batch // 2, batch - 1})
  weights = mirror.oracle_weights_from_device(dparams, cfg)
  oracle = mirror.make_oracle(cfg, weights, len(slots), faithful=True)
  ostate = mirror.mirror_state(engine, oracle, slots)
  sl = torch.as_tensor(slots)
  for step in range(3):
    n0 = engine.lib.mtx_launch_count()
    state, result = engine.generate(dparams, state)
    launches = int(engine.lib.mtx_launch_count() - n0)
    # up to 64 rows with one scale per token and kv head run the persistent step kernel (step_persistent_kernel<1 / 2>: quantising
    # QKV epilogue, fp16 attention over byte tiles); heads_and_dkv and larger steps take the per-kernel path
    assert (launches == 3) == (batch <= 64 and axis in ("dkv", "fp8-dkv")), launches
    ostate, odata = oracle.generate(ostate)
    torch.testing.assert_close(state["logits"].cpu()[sl], ostate["logits"], rtol=1e-1, atol=1e-1)
    state["tokens"][sl.to(state["tokens"].device)] = odata[:, :1].to(state["tokens"].device)
  # the appended rows are quantised rows: every row written this run has its scale > 0 and a byte of magnitude 127 or 128
  idx = int(state["cache"]["cache_ar_index"].item())
  P = cfg.max_prefill_predict_length
  last = state["cache"]["key"][:, slots[0], :, P + idx - 1].cpu()
  if axis.startswith("fp8-"):  # magnitude 448 (byte 0x7e / 0xfe) instead of 127 / 128
    last = (last.view(torch.float8_e4m3fn).to(torch.float32).abs() >= 448).to(torch.int32) * 127
  else:
    last = last.to(torch.int32) - 128
  if axis in ("dkv", "fp8-dkv"):
    assert (last.abs().amax(-1) >= 127).all()
  else:  # one scale per token: the maximum sits in ONE of the heads, and the heads share the scale
    assert (last.abs().amax(dim=(-2, -1)) >= 127).all()
    sc = state["cache"]["key_scale"][:, slots[0], :, P + idx - 1].cpu()
    assert (sc == sc[:, :1]).all() and (sc > 0).all()


@pytest.mark.parametrize("hq,hkv", [(8, 1), (4, 4), (6, 2)])
@pytest.mark.parametrize("fp8", [False, True])
def test_quantised_persistent_kernel_head_groupings(hq, hkv, fp8):
  """step_persistent_kernel<1 / 2> (quantising QKV epilogue, transposed fp16 attention over byte tiles) with 8, 1 and 3 query heads
  per kv head, ragged contexts that wrap the ring, against the oracle with the same quantiser."""
  cfg = pyconfig.initialize(
      None, base_num_decoder_layers=2, base_emb_dim=256, base_num_query_heads=hq, base_num_kv_heads=hkv, head_dim=64, base_mlp_dim=512,
      vocab_size=3000, per_device_batch_size=24, max_prefill_predict_length=128, max_target_length=256, weight_dtype="bfloat16",
      attention="dot_product", scan_layers=False, materialize_logits=True, quantize_kvcache=True, kv_quant_axis="dkv",
      kv_quant_dtype="fp8" if fp8 else "int8")
  rng = np.random.Generator(np.random.PCG64(hq * 10 + hkv))
  pl = rng.integers(1, 129, size=24)
  al = rng.integers(0, 127, size=24)
  al[pl < 128] = 0
  pl[0], al[0] = 128, 126
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=False)  # (launches are counted when they are enqueued, not when a graph replays)
  dparams = engine.load_params(make_params(cfg))
  state = engine.fill_synthetic_context(pl, al, seed=9)
  slots = [0, 7, 23]
  weights = mirror.oracle_weights_from_device(dparams, cfg)
  oracle = mirror.make_oracle(cfg, weights, len(slots), faithful=True)
  ostate = mirror.mirror_state(engine, oracle, slots)
  sl = torch.as_tensor(slots)
  for step in range(4):
    n0 = engine.lib.mtx_launch_count()
    state, result = engine.generate(dparams, state)
    assert int(engine.lib.mtx_launch_count() - n0) == 3
    ostate, odata = oracle.generate(ostate)
    torch.testing.assert_close(state["logits"].cpu()[sl], ostate["logits"], rtol=1e-1, atol=1e-1)
    state["tokens"][sl.to(state["tokens"].device)] = odata[:, :1].to(state["tokens"].device)
