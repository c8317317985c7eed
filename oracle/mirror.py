"""ORACLE helper -- copies the state of a CUDA ``MaxEngine`` into a ``DecodeOracle``.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (see the header of ``decode_ref.py``): only
``tests/``, ``__graft_entry__.smoke()`` and the checker legs of ``bench.py`` (``cpu_baseline``,
``--impl reference``, ``verify``) import it.

Rows (decode slots) never interact inside a step (attentions.py:587-588 masks per slot), so an
oracle holding a SUBSET of the engine's slots, with the same shared ring index, reproduces those
slots exactly.  That keeps the CPU side of a parity check at BASELINE scale (24 layers, V = 264,192,
P = 1024 / T = 3072 and larger) to seconds per step.
"""

from __future__ import annotations

import numpy as np
import torch

from maxtext_indextts2_b200 import pyconfig

from . import decode_ref as ref


def oracle_weights_from_device(dparams, cfg) -> ref.OracleWeights:
  """Packed K-major device tensors (include/mtx_b200.h ``mtx_weights``) -> matmul-ready fp32 oracle weights."""
  dparams = getattr(dparams, "source", None) or dparams  # the oracle stays on the UNFOLDED weights (maxengine.fold_norm_scales)
  t = dparams.tensors
  E, Hq, Hkv, D, M = cfg.emb_dim, cfg.num_query_heads, cfg.num_kv_heads, cfg.head_dim, cfg.mlp_dim
  f = lambda x: x.detach().to("cpu").to(torch.float32)
  layers = []
  for l in range(cfg.num_decoder_layers):
    wqkv = f(t["wqkv"][l])  # [(Hq+2Hkv)D, E]
    w01 = f(t["w01"][l]).reshape(M // 16, 2, 16, E)  # rows interleaved 16 at a time: (gate, value)
    layers.append(
        dict(
            attn_scale=f(t["attn_norm"][l]),
            wq=wqkv[: Hq * D].t().contiguous(),
            wk=wqkv[Hq * D : (Hq + Hkv) * D].t().contiguous(),
            wv=wqkv[(Hq + Hkv) * D :].t().contiguous(),
            wo=f(t["wo"][l]).t().contiguous(),
            mlp_scale=f(t["mlp_norm"][l]),
            w0=w01[:, 0].reshape(M, E).t().contiguous(),
            w1=w01[:, 1].reshape(M, E).t().contiguous(),
            wout=f(t["wout"][l]).t().contiguous(),
        )
    )
    if cfg.decoder_block == "gemma3":
      layers[-1].update(q_norm=f(t["q_norm"][l]), k_norm=f(t["k_norm"][l]), post_attn_scale=f(t["post_attn_norm"][l]),
                        post_ffw_scale=f(t["post_ffw_norm"][l]))
  logits = None if cfg.logits_via_embedding else f(t["logits"]).t().contiguous()
  return ref.OracleWeights(embedding=f(t["embedding"]), layers=layers, final_scale=f(t["final_norm"]), logits=logits)


def make_oracle(cfg, weights: ref.OracleWeights, slots: int, faithful: bool = True) -> ref.DecodeOracle:
  """A DecodeOracle over ready-made weights holding `slots` slots of the engine's configuration."""
  keys = cfg.get_keys()
  ocfg = pyconfig.HyperParameters({**keys, "per_device_batch_size": slots})
  o = ref.DecodeOracle.__new__(ref.DecodeOracle)
  o.cfg, o.faithful, o.w = ocfg, faithful, weights
  o.B, o.P, o.T = slots, cfg.max_prefill_predict_length, cfg.max_target_length
  o.R = o.T - o.P
  # (cfg carries quantize_kvcache: DecodeOracle.kv_quant follows it)
  o.scores_f32 = bool(cfg.float32_qk_product) or not faithful
  o.softmax_f32 = o.scores_f32 or bool(cfg.float32_logits)
  return o


def mirror_state(engine, oracle: ref.DecodeOracle, slots) -> dict:
  """decode_state of `oracle` equal to the engine's current state for the given slot ids.

  Engine layout (maxengine.py ``_alloc_state``): K/V [L, planes, Hkv, T, D] bf16, rows [0,P) prefill segment,
  [P,T) the AR ring; ``prefill_length[s]`` active prefill rows; ``cached_ar_lengths[s]`` rows appended since
  insert; ``cache_ar_index`` the shared ring index.  Oracle layout: the reference's (kvcache.py:343-484).
  """
  slots = [int(s) for s in slots]
  assert len(slots) == oracle.B
  P, R, L = oracle.P, oracle.R, oracle.cfg.num_decoder_layers
  state = oracle.init_decode_state()
  c = state["cache"]
  if getattr(engine, "_paged", False):
    return _mirror_paged(engine, oracle, slots, state)
  sl = torch.as_tensor(slots, device=engine._k.device)
  quant = bool(getattr(engine, "_kv_quant", False))
  for l in range(L):
    if quant:  # int8 cache: the oracle holds the dequantised values q * scale / 127.5 (kvcache.py:76-90)
      ks = engine._k_scale[l].index_select(0, sl).to("cpu")[..., None]
      vs = engine._v_scale[l].index_select(0, sl).to("cpu")[..., None]
      if getattr(engine, "_kv_fp8", False):  # float8_e4m3fn bytes: value * scale / 448
        k = engine._kq[l].index_select(0, sl).to("cpu").view(torch.float8_e4m3fn).to(torch.float32) * (ks / 448.0)
        v = engine._vq[l].index_select(0, sl).to("cpu").view(torch.float8_e4m3fn).to(torch.float32) * (vs / 448.0)
      else:
        k = (engine._kq[l].index_select(0, sl).to("cpu").to(torch.float32) - 128.0) * (ks / 127.5)
        v = (engine._vq[l].index_select(0, sl).to("cpu").to(torch.float32) - 128.0) * (vs / 127.5)
    else:
      k = engine._k[l].index_select(0, sl).to("cpu").to(torch.float32)  # [n, Hkv, T, D]
      v = engine._v[l].index_select(0, sl).to("cpu").to(torch.float32)
    c["prefill_key"][l] = k[:, :, :P].permute(0, 2, 1, 3).contiguous()
    c["prefill_value"][l] = v[:, :, :P].permute(0, 2, 1, 3).contiguous()
    c["ar_key"][l] = k[:, :, P:].permute(0, 2, 1, 3).contiguous()
    c["ar_value"][l] = v[:, :, P:].permute(0, 2, 1, 3).contiguous()
  plen = engine._prefill_len.cpu()[slots].to(torch.int64)
  alen = engine._ar_lengths.cpu()[slots].to(torch.int64)
  idx = int(engine._ar_index.item())
  c["prefill_segment_id"] = (torch.arange(P)[None, :] < plen[:, None]).to(torch.int32) * ref.ACTIVE
  # ring rows written since the slot's insert: the min(ar_lengths, R) rows before the shared index
  back = (idx - 1 - torch.arange(R)[None, :]) % R  # back[j] = ring row written j+1 steps ago
  seg = torch.zeros(len(slots), R, dtype=torch.int32)
  for i in range(len(slots)):
    n = int(min(int(alen[i]), R))
    seg[i, back[0, :n]] = ref.ACTIVE
  c["ar_segment_id"] = seg
  c["ar_index"] = idx
  c["ar_lengths"] = alen.to(torch.int32)
  state["next_pos"] = engine._next_pos.cpu()[slots].clone()
  state["generated_tokens"] = engine._generated.cpu()[slots].clone()
  state["tokens"] = engine._tokens.cpu()[slots].clone()
  return state


def _mirror_paged(engine, oracle: ref.DecodeOracle, slots, state: dict) -> dict:
  """attention=paged: a sequence is the tokens [0, sequence_lengths[s]) of page group s, gathered from the pools
  [L, Hkv, num_pages, tokens_per_page, D] through page_map (inference/page_manager.py:49-91).  The dense-cache oracle attends the
  same keys wherever they are stored: the first P tokens go to its prefill segment, the rest to ring rows ending at a shared
  ring index (no wrap: the caller keeps sequence_lengths - P below the ring size)."""
  P, R, L = oracle.P, oracle.R, oracle.cfg.num_decoder_layers
  c = state["cache"]
  ps = engine.page_state
  tpp = engine.page_manager.tokens_per_page
  lengths = [int(ps.sequence_lengths[s]) for s in slots]
  extra = [max(0, n - P) for n in lengths]
  idx = max(extra)
  assert idx < R, "the mirrored sequences do not fit the oracle's ring"
  for i, (s, n) in enumerate(zip(slots, lengths)):
    pages = torch.as_tensor(ps.page_map[s, : (n + tpp - 1) // tpp].astype(np.int64), device=engine._k_pages.device)
    k = engine._k_pages.index_select(2, pages).to("cpu").to(torch.float32)  # [L, Hkv, pages, tpp, D]
    v = engine._v_pages.index_select(2, pages).to("cpu").to(torch.float32)
    k = k.reshape(L, k.shape[1], -1, k.shape[-1])[:, :, :n].permute(0, 2, 1, 3)  # [L, n, Hkv, D]
    v = v.reshape(L, v.shape[1], -1, v.shape[-1])[:, :, :n].permute(0, 2, 1, 3)
    n0 = min(n, P)
    for l in range(L):
      c["prefill_key"][l][i, :n0] = k[l, :n0]
      c["prefill_value"][l][i, :n0] = v[l, :n0]
      if n > P:
        c["ar_key"][l][i, idx - extra[i] : idx] = k[l, P:]
        c["ar_value"][l][i, idx - extra[i] : idx] = v[l, P:]
    c["prefill_segment_id"][i, :n0] = ref.ACTIVE
    c["ar_segment_id"][i, idx - extra[i] : idx] = ref.ACTIVE
  c["ar_index"] = idx
  c["ar_lengths"] = torch.as_tensor(extra, dtype=torch.int32)
  state["next_pos"] = engine._next_pos.cpu()[slots].clone()
  state["generated_tokens"] = engine._generated.cpu()[slots].clone()
  state["tokens"] = engine._tokens.cpu()[slots].clone()
  return state


def bf16_ulp(x: float) -> float:
  """Spacing of bfloat16 at |x| (8 significand bits)."""
  x = abs(float(x))
  if x == 0.0:
    return 2.0**-133
  return 2.0 ** (int(np.floor(np.log2(x))) - 7)


def classify_mismatch(row_f32: torch.Tensor, got: int, want: int) -> dict:
  """A greedy mismatch judged on the fp32 oracle's logits of that row: the margin between the
  two candidates in bf16 ulps of the top logit.  `strict` = SURVEY's rule (below one ulp)."""
  top = float(row_f32.max())
  margin = abs(float(row_f32[want]) - float(row_f32[got]))
  ulp = bf16_ulp(top)
  return {"margin": margin, "ulps": margin / ulp, "top": top, "strict": margin < ulp}
