"""ORACLE -- CPU restatement of the reference's autoregressive decode path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; the product package never does and has no CPU path.

What it restates (file:line are relative to the reference tree
HyperBlaze456/maxtext-indextts2; the recipe is SURVEY.md Appendix A):

* one decode step        MaxText/maxengine.py:868-936  (``_generate_jit``)
* decode-state layout     MaxText/maxengine.py:1370-1427
* prefill                 MaxText/maxengine.py:400-530
* insert                  MaxText/maxengine.py:1045-1164
* embedding gather        MaxText/layers/embeddings.py:131-163
* RMSNorm                 MaxText/layers/normalizations.py:57-69
* dense projections       MaxText/layers/linears.py:188-232
* RoPE                    MaxText/layers/embeddings.py:270-315
* KV cache (two segments) MaxText/inference/kvcache.py:584-624, 626-718, 738-795
* attention               MaxText/layers/attentions.py:1206-1271 (dot), :1157-1201
                          (local max/exp/sum), :1376-1397 (merge), :587-588 (AR mask),
                          :96-118 (mask application), :1399-1466 (two-segment call)
* llama2 block            MaxText/layers/llama2.py:54-165
* gemma3 block            MaxText/layers/gemma3.py:36-197 (5 local : 1 global pattern, post norms),
                          attentions.py:2246-2265 (q/k RMSNorm before RoPE, query scalar after),
                          :2085-2088 (local RoPE base), :604-631 (sliding-window mask; in
                          AUTOREGRESSIVE mode the window is taken over CACHE INDICES of each segment,
                          next_pos = kv_seq_len - 1), linears.py:40-51,460 (gelu = flax nn.gelu, tanh form)
* gated MLP               MaxText/layers/linears.py:425-476
* output head             MaxText/layers/decoders.py:537-589
* sampling                MaxText/inference_utils.py:55-111

Two numeric modes share the code:

* ``faithful=True``  ("O-ref"): rounds to bfloat16 after every op the reference
  rounds after (activations are bf16 arrays there), fp32 accumulation inside dots.
* ``faithful=False`` ("O-f32"): the same graph with fp32 activations throughout
  (weights are still the bf16 values the reference casts them to at use).

PARITY STATUS.  The reference is Python/JAX and cannot be imported here (no jax,
flax, jetstream wheels; no network), and it stores no golden vectors for this path.
This oracle is therefore pinned by the reference's own *invariants*
(tests/test_oracle_invariants.py restates them): AR steps == full-sequence
forward (MaxText/tests/attention_test.py:361-406, model_test.py:119-191), RoPE vs
a complex-number implementation (MaxText/tests/llama_test.py:79-183), GQA decode
vs ``reference_gqa`` (MaxText/kernels/ragged_attention.py:122-161), state shapes
(MaxText/tests/maxengine_test.py:111-164).  RNG-dependent sampling
(``weighted`` / ``topk`` / ``nucleus``) draws from ``jax.random.categorical`` in the
reference, whose bit stream lives in jaxlib: for those strategies
**parity is unpinned** -- the algorithm (Gumbel-max over the same candidate set) is
restated, the random bits are this repo's Philox stream.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

DEFAULT_MASK_VALUE = -0.7 * float(np.finfo(np.float32).max)  # common_types.py:74
NEG_INF = -1.0e7  # inference_utils.py:20
ACTIVE = 1  # DECODING_ACTIVE_SEQUENCE_INDICATOR, common_types.py:70


def _bf16(x: torch.Tensor) -> torch.Tensor:
  return x.to(torch.bfloat16).to(torch.float32)


@dataclass
class OracleWeights:
  """bf16-valued fp32 copies of the parameters in matmul-ready 2-D shapes."""

  embedding: torch.Tensor  # [V, E]
  layers: list  # dicts: attn_scale, wq [E,Hq*D], wk, wv [E,Hkv*D], wo [Hq*D,E], mlp_scale, w0, w1 [E,M], wout [M,E]
  final_scale: torch.Tensor  # [E]
  logits: torch.Tensor | None  # [E, V] or None when tied


def prepare_weights(params: dict, config) -> OracleWeights:
  """Cast every kernel to the activation dtype as the reference does at use (linears.py:216)."""
  from maxtext_indextts2_b200.params import unscan_params  # layout helper only (scan_layers=True trees)

  p = unscan_params(params, config)["params"]
  E, Hq, Hkv, D = config.emb_dim, config.num_query_heads, config.num_kv_heads, config.head_dim
  f = lambda t: t.to(torch.bfloat16).to(torch.float32)
  layers = []
  gemma3 = config.decoder_block == "gemma3"
  for i in range(config.num_decoder_layers):
    lp = p["decoder"][f"layers_{i}"]
    sa = lp["self_attention"]
    layers.append(
        dict(
            attn_scale=f(lp["pre_self_attention_norm" if gemma3 else "pre_self_attention_layer_norm"]["scale"]),
            wq=f(sa["query"]["kernel"]).reshape(E, Hq * D),
            wk=f(sa["key"]["kernel"]).reshape(E, Hkv * D),
            wv=f(sa["value"]["kernel"]).reshape(E, Hkv * D),
            wo=f(sa["out"]["kernel"]).reshape(Hq * D, E),
            mlp_scale=f(lp["mlp"]["mlp_layer_norm"]["scale"]),
            w0=f(lp["mlp"]["wi_0"]["kernel"]),
            w1=f(lp["mlp"]["wi_1"]["kernel"]),
            wout=f(lp["mlp"]["wo"]["kernel"]),
        )
    )
    if gemma3:
      layers[-1].update(
          q_norm=f(sa["query_norm"]["scale"]),
          k_norm=f(sa["key_norm"]["scale"]),
          post_attn_scale=f(lp["post_self_attention_norm"]["scale"]),
          post_ffw_scale=f(lp["post_ffw_norm"]["scale"]),
      )
  logits = None if config.logits_via_embedding else f(p["decoder"]["logits_dense"]["kernel"])
  return OracleWeights(
      embedding=f(p["token_embedder"]["embedding"]),
      layers=layers,
      final_scale=f(p["decoder"]["decoder_norm"]["scale"]),
      logits=logits,
  )


class DecodeOracle:
  """Restates MaxEngine.prefill / insert / generate for the llama2 block and the gemma3 block."""

  @property
  def gemma3(self) -> bool:
    return self.cfg.decoder_block == "gemma3"

  def is_local(self, layer: int) -> bool:
    """gemma3.py:36-48: five sliding-window layers, then one global layer."""
    return self.gemma3 and layer % 6 != 5

  def query_scalar(self) -> float:
    """gemma3.py:51-58."""
    cfg = self.cfg
    if cfg.model_name in ("gemma3-4b", "gemma3-12b"):
      return cfg.head_dim**-0.5
    if cfg.model_name == "gemma3-27b":
      return (cfg.base_emb_dim // cfg.base_num_query_heads) ** -0.5
    raise ValueError(f"Unsupported model name: {cfg.model_name}")

  def __init__(self, config, params: dict, faithful: bool = True):
    self.cfg = config
    self.faithful = faithful
    self.w = prepare_weights(params, config)
    self.B = int(config.per_device_batch_size)
    self.P = config.max_prefill_predict_length
    self.T = config.max_target_length
    self.R = self.T - self.P  # AR ring length, kvcache.py:412
    # precision of the attention scores / softmax (attentions.py:1227-1262)
    self.scores_f32 = bool(config.float32_qk_product) or not faithful
    self.softmax_f32 = self.scores_f32 or bool(config.float32_logits)

  def kv_quant(self, x):
    """KVQuant.quantize + dequantisation (inference/kvcache.py:76-90; int8 or fp8; kv_quant_axis "dkv": one scale per token and kv
    head, or "heads_and_dkv"): scale = max|x| over the head's dims, q = int8(rint(x * 127.5 / scale)) (the conversion saturates: the row's maximum
    maps to rint(127.5) = 128 -> 127), cached value = q * scale / 127.5.  Identity unless quantize_kvcache."""
    if not getattr(self.cfg, "quantize_kvcache", False):
      return x
    # kvcache.py:66-73: "dkv" takes the maximum over the head's dims, "heads_and_dkv" over the kv heads as well (x [..., Hkv, D])
    scale = x.abs().amax(dim=(-2, -1) if self.cfg.kv_quant_axis == "heads_and_dkv" else -1, keepdim=True)
    if self.cfg.kv_quant_dtype == "fp8":
      # kvcache.py:38,86-88: value = float8_e4m3fn(x * (E4M3_MAX / scale)), E4M3_MAX = 448; dequantised value * scale / 448
      q = torch.where(scale > 0, (x * (448.0 / scale.clamp(min=1e-30))).to(torch.float8_e4m3fn).to(torch.float32), torch.zeros_like(x))
      return q * (scale / 448.0)
    q = torch.where(scale > 0, torch.clamp(torch.round(x * (127.5 / scale.clamp(min=1e-30))), -128.0, 127.0), torch.zeros_like(x))
    return q * (scale / 127.5)

  # -- small ops ---------------------------------------------------------------

  def r(self, x):
    return _bf16(x) if self.faithful else x

  def rms_norm(self, x, scale):
    """normalizations.py:57-69."""
    x32 = x.to(torch.float32)
    mean2 = torch.mean(x32 * x32, dim=-1, keepdim=True)
    y = self.r(x32 * torch.rsqrt(mean2 + self.cfg.normalization_layer_epsilon))
    return self.r(y * scale)

  def dense(self, x, w, out_f32: bool = False):
    """linears.py:188-232: bf16 operands, fp32 accumulate, bf16 result."""
    y = x @ w
    return y if out_f32 else self.r(y)

  def rope(self, x, pos, max_timescale=None):
    """embeddings.py:270-315.  x [B,T,H,D]; pos [B,T] int."""
    D = x.shape[-1]
    half = D // 2
    # timescale = min * (max/min) ** (2i/D) (embeddings.py:270-275), evaluated in fp64 and
    # rounded once to fp32 so that every libm gives the same table
    fraction = 2 * torch.arange(0, half, dtype=torch.float64) / D
    lo, hi = float(self.cfg.rope_min_timescale), float(max_timescale or self.cfg.rope_max_timescale)
    timescale = (lo * (hi / lo) ** fraction).to(torch.float32)
    sinusoid = pos.to(torch.float32)[:, :, None, None] / timescale
    sin = self.r(torch.sin(sinusoid))
    cos = self.r(torch.cos(sinusoid))
    a, b = x[..., :half], x[..., half:]
    first = self.r(self.r(a * cos) - self.r(b * sin))
    second = self.r(self.r(b * cos) + self.r(a * sin))
    return torch.cat((first, second), dim=-1)

  # -- attention ---------------------------------------------------------------

  def _local_attention(self, q, K, V, mask):
    """apply_attention_dot + compute_local_attention (attentions.py:1206-1271, 1157-1201).

    q [B,T,Hq,D]; K,V [B,S,Hkv,D]; mask bool [B,T,S] (True = attend).
    Returns unnormalised out [B,T,Hq,D], max [B,T,Hq,1], sum [B,T,Hq,1].
    """
    B, T, Hq, D = q.shape
    Hkv = K.shape[2]
    G = Hq // Hkv
    rs = (lambda t: t) if self.scores_f32 else self.r  # rounding of the scores
    rm = (lambda t: t) if self.softmax_f32 else self.r  # rounding inside the softmax
    qg = q.reshape(B, T, Hkv, G, D)
    s = rs(torch.einsum("btkgd,bskd->bkgts", qg, K))
    cap = self.cfg.attn_logits_soft_cap
    if cap:
      s = rs(rs(torch.tanh(rs(s / cap))) * cap)
    s = torch.where(mask[:, None, None, :, :], s, torch.tensor(DEFAULT_MASK_VALUE))
    m = torch.amax(s, dim=-1, keepdim=True)
    e = rm(torch.exp(rm(s - m)))
    l = rm(torch.sum(e, dim=-1, keepdim=True))
    o = rm(torch.einsum("bkgts,bskd->btkgd", e, V)).reshape(B, T, Hq, D)
    # moveaxis(-2,1) + reshape of attentions.py:1179-1185: [b,k,g,t,1] -> [b,t,k*g,1]
    m = m.permute(0, 3, 1, 2, 4).reshape(B, T, Hq, 1)
    l = l.permute(0, 3, 1, 2, 4).reshape(B, T, Hq, 1)
    return o, m, l

  def _normalize_attention(self, outs, maxes, sums):
    """attentions.py:1376-1397."""
    rm = (lambda t: t) if self.softmax_f32 else self.r
    gmax = maxes[0]
    for m in maxes[1:]:
      gmax = torch.maximum(gmax, m)
    gsum = 0
    for m, l in zip(maxes, sums):
      gsum = rm(gsum + rm(rm(torch.exp(rm(m - gmax))) * l))
    out = 0
    for m, o in zip(maxes, outs):
      w = rm(rm(torch.exp(rm(m - gmax))) / gsum)
      out = rm(out + rm(w * o))
    return out

  # -- the block ----------------------------------------------------------------

  def _mlp(self, lw, h):
    """linears.py:425-476 with mlp_activations [silu, linear] (llama2) or [gelu, linear] (gemma3)."""
    n = self.rms_norm(h, lw["mlp_scale"])
    a = self.dense(n, lw["w0"])
    if self.gemma3:
      # flax nn.gelu = jax.nn.gelu(approximate=True): x * 0.5 * (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) on a bf16 array; the
      # cdf factor is evaluated in fp32 and rounded once, then the product is rounded (as silu below)
      cdf = 0.5 * (1.0 + torch.tanh(0.7978845608028654 * (a + 0.044715 * a * a * a)))
      a = self.r(a * self.r(cdf))
    else:
      a = self.r(a * self.r(torch.sigmoid(a)))  # jax.nn.silu on a bf16 array
    b = self.dense(n, lw["w1"])
    return self.dense(self.r(a * b), lw["wout"])

  def _qkv(self, lw, n, pos, layer: int = 0):
    cfg = self.cfg
    B, T, _ = n.shape
    q = self.dense(n, lw["wq"]).reshape(B, T, cfg.num_query_heads, cfg.head_dim)
    k = self.dense(n, lw["wk"]).reshape(B, T, cfg.num_kv_heads, cfg.head_dim)
    v = self.dense(n, lw["wv"]).reshape(B, T, cfg.num_kv_heads, cfg.head_dim)
    if not self.gemma3:
      return self.rope(q, pos), self.rope(k, pos), v
    # attentions.py:2246-2265: RMSNorm over head_dim, RoPE (local layers: their own base, :2085-2088), query scalar
    q, k = self.rms_norm(q, lw["q_norm"]), self.rms_norm(k, lw["k_norm"])
    ts = None
    if self.is_local(layer) and float(cfg.local_rope_max_timescale) > 0:
      ts = float(cfg.local_rope_max_timescale)
    q, k = self.rope(q, pos, ts), self.rope(k, pos, ts)
    sc = self.query_scalar()
    if sc and sc != 1.0:
      q = self.r(q * sc)
    return q, k, v

  def _block(self, lw, x, a):
    """Everything of a decoder layer after the attention op: llama2.py:139-165 / gemma3.py:132-176.  a [B,T,Hq*D]."""
    o = self.dense(a, lw["wo"])
    if self.gemma3:
      o = self.rms_norm(o, lw["post_attn_scale"])
    h = self.r(o + x)
    m = self._mlp(lw, h)
    if self.gemma3:
      m = self.rms_norm(m, lw["post_ffw_scale"])
    return self.r(m + h)

  def _window_full(self, T: int):
    """attentions.py:624-631 for a full sequence (next_pos = 0): col in (row - W, row]."""
    W = int(self.cfg.sliding_window_size)
    row, col = torch.arange(T)[:, None], torch.arange(T)[None, :]
    return (col > row - W) & (col <= row)

  def _window_ar(self, S: int):
    """attentions.py:600-602,624-631 in AUTOREGRESSIVE mode: next_pos = kv_seq_len - 1 of the SEGMENT, so the window is the last
    sliding_window_size cache indices of the prefill segment and of the AR ring, whatever the true positions are."""
    W = int(self.cfg.sliding_window_size)
    col = torch.arange(S)
    return (col > S - 1 - W) & (col <= S - 1)

  def _output_head(self, x):
    """decoders.py:537-589."""
    cfg = self.cfg
    y = self.rms_norm(x, self.w.final_scale)
    if cfg.logits_via_embedding:
      logits = self.dense(y, self.w.embedding.t(), out_f32=bool(cfg.logits_dot_in_fp32))
      if cfg.normalize_embedding_logits:
        logits = logits / math.sqrt(y.shape[-1])
      if cfg.final_logits_soft_cap:
        logits = torch.tanh(logits / cfg.final_logits_soft_cap) * cfg.final_logits_soft_cap
    else:
      logits = self.dense(y, self.w.logits, out_f32=bool(cfg.logits_dot_in_fp32))
    return logits.to(torch.float32)

  def forward_full(self, tokens, segment_ids=None, return_kv: bool = False):
    """Full-sequence causal forward (TRAIN / PREFILL graph, no AR cache).

    tokens [B,T] int64. segment_ids [B,T] or None. Attention is normalised in place
    (attentions.py:1438-1441 ``out / sum``).
    """
    B, T = tokens.shape
    pos = torch.arange(T)[None, :].expand(B, T)
    x = self.w.embedding[tokens]
    causal = torch.tril(torch.ones(T, T, dtype=torch.bool))[None]
    if segment_ids is not None:
      mask = causal & (segment_ids[:, :, None] == segment_ids[:, None, :])
    else:
      mask = causal.expand(B, T, T)
    kvs = []
    rm = (lambda t: t) if self.softmax_f32 else self.r
    for li, lw in enumerate(self.w.layers):
      n = self.rms_norm(x, lw["attn_scale"])
      q, k, v = self._qkv(lw, n, pos, li)
      kvs.append((k, v))
      lmask = mask & self._window_full(T)[None] if self.is_local(li) else mask
      o, _, l = self._local_attention(q, k, v, lmask)
      a = rm(o / l)
      x = self._block(lw, x, self.r(a).reshape(B, T, -1))
    logits = self._output_head(x)
    return (logits, kvs) if return_kv else logits

  # -- engine-level API -----------------------------------------------------------

  def init_decode_state(self) -> dict:
    """maxengine.py:1370-1427: everything zero."""
    cfg, B, P, R = self.cfg, self.B, self.P, self.R
    Hkv, D, L = cfg.num_kv_heads, cfg.head_dim, cfg.num_decoder_layers
    cache = {
        "prefill_key": [torch.zeros(B, P, Hkv, D) for _ in range(L)],
        "prefill_value": [torch.zeros(B, P, Hkv, D) for _ in range(L)],
        "ar_key": [torch.zeros(B, R, Hkv, D) for _ in range(L)],
        "ar_value": [torch.zeros(B, R, Hkv, D) for _ in range(L)],
        "prefill_segment_id": torch.zeros(B, P, dtype=torch.int32),
        "ar_segment_id": torch.zeros(B, R, dtype=torch.int32),
        "ar_index": 0,
        "ar_lengths": torch.zeros(B, dtype=torch.int32),
    }
    return {
        "logits": torch.zeros(B, 1, cfg.vocab_size),
        "cache": cache,
        "next_pos": torch.zeros(B, 1, dtype=torch.int32),
        "generated_tokens": torch.zeros(B, 1, dtype=torch.int32),
        "tokens": torch.zeros(B, 1, dtype=torch.int32),
    }

  def prefill(self, padded_tokens, true_length: int, sampler=None):
    """maxengine.py:400-530.  padded_tokens [P'] ints, P' <= max_prefill_predict_length."""
    cfg = self.cfg
    toks = torch.as_tensor(padded_tokens, dtype=torch.int64)[None, :]
    Pp = toks.shape[1]
    seg = (torch.arange(Pp) < true_length).to(torch.int32)[None, :] * ACTIVE
    logits, kvs = self.forward_full(toks, seg, return_kv=True)
    selected = logits[:, true_length - 1 : true_length, :]
    sampler = sampler or (lambda lg: sampling(lg, "greedy"))
    first = sampler(selected).to(torch.int32).reshape(1, 1)
    P = self.P
    pk, pv = [], []
    for k, v in kvs:
      kk = torch.zeros(1, P, cfg.num_kv_heads, cfg.head_dim)
      vv = torch.zeros_like(kk)
      kk[:, :Pp], vv[:, :Pp] = k, v
      pk.append(kk)
      pv.append(vv)
    segp = torch.zeros(1, P, dtype=torch.int32)
    segp[:, :Pp] = seg
    prefix = {
        "logits": selected,
        "cache": {"prefill_key": pk, "prefill_value": pv, "prefill_segment_id": segp},
        "next_pos": torch.full((1, 1), true_length, dtype=torch.int32),
        "generated_tokens": torch.zeros(1, 1, dtype=torch.int32),
        "tokens": first,
    }
    return prefix, first

  def insert(self, prefix: dict, state: dict, slot: int) -> dict:
    """maxengine.py:1045-1164: AR data and the shared ring index are left untouched."""
    c, pc = state["cache"], prefix["cache"]
    for l in range(self.cfg.num_decoder_layers):
      # (kv_cache_prefill stores the quantised prefill keys / values, kvcache.py:610-617; what is inserted is that cache)
      c["prefill_key"][l][slot] = self.kv_quant(pc["prefill_key"][l][0])
      c["prefill_value"][l][slot] = self.kv_quant(pc["prefill_value"][l][0])
    c["prefill_segment_id"][slot] = pc["prefill_segment_id"][0]
    c["ar_segment_id"][slot] = 0
    c["ar_lengths"][slot] = 0
    state["logits"][slot] = prefix["logits"][0]
    state["next_pos"][slot] = prefix["next_pos"][0]
    state["generated_tokens"][slot] = prefix["generated_tokens"][0]
    state["tokens"][slot] = prefix["tokens"][0]
    return state

  def step_logits(self, state: dict) -> torch.Tensor:
    """Model.apply in AUTOREGRESSIVE mode (maxengine.py:884-893); mutates the cache."""
    cfg, c = self.cfg, state["cache"]
    B = state["tokens"].shape[0]
    tokens = state["tokens"].to(torch.int64)
    pos = state["next_pos"]
    idx = int(c["ar_index"])
    x = self.w.embedding[tokens]  # [B,1,E]
    # kvcache.py:774-777: mark the ring slot active before attention reads it
    c["ar_segment_id"][:, idx] = ACTIVE
    pmask = (c["prefill_segment_id"] == ACTIVE)[:, None, :]
    amask = (c["ar_segment_id"] == ACTIVE)[:, None, :]
    for l, lw in enumerate(self.w.layers):
      n = self.rms_norm(x, lw["attn_scale"])
      q, k, v = self._qkv(lw, n, pos, l)
      c["ar_key"][l][:, idx] = self.kv_quant(k[:, 0])  # kvcache.py:696-701 (quantised when quantize_kvcache: :658-718)
      c["ar_value"][l][:, idx] = self.kv_quant(v[:, 0])
      pm, am = pmask, amask
      if self.is_local(l):
        pm = pmask & self._window_ar(self.P)[None, None, :]
        am = amask & self._window_ar(self.R)[None, None, :]
      o_p, m_p, l_p = self._local_attention(q, c["prefill_key"][l], c["prefill_value"][l], pm)
      o_a, m_a, l_a = self._local_attention(q, c["ar_key"][l], c["ar_value"][l], am)
      a = self._normalize_attention([o_p, o_a], [m_p, m_a], [l_p, l_a])
      x = self._block(lw, x, self.r(a).reshape(B, 1, -1))
    c["ar_index"] = (idx + 1) % self.R  # kvcache.py:778
    c["ar_lengths"] += 1  # kvcache.py:779
    return self._output_head(x)

  def generate(self, state: dict, sampler=None):
    """maxengine.py:868-936.  Returns (state, ResultTokens.data [B,3])."""
    logits = self.step_logits(state)
    sampler = sampler or (lambda lg: sampling(lg, "greedy"))
    new_token = sampler(logits).to(torch.int32).reshape(-1, 1)
    state["logits"] = logits
    state["next_pos"] = state["next_pos"] + 1
    state["generated_tokens"] = state["generated_tokens"] + 1
    state["tokens"] = new_token
    data = torch.cat((new_token, torch.ones_like(new_token), state["generated_tokens"]), dim=1)
    return state, data


# ---------------------------------------------------------------------------------
# Op-level references used by the kernel parity tests
# ---------------------------------------------------------------------------------


def gqa_decode_ref(q, K, V, valid, softcap: float = 0.0, p_bf16: bool = True):
  """GQA decode attention over the valid rows, fp32 scores.

  Follows kernels/ragged_attention.py:122-161 (``reference_gqa``): fp32 logits,
  exp(s - max), probabilities cast to V's dtype before the PV product, output
  divided by the denominator.  q [B,Hq,D]; K,V [B,S,Hkv,D] (bf16-valued fp32);
  valid bool [B,S].  Returns out [B,Hq,D] fp32 (un-rounded), max [B,Hq], sum [B,Hq].
  """
  B, Hq, D = q.shape
  Hkv = K.shape[2]
  G = Hq // Hkv
  qg = q.reshape(B, Hkv, G, D).to(torch.float32)
  s = torch.einsum("bkgd,bskd->bkgs", qg, K.to(torch.float32))
  if softcap:
    s = torch.tanh(s / softcap) * softcap
  s = torch.where(valid[:, None, None, :], s, torch.tensor(DEFAULT_MASK_VALUE))
  m = s.amax(dim=-1)
  e = torch.exp(s - m[..., None])
  e = torch.where(valid[:, None, None, :], e, torch.zeros(()))
  l = e.sum(dim=-1)
  ep = _bf16(e) if p_bf16 else e
  o = torch.einsum("bkgs,bskd->bkgd", ep, V.to(torch.float32)) / l[..., None]
  return o.reshape(B, Hq, D), m.reshape(B, Hq), l.reshape(B, Hq)


# ---------------------------------------------------------------------------------
# Sampling (inference_utils.py:55-111)
# ---------------------------------------------------------------------------------

_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
  """Philox-4x32-10 (Salmon et al. 2011), vectorised over uint32 arrays."""
  c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
  c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
  k0 = np.uint32(k0)
  k1 = np.uint32(k1)
  mask = np.uint64(0xFFFFFFFF)
  with np.errstate(over="ignore"):
    for _ in range(10):
      p0 = c0.astype(np.uint64) * _PHILOX_M0
      p1 = c2.astype(np.uint64) * _PHILOX_M1
      hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
      hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
      c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
      k0 = np.uint32((int(k0) + int(_PHILOX_W0)) & 0xFFFFFFFF)
      k1 = np.uint32((int(k1) + int(_PHILOX_W1)) & 0xFFFFFFFF)
  return c0, c1, c2, c3


def gumbel_noise(seed: int, step: int, rows: np.ndarray, vocab: int) -> np.ndarray:
  """Gumbel(0,1) noise g[row, v] of this repo's sampler (fp32).

  counter = (v // 4, row, step, 0), key = (seed_lo, seed_hi); word v % 4 of the Philox
  output; u = ((x >> 9) + 0.5) * 2^-23 in (0,1); g = -log(-log(u)).
  """
  v4 = np.arange((vocab + 3) // 4, dtype=np.uint32)
  out = np.empty((len(rows), v4.shape[0] * 4), dtype=np.float32)
  for i, row in enumerate(rows):
    w = philox4x32_10(v4, np.uint32(row), np.uint32(step), np.uint32(0), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    x = np.stack(w, axis=1).reshape(-1)
    u = ((x >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0**-23)
    out[i] = -np.log(-np.log(u, dtype=np.float32), dtype=np.float32)
  return out[:, :vocab]


def nucleus_cutoff(logits: torch.Tensor, p: float) -> torch.Tensor:
  """inference_utils.py:96-99: the logit at the first sorted position whose cumulative
  softmax mass reaches p (temperature is NOT applied here).  logits [N,V] -> [N,1]."""
  srt = torch.sort(logits, dim=-1, descending=True).values
  cum = torch.cumsum(torch.softmax(srt, dim=-1), dim=-1)
  idx = torch.sum(cum < p, dim=-1, keepdim=True)
  idx = idx.clamp(max=logits.shape[-1] - 1)  # jnp.take_along_axis clamps out-of-range indices
  return torch.gather(srt, -1, idx)


def sampling(
    logits: torch.Tensor,
    algorithm: str,
    topk: int = 0,
    nucleus_topp: float = 0.0,
    temperature: float = 1.0,
    seed: int = 0,
    step: int = 0,
    row_offset: int = 0,
    return_scores: bool = False,
):
  """inference_utils.py:66-111.  logits [..., V] fp32 -> token ids [...].

  ``jax.random.categorical(rng, x)`` is restated as argmax(x + Gumbel) (its
  definition); the Gumbel bits come from :func:`gumbel_noise`.
  """
  lead = logits.shape[:-1]
  V = logits.shape[-1]
  lg = logits.reshape(-1, V).to(torch.float32)
  N = lg.shape[0]
  if algorithm == "greedy":
    out = torch.argmax(lg, dim=-1)  # first maximum wins, as jnp.argmax
    return out.reshape(lead)
  rows = np.arange(N, dtype=np.uint32) + np.uint32(row_offset)
  g = torch.from_numpy(gumbel_noise(seed, step, rows, V))
  if algorithm == "weighted":
    scores = lg / temperature + g
  elif algorithm == "nucleus":
    if nucleus_topp < 0:
      raise ValueError("Can't apply nucleus with parameter {nucleus_topp=} less zero")
    cutoff = nucleus_cutoff(lg, nucleus_topp)
    masked = torch.where(lg < cutoff, torch.full_like(lg, NEG_INF), lg)
    scores = masked / temperature + g
  elif algorithm == "topk":
    if topk <= 0:
      raise ValueError("Can't apply algorithm topk with parameter {topk=} less than or equal to zero")
    # lax.top_k keeps the k largest; ties resolved towards the lower index
    order = torch.sort(lg, dim=-1, descending=True, stable=True).indices[:, :topk]
    keep = torch.zeros_like(lg, dtype=torch.bool)
    keep.scatter_(1, order, True)
    scores = torch.where(keep, lg / temperature + g, torch.full_like(lg, -float("inf")))
  else:
    raise ValueError(f"Sampling {algorithm=} not supported!")
  out = torch.argmax(scores, dim=-1).reshape(lead)
  return (out, scores) if return_scores else out


def log_prob_of_chosen_token(logits: torch.Tensor, chosen: torch.Tensor) -> torch.Tensor:
  """inference_utils.py:55-63."""
  logps = torch.log_softmax(logits.to(torch.float32), dim=-1)
  return torch.gather(logps, -1, chosen.to(torch.int64)[..., None])[..., 0]
