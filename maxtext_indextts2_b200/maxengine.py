"""MaxEngine for the B200: the reference's engine API over the sm_100a decode kernels.

Mirrors ``MaxText/maxengine.py`` (class ``MaxEngine``, :100) for the decode path:

=====================  ==========================================  =====================
method                 reference                                    here
=====================  ==========================================  =====================
``load_params``        maxengine.py:218                             repack for K-major streaming, upload
``init_decode_state``  maxengine.py:1370-1453                       zero the device state
``prefill``            maxengine.py:533-574 (``_prefill_jit`` :400)  ``mtx_prefill_chunk``
``insert``             maxengine.py:1166-1192 (``_insert_jit`` :1045)  device copies
``bulk_insert``        maxengine.py:946                              loop over ``insert``
``generate``           maxengine.py:838-866 (``_generate_jit`` :868)  ``mtx_decode_step[_graph]``
=====================  ==========================================  =====================

Same signatures, state-dict keys (``logits, cache, next_pos, generated_tokens, tokens``),
``ResultTokens`` layout and config keys.  ``decode_state`` is donated exactly as in the
reference (``donate_argnums``, maxengine.py:868): it is updated in place and the returned
dict is the one to keep using.

Device memory, streams and (for the optional vocab-parallel mode) the process group come
from PyTorch; all arithmetic is in ``libmtx_b200.so``.  Nothing here computes on the CPU.
"""

from __future__ import annotations

import ctypes
import dataclasses
import math
from typing import Any, Optional

import numpy as np
import torch

from . import _lib
from . import gemma3
from . import page_manager
from . import parallel
from . import params as params_lib
from .common_types import DECODING_ACTIVE_SEQUENCE_INDICATOR


@dataclasses.dataclass
class SlotData:
  tokens: Any
  valid: Any
  lengths: Any
  log_prob: Any = None


@dataclasses.dataclass
class ResultTokens:
  """Same fields as JetStream's ``engine_api.ResultTokens`` (built at maxengine.py:916-928)."""

  data: Any  # [B, 3] int32: token, valid, length
  tokens_idx: tuple = (0, 1)
  valid_idx: tuple = (1, 2)
  length_idx: tuple = (2, 3)
  log_prob: Any = None
  samples_per_slot: int = 1

  def copy_to_host_async(self):
    if isinstance(self.data, torch.Tensor) and self.data.is_cuda:
      host = torch.empty(self.data.shape, dtype=self.data.dtype, pin_memory=True)
      host.copy_(self.data, non_blocking=True)
      self._host = host

  def convert_to_numpy(self) -> "ResultTokens":
    data = self.data.cpu().numpy() if isinstance(self.data, torch.Tensor) else np.asarray(self.data)
    lp = self.log_prob
    if isinstance(lp, torch.Tensor):
      lp = lp.cpu().numpy()
    return ResultTokens(data, self.tokens_idx, self.valid_idx, self.length_idx, lp, self.samples_per_slot)

  def get_result_at_slot(self, slot: int) -> SlotData:
    start, end = slot * self.samples_per_slot, (slot + 1) * self.samples_per_slot
    d = self.data
    return SlotData(
        tokens=d[start:end, self.tokens_idx[0] : self.tokens_idx[1]],
        valid=d[start:end, self.valid_idx[0] : self.valid_idx[1]],
        lengths=d[start:end, self.length_idx[0] : self.length_idx[1]][:, 0],
        log_prob=None if self.log_prob is None else self.log_prob[start:end],
    )


@dataclasses.dataclass
class ExistingPrefix:
  """maxengine.py:66-77: a prefix that has already been processed -- its cache (``prefix["cache"]`` of an earlier ``prefill``)
  and its tokens without padding."""

  cache: Any
  common_prefix_tokens: Any


class DeviceParams:
  """Weights in the layout of ``mtx_weights`` (include/mtx_b200.h), resident in HBM.

  `folded`: wqkv / w01 carry the RMSNorm scales of their input and attn_norm / mlp_norm are ones
  (see :func:`fold_norm_scales`); `source` is then the unfolded original (kept for checkers), else None."""

  def __init__(self, tensors: dict, folded: bool = False, source: "Optional[DeviceParams]" = None):
    self.tensors = tensors
    self.folded = folded
    self.source = source
    self.struct = _lib.Weights(**{k: v.data_ptr() for k, v in tensors.items()})

  def nbytes(self) -> int:
    seen, total = set(), 0
    for t in self.tensors.values():
      if t.data_ptr() not in seen:
        seen.add(t.data_ptr())
        total += t.numel() * t.element_size()
    return total


def fold_norm_scales(dp: DeviceParams) -> DeviceParams:
  """W' = bf16(W * diag(scale)) for the two projections that consume a normalised activation, scales -> 1.

  normalizations.py:57-69 computes ``bf16(bf16(x * rstd) * scale)`` and linears.py:216 multiplies that by ``bf16(W)``;
  ``rstd * dot(x, bf16(scale * W))`` is the same product with the per-feature factor rounded into the weight instead of
  into the activation (one bf16 rounding per term either way; the difference is measured by the parity tests, which
  keep the oracle on the UNFOLDED weights).  What it buys: the decode step no longer touches the activation tiles
  between their arrival and the MMA.  When every scale is exactly 1 (random init) this is the identity
  and nothing is copied."""
  if dp.folded:
    return dp
  t = dp.tensors
  an, mn = t["attn_norm"], t["mlp_norm"]
  if bool((an == 1).all()) and bool((mn == 1).all()):
    return DeviceParams(t, folded=True, source=None)
  new = dict(t)
  new["wqkv"] = (t["wqkv"].float() * an.float()[:, None, :]).to(torch.bfloat16).contiguous()
  w01 = torch.empty_like(t["w01"])
  for l in range(w01.shape[0]):  # layer by layer: the fp32 temporary of the whole tensor would be 1.3 GB
    w01[l] = (t["w01"][l].float() * mn[l].float()[None, :]).to(torch.bfloat16)
  new["w01"] = w01
  new["attn_norm"] = torch.ones_like(an)
  new["mlp_norm"] = torch.ones_like(mn)
  return DeviceParams(new, folded=True, source=dp)


def pack_params(params: dict, config, device, vocab_shard: Optional[tuple] = None) -> DeviceParams:
  """Reference parameter tree (params.py) -> K-major packed bf16 tensors on `device`.

  Every projection is stored [out_features, in_features] so a 128-row weight tile is one
  TMA box of contiguous 128-byte rows; wi_0 / wi_1 are interleaved 16 rows at a time so the
  gate and the value of one MLP feature land in the same warp of the GEMM epilogue.
  """
  p = params_lib.unscan_params(params, config)["params"]  # scan_layers=True checkpoints stack the layers on axis 1
  E, Hq, Hkv, D = config.emb_dim, config.num_query_heads, config.num_kv_heads, config.head_dim
  M, L = config.mlp_dim, config.num_decoder_layers
  bf = torch.bfloat16
  dev = lambda t: t.to(device=device, dtype=bf)
  wqkv, wo, w01, wout, an, mn = [], [], [], [], [], []
  gemma3 = config.decoder_block == "gemma3"
  qn, kn, pan, pfn = [], [], [], []
  for i in range(L):
    lp = p["decoder"][f"layers_{i}"]
    sa = lp["self_attention"]
    q = dev(sa["query"]["kernel"]).reshape(E, Hq * D).t()
    k = dev(sa["key"]["kernel"]).reshape(E, Hkv * D).t()
    v = dev(sa["value"]["kernel"]).reshape(E, Hkv * D).t()
    wqkv.append(torch.cat([q, k, v], dim=0).contiguous())
    wo.append(dev(sa["out"]["kernel"]).reshape(Hq * D, E).t().contiguous())
    w0 = dev(lp["mlp"]["wi_0"]["kernel"]).t().reshape(M // 16, 16, E)
    w1 = dev(lp["mlp"]["wi_1"]["kernel"]).t().reshape(M // 16, 16, E)
    w01.append(torch.stack([w0, w1], dim=1).reshape(2 * M, E).contiguous())
    wout.append(dev(lp["mlp"]["wo"]["kernel"]).t().contiguous())
    an.append(dev(lp["pre_self_attention_norm" if gemma3 else "pre_self_attention_layer_norm"]["scale"]))
    mn.append(dev(lp["mlp"]["mlp_layer_norm"]["scale"]))
    if gemma3:
      qn.append(dev(sa["query_norm"]["scale"]))
      kn.append(dev(sa["key_norm"]["scale"]))
      pan.append(dev(lp["post_self_attention_norm"]["scale"]))
      pfn.append(dev(lp["post_ffw_norm"]["scale"]))
  embedding = dev(p["token_embedder"]["embedding"]).contiguous()
  if config.logits_via_embedding:
    logits = embedding  # attend_on_embedding, embeddings.py:183-199
  else:
    logits = dev(p["decoder"]["logits_dense"]["kernel"]).t().contiguous()
  if vocab_shard is not None:
    lo, hi = vocab_shard
    logits = logits[lo:hi].contiguous()
  tensors = dict(
      embedding=embedding,
      attn_norm=torch.stack(an).contiguous(),
      wqkv=torch.stack(wqkv).contiguous(),
      wo=torch.stack(wo).contiguous(),
      mlp_norm=torch.stack(mn).contiguous(),
      w01=torch.stack(w01).contiguous(),
      wout=torch.stack(wout).contiguous(),
      final_norm=dev(p["decoder"]["decoder_norm"]["scale"]).contiguous(),
      logits=logits,
  )
  if gemma3:
    tensors.update(q_norm=torch.stack(qn).contiguous(), k_norm=torch.stack(kn).contiguous(),
                   post_attn_norm=torch.stack(pan).contiguous(), post_ffw_norm=torch.stack(pfn).contiguous())
  return DeviceParams(tensors)


def random_device_params(config, device, seed: int = 0, norm_jitter: float = 0.0) -> DeviceParams:
  """Random-init weights generated directly in HBM (benchmarks; same distributions as params.py).
  `norm_jitter` > 0 draws the RMSNorm scales from N(1, norm_jitter) instead of the init value 1 (parity tests)."""
  g = torch.Generator(device=device).manual_seed(seed)
  E, Hq, Hkv, D = config.emb_dim, config.num_query_heads, config.num_kv_heads, config.head_dim
  M, L, V = config.mlp_dim, config.num_decoder_layers, config.vocab_size
  bf = torch.bfloat16

  def rn(shape, std):
    out = torch.empty(shape, device=device, dtype=bf)
    flat = out.view(-1)
    step = 1 << 26
    for lo in range(0, flat.numel(), step):
      n = min(step, flat.numel() - lo)
      flat[lo : lo + n] = (torch.randn(n, device=device, generator=g) * std).to(bf)
    return out

  def scale(shape):
    if norm_jitter <= 0.0:
      return torch.ones(shape, device=device, dtype=bf)
    return (1.0 + norm_jitter * torch.randn(shape, device=device, generator=g)).to(bf)

  qkv_n = (Hq + 2 * Hkv) * D
  wqkv = rn((L, qkv_n, E), 1.0 / math.sqrt(E))
  wqkv[:, : Hq * D] /= math.sqrt(D)  # attentions.py:1900-1904
  embedding = rn((V, E), 1.0)
  tensors = dict(
      embedding=embedding,
      attn_norm=scale((L, E)),
      wqkv=wqkv,
      wo=rn((L, E, Hq * D), 1.0 / math.sqrt(Hq * D)),
      mlp_norm=scale((L, E)),
      w01=rn((L, 2 * M, E), 1.0 / math.sqrt(E)),
      wout=rn((L, E, M), 1.0 / math.sqrt(M)),
      final_norm=scale((E,)),
      logits=embedding if config.logits_via_embedding else rn((V, E), 1.0 / math.sqrt(E)),
  )
  if config.decoder_block == "gemma3":
    tensors.update(q_norm=scale((L, D)), k_norm=scale((L, D)), post_attn_norm=scale((L, E)), post_ffw_norm=scale((L, E)))
  return DeviceParams(tensors)


class MaxEngine:
  """The reference's ``MaxEngine`` for one B200 (one process per GPU)."""

  def __init__(self, config, devices: Any = None, use_cuda_graph: bool = True, vocab_shard: Optional[tuple] = None,
               gather: Any = None):
    """`vocab_shard` = (rank, world) and `gather` (candidates [5,B] -> [world,5,B]) default to the
    torch.distributed process group when config.vocab_parallelism > 1; tests inject their own."""
    self.config = config
    self.device = _lib.require_cuda()
    self.lib = _lib.load()
    self.use_cuda_graph = use_cuda_graph
    self.rng = None
    self._handle = ctypes.c_void_p()
    self._params: Optional[DeviceParams] = None
    self._bound_params = None
    self._state = None
    B = self.max_concurrent_decodes
    if B < 1 or B > 256:
      raise ValueError(f"per_device_batch_size={config.per_device_batch_size}: this engine holds 1..256 slots per GPU")
    self._chunk = max(1, min(256, int(config.prefill_chunk_size), config.max_prefill_predict_length))
    self._staging = B  # extra KV plane prefill writes into
    self._num_slots = B + 1
    if config.quantize_kvcache:  # the int8 cache holds the decode slots only; prefill has its own bf16 plane (index 0)
      self._staging, self._num_slots = 0, B
    # attention=paged (maxengine.py:131-136): the decode cache is the page pools, one page group per slot; prefill writes a
    # bf16 staging plane (index 0) like the int8 engine's
    self._paged = config.attention == "paged"
    self.page_manager, self._page_state, self._page_host_stale = None, None, False
    # pagedattn_device_state (a key of this implementation): the PageState lives on the device and every decode step begins with
    # update_decode_pages there (as the reference's jitted update does); the host copy is refreshed when somebody looks
    self._page_on_device = self._paged and bool(config.pagedattn_device_state)
    if self._paged:
      self._staging, self._num_slots = 0, B
      self.page_manager = page_manager.PageManager(config)
      self.page_state = self.page_manager.get_initial_page_state()
    self._vp_world, self._vp_rank, self._gather = 1, 0, gather
    if vocab_shard is not None:
      self._vp_rank, self._vp_world = int(vocab_shard[0]), int(vocab_shard[1])
    elif int(config.vocab_parallelism) > 1:
      import torch.distributed as dist

      if not dist.is_initialized() or dist.get_world_size() != int(config.vocab_parallelism):
        raise ValueError("vocab_parallelism needs an initialised process group of that size")
      self._vp_rank, self._vp_world = dist.get_rank(), dist.get_world_size()
      self._gather = parallel.all_gather_candidates
    self._v_lo, self._v_hi = parallel.vocab_shard(config.vocab_size, self._vp_world, self._vp_rank)
    if self._vp_world > 1 and config.decode_sampling_strategy == "topk" and int(config.decode_sampling_top_k) > 64:
      raise ValueError("vocab-parallel top-k gathers 64 candidates per shard: decode_sampling_top_k must be <= 64")
    if self._vp_world > 8:
      raise ValueError("at most 8 vocabulary shards")
    scale = 1.0
    if config.logits_via_embedding and config.normalize_embedding_logits:
      scale = 1.0 / math.sqrt(config.emb_dim)  # decoders.py:560-562
    self._cfg_struct = _lib.ModelConfig(
        num_layers=config.num_decoder_layers,
        emb_dim=config.emb_dim,
        num_q_heads=config.num_query_heads,
        num_kv_heads=config.num_kv_heads,
        head_dim=config.head_dim,
        mlp_dim=config.mlp_dim,
        vocab_size=self._v_hi - self._v_lo,
        vocab_offset=self._v_lo,
        max_prefill_len=config.max_prefill_predict_length,
        max_target_len=config.max_target_length,
        num_slots=self._num_slots,
        max_rows=max(B, self._chunk),
        rms_eps=config.normalization_layer_epsilon,
        rope_min_timescale=float(config.rope_min_timescale),
        rope_max_timescale=float(config.rope_max_timescale),
        attn_softcap=float(config.attn_logits_soft_cap or 0.0),
        # (decoders.py:552-565: the final soft cap sits inside the logits_via_embedding branch; an untied head ignores the key)
        final_softcap=float(config.final_logits_soft_cap or 0.0) if config.logits_via_embedding else 0.0,
        logits_scale=scale,
        logits_round_bf16=0 if config.logits_dot_in_fp32 else 1,
        embedding_rows=config.vocab_size,
        kv_quant=((2 if config.kv_quant_axis == "heads_and_dkv" else 1) + (2 if config.kv_quant_dtype == "fp8" else 0)) if config.quantize_kvcache else 0,
        norm_scales_folded=1 if config.fold_norm_scales else 0,
        decoder_block=1 if config.decoder_block == "gemma3" else 0,
        sliding_window=int(config.sliding_window_size) if config.decoder_block == "gemma3" else 0,
        local_rope_max_timescale=float(config.local_rope_max_timescale),
        query_scalar=float(gemma3.get_query_pre_attn_scalar(config)) if config.decoder_block == "gemma3" else 0.0,
        paged_num_pages=int(config.pagedattn_num_pages) if self._paged else 0,
        paged_tokens_per_page=int(config.pagedattn_tokens_per_page) if self._paged else 0,
        paged_max_pages_per_group=int(config.pagedattn_max_pages_per_group) if self._paged else 0,
        paged_device_state=1 if self._page_on_device else 0,
    )
    _lib.check(self.lib.mtx_engine_create(ctypes.byref(self._cfg_struct), ctypes.byref(self._handle)))
    ws = self.lib.mtx_engine_workspace_bytes(self._handle)
    self._ws_raw = torch.empty(ws + 1024, dtype=torch.uint8, device=self.device)
    self._ws_ptr = (self._ws_raw.data_ptr() + 1023) // 1024 * 1024
    self._ws_bytes = ws
    self._alloc_state()
    self._apply_sampling()

  def __del__(self):
    try:
      if self._handle:
        self.lib.mtx_engine_destroy(self._handle)
    except Exception:  # pragma: no cover - interpreter shutdown
      pass

  # -- properties (maxengine.py:1455-1487) -------------------------------------------------

  @property
  def max_concurrent_decodes(self) -> int:
    return int(self.config.per_device_batch_size)  # x mesh.size; one process per GPU here

  @property
  def max_prefill_length(self) -> int:
    return int(self.config.max_prefill_predict_length)

  @property
  def use_chunked_prefill(self) -> bool:
    return bool(self.config.use_chunked_prefill)

  @property
  def prefill_chunk_size(self) -> int:
    return int(self.config.prefill_chunk_size)

  @property
  def samples_per_slot(self) -> int:
    return 1

  # -- state --------------------------------------------------------------------------------

  def _alloc_state(self) -> None:
    cfg, dev = self.config, self.device
    B, S = self.max_concurrent_decodes, self._num_slots
    L, Hkv, T, D = cfg.num_decoder_layers, cfg.num_kv_heads, cfg.max_target_length, cfg.head_dim
    i32 = torch.int32
    z = lambda *shape, dtype=i32: torch.zeros(*shape, dtype=dtype, device=dev)
    self._kv_quant = bool(cfg.quantize_kvcache)
    self._kv_fp8 = self._kv_quant and cfg.kv_quant_dtype == "fp8"
    self._kv_zero = 0 if self._kv_fp8 else 128  # the byte of the value 0: e4m3 +0, or q + 128 with q = 0
    if self._kv_quant:
      # int8 decode cache (u = q + 128; or float8_e4m3fn bytes) + one fp32 scale per (layer, slot, kv head, row); prefill writes ONE
      # bf16 staging plane
      self._kq = torch.full((L, S, Hkv, T, D), self._kv_zero, dtype=torch.uint8, device=dev)
      self._vq = torch.full((L, S, Hkv, T, D), self._kv_zero, dtype=torch.uint8, device=dev)
      self._k_scale = z(L, S, Hkv, T, dtype=torch.float32)
      self._v_scale = z(L, S, Hkv, T, dtype=torch.float32)
      self._k = z(L, 1, Hkv, T, D, dtype=torch.bfloat16)
      self._v = z(L, 1, Hkv, T, D, dtype=torch.bfloat16)
      self._staging = 0
    elif self._paged:
      # PagedAttentionOp.key_pages / value_pages (paged_attention.py:152-160), all layers in one pool, + the device copy of the
      # PageState in one int32 buffer: [sequence_lengths | active_page | active_page_position | num_pages_used | has_active_page |
      # page_map | page_status] (the step reads the first three and the map; with pagedattn_device_state it updates all of it)
      self._kq = self._vq = self._k_scale = self._v_scale = None
      NP, TPP, MP = int(cfg.pagedattn_num_pages), int(cfg.pagedattn_tokens_per_page), int(cfg.pagedattn_max_pages_per_group)
      self._k_pages = z(L, Hkv, NP, TPP, D, dtype=torch.bfloat16)
      self._v_pages = z(L, Hkv, NP, TPP, D, dtype=torch.bfloat16)
      self._page_dev = z(5 * B + B * MP + NP)
      self._page_dev_small, self._page_dev_map = self._page_dev[: 3 * B], self._page_dev[5 * B : 5 * B + B * MP]
      self._uploaded_map = None
      self._k = z(L, 1, Hkv, T, D, dtype=torch.bfloat16)
      self._v = z(L, 1, Hkv, T, D, dtype=torch.bfloat16)
      self._staging = 0
    else:
      self._kq = self._vq = self._k_scale = self._v_scale = None
      self._k = z(L, S, Hkv, T, D, dtype=torch.bfloat16)
      self._v = z(L, S, Hkv, T, D, dtype=torch.bfloat16)
    self._tokens, self._next_pos, self._generated = z(B, 1), z(B, 1), z(B, 1)
    self._prefill_len, self._ar_lengths, self._ar_index = z(S), z(S), z(1)
    self._result = z(B, 3)
    self._log_prob = z(B, 1, dtype=torch.float32) if cfg.return_log_prob else None
    # top-k / nucleus read the logits back (two-pass sampler), so they are always materialised for them
    two_pass = cfg.decode_sampling_strategy in ("topk", "nucleus")
    V = self._v_hi - self._v_lo  # logits columns held by this process (all of them unless vocab-parallel)
    self._logits = z(B, 1, V, dtype=torch.float32) if (cfg.materialize_logits or two_pass) else None
    self._cand = z(130 * max(B, self._chunk), dtype=torch.float32)  # candidate payload of the vocab-parallel mode (5 or 130 floats per row)
    self._rng_state = z(4)
    self._first_token = z(1)
    self._first_log_prob = z(1, dtype=torch.float32) if cfg.return_log_prob else None
    self._prefill_logits = z(V, dtype=torch.float32)
    self._prefill_tokens = z(cfg.max_prefill_predict_length)
    self._state_struct = _lib.DecodeState(
        k_cache=self._k.data_ptr(),
        v_cache=self._v.data_ptr(),
        tokens=self._tokens.data_ptr(),
        next_pos=self._next_pos.data_ptr(),
        generated=self._generated.data_ptr(),
        prefill_len=self._prefill_len.data_ptr(),
        ar_lengths=self._ar_lengths.data_ptr(),
        ar_index=self._ar_index.data_ptr(),
        result=self._result.data_ptr(),
        log_prob=self._log_prob.data_ptr() if self._log_prob is not None else None,
        logits=self._logits.data_ptr() if self._logits is not None else None,
        rng_state=self._rng_state.data_ptr(),
        kq_cache=self._kq.data_ptr() if self._kq is not None else None,
        vq_cache=self._vq.data_ptr() if self._vq is not None else None,
        k_scale=self._k_scale.data_ptr() if self._k_scale is not None else None,
        v_scale=self._v_scale.data_ptr() if self._v_scale is not None else None,
    )
    if self._paged:
      MP = int(cfg.pagedattn_max_pages_per_group)
      base = self._page_dev.data_ptr()
      self._state_struct.k_pages = self._k_pages.data_ptr()
      self._state_struct.v_pages = self._v_pages.data_ptr()
      self._state_struct.page_lengths = base
      self._state_struct.active_page = base + 4 * B
      self._state_struct.active_page_pos = base + 8 * B
      self._state_struct.num_pages_used = base + 12 * B
      self._state_struct.has_active_page = base + 16 * B
      self._state_struct.page_map = base + 20 * B
      self._state_struct.page_status = base + 4 * (5 * B + B * MP)

  def _upload_page_state(self) -> None:
    """The PageState fields a step reads, host -> device (the reference passes page_state into the jitted step,
    maxengine.py:856-864).  The source is pageable memory, so the copy has left the host buffer when the call returns."""
    ps = self.page_state
    if self._page_on_device:  # the whole state (request boundaries only: prefill, insert, release)
      packed = np.concatenate((ps.sequence_lengths, ps.active_page, ps.active_page_position, ps.num_pages_used,
                               ps.has_active_page.astype(np.int32), ps.page_map.reshape(-1), ps.page_status))
      self._page_dev.copy_(torch.from_numpy(packed))
      return
    small = np.concatenate((ps.sequence_lengths, ps.active_page, ps.active_page_position))
    self._page_dev_small.copy_(torch.from_numpy(small))
    if ps.page_map is not self._uploaded_map:  # (the page manager returns the same array while no page was handed out)
      self._page_dev_map.copy_(torch.from_numpy(ps.page_map.reshape(-1)))
      self._uploaded_map = ps.page_map

  @property
  def page_state(self):
    """maxengine.py:133-136 `self.page_state`.  With the state on the device this reads it back (a stream synchronisation) if
    decode steps have run since the host copy was made."""
    if self._page_host_stale:
      B, MP, NP = self.max_concurrent_decodes, self.page_manager.max_pages_per_group, self.page_manager.num_pages
      flat = self._page_dev.cpu().numpy()
      f = lambda i: flat[i * B : (i + 1) * B].copy()
      self._page_state = page_manager.PageState(
          page_status=flat[5 * B + B * MP :].copy(), page_map=flat[5 * B : 5 * B + B * MP].reshape(B, MP).copy(), num_pages_used=f(3),
          sequence_lengths=f(0), active_page=f(1), has_active_page=f(4).astype(bool), active_page_position=f(2))
      self._page_host_stale = False
    return self._page_state

  @page_state.setter
  def page_state(self, value) -> None:
    self._page_state, self._page_host_stale = value, False

  def _advance_pages(self) -> None:
    """maxengine.py:847-849: the page state advances before the step reads it (one more token per active group, a new page for
    the groups that crossed a boundary): on the device, in the step's first kernel, or here on the host followed by an upload."""
    if self._page_on_device:
      self._page_host_stale = True
      return
    self.page_state = self.page_manager.update_decode_pages(self.page_state)
    self._upload_page_state()

  def release_pages(self, slot: int) -> None:
    """maxengine.py:1320-1328: hand the slot's pages back to the pool."""
    if not self._paged:
      return
    self.page_state = self.page_manager.release_pages(page_state=self.page_state, page_group_id=int(slot))
    if self._page_on_device:
      self._upload_page_state()

  def _apply_sampling(self) -> None:
    cfg = self.config
    _lib.check(
        self.lib.mtx_engine_set_sampling(
            self._handle,
            _lib.SAMPLING[cfg.decode_sampling_strategy],
            int(cfg.decode_sampling_top_k),
            float(cfg.decode_sampling_nucleus_p),
            float(cfg.decode_sampling_temperature),
        )
    )

  def _bind(self, params: DeviceParams) -> None:
    if self._bound_params is params:
      return
    if bool(self.config.fold_norm_scales) != bool(params.folded):
      raise ValueError("pass the DeviceParams returned by this engine's load_params (fold_norm_scales decides their layout)")
    _lib.check(
        self.lib.mtx_engine_bind(
            self._handle, ctypes.byref(params.struct), ctypes.byref(self._state_struct), ctypes.c_void_p(self._ws_ptr), self._ws_bytes
        )
    )
    self._bound_params = params

  def _stream(self):
    return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

  def _seed(self, rng) -> None:
    """Key the sampler's Philox stream (the reference threads a jax PRNG key instead)."""
    if rng is None:
      return
    seed = int(rng.sum().item()) if isinstance(rng, torch.Tensor) else int(np.asarray(rng).astype(np.uint64).sum())
    vals = torch.tensor([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=torch.int64).to(torch.int32)
    self._rng_state[1:3].copy_(vals.to(self.device), non_blocking=True)

  # -- API -----------------------------------------------------------------------------------

  def load_params(self, params: Optional[dict] = None, rng: Any = None, on_device_init: bool = False,
                  norm_jitter: float = 0.0) -> DeviceParams:
    """maxengine.py:218.  `params` is the reference-named tree (params.py); None = random init."""
    if params is None:
      if on_device_init:
        dp = random_device_params(self.config, self.device, self.config.init_weights_seed, norm_jitter)
        if self._vp_world > 1:
          dp.tensors["logits"] = dp.tensors["logits"][self._v_lo : self._v_hi].contiguous()
          dp = DeviceParams(dp.tensors)
      else:
        dp = pack_params(params_lib.init_params(self.config), self.config, self.device, self._shard_or_none())
    elif isinstance(params, DeviceParams):
      dp = params
    else:
      dp = pack_params(params, self.config, self.device, self._shard_or_none())
    if self.config.fold_norm_scales:
      dp = fold_norm_scales(dp)
    elif dp.folded and dp.source is not None:
      raise ValueError("these DeviceParams carry folded RMSNorm scales but fold_norm_scales=False")
    self._params = dp
    self._bind(dp)
    return dp

  def _shard_or_none(self):
    return (self._v_lo, self._v_hi) if self._vp_world > 1 else None

  def init_decode_state(self, rng: Any = None) -> dict:
    """maxengine.py:1370-1453: every field zero."""
    for t in (self._k, self._v, self._tokens, self._next_pos, self._generated, self._prefill_len, self._ar_lengths,
              self._ar_index, self._result):
      t.zero_()
    self._rng_state[0:1].zero_()
    if self._kv_quant:
      self._kq.fill_(self._kv_zero)
      self._vq.fill_(self._kv_zero)
      self._k_scale.zero_()
      self._v_scale.zero_()
    if self._logits is not None:
      self._logits.zero_()
    if self._paged:
      self._k_pages.zero_()
      self._v_pages.zero_()
      self.page_state = self.page_manager.get_initial_page_state()  # maxengine.py:1379-1381
      self._upload_page_state()
    self._seed(rng)
    self._state = {
        "logits": self._logits,
        "cache": {
            # [L, slots, Hkv, T, D]: rows [0,P) = cached_prefill_key, [P,T) = cached_ar_key (uint8 q + 128 with "key_scale" /
            # "value_scale" [L, slots, Hkv, T] when quantize_kvcache: KVTensor's qvalue / scale, kvcache.py:658-736)
            # (attention=paged: "key_pages" / "value_pages" [L, Hkv, num_pages, tokens_per_page, D], paged_attention.py:152-160)
            "key": self._kq if self._kv_quant else self._k,
            "value": self._vq if self._kv_quant else self._v,
            "key_pages": self._k_pages if self._paged else None,
            "value_pages": self._v_pages if self._paged else None,
            "key_scale": self._k_scale,
            "value_scale": self._v_scale,
            "prefill_length": self._prefill_len,  # == cache_prefill_segment_id.sum(-1)
            "cached_ar_lengths": self._ar_lengths,
            "cache_ar_index": self._ar_index,
        },
        "next_pos": self._next_pos,
        "generated_tokens": self._generated,
        "tokens": self._tokens,
    }
    return self._state

  def prefill(
      self,
      *,
      params: DeviceParams,
      padded_tokens: Any,
      true_length: int,
      existing_prefix: Any = None,
      images: Any = None,
      sampler: Any = None,
      rng: Any = None,
      request_id: Any = None,
      slot: Optional[int] = None,
      return_prompt_logp: bool = False,
  ):
    """maxengine.py:533-574.  Returns (prefix, ResultTokens) for one sequence."""
    start_position = 0
    if existing_prefix is not None:  # maxengine.py:434-440: this call continues a prefix processed by earlier calls
      if not self.use_chunked_prefill:
        raise ValueError("Using chunked prefill is needed for existing_prefix.")
      start_position = int(torch.as_tensor(existing_prefix.common_prefix_tokens).reshape(-1).numel())
    if return_prompt_logp:
      raise NotImplementedError("return_prompt_logp is outside the decode path")
    self._bind(params)
    self._seed(rng)
    cfg = self.config
    true_length = int(true_length)
    if self._paged:  # maxengine.py:549-554: the slot's pages are reserved before the prefill runs
      if slot is None:
        raise ValueError("attention=paged: prefill needs the slot (page group) the sequence will be inserted into")
      self.page_state = self.page_manager.update_prefill_pages(page_state=self.page_state, page_group_id=int(slot),
                                                               true_length=start_position + true_length)
      if self._page_on_device:
        self._upload_page_state()
    toks = torch.as_tensor(padded_tokens).reshape(-1)
    if true_length < 1 or true_length > toks.numel() or start_position + toks.numel() > cfg.max_prefill_predict_length:
      raise ValueError(f"true_length={true_length}, {toks.numel()} padded tokens after {start_position} prefix tokens, "
                       f"max_prefill_predict_length={cfg.max_prefill_predict_length}")
    n = toks.numel()
    self._prefill_tokens[:n].copy_(toks.to(torch.int32), non_blocking=True)
    if start_position:
      # the prefix's cache rows go back to the staging plane (a no-op copy when the previous call left them there); the new
      # positions [start_position, start_position + true_length) attend to them causally (kvcache.py:584-624 with previous_chunk)
      pk, pv = existing_prefix.cache["key"], existing_prefix.cache["value"]  # [L, Hkv, >= start_position, D]
      if int(pk.shape[2]) < start_position:
        raise ValueError(f"existing_prefix.cache holds {int(pk.shape[2])} rows for {start_position} common_prefix_tokens")
      self._k[:, self._staging, :, :start_position].copy_(pk[:, :, :start_position])
      self._v[:, self._staging, :, :start_position].copy_(pv[:, :, :start_position])
    want_logits = self._logits is not None or self._vp_world > 1
    for start in range(0, true_length, self._chunk):
      count = min(self._chunk, true_length - start)
      last = start + count == true_length
      _lib.check(
          self.lib.mtx_prefill_chunk(
              self._handle,
              ctypes.c_void_p(self._prefill_tokens.data_ptr() + 4 * start),
              count,
              start_position + start,
              self._staging,
              1 if last else 0,
              ctypes.c_void_p(self._first_token.data_ptr()),
              ctypes.c_void_p(self._prefill_logits.data_ptr()) if (want_logits and last) else None,
              ctypes.c_void_p(self._first_log_prob.data_ptr()) if (self._first_log_prob is not None and last) else None,
              self._stream(),
          )
      )
    if self._vp_world > 1:
      # first token (not on the decode hot path): the shards' logits of the last prompt position are gathered into the full
      # row and sampled with the configured strategy by the same kernel on every rank
      full = self._gather(self._prefill_logits.reshape(1, -1)).reshape(1, -1).contiguous()  # shards are in vocabulary order
      # (row_offset -1: the noise row of the prefill draw just made, as the unsharded prefill uses it)
      _lib.check(self.lib.mtx_sample_logits(self._handle, ctypes.c_void_p(full.data_ptr()), 1, full.shape[1], full.shape[1], -1,
                                            ctypes.c_void_p(self._first_token.data_ptr()),
                                            ctypes.c_void_p(self._first_log_prob.data_ptr()) if self._first_log_prob is not None else None,
                                            self._stream()))
    first = self._first_token.clone().reshape(1, 1)
    full_true_length = start_position + true_length  # maxengine.py:442
    prefix = {
        "logits": self._prefill_logits.clone().reshape(1, 1, -1) if want_logits else None,
        "cache": {
            "key": self._k[:, self._staging, :, :full_true_length].clone(),  # [L, Hkv, len, D]
            "value": self._v[:, self._staging, :, :full_true_length].clone(),
            "prefill_length": full_true_length,
        },
        "next_pos": torch.full((1, 1), full_true_length, dtype=torch.int32, device=self.device),
        "generated_tokens": torch.zeros((1, 1), dtype=torch.int32, device=self.device),
        "tokens": first,
    }
    data = torch.cat((first, torch.ones_like(first), torch.zeros_like(first)), dim=1)
    log_prob = self._first_log_prob.clone().reshape(1, 1) if self._first_log_prob is not None else None
    return prefix, ResultTokens(data=data, log_prob=log_prob)

  def insert(self, prefix: dict, decode_state: dict, slot: int, request_id: Any = None) -> dict:
    """maxengine.py:1045-1164: copy the prefill segment into `slot`, reset the slot's AR bookkeeping,
    leave the AR ring data and the shared ring index untouched."""
    if decode_state is not self._state:
      raise ValueError("decode_state must be the dict returned by this engine's init_decode_state (it is donated)")
    B = self.max_concurrent_decodes
    if not 0 <= int(slot) < B:
      raise ValueError(f"slot {slot} outside [0, {B})")
    n = int(prefix["cache"]["prefill_length"])
    k_src, v_src = prefix["cache"]["key"], prefix["cache"]["value"]  # [L, Hkv, n_src, D]
    if not (k_src.is_contiguous() and v_src.is_contiguous()):
      k_src, v_src = k_src.contiguous(), v_src.contiguous()
    if self._paged:
      # maxengine.py:1104-1131, 1189-1191: the prefix goes to the pages reserved for the slot (none if the pool was exhausted: the
      # group then stays empty, as in the reference) and the group is marked active
      ps = self.page_state
      active = ps.has_active_page.copy()
      active[slot] = True
      self.page_state = ps.replace(has_active_page=active)
      self._upload_page_state()
      n = min(n, int(ps.num_pages_used[slot]) * self.page_manager.tokens_per_page)
    # token / position ride along as launch arguments when they are host values; device tensors (the usual case) are copied
    if n > 0:
      _lib.check(
          self.lib.mtx_insert_prefix(
              self._handle, ctypes.c_void_p(k_src.data_ptr()), ctypes.c_void_p(v_src.data_ptr()), n, int(k_src.shape[2]), int(slot),
              n, 0, 0, self._stream()))
    self._next_pos[slot].copy_(prefix["next_pos"][0])
    self._generated[slot].copy_(prefix["generated_tokens"][0])
    self._tokens[slot].copy_(prefix["tokens"][0])
    if self._logits is not None and prefix["logits"] is not None:
      self._logits[slot].copy_(prefix["logits"][0])
    return decode_state

  def bulk_insert(self, prefix: dict, decode_state: dict, slots: list) -> dict:
    """maxengine.py:946-1043: the same prefix into several slots."""
    for s in slots:
      decode_state = self.insert(prefix, decode_state, s)
    return decode_state

  def generate(self, params: DeviceParams, decode_state: dict, sampler: Any = None, rng: Any = None):
    """maxengine.py:838-936: one token for every slot.  Returns (decode_state, ResultTokens)."""
    if decode_state is not self._state:
      raise ValueError("decode_state must be the dict returned by this engine's init_decode_state (it is donated)")
    self._bind(params)
    self._seed(rng)
    B = self.max_concurrent_decodes
    if self._paged:
      self._advance_pages()
    if self._vp_world > 1:
      # each rank scores its vocabulary shard; one all-gather of 5*B floats; identical commit everywhere
      cand = self.candidate_buffer(B)
      _lib.check(self.lib.mtx_decode_step_candidates(self._handle, B, ctypes.c_void_p(cand.data_ptr()), self._stream()))
      gathered = self._gather(cand).contiguous()
      _lib.check(self.lib.mtx_commit_candidates(self._handle, B, ctypes.c_void_p(gathered.data_ptr()), self._vp_world, self._stream()))
    else:
      fn = self.lib.mtx_decode_step_graph if self.use_cuda_graph else self.lib.mtx_decode_step
      _lib.check(fn(self._handle, B, self._stream()))
    result = ResultTokens(
        data=self._result.clone(),
        log_prob=self._log_prob.clone() if self._log_prob is not None else None,
    )
    return decode_state, result

  def generate_to_host(self, params: DeviceParams, decode_state: dict, host_result: torch.Tensor, host_tokens: Optional[torch.Tensor] = None,
                       host_log_prob: Optional[torch.Tensor] = None, sync: bool = False):
    """``generate`` for a serving loop that holds its tokens on the host (JetStream / OfflineEngine copy every step's
    ResultTokens to the host: offline_engine.py:612-614): one C call -- with pinned buffers one graph launch -- does the host ->
    device copy of this step's input tokens (`host_tokens` [B,1] int32 pinned, or None to continue from decode_state["tokens"]),
    the step, and the device -> host copy of ResultTokens.data into `host_result` [B,3] int32 pinned.  `sync=True` also waits
    for the stream inside that call (`host_result` is valid on return); otherwise the caller waits on the current stream.
    Returns (decode_state, ResultTokens over the host buffers)."""
    if decode_state is not self._state:
      raise ValueError("decode_state must be the dict returned by this engine's init_decode_state (it is donated)")
    key = (id(params), id(host_result), id(host_tokens), id(host_log_prob), sync)
    call = self._host_call if getattr(self, "_host_call_key", None) == key and self._bound_params is params else None
    if call is None:  # (validated once per set of buffers: a serving loop passes the same ones every step)
      if self._vp_world > 1:
        raise ValueError("generate_to_host is the batch-partitioned path; the vocab-parallel mode goes through generate()")
      self._bind(params)
      B = self.max_concurrent_decodes
      for t, n in ((host_result, 3 * B), (host_tokens, B), (host_log_prob, B)):
        if t is not None and (t.is_cuda or not t.is_contiguous() or t.numel() != n or t.element_size() != 4):
          raise ValueError("host buffers must be contiguous 4-byte CPU tensors of B*3 (result), B (tokens), B (log-probs) elements")
      fn = self.lib.mtx_decode_step_host_sync if sync else self.lib.mtx_decode_step_host
      args = (self._handle, B, ctypes.c_void_p(host_tokens.data_ptr()) if host_tokens is not None else None,
              ctypes.c_void_p(host_result.data_ptr()), ctypes.c_void_p(host_log_prob.data_ptr()) if host_log_prob is not None else None)
      result = ResultTokens(data=host_result, log_prob=host_log_prob)
      keep = (params, host_result, host_tokens, host_log_prob)  # the ids above stay unique while these are alive
      call = (fn, args, result, keep)
      self._host_call, self._host_call_key = call, key
    fn, args, result, _ = call
    if self._paged:
      self._advance_pages()
    rc = fn(*args, self._stream())
    if rc != 0:
      _lib.check(rc)
    return decode_state, result

  def candidate_buffer(self, rows: int) -> torch.Tensor:
    """This rank's payload of the vocab-parallel all-gather: [5, rows] (greedy / weighted: the shard's winner per row) or
    [rows, 130] (top-k / nucleus: its 64 best logits, their ids, max, sum exp)."""
    nf = int(self.lib.mtx_candidate_floats(self._handle))
    flat = self._cand[: rows * nf]
    return flat.view(5, rows) if nf == 5 else flat.view(rows, nf)

  def nucleus_truncated_rows(self) -> int:
    """Rows (since load_params) whose nucleus reached past the gathered candidates (vocab-parallel nucleus only)."""
    v = ctypes.c_longlong(0)
    _lib.check(self.lib.mtx_engine_counter(self._handle, 0, ctypes.byref(v)))
    return int(v.value)

  # -- helpers for tests / benchmarks ----------------------------------------------------------

  def fill_synthetic_context(self, prefill_lengths, ar_lengths, seed: int = 7) -> dict:
    """Benchmark set-up: random-normal K/V and per-slot context lengths without running prefill
    (SURVEY 8d "random-normal bf16 fill allowed for pure bandwidth runs")."""
    cfg = self.config
    B = self.max_concurrent_decodes
    P, R = cfg.max_prefill_predict_length, cfg.max_target_length - cfg.max_prefill_predict_length
    state = self.init_decode_state()
    g = torch.Generator(device=self.device).manual_seed(seed)
    if self._paged:
      # every slot gets the pages of a (prefill + generated)-token sequence; the pools hold random-normal rows
      for slot in range(B):
        self.page_state = self.page_manager.update_prefill_pages(self.page_state, slot, int(prefill_lengths[slot]) + int(ar_lengths[slot]))
        if not self.page_state.has_active_page[slot]:
          raise ValueError("pagedattn_num_pages is too small for the synthetic contexts")
      self._upload_page_state()
      for buf in (self._k_pages, self._v_pages):
        flat = buf.view(-1)
        step = 1 << 26
        for lo in range(0, flat.numel(), step):
          n = min(step, flat.numel() - lo)
          flat[lo : lo + n] = torch.randn(n, device=self.device, generator=g).to(torch.bfloat16)
    if self._kv_quant:
      # random int8 rows that are consistent with KVQuant.quantize: every row holds a +-127 (its max) and has scale ~ |N(0,1)| + 2
      for buf in (self._kq, self._vq):
        flat = buf.view(-1)
        step = 1 << 26
        for lo in range(0, flat.numel(), step):
          n = min(step, flat.numel() - lo)
          if self._kv_fp8:  # finite e4m3 values in [-448, 448] (the byte patterns 0x7f / 0xff are NaN)
            vals = (torch.rand(n, device=self.device, generator=g) * 2.0 - 1.0) * 448.0
            flat[lo : lo + n] = vals.to(torch.float8_e4m3fn).view(torch.uint8)
          else:
            flat[lo : lo + n] = torch.randint(1, 256, (n,), device=self.device, generator=g, dtype=torch.int32).to(torch.uint8)
        buf[..., 0] = 0x7E if self._kv_fp8 else 255  # the row's largest magnitude: e4m3 448, or q = 127
      for sc in (self._k_scale, self._v_scale):
        sc.copy_(torch.randn(sc.shape, device=self.device, generator=g).abs() + 2.0)
    for buf in (() if (self._kv_quant or self._paged) else (self._k, self._v)):
      flat = buf.view(-1)
      step = 1 << 26
      for lo in range(0, flat.numel(), step):
        n = min(step, flat.numel() - lo)
        flat[lo : lo + n] = torch.randn(n, device=self.device, generator=g).to(torch.bfloat16)
    pl = torch.as_tensor(prefill_lengths, dtype=torch.int32)
    al = torch.as_tensor(ar_lengths, dtype=torch.int32)
    if pl.numel() != B or al.numel() != B or int(pl.max()) > P or int(al.max()) >= R:
      raise ValueError("bad synthetic context lengths")
    self._prefill_len[:B].copy_(pl)
    self._ar_lengths[:B].copy_(al)
    self._ar_index.fill_(int(al.max()))
    self._next_pos.copy_((pl + al).reshape(B, 1))
    self._generated.copy_(al.reshape(B, 1))
    self._tokens.copy_(torch.randint(0, cfg.vocab_size, (B, 1), generator=torch.Generator().manual_seed(seed)).to(torch.int32))
    return state
