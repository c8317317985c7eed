"""Multi-GPU plumbing of the decode path: one process per GPU, ``torch.distributed`` for rendezvous.

Two ways to use N GPUs:

* **Batch-partitioned (default).**  Every rank owns ``per_device_batch_size`` slots, a full weight
  replica and its own KV cache; rows never interact inside a step, so there is NO collective on the
  data path (the reference replicates weights over its ``data`` mesh axis the same way).
  :func:`slots_for_rank` gives the slot range of a rank.

* **Vocab-parallel logits (optional).**  Rank r holds rows ``[r*V/N, (r+1)*V/N)`` of the logits
  matrix.  The reference's sharding rules (``vocab -> tensor``, configs/base.yml:351) make XLA
  all-gather the full ``[B, 1, V]`` fp32 logits (maxengine.py:894); here each rank contributes its
  shard's winner per row -- 5 floats -- and ONE ``all_gather`` over NVLink carries ``5*B`` floats
  per rank.  :func:`merge_candidates_reference` states the merge rule; on the GPU the merge is
  ``mtx_commit_candidates``.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def slots_for_rank(global_batch: int, world: int, rank: int) -> range:
  """Contiguous slot range of `rank` when `global_batch` requests are split over `world` GPUs."""
  if global_batch % world:
    raise ValueError(f"global batch {global_batch} is not divisible by {world} ranks")
  per = global_batch // world
  return range(rank * per, (rank + 1) * per)


def vocab_shard(vocab: int, world: int, rank: int) -> tuple:
  """Rows [lo, hi) of the logits matrix held by `rank`."""
  if vocab % world:
    raise ValueError(f"vocab {vocab} is not divisible by {world} shards")
  per = vocab // world
  return rank * per, (rank + 1) * per


def all_gather_candidates(cand: torch.Tensor, group=None) -> torch.Tensor:
  """cand [5, rows] fp32 of this rank -> [world, 5, rows], the single collective of the mode."""
  world = dist.get_world_size(group)
  out = torch.empty((world * cand.shape[0],) + tuple(cand.shape[1:]), dtype=cand.dtype, device=cand.device)
  dist.all_gather_into_tensor(out, cand.contiguous(), group=group)  # rank-major concatenation along dim 0
  return out.view((world,) + tuple(cand.shape))


def candidates_from_logits(logits: torch.Tensor, vocab_offset: int, scores: torch.Tensor | None = None) -> torch.Tensor:
  """The candidate record a shard emits for its logits slice [rows, V_shard] (host statement of what
  the logits epilogue + finalize kernel compute): best score, its global id, its logit, max, sum exp."""
  s = logits if scores is None else scores
  idx = torch.argmax(s, dim=-1)  # first maximum
  rows = torch.arange(logits.shape[0], device=logits.device)
  mx = logits.max(dim=-1).values
  out = torch.empty(5, logits.shape[0], dtype=torch.float32, device=logits.device)
  out[0] = s[rows, idx]
  out[1] = (idx + vocab_offset).to(torch.int32).view(torch.float32)
  out[2] = logits[rows, idx]
  out[3] = mx
  out[4] = torch.exp(logits - mx[:, None]).sum(-1)
  return out


def merge_candidates_reference(gathered: torch.Tensor):
  """[world, 5, rows] -> (token [rows] int32, log_prob [rows]).

  Highest score wins, the lowest vocabulary id on ties (shards are in vocabulary order, so this is
  jnp.argmax's first-maximum rule over the full row); log-softmax from the merged (max, sum exp).
  """
  score = gathered[:, 0]
  idx = gathered[:, 1].contiguous().view(torch.int32)
  raw, mx, sm = gathered[:, 2], gathered[:, 3], gathered[:, 4]
  best = score.max(dim=0).values
  is_best = score == best[None, :]
  big = torch.iinfo(torch.int32).max
  token = torch.where(is_best, idx, torch.full_like(idx, big)).min(dim=0).values
  pick = (idx == token[None, :]) & is_best
  chosen_raw = (raw * pick).sum(dim=0) / pick.sum(dim=0)
  m = mx.max(dim=0).values
  z = (sm * torch.exp(mx - m[None, :])).sum(dim=0)
  return token, chosen_raw - (m + torch.log(z))
