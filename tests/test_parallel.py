"""Multi-GPU logic.  CPU: world_size-2 gloo process group exercising the one collective of the
vocab-parallel mode and the batch partition; GPU: two vocabulary shards on one device, merged by
mtx_commit_candidates, against the unsharded engine."""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maxtext_indextts2_b200 import parallel
from oracle import decode_ref as ref


def _worker(rank, world, port, logits, ret):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    lo, hi = parallel.vocab_shard(logits.shape[1], world, rank)
    cand = parallel.candidates_from_logits(logits[:, lo:hi], lo)
    gathered = parallel.all_gather_candidates(cand)
    token, logp = parallel.merge_candidates_reference(gathered)
    ret[rank] = (token.clone(), logp.clone(), gathered.shape)
  finally:
    dist.destroy_process_group()


def test_vocab_parallel_merge_over_gloo_world_2():
  g = torch.Generator().manual_seed(0)
  logits = (torch.randn(6, 512, generator=g) * 3).to(torch.bfloat16).float()  # bf16-valued: ties are common
  logits[0, 300] = logits[0, 17] = logits[0].max() + 1  # a tie across the two shards: the lower id must win
  logits[1, 400] = logits[1, 401] = logits[1].max() + 1  # a tie inside shard 1
  ret = mp.Manager().dict()
  mp.spawn(_worker, args=(2, 29517, logits, ret), nprocs=2, join=True)
  want = ref.sampling(logits, "greedy")
  want_lp = ref.log_prob_of_chosen_token(logits[:, None, :], want[:, None])[:, 0]
  for rank in range(2):
    token, logp, shape = ret[rank]
    assert tuple(shape) == (2, 5, 6)
    assert torch.equal(token.long(), want)
    torch.testing.assert_close(logp, want_lp, rtol=1e-5, atol=1e-5)
  assert int(ret[0][0][0]) == 17 and int(ret[0][0][1]) == 400


def test_partitions():
  assert list(parallel.slots_for_rank(512, 8, 3)) == list(range(192, 256))
  assert parallel.vocab_shard(264192, 8, 7) == (231168, 264192)
  with pytest.raises(ValueError):
    parallel.slots_for_rank(10, 4, 0)
  with pytest.raises(ValueError):
    parallel.vocab_shard(1001, 2, 0)


@pytest.mark.gpu
@pytest.mark.parametrize(
    "strategy,kw",
    [("greedy", {}), ("weighted", {}), ("topk", dict(decode_sampling_top_k=7)), ("topk", dict(decode_sampling_top_k=64)),
     ("nucleus", dict(decode_sampling_nucleus_p=0.05)), ("nucleus", dict(decode_sampling_nucleus_p=0.95))],
)
def test_two_vocab_shards_on_one_gpu_match_the_unsharded_engine(strategy, kw):
  """Two vocabulary shards on one device, candidates exchanged by hand (what the NCCL all-gather carries), committed by
  mtx_commit_candidates on both: same tokens and log-probs as the unsharded engine.  top-k / nucleus travel as 64 candidates per
  shard; a nucleus that does not fit them (p = 0.95 over near-uniform random-init logits) is truncated and counted."""
  import ctypes

  from maxtext_indextts2_b200 import _lib, maxengine
  from tests.helpers import make_params, random_tokens, small_config

  cfg = small_config(per_device_batch_size=3, decode_sampling_strategy=strategy, decode_sampling_temperature=0.8, return_log_prob=True,
                     materialize_logits=True, **kw)
  params = make_params(cfg)
  full = maxengine.MaxEngine(cfg, use_cuda_graph=False)
  fp = full.load_params(params)
  box = {}
  shards = [maxengine.MaxEngine(cfg, use_cuda_graph=False, vocab_shard=(r, 2), gather=lambda c: box["gather"](c)) for r in range(2)]
  sp = [e.load_params(params) for e in shards]
  prompts = random_tokens((3, 16), cfg.vocab_size, seed=4)
  seed = np.array([5, 0], dtype=np.uint32)
  fstate = full.init_decode_state(rng=seed)
  sstates = [e.init_decode_state(rng=seed) for e in shards]
  for slot in range(3):
    n = 5 + 4 * slot
    prefix, _ = full.prefill(params=fp, padded_tokens=prompts[slot], true_length=n)
    fstate = full.insert(prefix, fstate, slot)
    for r, e in enumerate(shards):
      # the shard engines reuse the unsharded prefix (same K/V, same first token)
      sstates[r] = e.insert({**prefix, "logits": None}, sstates[r], slot)
  truncated_before = shards[0].nucleus_truncated_rows()
  B = 3
  for step in range(5):
    fstate, fres = full.generate(fp, fstate)
    cands = []
    for r, e in enumerate(shards):
      c = e.candidate_buffer(B)
      _lib.check(e.lib.mtx_decode_step_candidates(e._handle, B, ctypes.c_void_p(c.data_ptr()), e._stream()))
      cands.append(c.clone())
    gathered = torch.stack(cands).contiguous()  # [2, 5, B] or [2, B, 130]
    for r, e in enumerate(shards):
      _lib.check(e.lib.mtx_commit_candidates(e._handle, B, ctypes.c_void_p(gathered.data_ptr()), 2, e._stream()))
    torch.cuda.synchronize()
    exact = not (strategy == "nucleus" and kw["decode_sampling_nucleus_p"] > 0.5)
    for e in shards:
      if exact:
        assert torch.equal(e._result.cpu(), fres.data.cpu()), (step, e._result.cpu(), fres.data.cpu())
        torch.testing.assert_close(e._log_prob.cpu(), fres.log_prob.cpu(), rtol=1e-4, atol=1e-4)
      else:  # truncated nucleus: a token of the candidate set, identical on both shards
        ids = gathered[:, :, 64:128].contiguous().view(torch.int32).cpu()
        for b in range(B):
          assert int(e._result[b, 0]) in ids[:, b].reshape(-1).tolist()
        assert torch.equal(e._result.cpu(), shards[0]._result.cpu())
        for f in shards + [full]:
          f._tokens.copy_(shards[0]._tokens)  # keep the three engines on one history
    if strategy in ("greedy", "weighted"):
      tok, _ = parallel.merge_candidates_reference(gathered.cpu())
      assert torch.equal(tok, fres.data.cpu()[:, 0])
  if strategy == "nucleus":
    grew = shards[0].nucleus_truncated_rows() - truncated_before
    assert (grew == 0) == (kw["decode_sampling_nucleus_p"] < 0.5), grew


def _nccl_worker(rank, world, port, strategy, kw, ret):
  """One process per GPU: vocab-parallel decode over NCCL against the unsharded engine on rank 0."""
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  try:
    from maxtext_indextts2_b200 import maxengine
    from tests.helpers import make_params, random_tokens, small_config

    cfg = small_config(per_device_batch_size=4, decode_sampling_strategy=strategy, decode_sampling_temperature=0.9, return_log_prob=True,
                       vocab_parallelism=world, materialize_logits=True, **kw)
    params = make_params(cfg)
    engine = maxengine.MaxEngine(cfg, use_cuda_graph=False)
    dp = engine.load_params(params)
    seed = np.array([11, 0], dtype=np.uint32)
    state = engine.init_decode_state(rng=seed)
    prompts = random_tokens((4, 16), cfg.vocab_size, seed=6)
    for slot in range(4):
      prefix, _ = engine.prefill(params=dp, padded_tokens=prompts[slot], true_length=6 + slot)
      state = engine.insert(prefix, state, slot)
    toks = []
    for _ in range(6):
      state, res = engine.generate(dp, state)
      toks.append(res.data.cpu().clone())
    ret[rank] = torch.stack(toks)
  finally:
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("strategy,kw", [("greedy", {}), ("topk", dict(decode_sampling_top_k=8))])
def test_vocab_parallel_over_nccl_world_2(strategy, kw):
  """The one collective of the mode on real links: two ranks, two GPUs, NCCL all-gather of the candidate payload; every rank
  must emit the tokens of the unsharded engine.  Skipped on a box with a single GPU."""
  if torch.cuda.device_count() < 2:
    pytest.skip("needs two GPUs")
  from maxtext_indextts2_b200 import maxengine
  from tests.helpers import make_params, random_tokens, small_config

  ret = mp.Manager().dict()
  mp.spawn(_nccl_worker, args=(2, 29531, strategy, kw, ret), nprocs=2, join=True)
  cfg = small_config(per_device_batch_size=4, decode_sampling_strategy=strategy, decode_sampling_temperature=0.9, return_log_prob=True,
                     materialize_logits=True, **kw)
  full = maxengine.MaxEngine(cfg, use_cuda_graph=False)
  fp = full.load_params(make_params(cfg))
  state = full.init_decode_state(rng=np.array([11, 0], dtype=np.uint32))
  prompts = random_tokens((4, 16), cfg.vocab_size, seed=6)
  for slot in range(4):
    prefix, _ = full.prefill(params=fp, padded_tokens=prompts[slot], true_length=6 + slot)
    state = full.insert(prefix, state, slot)
  want = []
  for _ in range(6):
    state, res = full.generate(fp, state)
    want.append(res.data.cpu().clone())
  want = torch.stack(want)
  assert torch.equal(ret[0], want) and torch.equal(ret[1], want)
