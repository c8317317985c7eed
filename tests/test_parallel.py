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
@pytest.mark.parametrize("strategy", ["greedy", "weighted"])
def test_two_vocab_shards_on_one_gpu_match_the_unsharded_engine(strategy):
  from maxtext_indextts2_b200 import maxengine
  from tests.helpers import make_params, random_tokens, small_config

  cfg = small_config(per_device_batch_size=3, decode_sampling_strategy=strategy, decode_sampling_temperature=0.8, return_log_prob=True)
  params = make_params(cfg)
  full = maxengine.MaxEngine(cfg, use_cuda_graph=False)
  fp = full.load_params(params)
  box = {}
  shards = [maxengine.MaxEngine(cfg, use_cuda_graph=False, vocab_shard=(r, 2), gather=lambda c, r=r: box["gather"](r, c)) for r in range(2)]
  sp = [e.load_params(params) for e in shards]
  prompts = random_tokens((3, 16), cfg.vocab_size, seed=4)
  seed = np.array([5, 0], dtype=np.uint32)
  fstate = full.init_decode_state(rng=seed)
  sstates = [e.init_decode_state(rng=seed) for e in shards]
  # prefill: greedy first token through the merge, identical K/V on every shard
  box["gather"] = lambda r, c: torch.stack([box["cand0"], box["cand1"]])
  greedy_cfg = cfg if strategy == "greedy" else None
  for slot in range(3):
    n = 5 + 4 * slot
    prefix, _ = full.prefill(params=fp, padded_tokens=prompts[slot], true_length=n)
    fstate = full.insert(prefix, fstate, slot)
    for r, e in enumerate(shards):
      # the shard engines reuse the unsharded prefix (same K/V); first tokens come from the full engine
      sstates[r] = e.insert(prefix, sstates[r], slot)
  for step in range(5):
    fstate, fres = full.generate(fp, fstate)
    B = 3
    cands = []
    for r, e in enumerate(shards):
      c = e._cand[:, :B].contiguous()
      from maxtext_indextts2_b200 import _lib
      import ctypes

      _lib.check(e.lib.mtx_decode_step_candidates(e._handle, B, ctypes.c_void_p(c.data_ptr()), e._stream()))
      cands.append(c)
    gathered = torch.stack(cands).contiguous()
    for r, e in enumerate(shards):
      _lib.check(e.lib.mtx_commit_candidates(e._handle, B, ctypes.c_void_p(gathered.data_ptr()), 2, e._stream()))
    torch.cuda.synchronize()
    tok, logp = parallel.merge_candidates_reference(gathered.cpu())
    for e in shards:
      assert torch.equal(e._result.cpu(), fres.data.cpu())
      torch.testing.assert_close(e._log_prob.cpu(), fres.log_prob.cpu(), rtol=1e-4, atol=1e-4)
    assert torch.equal(tok, fres.data.cpu()[:, 0])
