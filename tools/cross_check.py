"""Randomised cross-check of the persistent step kernel against the per-kernel path at the real model scale
(24 layers): batches x context ranges, synthetic caches, three teacher-forced steps each.  Two correct summation
orders differ by ~0.02 on average and up to ~0.15 on single logits at this depth; anything beyond is a bug.

  python tools/cross_check.py            (on a B200; about a minute)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxtext_indextts2_b200 import maxengine, pyconfig  # noqa: E402
from tests.helpers import make_params  # noqa: E402


def engine_for(cfg, persistent):
  os.environ["MTX_PERSISTENT"] = "1" if persistent else "0"
  eng = maxengine.MaxEngine(cfg, use_cuda_graph=False)
  return eng, eng.load_params(on_device_init=True)


def main():
  P, T = 1024, 3072
  worst = 0.0
  for batch in (1, 3, 17, 40, 64):
    for lo, hi in ((1, 200), (64, 3000), (2900, 3071)):
      cfg = pyconfig.initialize(None, model_name="indextts2-t2s", per_device_batch_size=batch, max_prefill_predict_length=P,
                                max_target_length=T, materialize_logits=True, vocab_size=20480)
      rng = np.random.Generator(np.random.PCG64(1000 * batch + lo))
      total = rng.integers(lo, hi + 1, size=batch)
      pl = np.minimum(total, P)
      al = total - pl
      outs = []
      forced = None
      for persistent in (True, False):
        eng, dp = engine_for(cfg, persistent)
        state = eng.fill_synthetic_context(pl, al, seed=3)
        logits, toks = [], []
        for step in range(3):
          state, _ = eng.generate(dp, state)
          logits.append(state["logits"].float().cpu().clone())
          toks.append(state["tokens"].cpu().clone())
          if forced is not None:
            state["tokens"].copy_(forced[step])
        outs.append(logits)
        forced = forced or toks
        del eng, dp, state
        torch.cuda.empty_cache()
      dmax = max(float((a - b).abs().max()) for a, b in zip(*outs))
      dmean = max(float((a - b).abs().mean()) for a, b in zip(*outs))
      worst = max(worst, dmax)
      flag = "" if dmax <= 0.25 and dmean <= 0.04 else "   <-- OUTSIDE"
      print(f"batch {batch:3d} contexts [{lo},{hi}]: max |d| {dmax:.3f} mean |d| {dmean:.4f}{flag}", flush=True)
  print("worst", worst)
  return 0 if worst <= 0.25 else 1


if __name__ == "__main__":
  sys.exit(main())
