"""Continuous batching over MaxEngine: the caller in front of the decode step.

Mirrors the scheduling of the reference's ``InferenceWorker`` (MaxText/inference/offline_engine.py:473-715): a queue of
prompts, `max_concurrent_decodes` decode slots, prefill + insert whenever a slot is free (`prefill_done`, :565-597), batched
generate steps (`decode`, :599-631), token emission with EOS / length termination that frees the slot (`emit_token`,
`background_token_emission`, :637-715).  Differences: one thread (the reference emits from a background thread to overlap the
host work of a JAX dispatch; here a step is one C call and the tokens come back through pinned memory).  With
`attention=paged` a slot is a page group (`inference/page_manager.py`): prefill reserves the group's pages, and they go back to
the pool when the sequence ends (`MaxEngine.release_pages`, maxengine.py:1320-1328).

Host-side scheduling only: every token comes from MaxEngine (sm_100a kernels); nothing here computes on the CPU.
"""

from __future__ import annotations

import dataclasses
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from . import maxengine


@dataclasses.dataclass
class TokenOutput:
  token: int
  log_prob: Optional[float] = None


@dataclasses.dataclass
class CompletionOutput:
  """offline_engine.py: `CompletionOutput` (index, token ids, log-probs)."""

  index: int
  token_ids: np.ndarray
  logprobs: Optional[np.ndarray]


class OfflineEngine:
  def __init__(self, config, params: Any = None, engine: Optional[maxengine.MaxEngine] = None, eos_ids: Sequence[int] = (),
               min_decode_steps: int = 1, on_device_init: bool = False):
    self.config = config
    self.engine = engine or maxengine.MaxEngine(config)
    if isinstance(params, maxengine.DeviceParams):
      self.params = self.engine.load_params(params)
    else:
      self.params = self.engine.load_params(params, on_device_init=on_device_init)
    self.eos_ids = set(int(e) for e in eos_ids)
    self.min_decode_steps = max(1, int(min_decode_steps))
    self.slots = self.engine.max_concurrent_decodes
    self.decode_state = self.engine.init_decode_state()
    B = self.slots
    self._host_in = torch.zeros(B, 1, dtype=torch.int32).pin_memory()
    self._host_out = torch.zeros(B, 3, dtype=torch.int32).pin_memory()
    self._host_lp = torch.zeros(B, 1, dtype=torch.float32).pin_memory() if config.return_log_prob else None
    self.steps_run = 0
    self.prefills_run = 0

  def batch_inference(self, prompts: List[Sequence[int]], max_decode_length: int = 64) -> List[CompletionOutput]:
    """Decode every prompt to EOS or `max_decode_length` tokens (the first token comes from prefill), refilling slots as
    sequences finish.  Returns the completions in prompt order."""
    P = self.engine.max_prefill_length
    R = self.config.max_target_length - P
    if max_decode_length > R:
      raise ValueError(f"max_decode_length {max_decode_length} exceeds the AR ring of {R} rows")
    pending = list(range(len(prompts)))[::-1]
    slot_to_id: List[Optional[int]] = [None] * self.slots
    empty = list(range(self.slots))[::-1]
    tokens: List[List[TokenOutput]] = [[] for _ in prompts]
    stream = torch.cuda.current_stream(self.engine.device)

    def emit(pid: int, tok: int, lp: Optional[float]) -> bool:
      """offline_engine.py:681-715: append unless already finished; True when the sequence ends here."""
      seq = tokens[pid]
      if len(seq) == max_decode_length or (seq and seq[-1].token in self.eos_ids):
        return True
      seq.append(TokenOutput(tok, lp))
      return tok in self.eos_ids or len(seq) == max_decode_length

    while pending or any(s is not None for s in slot_to_id):
      # ---- prefill into every free slot (prefill_done, :565-597) ----
      while pending and empty:
        pid, slot = pending.pop(), empty.pop()
        ids = np.asarray(prompts[pid], dtype=np.int64)
        if not 1 <= ids.size <= P:
          raise ValueError(f"prompt {pid}: {ids.size} tokens, max_prefill_predict_length={P}")
        padded = torch.zeros(P, dtype=torch.int64)
        padded[: ids.size] = torch.from_numpy(ids)
        prefix, first = self.engine.prefill(params=self.params, padded_tokens=padded, true_length=int(ids.size), slot=slot)
        self.decode_state = self.engine.insert(prefix, self.decode_state, slot)
        self.prefills_run += 1
        lp = float(first.log_prob.reshape(-1)[0]) if first.log_prob is not None else None
        if emit(pid, int(first.data.reshape(-1)[0]), lp):
          self.engine.release_pages(slot)
          empty.append(slot)
        else:
          slot_to_id[slot] = pid
      if not any(s is not None for s in slot_to_id):
        continue
      # ---- decode (decode, :599-631): every slot advances, occupied or not ----
      for _ in range(self.min_decode_steps):
        self.decode_state, _ = self.engine.generate_to_host(self.params, self.decode_state, self._host_out, host_log_prob=self._host_lp)
        stream.synchronize()
        self.steps_run += 1
        out = self._host_out.numpy()
        lps = self._host_lp.numpy() if self._host_lp is not None else None
        for slot, pid in enumerate(slot_to_id):
          if pid is None:
            continue
          if emit(pid, int(out[slot, 0]), float(lps[slot, 0]) if lps is not None else None):
            slot_to_id[slot] = None
            self.engine.release_pages(slot)
            empty.append(slot)
    return [
        CompletionOutput(i, np.array([t.token for t in seq], dtype=np.int32),
                         np.array([t.log_prob for t in seq], dtype=np.float32) if seq and seq[0].log_prob is not None else None)
        for i, seq in enumerate(tokens)
    ]
