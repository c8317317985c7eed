"""Audio-token wire format: vocabulary ids of the expanded tokenizer <-> codec codes.

The reference expands the Gemma-3 vocabulary with the 8,192 codes of the MaskGCT semantic codec plus two LM-TTS
markers and padding (``vocab_expansion/extend_tokenizer.py:49-246``) and records the correspondence in a JSON file
(``audio_token_mapping_path``, configs/base.yml:496): ``embedding_to_audio`` / ``audio_to_embedding`` keyed by the
EMBEDDING index, which is the tokenizer index shifted down by one past the soft token at 262144
(``create_adjusted_embedding_index``, :29-46); markers get audio ids 8192 (``e_<BT>``) and 8193 (``e_<BA>``), padding -1.
What comes out of the decode step are embedding indices; what the codec (the step after this path) consumes are codes.

:func:`AudioTokenMap.from_json` reads the reference's file; :func:`AudioTokenMap.gemma3_indextts2` rebuilds the same table
from the construction rule (the file itself is an artefact of running the reference's script with the Gemma-3 tokenizer,
which needs network access): re-used ``<unusedN>`` tokens first, then the 1,950 added tokens, the two markers, 96 pads.
The lookups are device tensors, so a batch of sampled ids converts without leaving the GPU.
"""

from __future__ import annotations

import json

import torch

BEGIN_TEXT_AUDIO_ID = 8192   # e_<BT>, extend_tokenizer.py:162-166
BEGIN_AUDIO_AUDIO_ID = 8193  # e_<BA>
PADDING_AUDIO_ID = -1
NUM_CODES = 8192
SOFT_TOKEN_INDEX = 262144


class AudioTokenMap:
  """`embedding_to_audio` [vocab] int32 (-2 = not an audio token) and `audio_to_embedding` [8194] int32."""

  def __init__(self, embedding_to_audio: torch.Tensor, audio_to_embedding: torch.Tensor):
    self.embedding_to_audio = embedding_to_audio.to(torch.int32)
    self.audio_to_embedding = audio_to_embedding.to(torch.int32)

  @property
  def vocab_size(self) -> int:
    return int(self.embedding_to_audio.numel())

  def to(self, device) -> "AudioTokenMap":
    return AudioTokenMap(self.embedding_to_audio.to(device), self.audio_to_embedding.to(device))

  @classmethod
  def from_json(cls, path: str, vocab_size: int | None = None) -> "AudioTokenMap":
    """The reference's mapping file (extend_tokenizer.py:208-243), or its older ``audio_mappings`` form keyed by
    tokenizer index (then the soft-token shift is applied here)."""
    with open(path, "r", encoding="utf-8") as f:
      data = json.load(f)
    if "embedding_to_audio" in data:
      pairs = {int(k): int(v) for k, v in data["embedding_to_audio"].items()}
      size = vocab_size or int(data.get("stats", {}).get("adjusted_embedding_size", max(pairs) + 1))
    else:
      pairs = {}
      for k, v in data["audio_mappings"].items():
        idx = int(k)
        if idx == SOFT_TOKEN_INDEX:
          continue
        pairs[idx if idx < SOFT_TOKEN_INDEX else idx - 1] = int(v)
      size = vocab_size or max(pairs) + 1
    return cls._from_pairs(pairs, size)

  @classmethod
  def gemma3_indextts2(cls, vocab_size: int = 264192) -> "AudioTokenMap":
    """The table the reference's script produces for google/gemma-3-4b-pt: codes 0..98 on the unused tokens 6..104, codes
    99..6241 on 256001..262143, codes 6242..8191 on the 1,950 added tokens (embedding indices 262144..264093), then e_<BT>,
    e_<BA> and 96 padding rows up to `vocab_size` = 264,192 (a multiple of 256)."""
    pairs = {}
    code = 0
    for lo, hi in ((6, 104), (256001, 262143), (262144, 264093)):
      for idx in range(lo, hi + 1):
        pairs[idx] = code
        code += 1
    assert code == NUM_CODES
    pairs[264094] = BEGIN_TEXT_AUDIO_ID
    pairs[264095] = BEGIN_AUDIO_AUDIO_ID
    for idx in range(264096, vocab_size):
      pairs[idx] = PADDING_AUDIO_ID
    return cls._from_pairs(pairs, vocab_size)

  @classmethod
  def _from_pairs(cls, pairs: dict, vocab_size: int) -> "AudioTokenMap":
    e2a = torch.full((vocab_size,), -2, dtype=torch.int32)
    a2e = torch.full((NUM_CODES + 2,), -1, dtype=torch.int32)
    for idx, audio in pairs.items():
      if 0 <= idx < vocab_size:
        e2a[idx] = audio
        if audio >= 0:
          a2e[audio] = idx
    return cls(e2a, a2e)

  def to_codes(self, token_ids: torch.Tensor) -> torch.Tensor:
    """Sampled vocabulary ids -> codec codes in [0, 8192); markers 8192 / 8193; padding -1; text tokens -2."""
    return self.embedding_to_audio[token_ids.to(torch.int64)]

  def to_token_ids(self, codes: torch.Tensor) -> torch.Tensor:
    """Codec codes (and the two markers) -> vocabulary ids, e.g. the reference-audio prompt fed to prefill."""
    return self.audio_to_embedding[codes.to(torch.int64)]

  def is_audio(self, token_ids: torch.Tensor) -> torch.Tensor:
    c = self.to_codes(token_ids)
    return (c >= 0) & (c < NUM_CODES)
