"""Engine-level parity on the B200: MaxEngine (CUDA) against the CPU oracle on the same
random-init weights and synthetic prompts.

Stated tolerances:
* logits vs the dtype-faithful oracle: rtol = atol = 1e-1, the reference's own ceiling
  (MaxText/tests/model_test.py:191);
* logits vs the fp32 oracle: |d| <= 2^-5 * max|logit| (bf16 activations through the stack);
* greedy token ids: bit-exact, except at steps where the faithful oracle's top-2 margin is
  below `NEAR_TIE` of the top logit -- those are near-ties; the run is re-synchronised by
  teacher-forcing the oracle's token and the step is counted.
"""

import numpy as np
import pytest
import torch

from maxtext_indextts2_b200 import maxengine, pyconfig
from oracle import decode_ref as ref
from tests.helpers import make_params, random_tokens, small_config

pytestmark = pytest.mark.gpu

NEAR_TIE = 2**-6  # relative top-2 margin under which a greedy mismatch is a documented near-tie


def _assert_logits(got, want, deep=False):
  """rtol = atol = 1e-1 (MaxText/tests/model_test.py:191).  `deep`: the 24-layer model, where isolated entries of a
  264,192-wide row leave that band between any two correct bf16 evaluation orders (tests/test_baseline_configs_gpu.py
  measures it between the two CPU oracles): all but 5e-5 of the entries inside the band, every entry within 0.25."""
  if not deep:
    torch.testing.assert_close(got, want, rtol=1e-1, atol=1e-1)
    return
  d = (got - want).abs()
  assert d.max() <= 0.25 and (d > 0.1 + 0.1 * want.abs()).float().mean() <= 5e-5, f"max {d.max():.3f}"


def _prefill_both(engine, dparams, oracle, ostate, state, prompts, lengths, deep=False):
  for slot, (toks, n) in enumerate(zip(prompts, lengths)):
    padded = torch.zeros(oracle.P, dtype=torch.int64)
    padded[:n] = toks[:n]
    oprefix, ofirst = oracle.prefill(padded, n)
    ostate = oracle.insert(oprefix, ostate, slot)
    prefix, result = engine.prefill(params=dparams, padded_tokens=padded, true_length=n)
    assert result.data.shape == (1, 3)
    if prefix["logits"] is not None:
      _assert_logits(prefix["logits"].cpu()[0], oprefix["logits"][0], deep)
    got_first = int(prefix["tokens"].item())
    if got_first != int(ofirst):
      _assert_near_tie(oprefix["logits"][0, 0], got_first, int(ofirst))
      prefix["tokens"].fill_(int(ofirst))
    state = engine.insert(prefix, state, slot)
  return ostate, state


def _assert_near_tie(row_logits, got, want, margin=None):
  top = row_logits[want].item()
  other = row_logits[got].item()
  bound = NEAR_TIE * max(1.0, abs(top)) if margin is None else margin
  assert abs(top - other) <= bound, f"token {got} (logit {other}) vs oracle {want} (logit {top}) is not a near-tie"


def _lockstep(engine, dparams, oracle, ostate, state, steps, atol=1e-1, rtol=1e-1, tie_margin=None):
  near_ties = 0
  for step in range(steps):
    ostate, odata = oracle.generate(ostate)
    state, result = engine.generate(dparams, state)
    data = result.data.cpu()
    assert data.shape == odata.shape and data.dtype == torch.int32
    assert torch.equal(data[:, 1:], odata[:, 1:])  # valid flag and generated length
    if state["logits"] is not None:
      got = state["logits"].cpu()
      torch.testing.assert_close(got, ostate["logits"], rtol=rtol, atol=atol)
    for b in range(data.shape[0]):
      if data[b, 0] != odata[b, 0]:
        _assert_near_tie(ostate["logits"][b, 0], int(data[b, 0]), int(odata[b, 0]), tie_margin)
        near_ties += 1
        state["tokens"][b] = int(odata[b, 0])  # re-synchronise on the oracle's token
    assert torch.equal(state["next_pos"].cpu(), ostate["next_pos"])
  return near_ties


@pytest.mark.parametrize("use_graph", [False, True])
def test_greedy_decode_matches_oracle_small(use_graph):
  cfg = small_config(per_device_batch_size=3, max_prefill_predict_length=16, max_target_length=48, materialize_logits=True)
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=use_graph)
  dparams = engine.load_params(params)
  prompts = random_tokens((3, 16), cfg.vocab_size)
  ostate, state = _prefill_both(engine, dparams, oracle, oracle.init_decode_state(), engine.init_decode_state(), prompts, [16, 5, 1])
  near = _lockstep(engine, dparams, oracle, ostate, state, steps=32)
  assert near <= 3, f"{near} near-ties in 96 tokens"


def test_logits_close_to_fp32_oracle():
  cfg = small_config(per_device_batch_size=2, materialize_logits=True)
  params = make_params(cfg)
  f32 = ref.DecodeOracle(cfg, params, faithful=False)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=False)
  dparams = engine.load_params(params)
  prompts = random_tokens((2, 16), cfg.vocab_size, seed=9)
  ostate, state = _prefill_both(engine, dparams, f32, f32.init_decode_state(), engine.init_decode_state(), prompts, [7, 12])
  for _ in range(8):
    ostate, odata = f32.generate(ostate)
    state, _ = engine.generate(dparams, state)
    want = ostate["logits"]
    got = state["logits"].cpu()
    assert (got - want).abs().max() <= 2**-5 * want.abs().max()
    state["tokens"].copy_(odata[:, :1])  # teacher-force so both see the same history


def test_late_insert_and_ring_wrap():
  """Slots joining at different times share one ring index (maxengine.py:1060-1067, kvcache.py:778)."""
  cfg = small_config(per_device_batch_size=2, max_prefill_predict_length=8, max_target_length=20, materialize_logits=True)
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=True)
  dparams = engine.load_params(params)
  prompts = random_tokens((2, 8), cfg.vocab_size, seed=5)
  ostate, state = oracle.init_decode_state(), engine.init_decode_state()

  def add(slot, n):
    nonlocal ostate, state
    padded = torch.zeros(oracle.P, dtype=torch.int64)
    padded[:n] = prompts[slot, :n]
    oprefix, ofirst = oracle.prefill(padded, n)
    ostate = oracle.insert(oprefix, ostate, slot)
    prefix, _ = engine.prefill(params=dparams, padded_tokens=padded, true_length=n)
    prefix["tokens"].fill_(int(ofirst))
    state = engine.insert(prefix, state, slot)

  add(0, 3)
  # slot 1 is empty: the reference still runs it on zeros; only slot 0 is compared
  for _ in range(5):
    ostate, odata = oracle.generate(ostate)
    state, result = engine.generate(dparams, state)
    torch.testing.assert_close(state["logits"].cpu()[0], ostate["logits"][0], rtol=1e-1, atol=1e-1)
    state["tokens"].copy_(odata[:, :1])
  add(1, 6)
  assert int(state["cache"]["cache_ar_index"].item()) == 5 and int(state["cache"]["cached_ar_lengths"][1].item()) == 0
  for _ in range(9):  # 14 steps in total on a ring of 12: wraps, slot 0 loses its oldest AR rows in both
    ostate, odata = oracle.generate(ostate)
    state, result = engine.generate(dparams, state)
    torch.testing.assert_close(state["logits"].cpu(), ostate["logits"], rtol=1e-1, atol=1e-1)
    state["tokens"].copy_(odata[:, :1])
  assert int(state["cache"]["cache_ar_index"].item()) == 2


def test_tiny_audio_config_greedy_64_steps():
  """BASELINE config C1: MaxText decode.py-shaped run -- 4 layers, emb 256, audio-expanded vocabulary,
  batch 1, greedy, 37-token prompt (seed 1234), every one of the 64 decode steps compared."""
  cfg = pyconfig.initialize(None, model_name="tiny-audio", materialize_logits=True)
  params = make_params(cfg, perturb=False)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  engine = maxengine.MaxEngine(cfg)
  dparams = engine.load_params(params)
  rng = np.random.Generator(np.random.PCG64(1234))
  prompt = torch.from_numpy(rng.integers(0, 262144, size=(1, 64), dtype=np.int64))
  ostate, state = _prefill_both(engine, dparams, oracle, oracle.init_decode_state(), engine.init_decode_state(), prompt, [37])
  near = _lockstep(engine, dparams, oracle, ostate, state, steps=64)
  # every mismatch was checked to be a near-tie (top-2 margin <= 2^-6 of the top logit); with a 264k vocabulary of
  # random-init logits a handful per 64 steps is expected, their number moves with the summation order
  assert near <= 4


def test_weighted_sampling_follows_the_oracle_stream():
  cfg = small_config(per_device_batch_size=4, materialize_logits=True, decode_sampling_strategy="weighted",
                     decode_sampling_temperature=0.7, return_log_prob=True)
  params = make_params(cfg)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=True)
  dparams = engine.load_params(params)
  state = engine.init_decode_state(rng=np.array([123, 0], dtype=np.uint32))
  prompts = random_tokens((4, 16), cfg.vocab_size, seed=2)
  for slot in range(4):
    prefix, _ = engine.prefill(params=dparams, padded_tokens=prompts[slot], true_length=9 + slot)
    state = engine.insert(prefix, state, slot)
  for step in range(6):
    state, result = engine.generate(dparams, state)
    logits = state["logits"].cpu()
    toks, scores = ref.sampling(logits, "weighted", temperature=0.7, seed=123, step=step, return_scores=True)
    got = result.data.cpu()[:, 0]
    for b in range(4):
      if int(got[b]) != int(toks[b, 0]):
        s = scores[b]
        assert abs(s[int(got[b])] - s[int(toks[b, 0])]) < 1e-3  # ulp-level difference of logf on the two sides
    lp = ref.log_prob_of_chosen_token(logits, got.reshape(4, 1).long())
    torch.testing.assert_close(result.log_prob.cpu(), lp, rtol=1e-3, atol=1e-3)


def test_engine_api_shapes_and_errors():
  """MaxText/tests/maxengine_test.py:111-164 (shapes / dtypes) and the error behaviour of SURVEY 8b."""
  cfg = small_config(per_device_batch_size=2)
  engine = maxengine.MaxEngine(cfg)
  dparams = engine.load_params()  # random init, as decode.py does without load_parameters_path
  state = engine.init_decode_state()
  assert set(state) == {"logits", "cache", "next_pos", "generated_tokens", "tokens"}
  assert state["tokens"].shape == (2, 1) and state["tokens"].dtype == torch.int32
  assert engine.max_concurrent_decodes == 2 and engine.max_prefill_length == 16 and engine.samples_per_slot == 1
  prefix, result = engine.prefill(params=dparams, padded_tokens=torch.arange(16), true_length=4)
  assert prefix["next_pos"].tolist() == [[4]] and prefix["generated_tokens"].tolist() == [[0]]
  assert result.data.shape == (1, 3) and result.data.cpu()[0, 1:].tolist() == [1, 0]
  state = engine.bulk_insert(prefix, state, [0, 1])
  state, result = engine.generate(dparams, state)
  assert result.data.shape == (2, 3) and result.data.dtype == torch.int32
  slot0 = result.convert_to_numpy().get_result_at_slot(0)
  assert slot0.tokens.shape == (1, 1) and slot0.valid[0, 0] == 1 and slot0.lengths[0] == 1
  assert state["next_pos"].cpu().tolist() == [[5], [5]]
  with pytest.raises(ValueError):
    engine.insert(prefix, state, 7)
  with pytest.raises(ValueError):
    engine.prefill(params=dparams, padded_tokens=torch.arange(16), true_length=4, existing_prefix=object())
  with pytest.raises(ValueError):
    pyconfig.initialize(None, decode_sampling_strategy="beam")
  with pytest.raises(ValueError):
    pyconfig.initialize(None, not_a_key=1)


@pytest.mark.parametrize(
    "strategy,kw",
    [("topk", dict(decode_sampling_top_k=5)), ("topk", dict(decode_sampling_top_k=1)), ("nucleus", dict(decode_sampling_nucleus_p=0.8))],
)
def test_topk_and_nucleus_follow_the_oracle_stream(strategy, kw):
  """inference_utils.py:87-111 as CUDA radix-select kernels; same Philox stream as the oracle."""
  temp = 0.9
  cfg = small_config(per_device_batch_size=4, decode_sampling_strategy=strategy, decode_sampling_temperature=temp,
                     return_log_prob=True, **kw)
  params = make_params(cfg)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=True)
  dparams = engine.load_params(params)
  state = engine.init_decode_state(rng=np.array([77, 0], dtype=np.uint32))
  assert state["logits"] is not None  # the two-pass sampler materialises them
  prompts = random_tokens((4, 16), cfg.vocab_size, seed=3)
  for slot in range(4):
    prefix, _ = engine.prefill(params=dparams, padded_tokens=prompts[slot], true_length=6 + 2 * slot)
    state = engine.insert(prefix, state, slot)
  k = int(cfg.decode_sampling_top_k)
  p = float(cfg.decode_sampling_nucleus_p)
  for step in range(6):
    state, result = engine.generate(dparams, state)
    logits = state["logits"].cpu()
    toks, scores = ref.sampling(logits, strategy, topk=k, nucleus_topp=p, temperature=temp, seed=77, step=step, return_scores=True)
    got = result.data.cpu()[:, 0]
    for b in range(4):
      g, w = int(got[b]), int(toks[b, 0])
      if strategy == "topk":
        assert g in torch.topk(logits[b, 0], k).indices.tolist()
        if k == 1:
          assert g == int(torch.argmax(logits[b, 0]))
      else:
        cut = ref.nucleus_cutoff(logits[b], p)[0, 0]
        assert logits[b, 0, g] >= cut - 1e-6  # inside the nucleus
      if g != w:
        assert abs(scores[b][g] - scores[b][w]) < 1e-3, (step, b, g, w)
    lp = ref.log_prob_of_chosen_token(logits, got.reshape(4, 1).long())
    torch.testing.assert_close(result.log_prob.cpu(), lp, rtol=1e-3, atol=1e-3)


def test_topk_ties_keep_the_lowest_index():
  """lax.top_k (inference_utils.py:103) keeps the lower index of equal logits.  Every logit appears twice here (the output
  projection's columns come in identical pairs) and k is odd, so the cut-off value always has one entry kept and one dropped;
  a high temperature makes the draw visit the whole kept set."""
  temp, k = 6.0, 5
  cfg = small_config(per_device_batch_size=4, decode_sampling_strategy="topk", decode_sampling_top_k=k,
                     decode_sampling_temperature=temp)
  params = make_params(cfg)
  if cfg.logits_via_embedding:
    emb = params["params"]["token_embedder"]["embedding"]
    emb[1::2] = emb[0::2]
  else:
    w = params["params"]["decoder"]["logits_dense"]["kernel"]
    w[:, 1::2] = w[:, 0::2]
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=True)
  dparams = engine.load_params(params)
  state = engine.init_decode_state(rng=np.array([5, 0], dtype=np.uint32))
  prompts = random_tokens((4, 16), cfg.vocab_size, seed=11)
  for slot in range(4):
    prefix, _ = engine.prefill(params=dparams, padded_tokens=prompts[slot], true_length=9 + slot)
    state = engine.insert(prefix, state, slot)
  picked_last = 0
  for step in range(16):
    state, result = engine.generate(dparams, state)
    logits = state["logits"].cpu()
    assert torch.equal(logits[..., 0::2], logits[..., 1::2])
    toks, scores = ref.sampling(logits, "topk", topk=k, temperature=temp, seed=5, step=step, return_scores=True)
    got = result.data.cpu()[:, 0]
    for b in range(4):
      g, w = int(got[b]), int(toks[b, 0])
      assert scores[b][g] > -float("inf"), (step, b, g)  # inside the oracle's kept set
      if g != w:
        assert abs(scores[b][g] - scores[b][w]) < 1e-3, (step, b, g, w)
      kept = torch.nonzero(scores[b] > -float("inf"))[:, 0]
      picked_last += int(g == int(kept[torch.argmin(logits[b, 0, kept])]))
  assert picked_last > 0  # the tied entry itself was drawn at least once


def test_indextts2_scale_logits_and_greedy_tokens():
  """BASELINE config C2 shape (24 layers, emb 1280, 20/4 heads x 64, mlp 5120, V = 264,192) with short
  P/T so the CPU oracle finishes in about a minute: prefill + 6 decode steps, 4 slots."""
  cfg = pyconfig.initialize(None, model_name="indextts2-t2s", per_device_batch_size=4, max_prefill_predict_length=64,
                            max_target_length=128, materialize_logits=True)
  params = make_params(cfg, perturb=True)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  engine = maxengine.MaxEngine(cfg)
  dparams = engine.load_params(params)
  rng = np.random.Generator(np.random.PCG64(7))
  prompts = torch.from_numpy(rng.integers(0, cfg.vocab_size, size=(4, 64), dtype=np.int64))
  ostate, state = _prefill_both(engine, dparams, oracle, oracle.init_decode_state(), engine.init_decode_state(), prompts, [40, 17, 64, 5], deep=True)
  # 24 layers deep, the bf16 rounding noise of the two implementations (bf16 softmax in the oracle, fp32
  # here) reaches ~6 bf16 ulps on isolated logits: stated tolerance atol = 0.2 (logits span about +-4),
  # and all but 1e-5 of the entries inside the reference's 1e-1
  # (a greedy mismatch is a near-tie here when the oracle's margin is inside that same logit tolerance: two correct
  # implementations of this depth differ by 0.02 on average and up to 0.15 on single logits, measured between the
  # persistent kernel, its whole-tile variant and the per-kernel path)
  near = _lockstep(engine, dparams, oracle, ostate, state, steps=6, atol=0.2, tie_margin=0.2)
  d = (state["logits"].cpu() - ostate["logits"]).abs()
  assert (d > 0.1 + 0.1 * ostate["logits"].abs()).float().mean() < 1e-5
  assert near <= 2


def test_generate_to_host_matches_generate():
  """mtx_decode_step_host (host tokens in, ResultTokens.data out to pinned host memory, one graph replay in between) produces what
  generate() produces; feeding the tokens back from the host buffer continues the same sequence."""
  cfg = small_config(per_device_batch_size=3, max_prefill_predict_length=16, max_target_length=48, return_log_prob=True)
  params = make_params(cfg)
  prompts = random_tokens((3, 16), cfg.vocab_size, seed=8)
  runs = []
  for use_host in (False, True, "sync", "pageable"):
    engine = maxengine.MaxEngine(cfg)
    dparams = engine.load_params(params)
    state = engine.init_decode_state()
    for slot, n in enumerate((16, 4, 9)):
      prefix, _ = engine.prefill(params=dparams, padded_tokens=prompts[slot], true_length=n)
      state = engine.insert(prefix, state, slot)
    pin = (lambda t: t) if use_host == "pageable" else (lambda t: t.pin_memory())  # pinned: the copies are graph nodes
    host_out = pin(torch.zeros(3, 3, dtype=torch.int32))
    host_lp = pin(torch.zeros(3, 1, dtype=torch.float32))
    host_in = pin(state["tokens"].cpu())
    toks = []
    for step in range(6):
      if use_host:
        state, res = engine.generate_to_host(dparams, state, host_out, host_tokens=host_in, host_log_prob=host_lp, sync=use_host == "sync")
        if use_host != "sync":
          torch.cuda.synchronize()
        assert res.data is host_out
        host_in[:, 0] = host_out[:, 0]
        toks.append((host_out.clone(), host_lp.clone()))
      else:
        state, res = engine.generate(dparams, state)
        toks.append((res.data.cpu(), res.log_prob.cpu()))
    runs.append(toks)
    with pytest.raises(ValueError):
      engine.generate_to_host(dparams, state, torch.zeros(2, 3, dtype=torch.int32))
  for other in runs[1:]:
    for (da, la), (db, lb) in zip(runs[0], other):
      assert torch.equal(da, db)
      torch.testing.assert_close(la, lb, rtol=0, atol=0)


@pytest.mark.parametrize(
    "kw",
    [
        dict(logits_via_embedding=True),                                  # tied logits, / sqrt(E) (decoders.py:552-562)
        dict(logits_via_embedding=True, normalize_embedding_logits=False),
        dict(final_logits_soft_cap=5.0),                                  # decoders.py:552-565: ignored by an untied head
        dict(logits_via_embedding=True, final_logits_soft_cap=5.0),       # decoders.py:563-565
        dict(attn_logits_soft_cap=8.0),                                   # attentions.py:1231-1236
        dict(logits_dot_in_fp32=True),                                    # decoders.py:557,571
        dict(logits_via_embedding=True, final_logits_soft_cap=5.0, attn_logits_soft_cap=8.0, logits_dot_in_fp32=True),
    ],
    ids=["tied", "tied-unnormalised", "final-softcap-untied", "final-softcap", "attn-softcap", "logits-fp32", "all"],
)
@pytest.mark.parametrize("batch", [3, 70])
def test_output_head_and_softcap_variants_match_oracle(kw, batch):
  """The output-head variants of decoders.py:537-589 and the attention soft cap, on the persistent kernel (batch 3) and on the
  per-kernel path (batch 70), against the faithful oracle."""
  cfg = small_config(per_device_batch_size=batch, max_prefill_predict_length=16, max_target_length=48, materialize_logits=True, **kw)
  params = make_params(cfg)
  slots = [0, 1, 2]
  oracle = ref.DecodeOracle(small_config(per_device_batch_size=3, max_prefill_predict_length=16, max_target_length=48, **kw), params, faithful=True)
  engine = maxengine.MaxEngine(cfg)
  dparams = engine.load_params(params)
  prompts = random_tokens((3, 16), cfg.vocab_size, seed=21)
  ostate, state = oracle.init_decode_state(), engine.init_decode_state()
  # (un-normalised tied logits are sqrt(E) times larger and so is their bf16 noise: compared on the normalised scale)
  sc = cfg.emb_dim**-0.5 if cfg.logits_via_embedding and not cfg.normalize_embedding_logits else 1.0
  for slot, n in zip(slots, (16, 7, 11)):
    padded = torch.zeros(16, dtype=torch.int64)
    padded[:n] = prompts[slot, :n]
    oprefix, _ = oracle.prefill(padded, n)
    ostate = oracle.insert(oprefix, ostate, slot)
    prefix, _ = engine.prefill(params=dparams, padded_tokens=padded, true_length=n)
    _assert_logits(prefix["logits"].cpu()[0] * sc, oprefix["logits"][0] * sc)
    prefix["tokens"].fill_(int(oprefix["tokens"].item()))
    state = engine.insert(prefix, state, slot)
  for step in range(8):
    ostate, odata = oracle.generate(ostate)
    state, result = engine.generate(dparams, state)
    _assert_logits(state["logits"].cpu()[:3] * sc, ostate["logits"] * sc)
    state["tokens"][:3] = odata[:, :1].to(state["tokens"].device)  # teacher-force the oracle's tokens


def test_chunked_prefill_with_existing_prefix_matches_one_call():
  """maxengine.py:434-440: a prompt processed by three prefill calls (existing_prefix carrying the cache and the tokens done so far)
  against the same prompt in one call and against the oracle's prefill: same cache rows and last-position logits up to the
  summation order of the attention (different chunk boundaries), same next_pos; the prefix decodes like the one-call prefix."""
  cfg = small_config(per_device_batch_size=2, max_prefill_predict_length=32, max_target_length=48, materialize_logits=True,
                     use_chunked_prefill=True, prefill_chunk_size=8)
  params = make_params(cfg)
  oracle = ref.DecodeOracle(cfg, params, faithful=True)
  engine = maxengine.MaxEngine(cfg)
  dparams = engine.load_params(params)
  toks = random_tokens((1, 32), cfg.vocab_size, seed=77)[0]
  n = 27
  whole, _ = engine.prefill(params=dparams, padded_tokens=toks, true_length=n)
  # 11 + 9 + 7 tokens, the middle chunk padded
  prefix, _ = engine.prefill(params=dparams, padded_tokens=toks[:11], true_length=11)
  assert int(prefix["next_pos"]) == 11 and prefix["cache"]["key"].shape[2] == 11
  padded = torch.zeros(12, dtype=torch.int64)
  padded[:9] = toks[11:20]
  prefix, _ = engine.prefill(params=dparams, padded_tokens=padded, true_length=9,
                             existing_prefix=maxengine.ExistingPrefix(cache=prefix["cache"], common_prefix_tokens=toks[:11]))
  assert int(prefix["next_pos"]) == 20
  prefix, result = engine.prefill(params=dparams, padded_tokens=toks[20:27], true_length=7,
                                  existing_prefix=maxengine.ExistingPrefix(cache=prefix["cache"], common_prefix_tokens=toks[:20]))
  assert int(prefix["next_pos"]) == n and int(prefix["cache"]["prefill_length"]) == n and result.data.shape == (1, 3)
  for name in ("key", "value"):
    a, b = prefix["cache"][name].float().cpu(), whole["cache"][name].float().cpu()
    assert a.shape == b.shape
    assert (a - b).abs().max() <= 2**-6 * b.abs().max()
  oprefix, ofirst = oracle.prefill(toks, n)
  _assert_logits(prefix["logits"].cpu()[0], oprefix["logits"][0])
  _assert_logits(whole["logits"].cpu()[0], oprefix["logits"][0])
  # and it decodes: insert + 4 lock-step steps
  prefix["tokens"].fill_(int(ofirst))
  state = engine.insert(prefix, engine.init_decode_state(), 1)
  ostate = oracle.insert(oprefix, oracle.init_decode_state(), 1)
  for _ in range(4):
    ostate, odata = oracle.generate(ostate)
    state, _ = engine.generate(dparams, state)
    torch.testing.assert_close(state["logits"].cpu()[1], ostate["logits"][1], rtol=1e-1, atol=1e-1)
    state["tokens"].copy_(odata[:, :1])
  with pytest.raises(ValueError, match="chunked prefill"):
    plain = maxengine.MaxEngine(small_config())
    plain.prefill(params=plain.load_params(make_params(small_config())), padded_tokens=toks[:4], true_length=4,
                  existing_prefix=maxengine.ExistingPrefix(cache=prefix["cache"], common_prefix_tokens=toks[:4]))


def test_rebind_state_follows_moved_buffers():
  """mtx_engine_rebind_state (what an XLA FFI handler calls when its operand buffers moved): a no-op for the same buffers; after
  the KV cache has moved to new allocations the next steps continue bit-identically; quantised / paged engines refuse."""
  import ctypes

  from maxtext_indextts2_b200 import _lib

  cfg = small_config(per_device_batch_size=2, materialize_logits=True)
  params = make_params(cfg)
  prompts = random_tokens((2, 16), cfg.vocab_size, seed=4)

  def run(move):
    engine = maxengine.MaxEngine(cfg, use_cuda_graph=False)
    dparams = engine.load_params(params)
    state = engine.init_decode_state()
    for slot, n in enumerate((9, 16)):
      prefix, _ = engine.prefill(params=dparams, padded_tokens=prompts[slot], true_length=n)
      state = engine.insert(prefix, state, slot)
    out = []
    for step in range(6):
      if step == 3:
        _lib.check(engine.lib.mtx_engine_rebind_state(engine._handle, ctypes.byref(engine._state_struct)))  # nothing moved
        if move:
          engine._k, engine._v = engine._k.clone(), engine._v.clone()
          engine._state_struct.k_cache, engine._state_struct.v_cache = engine._k.data_ptr(), engine._v.data_ptr()
          _lib.check(engine.lib.mtx_engine_rebind_state(engine._handle, ctypes.byref(engine._state_struct)))
      state, result = engine.generate(dparams, state)
      out.append((result.data.cpu().clone(), state["logits"].cpu().clone()))
    return out

  for (d0, l0), (d1, l1) in zip(run(False), run(True)):
    assert torch.equal(d0, d1) and torch.equal(l0, l1)
  quant = maxengine.MaxEngine(small_config(quantize_kvcache=True, kv_quant_axis="dkv"))
  quant.load_params(make_params(cfg))
  assert quant.lib.mtx_engine_rebind_state(quant._handle, ctypes.byref(quant._state_struct)) == 0  # unchanged: still a no-op
  quant._state_struct.tokens = quant._next_pos.data_ptr()
  assert quant.lib.mtx_engine_rebind_state(quant._handle, ctypes.byref(quant._state_struct)) == _lib.MTX_ERR_UNSUPPORTED
