#!/usr/bin/env python3
"""Decode-step benchmark: audio tokens/s at batch 64 on the IndexTTS2-scale model (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the reference decode arithmetic on the host CPU

One "step" = one autoregressive decode step (MaxEngine.generate) for `batch` slots per GPU, i.e.
`batch` audio tokens per GPU.  Work is partitioned by request batch: every rank owns `batch` slots
and a full weight replica; there is no data-path collective (weak scaling).

Workload: 24 layers, emb 1280, 20/4 heads x 64, mlp 5120, vocabulary 264,192
(text + audio codec tokens), bf16, P = 1024, T = 3072, contexts uniform in [512, 1536] (seed 7),
random-init weights, random-normal KV fill (synthetic).
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

# One string for both arms: the driver divides the two lines only if their metric (and unit) agree.
METRIC = "audio tokens/s at batch 64 decode (whole job)"
ORACLE_KIND = "port (parity unpinned)"  # the reference is pure Python/JAX: nothing of it could be compiled or imported here

KERNEL_CLASSES = ["prepare", "rmsnorm", "qkv_rope_append", "attention", "out_proj", "mlp_up", "mlp_down", "logits_sample", "finalize", "persistent_step"]
NCLS = len(KERNEL_CLASSES)


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=64)
  ap.add_argument("--warmup", type=int, default=4)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--batch", type=int, default=64, help="decode slots per GPU")
  ap.add_argument("--model", default="indextts2-t2s")
  ap.add_argument("--layers", type=int, default=0, help="override base_num_decoder_layers (e.g. gemma3-27b geometry with 12 of its 62 layers)")
  ap.add_argument("--context-min", type=int, default=512)
  ap.add_argument("--context-max", type=int, default=1536)
  ap.add_argument("--prefill-len", type=int, default=0, help="override max_prefill_predict_length (BASELINE configs[3]: 4096)")
  ap.add_argument("--target-len", type=int, default=0, help="override max_target_length (BASELINE configs[3]: 5632)")
  ap.add_argument("--no-graph", action="store_true")
  ap.add_argument("--kv-int8", action="store_true", help="quantize_kvcache=True, kv_quant_dtype=int8, kv_quant_axis=dkv (SURVEY 8f-2; not the judged line)")
  ap.add_argument("--kv-fp8", action="store_true", help="quantize_kvcache=True, kv_quant_dtype=fp8 (float8_e4m3fn bytes)")
  ap.add_argument("--kv-axis", default="dkv", choices=["dkv", "heads_and_dkv"], help="kv_quant_axis of --kv-int8 / --kv-fp8")
  ap.add_argument("--paged", type=int, default=0, help="attention=paged with this many tokens per page (SURVEY 8f-4; not the judged line): "
                  "the page manager runs on the host before every step and its state is uploaded, inside the timed region")
  ap.add_argument("--no-fold", action="store_true", help="keep the RMSNorm scales out of the weights (fold_norm_scales=False)")
  ap.add_argument("--skip-cpu-baseline", action="store_true")
  ap.add_argument("--cpu-slots", type=int, default=0, help="slots in the CPU baseline sample (0 = all slots of the batch)")
  ap.add_argument("--cpu-steps", type=int, default=8, help="timed steps of the cpu_baseline leg of the GPU arm (after 2 warm-ups)")
  ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall-clock bound of a CPU leg; steps are cut (and reported) past it")
  ap.add_argument("--no-verify", action="store_true", help="skip the oracle check of one step of the timed state")
  ap.add_argument("--prompt-len", type=int, default=4000, help="--mode prefill: prompt tokens (BASELINE configs[3]: 4000 of P = 4096)")
  ap.add_argument("--mode", default="batch", choices=["batch", "vocab-parallel", "prefill"],
                  help="batch: request-batch partitioned, no collective (default); vocab-parallel: every GPU decodes the same --batch slots, "
                       "the logits projection is sharded over the GPUs and one NCCL all-gather carries the per-shard candidates")
  ap.add_argument("--sampling", default="greedy", choices=["greedy", "weighted", "topk", "nucleus"], help="decode_sampling_strategy")
  ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                  help="weak: --batch slots per GPU (default, the judged line); strong: --batch slots in total, split over the GPUs")
  return ap.parse_args()


def make_config(args):
  from maxtext_indextts2_b200 import pyconfig

  kw = {}
  if args.prefill_len:
    kw["max_prefill_predict_length"] = args.prefill_len
  if args.target_len:
    kw["max_target_length"] = args.target_len
  if args.layers:
    kw["base_num_decoder_layers"] = args.layers
  if args.no_fold:
    kw["fold_norm_scales"] = False
  if args.kv_int8 or args.kv_fp8:
    kw.update(quantize_kvcache=True, kv_quant_dtype="fp8" if args.kv_fp8 else "int8", kv_quant_axis=args.kv_axis)
  if args.paged:
    T = kw.get("max_target_length", 0) or 3072
    kw.update(attention="paged", pagedattn_tokens_per_page=args.paged, pagedattn_num_pages=args.batch * ((T + args.paged - 1) // args.paged) + 1)
  if args.sampling != "greedy":
    kw["decode_sampling_strategy"] = args.sampling
    kw["decode_sampling_top_k"] = 64
    kw["decode_sampling_nucleus_p"] = 0.9
  if args.mode == "vocab-parallel":
    kw["vocab_parallelism"] = int(os.environ.get("WORLD_SIZE", "1"))
  return pyconfig.initialize(None, model_name=args.model, per_device_batch_size=args.batch, **kw)


def context_lengths(args, cfg, rank=0):
  rng = np.random.Generator(np.random.PCG64(7 + rank))
  P = cfg.max_prefill_predict_length
  total = rng.integers(args.context_min, args.context_max + 1, size=args.batch)
  prefill = np.minimum(total, P)
  # part of the longer contexts was decoded, the rest prefetched; short ones are all prompt
  ar = total - prefill
  return prefill.astype(np.int64), ar.astype(np.int64)


def algorithmic_bytes(cfg, batch, ctx_sum, ctx_sum_local=None):
  """weights once + valid KV rows + KV append + embedding rows + outputs.  `ctx_sum_local`: rows a sliding-window
  layer reads (gemma3: five layers of six); the global layers read `ctx_sum`."""
  E, Hq, Hkv, D = cfg.emb_dim, cfg.num_query_heads, cfg.num_kv_heads, cfg.head_dim
  M, V, L = cfg.mlp_dim, cfg.vocab_size, cfg.num_decoder_layers
  w_layers = 2 * L * (E * Hq * D + 2 * E * Hkv * D + Hq * D * E + 3 * E * M)
  w_norm = 2 * (2 * L + 1) * E
  w_logits = 2 * E * V
  kv_row = L * 2 * Hkv * D * 2
  if cfg.quantize_kvcache:
    kv_row = L * 2 * Hkv * (D + 4)  # one byte per element + one fp32 scale per (token, kv head)
  kv_read = int(ctx_sum) * kv_row
  if cfg.decoder_block == "gemma3" and ctx_sum_local is not None:
    n_local = sum(1 for l in range(L) if l % 6 != 5)
    kv_read = (int(ctx_sum) * (L - n_local) + int(ctx_sum_local) * n_local) * (kv_row // L)
    w_norm = 2 * (4 * L + 1) * E + 2 * 2 * L * D
  return {
      "weights": w_layers + w_norm + w_logits,
      "kv_read": kv_read,
      "kv_write": batch * kv_row,
      "misc": batch * E * 2 + batch * 8,
      "logits_weights": w_logits,
      "attention_per_layer": kv_read // L,  # (mean over the layers: gemma3's local layers read fewer rows than its global ones)
  }


class ClockSampler:
  """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""

  QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
           "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.index = index
    self.proc = None
    self.lines = []

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20"],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append((time.monotonic(), line.strip()))

  def mark(self):
    """Start of the region whose samples count (nvidia-smi needs up to a second to deliver its first line, longer when eight
    ranks start one each: the sampler is started early and the samples before the mark are dropped)."""
    self.t_mark = time.monotonic()

  def samples_since_mark(self):
    return sum(1 for t, _ in self.lines if t >= getattr(self, "t_mark", 0.0))

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for t_line, line in self.lines:
      if t_line < getattr(self, "t_mark", 0.0):
        continue
      parts = [p.strip() for p in line.split(",")]
      if len(parts) < 7:
        continue
      try:
        sm.append(float(parts[0]))
        mx.append(float(parts[1]))
      except ValueError:
        continue
      for name, val in zip(names, parts[3:7]):
        if val.lower().startswith("active"):
          reasons.add(name)
    return {
        "sm_mhz": statistics.median(sm) if sm else None,
        "sm_max_mhz": max(mx) if mx else None,
        "samples": len(sm),
        "reasons": sorted(reasons),
    }


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference decode step, timed on the host cores
# ---------------------------------------------------------------------------------------------


def cpu_reference_run(args, cfg, slots, steps, warmup, budget_s):
  """Times oracle/decode_ref.py (dtype-faithful restatement of MaxEngine.generate) on `slots` slots of the workload.

  The reference's own JAX path cannot run here (no jax in the image); BASELINE.md section 4.  Weights and cache are
  random (one random layer repeated: the values do not matter for the time, only shapes and dtypes do)."""
  from oracle import decode_ref as ref
  from oracle import mirror

  threads = os.cpu_count() or 1
  torch.set_num_threads(threads)
  E, Hq, Hkv, D = cfg.emb_dim, cfg.num_query_heads, cfg.num_kv_heads, cfg.head_dim
  M, V, L = cfg.mlp_dim, cfg.vocab_size, cfg.num_decoder_layers
  g = torch.Generator().manual_seed(0)
  bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
  rn = lambda shape, std: bf(torch.randn(shape, generator=g) * std)
  layer = dict(
      attn_scale=torch.ones(E), wq=rn((E, Hq * D), E**-0.5 / D**0.5), wk=rn((E, Hkv * D), E**-0.5), wv=rn((E, Hkv * D), E**-0.5),
      wo=rn((Hq * D, E), (Hq * D) ** -0.5), mlp_scale=torch.ones(E), w0=rn((E, M), E**-0.5), w1=rn((E, M), E**-0.5),
      wout=rn((M, E), M**-0.5))
  layers = [{k: v.clone() for k, v in layer.items()} for _ in range(L)]
  block = rn((4096, E), 1.0)
  table = block.repeat((V + 4095) // 4096, 1)[:V].contiguous()
  weights = ref.OracleWeights(embedding=table, layers=layers, final_scale=torch.ones(E), logits=(table * E**-0.5).t().contiguous())
  oracle = mirror.make_oracle(cfg, weights, slots, faithful=True)
  state = oracle.init_decode_state()
  prefill, ar = context_lengths(args, cfg)
  prefill, ar = prefill[:slots], ar[:slots]
  c = state["cache"]
  for name in ("prefill_key", "prefill_value", "ar_key", "ar_value"):
    first = bf(torch.randn(c[name][0].shape, generator=g))
    for l in range(L):
      c[name][l] = first.clone()
  c["prefill_segment_id"] = (torch.arange(oracle.P)[None, :] < torch.from_numpy(prefill)[:, None]).to(torch.int32)
  idx = int(ar.max())
  c["ar_segment_id"] = ((torch.arange(oracle.R)[None, :] >= idx - torch.from_numpy(ar)[:, None]) & (torch.arange(oracle.R)[None, :] < idx)).to(torch.int32)
  c["ar_index"] = idx
  c["ar_lengths"] = torch.from_numpy(ar).to(torch.int32)
  state["next_pos"] = torch.from_numpy(prefill + ar).to(torch.int32).reshape(slots, 1)
  state["tokens"] = torch.randint(0, V, (slots, 1), generator=g).to(torch.int32)
  t_begin = time.perf_counter()
  done_warm = 0
  for _ in range(warmup):
    state, _ = oracle.generate(state)
    done_warm += 1
    if time.perf_counter() - t_begin > budget_s / 3:
      break
  done = 0
  t0 = time.perf_counter()
  for _ in range(steps):
    state, _ = oracle.generate(state)
    done += 1
    if time.perf_counter() - t_begin > budget_s:
      break
  dt = time.perf_counter() - t0
  cut = "" if (done == steps and done_warm == warmup) else f" (cut from {steps} steps / {warmup} warm-ups by the {budget_s:.0f} s bound)"
  return {
      "value": slots * done / dt,
      "unit": "audio tokens/s",
      "cores": threads,
      "kind": ORACLE_KIND,
      "ms_per_step": 1e3 * dt / done,
      "steps_run": done,
      "warmup_run": done_warm,
      "slots": slots,
      "sample": f"{slots} of {args.batch} slots of the same workload, {done} decode steps after {done_warm} warm-up{cut}; "
                "oracle/decode_ref.py (torch CPU, dtype-faithful restatement of MaxEngine.generate; the reference is pure Python/JAX and "
                "jax is not installable in this image, so nothing of it was compiled or imported: parity of the port is pinned by the "
                "reference's invariants only)",
  }


def run_reference_arm(args):
  """`--impl reference`: the reference decode arithmetic on the host cores, same workload, metric and unit as the GPU arm.
  All slots of one GPU's batch, --steps / --warmup honoured (bounded by --cpu-budget-s; what really ran is reported)."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  cfg = make_config(args)
  slots = args.cpu_slots or args.batch
  res = cpu_reference_run(args, cfg, slots, max(1, args.steps), max(0, args.warmup), args.cpu_budget_s)
  line = {
      "impl": "reference",
      "metric": METRIC,
      "value": res["value"],
      "unit": "audio tokens/s",
      "n_gpus": args.gpus,
      "steps": res["steps_run"],
      "warmup": res["warmup_run"],
      "steps_requested": args.steps,
      "warmup_requested": args.warmup,
      "ms_per_step": res["ms_per_step"],
      "higher_is_better": True,
      "scaling": "weak",
      "vs_baseline": None,
      "dtype": "bf16",
      "data": "synthetic",
      "config": workload_config(args, cfg, args.gpus),
      "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
      "e2e": {"value": res["value"], "unit": "audio tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
      "note": "one host runs ONE batch of the job's slots however many GPUs the job has: the CPU figure does not scale with --gpus",
  }
  print(json.dumps(line))


def workload_config(args, cfg, world=None):
  world = world or args.gpus
  per_gpu = args.batch // world if args.scaling == "strong" else args.batch
  return {
      "workload": (f"IndexTTS2-scale text-to-semantic GPT decode step (BASELINE configs[1])" if cfg.decoder_block == "llama2" else
                   f"{cfg.model_name} geometry, gemma3 block (not a BASELINE config), decode step") +
                  f": L={cfg.num_decoder_layers} E={cfg.emb_dim} "
                  f"Hq={cfg.num_query_heads} Hkv={cfg.num_kv_heads} D={cfg.head_dim} M={cfg.mlp_dim} V={cfg.vocab_size}, greedy",
      "batch_per_gpu": per_gpu,
      "global_batch": per_gpu * world,
      "context": f"uniform[{args.context_min},{args.context_max}] valid rows per slot, P={cfg.max_prefill_predict_length} T={cfg.max_target_length}",
      "parallelism": f"request-batch partitioned x{world}, no collective",
      "kv_cache": (f"{cfg.kv_quant_dtype} + fp32 scale ({cfg.kv_quant_axis})" if cfg.quantize_kvcache else "bf16") +
                  (f", paged: {cfg.pagedattn_num_pages} pages of {cfg.pagedattn_tokens_per_page} tokens, host page manager + state upload "
                   f"every step inside the timed region" if cfg.attention == "paged" else ""),
      "l2": "working set per step (weights 1.81 GB + KV 1.6 GB) exceeds the 126 MB L2; no flush needed",
  }


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------


def ncu_traffic(kernel):
  """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, or None."""
  try:
    table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    return table.get(kernel)
  except (OSError, ValueError):
    return None


def persistent_phase_times(lib, engine, B, sptr, L):
  """One traced step of the persistent kernel: microseconds per phase (mean over layers), measured from the
  release of the previous grid barrier to the release of the phase's own barrier (all CTAs, %globaltimer)."""
  from maxtext_indextts2_b200 import _lib

  words = int(lib.mtx_step_trace_words(engine._handle))
  nb_alloc = max(200, 2 + 5 * L + 2)
  tr = torch.zeros(words, dtype=torch.int64, device="cuda")
  lib.mtx_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
  try:
    _lib.check(lib.mtx_decode_step(engine._handle, B, sptr))
    torch.cuda.synchronize()
  finally:
    lib.mtx_debug_set_trace(None)
  t = tr.cpu().numpy()
  if t[0] == 0:
    return None
  g = int(np.argmax(t < t[0] - 100000))  # row 0 = end stamps (largest), row 1 = start stamps (smallest)
  if g <= 0:
    return None
  a = t[: 2 * nb_alloc * g].reshape(-1, g)
  start, end = a[1], a[0]
  rel = [start.max()]
  nb = 2 + 5 * L
  for k in range(1, nb + 1):
    if a[2 * k + 1].min() <= 0:
      return None
    rel.append(a[2 * k + 1].max())
  rel = np.array(rel, dtype=np.float64) / 1e3
  names = ["qkv_rope_append", "attention", "out_proj", "mlp_up", "mlp_down"]
  out = {"ctas": g, "embed_gather": round(float(rel[1] - rel[0]), 2)}
  for j, nm in enumerate(names):
    d = [rel[2 + 5 * l + j] - rel[1 + 5 * l + j] for l in range(L)]
    out[nm + "_per_layer"] = round(float(np.mean(d)), 2)
  out["final_norm"] = round(float(rel[nb] - rel[nb - 1]), 2)
  out["logits_sample"] = round(float(end.max() / 1e3 - rel[nb]), 2)
  out["kernel_total"] = round(float(end.max() / 1e3 - start.min() / 1e3), 2)
  out["unit"] = "us"
  return out


def verify_step(engine, dparams, cfg, B):
  """Correctness gate of the timed run: ONE more decode step of the state the benchmark just timed, replayed for four slots
  through the CPU oracle (as the checker; weights downloaded from the device, KV rows of those slots mirrored).  The GPU's
  greedy token must be the dtype-faithful oracle's argmax, or differ from it only at a near-tie (fp32-oracle margin between the
  two candidates in bf16 ulps of the top logit is reported)."""
  from oracle import mirror

  slots = sorted({0, B // 3, (2 * B) // 3, B - 1})
  weights = mirror.oracle_weights_from_device(dparams, cfg)
  faithful = mirror.make_oracle(cfg, weights, len(slots), faithful=True)
  f32 = mirror.make_oracle(cfg, weights, len(slots), faithful=False)
  fstate = mirror.mirror_state(engine, faithful, slots)
  gstate = mirror.mirror_state(engine, f32, slots)
  state, result = engine.generate(dparams, engine._state)
  torch.cuda.synchronize()
  got = result.data.cpu()[slots, 0].tolist()
  fstate, fdata = faithful.generate(fstate)
  gstate, _ = f32.generate(gstate)
  want = fdata[:, 0].tolist()
  ulps, ok = [], True
  for i, (g, w) in enumerate(zip(got, want)):
    if g != w:
      c = mirror.classify_mismatch(gstate["logits"][i, 0], g, w)
      ulps.append(round(c["ulps"], 2))
      ok = ok and c["ulps"] <= 8.0
  return {
      "ok": bool(ok),
      "slots": slots,
      "tokens_equal": sum(int(g == w) for g, w in zip(got, want)),
      "near_tie_ulps": ulps,
      "rule": "GPU greedy token == argmax of the dtype-faithful oracle on the same weights and KV rows, or fp32-oracle margin <= 8 bf16 ulps of the top logit",
      "oracle": ORACLE_KIND,
  }


def run_vocab_parallel(args, world, rank, local_rank):
  """--mode vocab-parallel: every rank decodes the same `batch` slots (weights of the layers and the KV cache replicated), rank r
  scores vocabulary rows [r V/N, (r+1) V/N) and ONE NCCL all-gather per step carries the candidates (BASELINE
  configs[4]).  Strong scaling of the logits projection only: the whole-job throughput is `batch` tokens per step."""
  import torch.distributed as dist

  from maxtext_indextts2_b200 import maxengine

  cfg = make_config(args)
  B = int(cfg.per_device_batch_size)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=False)
  dparams = engine.load_params(on_device_init=True)
  prefill, ar = context_lengths(args, cfg, 0)  # the same slots on every rank
  state = engine.fill_synthetic_context(prefill, ar)
  stream = torch.cuda.current_stream()

  def barrier():
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()

  n0 = engine.lib.mtx_launch_count()
  state, _ = engine.generate(dparams, state)
  torch.cuda.synchronize()
  launches_per_step = int(engine.lib.mtx_launch_count() - n0)
  for _ in range(max(3, args.warmup)):
    state, _ = engine.generate(dparams, state)
  barrier()
  sampler = ClockSampler(local_rank)
  sampler.start()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  ev0.record(stream)
  for _ in range(args.steps):
    state, result = engine.generate(dparams, state)
  ev1.record(stream)
  barrier()
  step_ms = ev0.elapsed_time(ev1) / args.steps
  # end to end: host tokens in, sampled tokens out, one sync per step
  host_in = torch.zeros(B, 1, dtype=torch.int32).pin_memory()
  host_out = torch.zeros(B, 3, dtype=torch.int32).pin_memory()
  host_in.copy_(state["tokens"].cpu())
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record(stream)
  for _ in range(args.steps):
    state["tokens"].copy_(host_in, non_blocking=True)
    state, result = engine.generate(dparams, state)
    host_out.copy_(result.data, non_blocking=True)
    stream.synchronize()
    host_in[:, 0] = host_out[:, 0]
  e1.record(stream)
  barrier()
  e2e_ms = e0.elapsed_time(e1) / args.steps
  clocks = sampler.stop()
  # the collective alone
  cand = engine.candidate_buffer(B)
  for _ in range(5):
    engine._gather(cand)
  barrier()
  g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  g0.record(stream)
  for _ in range(50):
    engine._gather(cand)
  g1.record(stream)
  barrier()
  gather_us = g0.elapsed_time(g1) / 50 * 1e3
  tok = result.data[:, 0].clone().to(torch.int64)
  lo, hi = tok.clone(), tok.clone()
  dist.all_reduce(lo, op=dist.ReduceOp.MIN)
  dist.all_reduce(hi, op=dist.ReduceOp.MAX)
  times = torch.tensor([step_ms, e2e_ms, gather_us], dtype=torch.float64, device="cuda")
  dist.all_reduce(times, op=dist.ReduceOp.MAX)
  step_ms, e2e_ms, gather_us = times.tolist()
  if rank != 0:
    return None
  V = cfg.vocab_size
  return {
      "metric": METRIC,
      "mode": "vocab-parallel",
      "value": B * 1e3 / step_ms,
      "unit": "audio tokens/s",
      "n_gpus": world,
      "steps": args.steps,
      "warmup": max(3, args.warmup),
      "ms_per_step": step_ms,
      "higher_is_better": True,
      "scaling": "strong",
      "vs_baseline": None,
      "dtype": "bf16",
      "data": "synthetic",
      "config": {**workload_config(args, cfg, 1), "parallelism": f"vocab-parallel logits x{world}: V/N = {V // world} rows per GPU, layers and KV replicated, "
                 f"one NCCL all-gather of the per-shard candidates per step ({args.sampling} sampling)"},
      "e2e": {"value": B * 1e3 / e2e_ms, "unit": "audio tokens/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": B * 4, "d2h_bytes_per_step": B * 12,
              "api": "MaxEngine.generate (vocab_parallelism=N) with pinned host token buffers, one stream sync per step"},
      "gpu_launches": launches_per_step * args.steps,
      "launches_per_step": launches_per_step,
      "cuda_graph": False,
      "clocks": clocks,
      "allgather": {"us": gather_us, "payload_bytes_per_rank": int(cand.numel() * 4), "floats_per_row": int(cand.numel() // B),
                    "what": "dist.all_gather_into_tensor of the candidate payload, 50 back to back, max over ranks"},
      "tokens_identical_on_all_ranks": bool(torch.equal(lo, hi)),
  }


def run_prefill(args):
  """--mode prefill: MaxEngine.prefill of one `--prompt-len`-token prompt (BASELINE configs[3]: the 4k text + reference-audio prompt,
  timed separately from the decode steps), chunks of prefill_chunk_size rows through the tcgen05 GEMMs and the causal attention
  kernel, then MaxEngine.insert into slot 0.  Not the judged metric: a second line for profiles/."""
  from maxtext_indextts2_b200 import maxengine

  if not args.prefill_len:
    args.prefill_len, args.target_len = 4096, 5632
  args.batch = max(1, min(args.batch, 8))
  cfg = make_config(args)
  engine = maxengine.MaxEngine(cfg)
  dparams = engine.load_params(on_device_init=True)
  state = engine.init_decode_state()
  n = min(args.prompt_len, cfg.max_prefill_predict_length)
  tokens = torch.randint(0, cfg.vocab_size, (cfg.max_prefill_predict_length,), generator=torch.Generator().manual_seed(1))
  stream = torch.cuda.current_stream()
  for _ in range(max(2, args.warmup)):
    prefix, _ = engine.prefill(params=dparams, padded_tokens=tokens, true_length=n)
  torch.cuda.synchronize()
  steps = max(1, min(args.steps, 20))
  n0 = engine.lib.mtx_launch_count()
  e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
  pre_ms, ins_ms = 0.0, 0.0
  for _ in range(steps):
    e0.record(stream)
    prefix, first = engine.prefill(params=dparams, padded_tokens=tokens, true_length=n)
    e1.record(stream)
    state = engine.insert(prefix, state, 0)
    e2.record(stream)
    torch.cuda.synchronize()
    pre_ms += e0.elapsed_time(e1)
    ins_ms += e1.elapsed_time(e2)
  launches = int(engine.lib.mtx_launch_count() - n0) // steps
  E, Hq, Hkv, D, M, L = cfg.emb_dim, cfg.num_query_heads, cfg.num_kv_heads, cfg.head_dim, cfg.mlp_dim, cfg.num_decoder_layers
  gemm_flops = 2.0 * n * L * (E * Hq * D + 2 * E * Hkv * D + Hq * D * E + 3 * E * M)
  attn_flops = 4.0 * (n * (n + 1) / 2) * L * Hq * D
  line = {
      "metric": "prefill ms for one prompt (BASELINE configs[3], timed separately from decode)",
      "mode": "prefill",
      "value": pre_ms / steps,
      "unit": "ms",
      "higher_is_better": False,
      "n_gpus": 1,
      "steps": steps,
      "insert_ms": ins_ms / steps,
      "prompt_tokens": n,
      "prompt_tokens_per_s": n / (pre_ms / steps / 1e3),
      "tflops": (gemm_flops + attn_flops) / (pre_ms / steps / 1e3) / 1e12,
      "gpu_launches_per_prefill": launches,
      "dtype": "bf16",
      "data": "synthetic",
      "config": {"workload": f"L={L} E={E} Hq={Hq} Hkv={Hkv} D={D} M={M} V={cfg.vocab_size}, prompt {n} of P={cfg.max_prefill_predict_length}, chunk {cfg.prefill_chunk_size}"},
  }
  print(json.dumps(line), flush=True)


def main():
  args = parse_args()
  if args.impl == "reference":
    run_reference_arm(args)
    return
  if args.mode == "prefill":
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    run_prefill(args)
    return

  import torch.distributed as dist

  from maxtext_indextts2_b200 import _lib, maxengine

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if world != args.gpus and world > 1:
    raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
  torch.cuda.set_device(local_rank)
  # Libraries (NCCL's version banner) write to file descriptor 1; rank 0 must print exactly ONE JSON line there.
  # Everything until that line goes to stderr.
  sys.stdout.flush()
  saved_stdout = os.dup(1)
  os.dup2(2, 1)
  if world > 1:
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
      os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
  if _lib.needs_build():
    _lib.build()
  lib = _lib.load()

  if args.mode == "vocab-parallel":
    if world < 2:
      raise SystemExit("--mode vocab-parallel needs torchrun with at least 2 ranks")
    line = run_vocab_parallel(args, world, rank, local_rank)
    if rank == 0:
      sys.stdout.flush()
      os.dup2(saved_stdout, 1)
      print(json.dumps(line), flush=True)
      os.dup2(2, 1)
    dist.barrier()
    dist.destroy_process_group()
    return

  if args.scaling == "strong":
    if args.batch % world:
      raise SystemExit(f"--scaling strong: --batch {args.batch} is not divisible by {world} GPUs")
    args.batch_total, args.batch = args.batch, args.batch // world
  cfg = make_config(args)
  if args.scaling == "strong":
    args.batch = args.batch_total  # workload_config / context draw are stated on the global batch
  B = int(cfg.per_device_batch_size)
  engine = maxengine.MaxEngine(cfg, use_cuda_graph=not args.no_graph)
  dparams = engine.load_params(on_device_init=True)
  prefill, ar = context_lengths(args, cfg, rank if args.scaling == "weak" else 0)
  if args.scaling == "strong":  # one draw for the whole job; this rank owns slots [rank * B, (rank + 1) * B)
    prefill, ar = prefill[rank * B : (rank + 1) * B], ar[rank * B : (rank + 1) * B]
  state = engine.fill_synthetic_context(prefill, ar)
  ctx_sum = int((prefill + ar).sum()) + B  # the appended row is read too
  ctx_sum_local = None
  if cfg.decoder_block == "gemma3":
    # rows inside the cache-index window of a local layer (attentions.py:600-602,624-631): prefill rows [P - W, P) and ring
    # indices [R - W, R); the ring rows of slot s are the al[s] + 1 indices ending at the shared ring index
    P_, R_, W_ = cfg.max_prefill_predict_length, cfg.max_target_length - cfg.max_prefill_predict_length, int(cfg.sliding_window_size)
    idx = int(ar.max())
    tot = 0
    for s_ in range(B):
      tot += max(0, int(prefill[s_]) - max(0, P_ - W_))
      ring = (idx - np.arange(int(ar[s_]) + 1)) % R_
      tot += int((ring >= max(0, R_ - W_)).sum())
    ctx_sum_local = tot
  step_c = lib.mtx_decode_step if args.no_graph else lib.mtx_decode_step_graph
  stream = torch.cuda.current_stream()
  sptr = ctypes.c_void_p(stream.cuda_stream)

  def step_fn(handle, rows, sp):
    if args.paged:  # what MaxEngine.generate does before the step: PageManager.update_decode_pages + upload of the page state
      engine._advance_pages()
    return step_c(handle, rows, sp)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  sampler = ClockSampler(local_rank)
  sampler.start()  # (early: see ClockSampler.mark)
  # ---- kernel launches per step (counted on one eager step) and per-class times ----
  n0 = lib.mtx_launch_count()
  if args.paged:
    engine._advance_pages()
  _lib.check(lib.mtx_decode_step(engine._handle, B, sptr))
  torch.cuda.synchronize()
  launches_per_step = int(lib.mtx_launch_count() - n0)

  for _ in range(max(3, args.warmup)):
    _lib.check(step_fn(engine._handle, B, sptr))
  barrier()

  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  sampler.mark()
  ev0.record(stream)
  for _ in range(args.steps):
    _lib.check(step_fn(engine._handle, B, sptr))
  ev1.record(stream)
  barrier()
  elapsed_ms = ev0.elapsed_time(ev1)

  # ---- end to end through the public API with HOST buffers: this step's tokens in (pinned), result tokens out (pinned) ----
  host_in = torch.zeros(B, 1, dtype=torch.int32).pin_memory()
  host_out = torch.zeros(B, 3, dtype=torch.int32).pin_memory()
  # The same workload as the device-timed loop: its contexts (a step gets ~0.2 us slower per step it has run at batch 64: every
  # step appends a row to every slot), the same number of untimed steps before the timed ones.
  state = engine.fill_synthetic_context(prefill, ar)
  host_in.copy_(state["tokens"].cpu())
  for _ in range(1 + max(3, args.warmup)):
    state, result = engine.generate_to_host(dparams, state, host_out, host_tokens=host_in)
    torch.cuda.synchronize()
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(stream)
  hin, hout = host_in.numpy(), host_out.numpy()  # (views of the pinned buffers)
  for _ in range(args.steps):
    # H2D of this step's input tokens, the step, D2H of the sampled tokens, and the wait for them: one call per step (the
    # caller needs the tokens before the next step: detokenise / stop check)
    state, result = engine.generate_to_host(dparams, state, host_out, host_tokens=host_in, sync=True)
    hin[:, 0] = hout[:, 0]
  e1.record(stream)
  barrier()
  e2e_ms = e0.elapsed_time(e1)
  clock_window = "timed regions"
  if sampler.samples_since_mark() < 2:
    # the two timed loops were over before nvidia-smi delivered two samples (short runs): the same steps, untimed, until it has
    clock_window = "the same decode steps repeated (untimed) right after the timed regions, until two samples arrived"
    t_burst = time.monotonic()
    while sampler.samples_since_mark() < 2 and time.monotonic() - t_burst < 5.0:
      for _ in range(50):
        _lib.check(step_fn(engine._handle, B, sptr))
      torch.cuda.synchronize()
  clocks = sampler.stop()
  clocks["window"] = clock_window

  # ---- per-kernel-class device times of one eager step (CUDA events on the launching stream) ----
  class_ms = (ctypes.c_float * NCLS)()
  class_n = (ctypes.c_int32 * NCLS)()
  acc = np.zeros(NCLS)
  prof_steps = 3
  for _ in range(prof_steps):
    if args.paged:
      engine._advance_pages()
    _lib.check(lib.mtx_profile_decode_step(engine._handle, B, sptr, class_ms, class_n))
    acc += np.array(list(class_ms))
  acc /= prof_steps
  counts = list(class_n)
  phases = persistent_phase_times(lib, engine, B, sptr, cfg.num_decoder_layers) if counts[NCLS - 1] > 0 else None

  times = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device="cuda")
  if world > 1:
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
  elapsed_ms, e2e_ms = times.tolist()

  if rank == 0:
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
      pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    ab = algorithmic_bytes(cfg, B, ctx_sum + B * args.steps / 2, ctx_sum_local)
    step_bytes = ab["weights"] + ab["kv_read"] + ab["kv_write"] + ab["misc"]
    ms_per_step = elapsed_ms / args.steps
    tokens_per_s = world * B * args.steps / (elapsed_ms / 1e3)
    step_gbs = step_bytes / (ms_per_step * 1e-3) / 1e9

    # dominant kernel class by device time
    dom = int(np.argmax(acc))
    L = cfg.num_decoder_layers
    E, Hq, Hkv, D, M, V = cfg.emb_dim, cfg.num_query_heads, cfg.num_kv_heads, cfg.head_dim, cfg.mlp_dim, cfg.vocab_size
    per_launch_bytes = {
        "persistent_step": step_bytes,
        "attention": ab["attention_per_layer"] + B * Hq * D * 2 * 2,
        "logits_sample": 2 * E * V + B * E * 2,
        "mlp_up": 2 * 2 * M * E + B * E * 2 + B * M * 2,
        "mlp_down": 2 * M * E + B * M * 2 + 2 * B * E * 2,
        "qkv_rope_append": 2 * (Hq + 2 * Hkv) * D * E + B * E * 2 + B * (Hq + 2 * Hkv) * D * 2,
        "out_proj": 2 * Hq * D * E + 3 * B * E * 2,
    }
    name = KERNEL_CLASSES[dom]
    n_launch = max(1, counts[dom])
    dur_ms = acc[dom] / n_launch
    achieved = per_launch_bytes.get(name, 0) / (dur_ms * 1e-3) / 1e9 if dur_ms > 0 else 0.0
    roofline = {
        "bound": "hbm",
        "kernel": name,
        "achieved": achieved,
        "peak": hbm_peak,
        "unit": "GB/s",
        "frac": achieved / hbm_peak,
        "traffic": ncu_traffic(name),
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": per_launch_bytes.get(name, 0),
        "launch_ms": dur_ms,
        "whole_step": {"algorithmic_bytes": step_bytes, "achieved_gbs": step_gbs, "frac": step_gbs / hbm_peak},
        "class_ms_per_step": {KERNEL_CLASSES[i]: round(float(acc[i]), 4) for i in range(NCLS)},
        "class_launches_per_step": {KERNEL_CLASSES[i]: counts[i] for i in range(NCLS)},
        "persistent_step_phases": phases,
        "class_gbs": {k: round(per_launch_bytes[k] * counts[KERNEL_CLASSES.index(k)] / (acc[KERNEL_CLASSES.index(k)] * 1e-3) / 1e9, 1)
                      for k in per_launch_bytes if acc[KERNEL_CLASSES.index(k)] > 0},
    }
    verify = None
    if not args.no_verify:  # (attention=paged: the mirror gathers the slots' rows from the page pools)
      verify = verify_step(engine, dparams, cfg, B)
    cpu = None
    if not args.skip_cpu_baseline and world == 1:
      cpu = cpu_reference_run(args, cfg, args.cpu_slots or B, args.cpu_steps, 2, min(args.cpu_budget_s, 40.0))
      cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "metric": METRIC,
        "value": tokens_per_s,
        "unit": "audio tokens/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "bf16" if not cfg.quantize_kvcache else "bf16 (int8 KV cache)",
        "data": "synthetic",
        "config": workload_config(args, cfg, world),
        "e2e": {
            "value": world * B * args.steps / (e2e_ms / 1e3),
            "unit": "audio tokens/s",
            "ms_per_step": e2e_ms / args.steps,
            "h2d_bytes_per_step": B * 4,
            "d2h_bytes_per_step": B * 3 * 4,
            "api": "MaxEngine.generate_to_host(sync=True) (mtx_decode_step_host_sync): pinned host token buffer in, ResultTokens.data to a pinned host buffer, the copies are nodes of the step's CUDA graph, one stream sync per step",
        },
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "cuda_graph": not args.no_graph,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "verify": verify,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
