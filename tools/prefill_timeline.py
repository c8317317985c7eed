"""Debug: per-kernel start / end (first CTA, %globaltimer) of one 4k-token prefill: where the time of a chunk goes.
  python tools/prefill_timeline.py [prompt_len]"""
import collections, ctypes, sys, torch, numpy as np
sys.path.insert(0, '.')
from maxtext_indextts2_b200 import _lib, maxengine, pyconfig
lib = _lib.load()
lib.mtx_debug_set_timeline.argtypes = [ctypes.c_void_p]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cfg = pyconfig.initialize(None, model_name="indextts2-t2s", per_device_batch_size=1, max_prefill_predict_length=4096, max_target_length=5632)
eng = maxengine.MaxEngine(cfg)
dp = eng.load_params(on_device_init=True)
eng.init_decode_state()
tokens = torch.randint(0, cfg.vocab_size, (4096,))
for _ in range(2):
  eng.prefill(params=dp, padded_tokens=tokens, true_length=n)
torch.cuda.synchronize()
tl = torch.zeros(4 + 3 * 3000 + 16, dtype=torch.int64, device='cuda')
_lib.check(lib.mtx_debug_set_timeline(ctypes.c_void_p(tl.data_ptr())))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.prefill(params=dp, padded_tokens=tokens, true_length=n); e1.record()
torch.cuda.synchronize()
_lib.check(lib.mtx_debug_set_timeline(None))
t = tl.cpu().numpy()
cnt = min(int(t[0]), 1000); e = t[4:4 + 3 * cnt].reshape(cnt, 3)
e = e[e[:, 2] > 0]
e = e[np.argsort(e[:, 1])]
names = {0: 'prepare', 1: 'rmsnorm', 3: 'attn_decode', 8: 'finalize', 31: 'rows_qkv', 32: 'rows_residual', 33: 'rows_swiglu', 34: 'rows_logits', 40: 'attn_prefill'}
print(f"prefill {n} tokens: {e0.elapsed_time(e1):.2f} ms; {cnt} timeline entries (first ~{cnt // 122} chunks)")
agg, gaps = collections.defaultdict(list), collections.defaultdict(list)
pe = None
for k, s, en in e:
  agg[int(k)].append((en - s) / 1e3)
  if pe is not None: gaps[int(k)].append((s - pe) / 1e3)
  pe = en
for k in sorted(agg):
  print(f"{str(names.get(k, k)):14s} n {len(agg[k]):4d} mean dur {np.mean(agg[k]):7.2f} us  max {np.max(agg[k]):7.2f}  mean gap after previous kernel's end {np.mean(gaps[k]) if gaps[k] else 0:7.2f} us")
print("span of the recorded kernels:", (e[:, 2].max() - e[0, 1]) / 1e3, "us; sum of durations", sum(sum(v) for v in agg.values()), "us")
# attention duration by chunk
att = [(s, (en - s) / 1e3) for k, s, en in e if int(k) in (40, 3)]
print("attention duration of every 24th launch (one per chunk):", [round(d, 1) for _, d in att[::24]])
