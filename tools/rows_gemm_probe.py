"""Probe of gemm_rows.cuh at 256 rows: the four GEMM shapes of one IndexTTS2-scale layer back to back (programmatic
dependent launch, no graph), steady-state time per sequence, and the clock64 timeline of CTA (0,0,0) of each shape."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from maxtext_indextts2_b200 import _lib
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rt = 128 if rows <= 128 else 256
shapes = [("qkv", 1792, 1280, 0), ("out", 1280, 1280, 0), ("up", 10240, 1280, 0), ("down", 1280, 5120, 0)]
bufs = {}
for name, n, k, s in shapes:
  bufs[name] = (torch.randn(rt, k, device='cuda').to(torch.bfloat16), torch.randn(n, k, device='cuda').to(torch.bfloat16) / k**0.5,
                torch.zeros(rows, n, dtype=torch.bfloat16, device='cuda'))
def run(name, n, k, s):
  x, w, o = bufs[name]
  _lib.check(lib.mtx_linear(P(x), P(w), P(o), rows, n, k, s, st))
for sp in (0, 2, 4, 8):
  for _ in range(5):
    for name, n, k, s in shapes: run(name, n, k, sp if name != "up" else 0)
  torch.cuda.synchronize()
  for name, n, k, s in shapes:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): run(name, n, k, sp if name != "up" else 0)
    e1.record(); torch.cuda.synchronize()
    print(f"rows {rows} splits {sp or 'auto'} {name:5s} n={n} k={k}: {e0.elapsed_time(e1) * 5:.2f} us per launch (200 back to back)")
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(100):
    for name, n, k, s in shapes: run(name, n, k, sp if name != "up" else 0)
  e1.record(); torch.cuda.synchronize()
  print(f"rows {rows} splits {sp or 'auto'}: {e0.elapsed_time(e1) * 10:.2f} us per 4-GEMM layer sequence")
trace = torch.zeros(256, dtype=torch.int64, device='cuda')
for sp in (1, 4, 8):
  for name, n, k, s in shapes:
    for it in range(3):
      trace.zero_(); lib.mtx_debug_set_trace(P(trace))
      run(name, n, k, sp); torch.cuda.synchronize()
    t = trace.cpu().numpy()[:11]
    print(f"splits {sp}", name, "clock64 stamps [setup, past griddep, first stage, mma committed, acc ready, parked, cluster1, reduced, epilogue, cluster2, end]:", t.tolist(), "(cycles; 1.9 GHz)")
    lib.mtx_debug_set_trace(None)
