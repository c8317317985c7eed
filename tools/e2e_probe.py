"""Where the end-to-end overhead of a host-driven decode step goes: graph replay + sync only, + result copy, + token copy."""
import ctypes, sys, time, torch
sys.path.insert(0, '.')
import bench
from maxtext_indextts2_b200 import _lib, maxengine
lib = _lib.load()
args = bench.parse_args()
cfg = bench.make_config(args)
eng = maxengine.MaxEngine(cfg)
dp = eng.load_params(on_device_init=True)
pl, al = bench.context_lengths(args, cfg)
state = eng.fill_synthetic_context(pl, al)
B = args.batch
stream = torch.cuda.current_stream()
st = ctypes.c_void_p(stream.cuda_stream)
hin = torch.zeros(B, 1, dtype=torch.int32).pin_memory()
hout = torch.zeros(B, 3, dtype=torch.int32).pin_memory()
N = 60
def timed(fn, name):
    eng.fill_synthetic_context(pl, al)  # (the same contexts for every variant: a step gets slower as they grow)
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    for _ in range(N): fn()
    e1.record(stream); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{name:55s} {e0.elapsed_time(e1) / N * 1e3:8.1f} us/step (events)  {(t1 - t0) / N * 1e6:8.1f} us/step (wall)")
h = eng._handle
timed(lambda: lib.mtx_decode_step_graph(h, B, st), "graph replays back to back (no sync)")
def a():
    lib.mtx_decode_step_graph(h, B, st); stream.synchronize()
timed(a, "graph replay + stream sync")
def b():
    lib.mtx_decode_step_host_sync(h, B, None, ctypes.c_void_p(hout.data_ptr()), None, st)
timed(b, "step_host_sync: result D2H only")
def c():
    lib.mtx_decode_step_host_sync(h, B, ctypes.c_void_p(hin.data_ptr()), ctypes.c_void_p(hout.data_ptr()), None, st)
timed(c, "step_host_sync: tokens H2D + result D2H")
hi, ho = hin.numpy(), hout.numpy()
def d():
    eng.generate_to_host(dp, eng._state, hout, host_tokens=hin, sync=True); hi[:, 0] = ho[:, 0]
timed(d, "MaxEngine.generate_to_host(sync=True) + feed back")
