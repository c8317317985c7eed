"""Debug: per-kernel start/end inside one CUDA-graph replay of the batch-64 decode step."""
import ctypes, sys, torch, numpy as np
sys.path.insert(0, '.')
import bench
from maxtext_indextts2_b200 import _lib, maxengine
lib = _lib.load()
lib.mtx_debug_set_timeline.argtypes = [ctypes.c_void_p]
args = bench.parse_args()
cfg = bench.make_config(args)
eng = maxengine.MaxEngine(cfg)
dp = eng.load_params(on_device_init=True)
pl, al = bench.context_lengths(args, cfg)
state = eng.fill_synthetic_context(pl, al)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(5):
    _lib.check(lib.mtx_decode_step_graph(eng._handle, args.batch, st))
torch.cuda.synchronize()
tl = torch.zeros(4096, dtype=torch.int64, device='cuda')
_lib.check(lib.mtx_debug_set_timeline(ctypes.c_void_p(tl.data_ptr())))
_lib.check(lib.mtx_decode_step_graph(eng._handle, args.batch, st))
torch.cuda.synchronize()
_lib.check(lib.mtx_debug_set_timeline(None))
t = tl.cpu().numpy()
n = int(t[0]); e = t[4:4+3*n].reshape(n, 3)
names = {0:'prepare',1:'rmsnorm',3:'attention',8:'finalize',10:'gemm_store',11:'qkv',12:'residual',13:'swiglu',14:'logits'}
e = e[np.argsort(e[:,1])]
t0 = e[0,1]
print('entries', n, 'total us', (e[:,2].max()-t0)/1e3)
prev_end = t0
for i,(k,s,en) in enumerate(e[:40]):
    print(f"{i:3d} {names.get(int(k),k):10s} start {(s-t0)/1e3:8.2f} dur {(en-s)/1e3:7.2f} gap_from_prev_end {(s-prev_end)/1e3:7.2f}")
    prev_end = en
# aggregate: per kind mean duration and mean gap before
import collections
agg = collections.defaultdict(list); gaps = collections.defaultdict(list)
pe = None
for k,s,en in e:
    agg[int(k)].append((en-s)/1e3)
    if pe is not None: gaps[int(k)].append((s-pe)/1e3)
    pe = en
for k in agg: print(names.get(k,k), 'n', len(agg[k]), 'mean dur', np.mean(agg[k]).round(2), 'mean start-after-prev-end', np.mean(gaps[k]).round(2) if gaps[k] else None)
