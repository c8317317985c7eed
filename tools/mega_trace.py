"""Debug: barrier arrive / release times of every CTA of the persistent step kernel (batch-64 bench config)."""
import ctypes, sys, torch, numpy as np
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
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(5):
    _lib.check(lib.mtx_decode_step(eng._handle, args.batch, st))
torch.cuda.synchronize()
NB, NC = 200, 320
tr = torch.zeros(2 * NB * NC, dtype=torch.int64, device='cuda')
lib.mtx_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
_lib.check(lib.mtx_decode_step(eng._handle, args.batch, st))
torch.cuda.synchronize()
lib.mtx_debug_set_trace(None)
t = tr.cpu().numpy()
# rows of gridDim.x entries: row 0 = end stamps (largest values), row 1 = start stamps (smallest)
g = int(np.argmax(t < t[0] - 100000))
print("grid", g)
nb = 0
a = t[: (len(t) // g) * g].reshape(-1, g)
start = a[1]; end = a[0]
t0 = start.min()
print(f"start spread {(start.max()-t0)/1e3:.2f} us; end {(end.max()-t0)/1e3:.2f} us")
names = ["norm1", "qkv", "attn", "oproj", "norm2", "up", "down"]
prev_rel = start
rows = []
k = 1
while 2 * k + 1 < a.shape[0] and a[2 * k].min() > 0:
    arr, rel = a[2 * k], a[2 * k + 1]
    rows.append(((arr.min() - prev_rel.max()) / 1e3, (arr.max() - prev_rel.max()) / 1e3, (rel.min() - arr.max()) / 1e3, (rel.max() - arr.max()) / 1e3))
    prev_rel = rel
    k += 1
rows = np.array(rows)
print("barriers", len(rows))
L = (len(rows) - 1) // 7
for i in range(min(len(rows), 15)):
    nm = names[i % 7] if i < 7 * L else "final_norm"
    print(f"{i:3d} {nm:8s} first-arrive {rows[i,0]:7.2f} last-arrive {rows[i,1]:7.2f}  release first {rows[i,2]:6.2f} last {rows[i,3]:6.2f}")
for j, nm in enumerate(names):
    sel = rows[j:7 * L:7]
    print(f"{nm:8s} mean phase(last-arrive after prev release) {sel[:,1].mean():7.2f}  first-arrive {sel[:,0].mean():7.2f}  barrier latency first {sel[:,2].mean():6.2f} last {sel[:,3].mean():6.2f}")
print("final_norm", rows[7 * L])
print(f"logits phase: {(end.max() - prev_rel.max())/1e3:.2f} us")
# ---- distribution of arrival times for the phases of layer 1, and worker event logs ----
k0 = 1 + 7
for j, nm in enumerate(names):
    k = k0 + j
    prev = a[2 * (k - 1) + 1].max()
    arr = np.sort((a[2 * k] - prev) / 1e3)
    print(nm, "arrive pct 0/10/50/90/100:", np.percentile(arr, [0, 10, 50, 90, 100]).round(2), "argmax cta", int(np.argmax(a[2 * k])))
ev = t[2 * 200 * g:2 * 200 * g + g * 64].reshape(g, 32, 2)
rel = {1: a[2 * (k0 + 0) + 1].max(), 2: a[2 * (k0 + 2) + 1].max(), 3: a[2 * (k0 + 4) + 1].max()}  # qkv after norm1 barrier, oproj after attn, up after norm2
rel_down = a[2 * (k0 + 5) + 1].max()
for c in list(range(0, 16)) + [g - 1]:
    out = []
    seen2 = 0
    for i in range(32):
        eid, tm = int(ev[c, i, 0]), ev[c, i, 1]
        if tm == 0: break
        epi = eid // 10
        base = rel.get(epi, 0)
        if epi == 2:
            # two residual phases per layer: oproj first then down
            seen2 += (eid % 10 == 0)
            base = rel[2] if seen2 <= 1 else rel_down
        out.append(f"{eid}:{(tm - base) / 1e3:.2f}")
    print("cta", c, " ".join(out))
