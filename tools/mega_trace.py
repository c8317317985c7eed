"""Debug: grid-barrier arrive / release times of every CTA of the persistent step kernel (batch-64 bench config)."""
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
tr = torch.zeros(2 * NB * NC + NC * 64, dtype=torch.int64, device='cuda')
lib.mtx_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_lib.check(lib.mtx_decode_step(eng._handle, args.batch, st))
e1.record()
torch.cuda.synchronize()
lib.mtx_debug_set_trace(None)
print("step (prepare + persistent + finalize), events: %.1f us" % (e0.elapsed_time(e1) * 1e3))
t = tr.cpu().numpy()
# rows of gridDim.x entries: row 0 = end stamps (largest values), row 1 = start stamps (smallest)
g = int(np.argmax(t < t[0] - 100000))
print("grid", g)
a = t[: (len(t) // g) * g].reshape(-1, g)
start, end = a[1], a[0]
t0 = start.min()
print(f"start spread {(start.max()-t0)/1e3:.2f} us; end {(end.max()-t0)/1e3:.2f} us")
names = ["qkv", "attn", "oproj", "up", "down"]
prev_rel = start
rows = []
k = 1
while 2 * k + 1 < a.shape[0] and a[2 * k].min() > 0:
    arr, rel = a[2 * k], a[2 * k + 1]
    rows.append(((arr.min() - prev_rel.max()) / 1e3, np.median(arr - prev_rel.max()) / 1e3, (arr.max() - prev_rel.max()) / 1e3,
                 (rel.min() - arr.max()) / 1e3, (rel.max() - arr.max()) / 1e3))
    prev_rel = rel
    k += 1
rows = np.array(rows)
nb = len(rows)
L = (nb - 2) // 5
print("barriers", nb, "layers", L)
def show(i, nm):
    print(f"{i:3d} {nm:8s} arrive first {rows[i,0]:7.2f} median {rows[i,1]:7.2f} last {rows[i,2]:7.2f} | release first {rows[i,3]:6.2f} last {rows[i,4]:6.2f}")
show(0, "embed")
for i in range(1, min(nb, 11)):
    show(i, names[(i - 1) % 5])
print("per phase means over layers (us): first / median / last arrival after previous release, barrier latency (last release - last arrival)")
tot = 0.0
for j, nm in enumerate(names):
    sel = rows[1 + j:1 + 5 * L:5]
    print(f"  {nm:6s} {sel[:,0].mean():7.2f} {sel[:,1].mean():7.2f} {sel[:,2].mean():7.2f}   barrier {sel[:,4].mean():6.2f}")
    tot += sel[:, 2].mean() + sel[:, 4].mean()
print(f"  layer total {tot:.2f} us -> {tot * L:.1f} us for {L} layers")
show(nb - 1, "fin.norm")
print(f"logits phase: {(end.max() - prev_rel.max())/1e3:.2f} us")

# ---- event logs of layer 1 (times relative to the release of the barrier before that layer's QKV phase) ----
evs = t[2 * 200 * g:2 * 200 * g + g * 3 * 64].reshape(g, 3, 32, 2)
base = a[2 * 6 + 1].max()   # barrier 6 = after the MLP-down phase of layer 0
rel_names = {7: "qkv", 8: "attn", 9: "oproj", 10: "up", 11: "down"}
print("releases (us after base):", {nm: round((a[2 * k + 1].max() - base) / 1e3, 2) for k, nm in rel_names.items()})
print("logits phase events are relative to the final-norm barrier release:", round((prev_rel.max() - base) / 1e3, 2))
for c in [0, 1, 50, 100, 139, 147]:
    for role, rn in enumerate(["epi", "xform/attn", "producer"]):
        out = []
        for i in range(32):
            eid, tm = int(evs[c, role, i, 0]), evs[c, role, i, 1]
            if tm == 0: break
            out.append(f"{eid}:{(tm - base) / 1e3:.2f}")
        print("cta", c, rn, " ".join(out))

# ---- per-CTA lateness at the attention barrier, averaged over layers ----
late = np.zeros(g)
for l in range(L):
    k = 3 + 5 * l
    arr = a[2 * k]
    late += (arr - arr.min()) / 1e3
late /= L
order = np.argsort(late)
print("attention arrival lateness per CTA (us, mean over layers): min %.2f median %.2f max %.2f" % (late.min(), np.median(late), late.max()))
print("  earliest CTAs:", [(int(c), round(float(late[c]), 1)) for c in order[:8]])
print("  latest CTAs:  ", [(int(c), round(float(late[c]), 1)) for c in order[-12:]])
for nm, kk in (("qkv", 2), ("oproj", 4), ("up", 5), ("down", 6)):
    lt = np.zeros(g)
    for l in range(L):
        arr = a[2 * (kk + 5 * l)]
        lt += (arr - arr.min()) / 1e3
    lt /= L
    o2 = np.argsort(lt)
    print(f"{nm} lateness: median {np.median(lt):.2f} max {lt.max():.2f} latest CTAs", [(int(c), round(float(lt[c]), 1)) for c in o2[-6:]])

# ---- attention phase of layer 1, all CTAs: when the tile loop ends, when the CTA's warps have met, when the
# merges are done (events 500 / 610 / 611 / 501 of the first and the last attention warp; MTX_PK_EVENTS build) ----
def _first(c, role, eid):
    for i in range(32):
        if int(evs[c, role, i, 0]) == eid and evs[c, role, i, 1] != 0:
            return (evs[c, role, i, 1] - base) / 1e3
    return np.nan
if np.any(evs[:, 1, :, 0] == 610):
    rel_qkv = (a[2 * 7 + 1].max() - base) / 1e3
    for role, rn in ((1, "first attention warp"), (2, "last attention warp")):
        st = {e: np.array([_first(c, role, e) for c in range(g)]) - rel_qkv for e in (500, 610, 611, 501)}
        def q(x):
            x = x[~np.isnan(x)]
            return "min %.2f med %.2f p90 %.2f max %.2f" % (x.min(), np.median(x), np.percentile(x, 90), x.max()) if len(x) else "-"
        print(f"attention, {rn} (us after the QKV barrier release):")
        print("  loop start   ", q(st[500]))
        print("  loop end     ", q(st[610]))
        print("  warps met    ", q(st[611]))
        print("  merges done  ", q(st[501]))
    arr = (a[2 * 8] - base) / 1e3 - rel_qkv
    print("  barrier arrival min %.2f med %.2f max %.2f; release %.2f" % (arr.min(), np.median(arr), arr.max(), (a[2 * 8 + 1].max() - base) / 1e3 - rel_qkv))

# ---- epilogue-warp events of layer 1 over all CTAs: id -> distribution of the n-th occurrence (us after base) ----
if np.any(evs[:, 0, :, 0] != 0):
    from collections import defaultdict
    occ = defaultdict(list)
    for c in range(g):
        seen = defaultdict(int)
        for i in range(32):
            eid, tm = int(evs[c, 0, i, 0]), evs[c, 0, i, 1]
            if tm == 0: break
            occ[(eid, seen[eid])].append((tm - base) / 1e3)
            seen[eid] += 1
    print("epilogue-warp events (id, occurrence): CTAs, min / median / max us after base")
    for key in sorted(occ, key=lambda k: np.median(occ[k])):
        v = np.array(occ[key])
        print(f"  {key[0]:4d}#{key[1]}: n={len(v):3d}  {v.min():7.2f} {np.median(v):7.2f} {v.max():7.2f}")

# per-CTA arrays for offline analysis
try:
    np.save("gpurun_out/attn_lateness.npy", late)
    if np.any(evs[:, 1, :, 0] == 610):
        arrs = {}
        for role in (1, 2):
            for e in (500, 610, 611, 501):
                arrs[f"r{role}_{e}"] = np.array([_first(c, role, e) for c in range(g)]) - rel_qkv
        np.savez("gpurun_out/attn_events.npz", **arrs)
except Exception as ex:  # noqa
    print("could not save arrays:", ex)

# ---- last attention warp, every other layer: loop end (610) and merges done (501) relative to that layer's QKV
# barrier release, per CTA mean over the logged layers ----
if np.any(evs[:, 2, :, 0] == 610):
    le = np.full((g, 16), np.nan); md = np.full((g, 16), np.nan)
    for c in range(g):
        k610 = k501 = 0
        for i in range(32):
            eid, tm = int(evs[c, 2, i, 0]), evs[c, 2, i, 1]
            if tm == 0: break
            if eid == 610 and k610 < 16:
                lay = 2 * k610
                le[c, k610] = (tm - a[2 * (2 + 5 * lay) + 1].max()) / 1e3; k610 += 1
            if eid == 501 and k501 < 16:
                lay = 2 * k501
                md[c, k501] = (tm - a[2 * (2 + 5 * lay) + 1].max()) / 1e3; k501 += 1
    lem, mdm = np.nanmean(le, axis=1), np.nanmean(md, axis=1)
    print("last attention warp over %d layers: loop end mean-per-CTA min %.2f med %.2f max %.2f | merges done min %.2f med %.2f max %.2f" % (
        int(np.sum(~np.isnan(le[0]))), np.nanmin(lem), np.nanmedian(lem), np.nanmax(lem), np.nanmin(mdm), np.nanmedian(mdm), np.nanmax(mdm)))
    print("  per-layer spread of loop end (max - median over CTAs):", np.round(np.nanmax(le, axis=0) - np.nanmedian(le, axis=0), 2))
    print("  slowest CTAs by mean loop end:", [(int(c), round(float(lem[c]), 2)) for c in np.argsort(lem)[-10:]])
    print("  fastest CTAs by mean loop end:", [(int(c), round(float(lem[c]), 2)) for c in np.argsort(lem)[:10]])
    np.savez("gpurun_out/attn_all_layers.npz", loop_end=le, merges_done=md)

# ---- the CTAs whose first attention warp finishes its merges last (layer 1): full event logs ----
if np.any(evs[:, 1, :, 0] == 501):
    t501 = np.array([_first(c, 1, 501) for c in range(g)])
    for c in np.argsort(np.nan_to_num(t501))[-6:]:
        for role, rn in ((1, "first attn warp"), (2, "last attn warp")):
            out = []
            for i in range(32):
                eid, tm = int(evs[c, role, i, 0]), evs[c, role, i, 1]
                if tm == 0: break
                if (tm - base) / 1e3 > rel_qkv - 1 and eid >= 500: out.append(f"{eid}:{(tm - base) / 1e3 - rel_qkv:.2f}")
            print("late cta", int(c), rn, " ".join(out[:14]))
