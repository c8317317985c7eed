"""Debug: clock64 timeline of the GEMM / attention kernels (one CTA)."""
import ctypes, sys, torch, numpy as np
sys.path.insert(0, '.')
from maxtext_indextts2_b200 import _lib
lib = _lib.load()
lib.mtx_debug_set_trace.argtypes = [ctypes.c_void_p]
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
trace = torch.zeros(256, dtype=torch.int64, device='cuda')

def gemm(rows, n, k, splits):
    rt = 16
    while rt < rows: rt *= 2
    x = torch.randn(rt, k, device='cuda').to(torch.bfloat16); w = torch.randn(n, k, device='cuda').to(torch.bfloat16)
    out = torch.zeros(rows, n, dtype=torch.bfloat16, device='cuda')
    for it in range(3):
        trace.zero_(); lib.mtx_debug_set_trace(P(trace))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(lib.mtx_linear(P(x), P(w), P(out), rows, n, k, splits, st)); e1.record()
        torch.cuda.synchronize()
    t = trace.cpu().numpy()
    print(f"GEMM rows={rows} n={n} k={k} splits={splits}: {e0.elapsed_time(e1)*1e3:.1f} us; W bytes {n*k*2/1e6:.1f} MB")
    print("  producer: start-of-issue", t[0], "after griddep", t[1], "refill issue times", t[2:2+24][t[2:26]>0])
    print("  mma: full-arrived times", t[40:40+24][t[40:64]>0], "all committed", t[39])
    print("  epilogue: tmem_full", t[80], "parked", t[82], "after cluster sync1", t[83], "reduced", t[84], "epilogue done", t[85], "after sync2", t[86], "cta end", t[81])
    lib.mtx_debug_set_trace(None)

def attn(B, Hq, Hkv, D, P_, T, ctx):
    q = torch.randn(B, Hq*D, device='cuda').to(torch.bfloat16)
    K = torch.randn(B, Hkv, T, D, device='cuda').to(torch.bfloat16); V = torch.randn_like(K)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device='cuda')
    plane, len0, rf, rl = i32(list(range(B))), i32([min(c, P_) for c in ctx]), i32([0]*B), i32([max(0, c-P_) for c in ctx])
    out = torch.zeros(B, Hq*D, dtype=torch.bfloat16, device='cuda')
    scratch = torch.empty(lib.mtx_attention_scratch_bytes(B, Hkv, Hq, D, P_, T), dtype=torch.uint8, device='cuda')
    for it in range(3):
        trace.zero_(); lib.mtx_debug_set_trace(P(trace))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(lib.mtx_decode_attention(P(q), P(K), P(V), P(plane), P(len0), P(rf), P(rl), P(out), B, B, Hq, Hkv, D, P_, T, 0.0, P(scratch), st)); e1.record()
        torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(-1, 8)
    mb = sum(ctx) * Hkv * D * 2 * 2 / 1e6
    print(f"ATTN B={B} ctx_sum={sum(ctx)}: {e0.elapsed_time(e1)*1e3:.1f} us (incl. worklist kernel + memset); {mb:.1f} MB")
    print("  per item [start, K arrived, S done, V arrived, PV done, smem merged, before end sync, after end sync]")
    for row in t[:6]:
        if row.any(): print("   ", row)
    lib.mtx_debug_set_trace(None)

gemm(64, 10240, 1280, 1)
gemm(64, 1792, 1280, 8)
gemm(64, 1280, 5120, 16)
gemm(64, 1280, 1280, 16)
gemm(64, 264192, 1280, 1)
rng = np.random.default_rng(7)
attn(64, 20, 4, 64, 1024, 3072, rng.integers(512, 1537, size=64).tolist())
attn(8, 20, 4, 64, 1024, 3072, [1024]*8)
