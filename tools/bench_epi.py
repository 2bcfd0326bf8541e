"""GEMM epilogue dissection with CUDA-graph replays (no host launch overhead in the timing)."""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
lib = _lib.load(); dev = "cuda"
def graph_time(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
shapes = [(81920, 320, 320), (81920, 1152, 320), (20480, 640, 640)]
for M, N, K in shapes:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev).to(torch.bfloat16)
    for dbg in (0,):  # 1 = no TMA loads, 2 = no MMA, 16 = no TMA stores
        lib.lavie_debug_set(2, dbg)
        t0 = graph_time(lambda: ops.gemm(a, w))
        t2 = graph_time(lambda: ops.gemm(a, w, bias=b, residual=r))
        print(f"M={M} N={N} K={K} debug={dbg:3d}: plain {t0*1e3:6.1f} us | +bias+res {t2*1e3:6.1f} us")
    lib.lavie_debug_set(2, 0)
    t = graph_time(lambda: a @ w.t())
    print(f"M={M} N={N} K={K} cuBLAS plain {t*1e3:6.1f} us")
