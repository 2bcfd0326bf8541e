"""GEGLU projection GEMMs of the three resolutions, CUDA-graph timed."""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, packing
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
dev = "cuda"
for M, C in [(81920, 320), (20480, 640), (5120, 1280)]:
    a = torch.randn(M, C, device=dev).to(torch.bfloat16)
    w = (torch.randn(8 * C, C, device=dev) * C ** -0.5).to(torch.bfloat16)
    b = torch.randn(8 * C, device=dev)
    t = graph_time(lambda: ops.gemm(a, w, bias=b, geglu=True))
    print(f"geglu M={M} N={8*C} K={C}: {t*1e3:7.1f} us  {2*M*8*C*C/t/1e9:7.1f} TF/s")
