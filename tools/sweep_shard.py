"""Planner choice vs forced tile width at the per-rank shapes of the 8-GPU run (batch 1, 4 frames)."""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
from lavie_b200.packing import pack_conv3x3
lib = _lib.load(); dev = "cuda"
def graph_time(fn, n=20):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
NF = 4
for H, W, C, N in [(40, 64, 320, 320), (40, 64, 640, 320), (20, 32, 640, 640), (20, 32, 1280, 640), (10, 16, 1280, 1280),
                   (10, 16, 2560, 1280), (5, 8, 1280, 1280), (5, 8, 2560, 1280)]:
    x = torch.randn(NF * H * W, C, device=dev).to(torch.bfloat16)
    w = pack_conv3x3(torch.randn(N, C, 3, 3, device=dev) * (9 * C) ** -0.5)
    t0 = graph_time(lambda: ops.conv3x3(x, NF, H, W, w))
    res = []
    for bn in (64, 128, 160, 192, 256, 320):
        res.append((graph_time(lambda: ops.conv3x3(x, NF, H, W, w, block_n=bn)), bn))
    res.sort()
    print(f"conv M={NF*H*W} C={C} N={N}: planner {t0*1e3:6.1f} us | best " + ", ".join(f"bn{bn}: {t*1e3:.1f}" for t, bn in res[:3]), flush=True)
for M, N, K in [(10240, 320, 320), (10240, 1152, 320), (10240, 320, 1280), (2560, 640, 640), (2560, 1920, 640),
                (640, 1280, 1280), (640, 3840, 1280), (160, 1280, 1280), (640, 1280, 5120)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    t0 = graph_time(lambda: ops.gemm(a, w, bias=b))
    res = []
    for bn in (64, 128, 160, 192, 256, 320):
        res.append((graph_time(lambda: ops.gemm(a, w, bias=b, block_n=bn)), bn))
    res.sort()
    print(f"gemm M={M} N={N} K={K}: planner {t0*1e3:6.1f} us | best " + ", ".join(f"bn{bn}: {t*1e3:.1f}" for t, bn in res[:3]), flush=True)
