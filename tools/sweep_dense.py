"""Dense GEMM shapes of the model (+bias +residual): planner choice vs forced BLOCK_N, CUDA-graph timed."""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
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
shapes = [(81920, 320, 320, True), (81920, 320, 1280, True), (20480, 640, 640, True), (20480, 640, 2560, True),
          (5120, 1280, 1280, True), (5120, 1280, 5120, True), (81920, 1152, 320, False), (20480, 1920, 640, False),
          (5120, 3840, 1280, False), (81920, 320, 640, True), (1280, 1280, 1280, True)]
for M, N, K, res in shapes:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).to(torch.bfloat16) if res else None
    t0 = graph_time(lambda: ops.gemm(a, w, bias=b, residual=r))
    out = []
    for bn in (128, 160, 192, 256, 320):
        t = graph_time(lambda: ops.gemm(a, w, bias=b, residual=r, block_n=bn))
        out.append(f"bn{bn}: {t*1e3:5.1f}")
    print(f"M={M} N={N} K={K} res={int(res)}: planner {t0*1e3:5.1f} us ({2.0*M*N*K/t0/1e9:5.0f} TF/s) | " + "  ".join(out), flush=True)
