"""conv_in / conv_out (4 <-> 320 channels at 16x40x64) timing, CUDA-graph timed."""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops
dev = "cuda"
def graph_time(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B, Fr, H, W = 2, 16, 40, 64
x = torch.randn(B, 4, Fr, H, W, device=dev)
w = torch.randn(320, 4, 3, 3, device=dev) / 6
b = torch.randn(320, device=dev)
y = ops.conv_in(x, w, b)
t_in = graph_time(lambda: ops.conv_in(x, w, b))
ss = ops.groupnorm_scale_shift(y, B, Fr * H * W, torch.ones(320, device=dev), torch.zeros(320, device=dev), 1e-5)
wo = (torch.randn(4, 320, 3, 3, device=dev) * (9 * 320) ** -0.5).permute(0, 2, 3, 1).contiguous()
bo = torch.randn(4, device=dev)
t_out = graph_time(lambda: ops.conv_out(y, ss, B, Fr, H, W, wo, bo))
print(f"conv_in {t_in*1e3:.1f} us   conv_out {t_out*1e3:.1f} us")
