"""Does mainloop throughput follow L2->SM bytes per flop?  Same A operand, output tile width BN swept."""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
from lavie_b200.packing import pack_conv3x3
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
NF, H, W, C, N = 32, 40, 64, 320, 1280
x = torch.randn(NF * H * W, C, device=dev).to(torch.bfloat16)
w = pack_conv3x3(torch.randn(N, C, 3, 3, device=dev) * (9 * C) ** -0.5)
a = torch.randn(NF * H * W, 9 * C, device=dev).to(torch.bfloat16)
w2 = (torch.randn(N, 9 * C, device=dev) * 0.01).to(torch.bfloat16)
fl = 2.0 * NF * H * W * N * 9 * C
for bn in (160, 256, 320):
    ms = graph_time(lambda: ops.conv3x3(x, NF, H, W, w, block_n=bn))
    ms2 = graph_time(lambda: ops.gemm(a, w2, block_n=bn))
    bpc = (16384 + bn * 64) / (bn * 1.994)
    print(f"bn={bn:3d} ({bpc:5.1f} B/clk/SM at peak): conv {ms*1e3:7.1f} us {fl/ms/1e9:7.1f} TF/s | dense same shape {ms2*1e3:7.1f} us {fl/ms2/1e9:7.1f} TF/s")

# the N = 320 / 640 layers of the model
for NF, H, W, C, N in [(32, 40, 64, 320, 320), (32, 40, 64, 640, 320), (32, 20, 32, 640, 640), (32, 20, 32, 1280, 640)]:
    x = torch.randn(NF * H * W, C, device=dev).to(torch.bfloat16)
    w = pack_conv3x3(torch.randn(N, C, 3, 3, device=dev) * (9 * C) ** -0.5)
    fl = 2.0 * NF * H * W * N * 9 * C
    for bn in (0, 160, 256, 320):
        ms = graph_time(lambda: ops.conv3x3(x, NF, H, W, w, block_n=bn))
        print(f"conv {H}x{W} C={C} N={N} bn={bn:3d}: {ms*1e3:7.1f} us {fl/ms/1e9:7.1f} TF/s")
