import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
from lavie_b200.packing import pack_conv3x3
lib = _lib.load(); dev = "cuda"
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
for NF, H, W, C, N in [(32, 40, 64, 320, 320), (32, 40, 64, 640, 320), (32, 20, 32, 640, 640), (32, 10, 16, 1280, 1280),
                       (32, 5, 8, 1280, 1280), (32, 5, 8, 2560, 1280)]:
    x = torch.randn(NF * H * W, C, device=dev).to(torch.bfloat16)
    w = pack_conv3x3(torch.randn(N, C, 3, 3, device=dev) * (9 * C) ** -0.5)
    fl = 2.0 * NF * H * W * N * 9 * C
    res = []
    for dbg in (1024, 0, 2):
        lib.lavie_debug_set(2, dbg)
        t = graph_time(lambda: ops.conv3x3(x, NF, H, W, w))
        res.append(f"{t*1e3:7.1f} us {fl/t/1e9:7.1f} TF/s")
    lib.lavie_debug_set(2, 0)
    print(f"conv M={NF*H*W:6d} N={N:5d} K={9*C:6d}: tiled-4D {res[0]} | im2col-mode {res[1]} | im2col-mode TMA-only {res[2]}")
