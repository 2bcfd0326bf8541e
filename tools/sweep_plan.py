"""Sweep (BLOCK_N, split-K) for the small-M convolutions and compare with the planner's own choice (bn=0, splits=0)."""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
from lavie_b200.packing import pack_conv3x3
lib = _lib.load(); dev = "cuda"
def graph_time(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
shapes = [(32, 10, 16, 1280, 1280), (32, 10, 16, 2560, 1280), (32, 5, 8, 1280, 1280), (32, 5, 8, 2560, 1280),
          (32, 20, 32, 1920, 640), (32, 20, 32, 640, 1280), (32, 20, 32, 640, 640), (32, 20, 32, 1280, 640),
          (32, 40, 64, 320, 320), (32, 40, 64, 640, 320), (32, 10, 16, 640, 1280)]
if len(sys.argv) > 1 and int(sys.argv[1]) >= 0:
    shapes = shapes[int(sys.argv[1]):int(sys.argv[1]) + 1]
for NF, H, W, C, N in shapes:
    x = torch.randn(NF * H * W, C, device=dev).to(torch.bfloat16)
    w = pack_conv3x3(torch.randn(N, C, 3, 3, device=dev) * (9 * C) ** -0.5)
    fl = 2.0 * NF * H * W * N * 9 * C
    lib.lavie_debug_set(1, 0)
    t0 = graph_time(lambda: ops.conv3x3(x, NF, H, W, w))
    lib.lavie_debug_set(5, 1)
    t_single = graph_time(lambda: ops.conv3x3(x, NF, H, W, w))
    lib.lavie_debug_set(5, 0)
    print(f"conv M={NF*H*W} C={C} N={N}: planner {t0*1e3:6.1f} us ({fl/t0/1e9:6.0f} TF/s), single-launch plan {t_single*1e3:6.1f} us", flush=True)
    if len(sys.argv) <= 2:
        continue
    res = []
    for bn in (128, 160, 192, 256, 320):
        for s in (1, 2, 3, 4, 5, 6, 8):
            lib.lavie_debug_set(1, s)
            try:
                t = graph_time(lambda: ops.conv3x3(x, NF, H, W, w, block_n=bn))
            except Exception as e:
                continue
            res.append((t, bn, s))
    lib.lavie_debug_set(1, 0)
    res.sort()
    print(f"conv M={NF*H*W} C={C} N={N}: planner {t0*1e3:6.1f} us ({fl/t0/1e9:6.0f} TF/s); best " +
          ", ".join(f"bn={bn} s={s}: {t*1e3:.1f}" for t, bn, s in res[:4]), flush=True)
