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
NF, H, W, C, N = 32, 40, 64, 320, 320
x = torch.randn(NF * H * W, C, device=dev).to(torch.bfloat16)
w = pack_conv3x3(torch.randn(N, C, 3, 3, device=dev) * (9 * C) ** -0.5)
col = ops.im2col3x3(x, NF, H, W, 1)            # [81920, 2880]
wide = torch.randn(81920, 2880 * 2, device=dev).to(torch.bfloat16)[:, :2880]   # pitch 11520 B
for dbg in (0, 2, 1):
    lib.lavie_debug_set(2, dbg)
    for bn in (160, 256):
        t_conv = graph_time(lambda: ops.conv3x3(x, NF, H, W, w, block_n=bn))
        t_gemm = graph_time(lambda: ops.gemm(col, w, block_n=bn))
        t_wide = graph_time(lambda: ops.gemm(wide, w, block_n=bn))
        print(f"debug={dbg} bn={bn}: conv(4D, 640B pitch) {t_conv*1e3:6.1f} us | gemm on im2col (2D, 5760B pitch) {t_gemm*1e3:6.1f} us | gemm 2D 11520B pitch {t_wide*1e3:6.1f} us")
lib.lavie_debug_set(2, 0)
