"""Micro-benchmark of the GEMM / conv kernel on the step's dominant shapes (optimisation tool)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
from lavie_b200.packing import pack_conv3x3

dev = "cuda"
lib = _lib.load()
dbg = getattr(lib, "lavie_debug_set", None)
if dbg is not None:
    dbg.argtypes = [ctypes.c_int, ctypes.c_int]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


CONVS = [(32, 40, 64, 320, 320), (32, 40, 64, 640, 320), (32, 20, 32, 640, 640), (32, 20, 32, 1280, 640),
         (32, 10, 16, 1280, 1280), (32, 5, 8, 1280, 1280), (32, 5, 8, 2560, 1280)]
GEMMS = [(81920, 320, 320, True), (81920, 1152, 320, False), (81920, 320, 1280, True), (20480, 640, 640, True),
         (20480, 1920, 640, False), (5120, 1280, 1280, True), (5120, 1280, 5120, True), (5120, 3840, 1280, False)]


def run(label):
    print(f"== {label}")
    for NF, H, W, C, N in CONVS:
        x = torch.randn(NF * H * W, C, device=dev).to(torch.bfloat16)
        w = pack_conv3x3(torch.randn(N, C, 3, 3, device=dev) * (9 * C) ** -0.5)
        b = torch.randn(N, device=dev)
        ms = timeit(lambda: ops.conv3x3(x, NF, H, W, w, bias=b))
        fl = 2.0 * NF * H * W * N * 9 * C
        print(f"conv M={NF*H*W:6d} N={N:5d} K={9*C:6d}: {ms*1e3:8.1f} us {fl/ms/1e9:8.1f} TF/s")
    for M, N, K, res in GEMMS:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
        b = torch.randn(N, device=dev)
        r = torch.randn(M, N, device=dev).to(torch.bfloat16) if res else None
        ms = timeit(lambda: ops.gemm(a, w, bias=b, residual=r))
        print(f"gemm M={M:6d} N={N:5d} K={K:6d} res={int(res)}: {ms*1e3:8.1f} us {2.0*M*N*K/ms/1e9:8.1f} TF/s")


if dbg is not None and len(sys.argv) > 1:
    for v in sys.argv[1:]:
        dbg(2, int(v))
        run(f"debug={v} (1 = no TMA, 2 = no MMA)")
else:
    run("default")
