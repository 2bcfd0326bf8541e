import ctypes, sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
from lavie_b200.packing import pack_conv3x3
lib = _lib.load()
dev = "cuda"
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
NF, H, W, C, N = 32, 40, 64, 320, 320
x = torch.randn(NF * H * W, C, device=dev).to(torch.bfloat16)
w = pack_conv3x3(torch.randn(N, C, 3, 3, device=dev) * (9 * C) ** -0.5)
a = torch.randn(20480, 5120, device=dev).to(torch.bfloat16)
w2 = (torch.randn(1280, 5120, device=dev) * 0.01).to(torch.bfloat16)
for dbg, rot in ((2, 7), (6, 7), (10, 7), (14, 7)):
    lib.lavie_debug_set(2, dbg)
    lib.lavie_debug_set(0, rot)
    print('k_rot', rot)
    for bn in (160, 256):
        ms = timeit(lambda: ops.conv3x3(x, NF, H, W, w, block_n=bn))
        ms2 = timeit(lambda: ops.gemm(a, w2, block_n=bn))
        print(f"debug={dbg} bn={bn:3d}: conv L0 {ms*1e3:7.1f} us {2.0*NF*H*W*N*9*C/ms/1e9:7.1f} TF/s | gemm 20480x1280x5120 {ms2*1e3:7.1f} us {2.0*20480*1280*5120/ms2/1e9:7.1f} TF/s")
