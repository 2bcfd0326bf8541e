import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
lib = _lib.load(); dev = "cuda"
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for dbg in (0, 16):
    lib.lavie_debug_set(2, dbg)
    for M, N, K in [(256, 160, 64), (18944, 160, 64), (18944, 160, 320), (18944*2, 160, 320), (18944*4, 160, 320), (18944*4, 160, 64), (18944*4, 160, 1280), (18944*4, 320, 320), (18944 * 4, 256, 320), (18944 * 4, 128, 320)]:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
        t = timeit(lambda: ops.gemm(a, w))
        tiles = ((M + 255) // 256) * ((N + 159) // 160)
        print(f"debug={dbg:2d} M={M:6d} N={N:4d} K={K:5d}: {t*1e3:7.1f} us  pair-tiles={tiles} per-pair={tiles/74:.2f}")
