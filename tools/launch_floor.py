"""Fixed cost per launch inside a CUDA graph: tiny problems, 100 dependent launches, PDL on / off."""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
from lavie_b200.packing import pack_conv3x3
lib = _lib.load(); dev = "cuda"
def graph_time(fn, n=100):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
x = torch.randn(256, 320, device=dev).to(torch.bfloat16)
w = (torch.randn(320, 320, device=dev) * 0.05).to(torch.bfloat16)
b = torch.randn(320, device=dev)
gamma = torch.ones(320, device=dev); beta = torch.zeros(320, device=dev)
xc = torch.randn(2 * 8 * 16, 64, device=dev).to(torch.bfloat16)
wc = pack_conv3x3(torch.randn(64, 64, 3, 3, device=dev) * 0.05)
ss = ops.groupnorm_scale_shift(x, 1, 256, gamma, beta, 1e-5)
heads, d, pitch = 8, 40, 48
q = torch.zeros(128, heads * pitch, device=dev).to(torch.bfloat16)
cases = {
    "layernorm 256x320": lambda: ops.layernorm(x, gamma, beta),
    "gn scale_shift 256x320": lambda: ops.groupnorm_scale_shift(x, 1, 256, gamma, beta, 1e-5),
    "gn apply 256x320": lambda: ops.groupnorm_apply(x, ss, 1, 256, True),
    "gemm 256x320x320 +bias+res": lambda: ops.gemm(x, w, bias=b, residual=x),
    "conv3x3 2x8x16 64->64": lambda: ops.conv3x3(xc, 2, 8, 16, wc),
    "attention 1x8 heads 128x128 d40": lambda: ops.attention(q, q, q, 1, heads, 128, 128, d, pitch, 1),
}
for pdl in (1, 0):
    lib.lavie_debug_set(3, pdl)
    print(f"PDL {'on' if pdl else 'off'}: " + " | ".join(f"{k}: {graph_time(fn):.2f} us" for k, fn in cases.items()))
lib.lavie_debug_set(3, 1)
