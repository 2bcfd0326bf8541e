import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
from lavie_b200.packing import pack_conv3x3
lib = _lib.load(); dev = "cuda"
NF, H, W, C, N = 32, 40, 64, 320, 320
x = torch.randn(NF * H * W, C, device=dev).to(torch.bfloat16)
w = pack_conv3x3(torch.randn(N, C, 3, 3, device=dev) * (9 * C) ** -0.5)
a = torch.randn(20480, 5120, device=dev).to(torch.bfloat16)
w2 = (torch.randn(1280, 5120, device=dev) * 0.01).to(torch.bfloat16)
lib.lavie_debug_set(2, int(sys.argv[1]) if len(sys.argv) > 1 else 2)
for _ in range(2):
    ops.conv3x3(x, NF, H, W, w, block_n=160)
    ops.gemm(a, w2, block_n=256)
torch.cuda.synchronize(); print("done")
