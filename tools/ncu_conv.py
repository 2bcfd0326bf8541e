"""Small driver for ncu captures: a few launches of the dominant conv / GEMM shapes of the full-size step."""
import sys
import torch
sys.path.insert(0, ".")
from lavie_b200 import ops
from lavie_b200.packing import pack_conv3x3

torch.manual_seed(0)
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def conv(NF, H, W, C, N):
    x = torch.randn(NF * H * W, C, device=dev).to(torch.bfloat16)
    w = pack_conv3x3((torch.randn(N, C, 3, 3, device=dev) * (9 * C) ** -0.5))
    b = torch.randn(N, device=dev)
    for _ in range(3):
        ops.conv3x3(x, NF, H, W, w, bias=b, stats=True)     # as in the step: the epilogue also emits the GroupNorm statistics


def gemm(M, N, K, res=True):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).to(torch.bfloat16) if res else None
    for _ in range(3):
        ops.gemm(a, w, bias=b, residual=r)


if which in ("all", "conv"):
    conv(32, 40, 64, 320, 320)      # L0 resnet conv: M=81920 N=320 K=2880
    conv(32, 20, 32, 640, 640)      # L1
    conv(32, 5, 8, 1280, 1280)      # L3: M=1280 (few tiles)
if which in ("all", "gemm"):
    gemm(81920, 320, 320)           # to_out + residual at L0 (epilogue heavy)
    gemm(5120, 1280, 5120)          # ff2 at L2
torch.cuda.synchronize()
print("done")
