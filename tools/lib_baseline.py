"""Library reference points (cuBLAS / cuDNN via stock PyTorch) for the step's dominant shapes.  NOT the product
path: only tells what the vendor libraries reach on the same problem sizes on this GPU."""
import torch, sys
import torch.nn.functional as F
dev = "cuda"

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for M, N, K in [(81920, 320, 2880), (81920, 320, 320), (81920, 1152, 320), (20480, 640, 5760), (5120, 1280, 11520),
                (1280, 1280, 11520), (5120, 3840, 1280), (8192, 8192, 8192)]:
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16); w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: a @ w.t())
    print(f"cublas gemm M={M:6d} N={N:5d} K={K:6d}: {ms*1e3:8.1f} us {2.0*M*N*K/ms/1e9:8.1f} TF/s")
for NF, H, W, C, N in [(32, 40, 64, 320, 320), (32, 20, 32, 640, 640), (32, 10, 16, 1280, 1280), (32, 5, 8, 1280, 1280)]:
    x = torch.randn(NF, C, H, W, device=dev, dtype=torch.bfloat16).to(memory_format=torch.channels_last)
    w = torch.randn(N, C, 3, 3, device=dev, dtype=torch.bfloat16).to(memory_format=torch.channels_last)
    ms = timeit(lambda: F.conv2d(x, w, padding=1))
    print(f"cudnn conv NHWC M={NF*H*W:6d} N={N:5d} K={9*C:6d}: {ms*1e3:8.1f} us {2.0*NF*H*W*N*9*C/ms/1e9:8.1f} TF/s")
q = torch.randn(32, 8, 2560, 40, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: F.scaled_dot_product_attention(q, q, q))
print(f"sdpa B=32 h=8 S=2560 d=40: {ms*1e3:8.1f} us {4.0*256*2560*2560*40/ms/1e9:8.1f} TF/s")
