"""Per-tile timeline of CTA 0 of the GEMM (debug bit 512): epilogue warp 0 and the MMA issuer stamp clock64 into the
split-K workspace.  Usage: python tools/gemm_timeline.py M N K [res]"""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
lib = _lib.load(); dev = "cuda"
M, N, K = (int(x) for x in sys.argv[1:4])
res = len(sys.argv) > 4
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
r = torch.randn(M, N, device=dev).to(torch.bfloat16) if res else None
ws = ops._workspace(a.device) if hasattr(ops, "_workspace") else None
for _ in range(3): ops.gemm(a, w, residual=r)
torch.cuda.synchronize()
lib.lavie_debug_set(2, 512)
ops.gemm(a, w, residual=r)
torch.cuda.synchronize()
lib.lavie_debug_set(2, 0)
ws = ops._workspace(a.device)
t = ws.view(torch.int64)[:16384].cpu()
epi, mma = t[:8192].view(-1, 8), t[8192:].view(-1, 8)
base = int(mma[0, 0])
print("tile | epilogue: start  wait_full  full  ld0_done  chunks_done  arrived | mma: start  empty_ok  first_full  last_full  issued")
for i in range(12):
    e = [int(x) - base for x in epi[i, :6]]; m = [int(x) - base for x in mma[i, :5]]
    print(f"{i:3d} | " + " ".join(f"{x:7d}" for x in e) + " | " + " ".join(f"{x:7d}" for x in m))
