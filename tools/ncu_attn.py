import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops
dev = "cuda"
batch, S, heads, d, pitch = 32, 2560, 8, 40, 48
hp = heads * pitch
qkv = torch.zeros(batch * S, 3, heads, pitch, device=dev)
qkv[..., :d] = torch.randn(batch * S, 3, heads, d, device=dev)
qkv = qkv.reshape(batch * S, 3 * hp).to(torch.bfloat16)
for _ in range(2):
    o = ops.attention(qkv[:, :hp], qkv[:, hp:2 * hp], qkv[:, 2 * hp:], batch, heads, S, S, d, pitch)
torch.cuda.synchronize(); print("done")
