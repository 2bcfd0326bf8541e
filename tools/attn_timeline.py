"""Per-key-tile clock64 timeline of one mid-grid CTA of the attention kernel (thread 0 = issuing warp).
Usage: python tools/attn_timeline.py [Sq Sk d pitch]"""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
lib = _lib.load(); dev = "cuda"
Sq, Sk, d, pitch = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (2560, 2560, 40, 48)
batch, heads = 32, 8
def mk(rows):
    t = torch.zeros(rows, heads, pitch, device=dev); t[..., :d] = torch.randn(rows, heads, d, device=dev)
    return t.reshape(rows, heads * pitch).to(torch.bfloat16)
q, k, v = mk(batch * Sq), mk(batch * Sk), mk(batch * Sk)
for _ in range(3): ops.attention(q, k, v, batch, heads, Sq, Sk, d, pitch, 1)
buf = torch.zeros(8192, dtype=torch.int64, device=dev)
lib.lavie_debug_buffer(buf.data_ptr())
ops.attention(q, k, v, batch, heads, Sq, Sk, d, pitch, 1)
torch.cuda.synchronize()
lib.lavie_debug_buffer(None)
t = buf.cpu().view(-1, 8)
base = int(t[0, 0])
print("tile | softmax warp 0: top  S_ready  S_in_regs  exps_done  P_written | control warp: QK(j+1)_issued  PV(j)_issued  refills_issued | deltas: wait_S ld exps P")
n = (Sk + 63) // 64
for j in list(range(min(n, 6))) + list(range(max(6, n - 4), n)):
    r = [int(x) - base for x in t[j, :8]]
    dl = [r[i + 1] - r[i] for i in range(4)]
    print(f"{j:3d} | " + " ".join(f"{x:7d}" for x in r[:5]) + " | " + " ".join(f"{x:7d}" for x in r[5:8]) + " | " + " ".join(f"{x:5d}" for x in dl))
tot = int(t[n - 1, 4]) - base
print(f"per tile: {tot / n:.0f} cycles")
