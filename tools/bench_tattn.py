"""Temporal attention at the three resolutions of the full-size step, CUDA-graph timed."""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops
from lavie_b200.packing import rope_table, rel_pos_bias_table
dev = "cuda"
def graph_time(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B, F, heads = 2, 16, 8
freqs = 1.0 / (10000.0 ** (torch.arange(0, 32, 2).float() / 32))
rope = rope_table(freqs, F).to(dev)
bias = rel_pos_bias_table(torch.randn(32, heads), F).to(dev)
for HW, d, pitch in [(2560, 40, 48), (640, 80, 80), (160, 160, 160)]:
    # several buffers so that the data does not sit in L2
    bufs = [torch.randn(B * F * HW, 3 * heads * pitch, device=dev).to(torch.bfloat16) for _ in range(4)]
    k = [0]
    def run():
        k[0] = (k[0] + 1) % len(bufs)
        return ops.temporal_attention(bufs[k[0]], B, F, HW, heads, d, pitch, rope, bias)
    t = graph_time(run, 12)
    by = B * F * HW * heads * (3 * pitch + d) * 2
    print(f"tattn HW={HW} d={d}: {t*1e3:7.1f} us  {by/t/1e6:7.0f} GB/s")
