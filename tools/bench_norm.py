import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
lib = _lib.load()
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
for target in (148, 222, 296, 444):
  lib.lavie_debug_set(6, target)
  print('GN target CTAs', target)
  # rotate over several buffers so the data does not sit in L2 (126 MB)
  for rows, C, samples in [(81920, 320, 2), (81920, 320, 32), (20480, 640, 2), (5120, 1280, 2), (81920, 640, 2)]:
      xs = [torch.randn(rows, C, device=dev).to(torch.bfloat16) for _ in range(6)]
      g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
      k = [0]
      def nxt():
          k[0] = (k[0] + 1) % len(xs); return xs[k[0]]
      t_ln = graph_time(lambda: ops.layernorm(nxt(), g, b), 12)
      ss = ops.groupnorm_scale_shift(xs[0], samples, rows // samples, g, b, 1e-5)
      t_ap = graph_time(lambda: ops.groupnorm_apply(nxt(), ss, samples, rows // samples, True), 12)
      t_st = graph_time(lambda: ops.groupnorm_scale_shift(nxt(), samples, rows // samples, g, b, 1e-5), 12)
      by = rows * C * 2
      print(f"rows={rows} C={C} samples={samples}: LN {t_ln*1e3:6.1f} us {2*by/t_ln/1e9:6.0f} GB/s | GN apply {t_ap*1e3:6.1f} us {2*by/t_ap/1e9:6.0f} GB/s | GN stats+finalize {t_st*1e3:6.1f} us {by/t_st/1e9:6.0f} GB/s")
