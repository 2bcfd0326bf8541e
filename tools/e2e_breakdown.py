"""Where does the host-driven (e2e) step lose time against back-to-back replays?"""
import sys, time
import torch
sys.path.insert(0, ".")
from lavie_b200 import UNet3DConditionModel, ops
from lavie_b200.pipeline import CFGDenoiser, DDIMSchedule
from lavie_b200.synthetic import synthetic_inputs, synthetic_state_dict

m = UNet3DConditionModel(); m.load_state_dict(synthetic_state_dict()); m = m.to("cuda").eval()
sample, t, text = synthetic_inputs(2, 16, 40, 64)
sched = DDIMSchedule(50)
den = CFGDenoiser(m, 7.5, sched)
lat = sample[:1].cuda().float().contiguous()
txt = text.cuda()
for _ in range(3):
    x = den.step(lat, 500, txt)
torch.cuda.synchronize()
def wall(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def wall_sync_each(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print(f"device-resident step, back-to-back : {wall(lambda: den.step(lat, 500, txt)):.3f} ms")
print(f"device-resident step, sync each    : {wall_sync_each(lambda: den.step(lat, 500, txt)):.3f} ms")
g = list(m._graphs.values())[0]
print(f"bare graph replay, back-to-back    : {wall(lambda: g['graph'].replay()):.3f} ms")
print(f"bare graph replay, sync each       : {wall_sync_each(lambda: g['graph'].replay()):.3f} ms")
t0 = time.perf_counter()
for _ in range(10): g['graph'].replay()
host = (time.perf_counter() - t0) / 10 * 1e3
torch.cuda.synchronize()
print(f"host time inside replay() call     : {host:.3f} ms")
lat_host = lat.cpu().pin_memory(); txt_host = text.contiguous().pin_memory(); out_host = torch.empty_like(lat_host).pin_memory()
def e2e():
    lat_d = lat_host.to("cuda", non_blocking=True)
    new = den.step(lat_d, 500, txt_host)
    out_host.copy_(new, non_blocking=True)
    torch.cuda.synchronize()
    lat_host.copy_(out_host)
print(f"e2e step (host buffers)            : {wall(e2e):.3f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): e2e()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
