"""Per-shape breakdown (CUDA events, GPU kept busy while the host enqueues) of the per-rank workload of the sharded runs on
ONE GPU, without collectives: python tools/profile_shard_tags.py [B F] -- e.g. 1 4 = the 8-GPU shard, 1 8 = the 4-GPU shard."""
import sys, collections
import torch
sys.path.insert(0, ".")
from lavie_b200 import UNet3DConditionModel, ops
from lavie_b200.synthetic import synthetic_inputs, synthetic_state_dict
B, F = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1, 4)
m = UNet3DConditionModel(use_cuda_graph=False)
m.load_state_dict(synthetic_state_dict())
m = m.to("cuda").eval()
sample, t, text = synthetic_inputs(B, F, 40, 64)
s, e = sample.cuda(), text.cuda()
for _ in range(2):
    m(s, t, encoder_hidden_states=e)
ops.PROFILE = []
torch.cuda._sleep(150_000_000)
m(s, t, encoder_hidden_states=e)
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
agg = collections.OrderedDict()
for name, flops, nbytes, e0, e1, tag in prof:
    a = agg.setdefault((name, tag), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += flops; a[3] += nbytes
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.2f} ms over {len(prof)} launches (B={B}, F={F})")
for (name, tag), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{a[1]:8.3f} ms {100*a[1]/tot:5.1f}% x{a[0]:3d} {a[1]/a[0]*1e3:8.1f} us/launch {a[2]/(a[1]*1e-3)/1e12 if a[1] else 0:7.1f} TF/s "
          f"{a[3]/(a[1]*1e-3)/1e9 if a[1] else 0:7.0f} GB/s  {name} {tag}")
m.use_cuda_graph = True
for _ in range(3): m(s, t, encoder_hidden_states=e)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(20): m(s, t, encoder_hidden_states=e)
ev1.record(); torch.cuda.synchronize()
print(f"graph replay: {ev0.elapsed_time(ev1)/20:.3f} ms/forward")
