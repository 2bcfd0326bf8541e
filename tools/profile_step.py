"""Per-launch CUDA-event breakdown of one eager full-size step (bring-up / optimisation tool)."""
import sys, json, collections
import torch
sys.path.insert(0, ".")
from lavie_b200 import UNet3DConditionModel, ops
from lavie_b200.synthetic import synthetic_inputs, synthetic_state_dict

m = UNet3DConditionModel(use_cuda_graph=False)
m.load_state_dict(synthetic_state_dict())
m = m.to("cuda").eval()
sample, t, text = synthetic_inputs(2, 16, 40, 64)
s, e = sample.cuda(), text.cuda()
for _ in range(2):
    m(s, t, encoder_hidden_states=e)
ops.PROFILE = []
m(s, t, encoder_hidden_states=e)
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
agg = collections.OrderedDict()
for name, flops, nbytes, e0, e1, tag in prof:
    a = agg.setdefault((name, tag), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += flops; a[3] += nbytes
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.2f} ms over {len(prof)} launches")
for (name, tag), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = a[2] / (a[1] * 1e-3) / 1e12 if a[1] > 0 else 0
    gb = a[3] / (a[1] * 1e-3) / 1e9 if a[1] > 0 else 0
    print(f"{a[1]:8.3f} ms {100*a[1]/tot:5.1f}% x{a[0]:3d} {a[1]/a[0]*1e3:8.1f} us/launch {tf:7.1f} TF/s {gb:7.0f} GB/s  {name} {tag}")
