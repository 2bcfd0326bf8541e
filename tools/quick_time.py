"""Bring-up timing of one full-size denoiser step (not the bench; see bench.py)."""
import sys, time
import torch
sys.path.insert(0, ".")
from lavie_b200 import UNet3DConditionModel
from lavie_b200.synthetic import synthetic_inputs, synthetic_state_dict

m = UNet3DConditionModel()
m.load_state_dict(synthetic_state_dict())
m = m.to("cuda").eval()
sample, t, text = synthetic_inputs(2, 16, 40, 64)
s, e = sample.cuda(), text.cuda()
for graph in (False, True):
    m.use_cuda_graph = graph
    for _ in range(3):
        out = m(s, t, encoder_hidden_states=e).sample
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    t0 = time.time()
    ev0.record()
    for _ in range(n):
        out = m(s, t, encoder_hidden_states=e).sample
    ev1.record()
    torch.cuda.synchronize()
    print(f"graph={graph}: {ev0.elapsed_time(ev1)/n:.2f} ms/step (wall {1e3*(time.time()-t0)/n:.2f} ms), out std {float(out.std()):.4f}")
print("mem GB", torch.cuda.max_memory_allocated()/2**30)
