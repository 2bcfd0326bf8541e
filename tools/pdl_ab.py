"""A/B of programmatic dependent launch on one full-size denoiser step (CUDA-graph replay, device-timed)."""
import sys
import torch
sys.path.insert(0, ".")
from lavie_b200 import UNet3DConditionModel, _lib
from lavie_b200.synthetic import synthetic_inputs, synthetic_state_dict

sd = synthetic_state_dict()
sample, t, text = synthetic_inputs(2, 16, 40, 64)
s, e = sample.cuda(), text.cuda()
outs = {}
for pdl in (0, 1):
    _lib.load().lavie_debug_set(3, pdl)
    m = UNet3DConditionModel()
    m.load_state_dict(sd)
    m = m.to("cuda").eval()
    for _ in range(3):
        out = m(s, t, encoder_hidden_states=e).sample
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    ev0.record()
    for _ in range(n):
        out = m(s, t, encoder_hidden_states=e).sample
    ev1.record()
    torch.cuda.synchronize()
    outs[pdl] = out.float().clone()
    print(f"pdl={pdl}: {ev0.elapsed_time(ev1)/n:.3f} ms/step", flush=True)
    del m
print("max abs diff pdl on/off:", float((outs[0] - outs[1]).abs().max()))
