"""Is the full-size step time-bound or ENERGY-bound (1000 W software power cap)?  Replays the captured step with an
extra low-power spin kernel of X us appended to every step: if the step time grows by less than X, the GPU was converting
the idle time into higher clocks for the rest of the step, i.e. removing low-power kernels cannot shorten the step."""
import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import UNet3DConditionModel
from lavie_b200.synthetic import synthetic_inputs, synthetic_state_dict
m = UNet3DConditionModel()
m.load_state_dict(synthetic_state_dict())
m = m.to("cuda").eval()
sample, t, text = synthetic_inputs(2, 16, 40, 64)
s, e = sample.cuda(), text.cuda()
for _ in range(5):
    m(s, t, encoder_hidden_states=e)
torch.cuda.synchronize()
cycles_per_us = 1965                       # _sleep counts SM clocks; close enough for a relative probe
for spin_us in (0, 500, 1000, 2000, 0):
    for _ in range(5):
        m(s, t, encoder_hidden_states=e)
        if spin_us: torch.cuda._sleep(spin_us * cycles_per_us)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40
    e0.record()
    for _ in range(n):
        m(s, t, encoder_hidden_states=e)
        if spin_us: torch.cuda._sleep(spin_us * cycles_per_us)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"spin {spin_us:5d} us per step: {ms:.3f} ms per step  (step minus spin: {ms - spin_us / 1000:.3f} ms)")
