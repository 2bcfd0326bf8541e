"""Pin the oracle's DDIM update to the reference's own scheduler code.

Run ONLY inside the build container:   python tests/golden/make_golden_ddim.py

Imports the UNMODIFIED ``DDIMScheduler`` the reference vendors at /root/reference/vsr/diffusion/scheduling_ddim.py
(stand-ins for the diffusers plumbing it imports live in tests/golden/shims/), configures it the way the base pipeline
does (SD-1.4 scheduler_config: steps_offset 1, clip_sample False, set_alpha_to_one False; betas from
base/configs/sample.yaml:23-25) and records ``step()`` (:292-414) on seeded tensors for several timesteps, including
the last one (prev_timestep < 0 -> final_alpha_cumprod).  The vendored ``set_timesteps`` (:267-290) is a MODIFIED
linspace variant; the base pipeline uses stock diffusers 0.16 (``(arange(n) * ratio)[::-1] + offset``, the commented
block :243-265), so only the per-step arithmetic is pinned here and ``timestep`` values are passed explicitly.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "shims"))

import torch  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_scheduling_ddim", "/root/reference/vsr/diffusion/scheduling_ddim.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)


def main():
    sch = mod.DDIMScheduler(num_train_timesteps=1000, beta_start=1e-4, beta_end=2e-2, beta_schedule="linear",
                            clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    sch.set_timesteps(50)
    g = torch.Generator().manual_seed(1234)
    cases = []
    for t in (981, 961, 501, 21, 1):
        eps = torch.randn(1, 4, 3, 8, 8, generator=g)
        x = torch.randn(1, 4, 3, 8, 8, generator=g)
        prev = sch.step(eps, t, x, eta=0.0).prev_sample
        cases.append({"t": t, "model_output": eps, "sample": x, "prev_sample": prev})
    torch.save({"cases": cases, "alphas_cumprod": sch.alphas_cumprod, "num_inference_steps": 50},
               os.path.join(HERE, "ddim_steps.pt"))
    print("ddim_steps.pt:", [c["t"] for c in cases])


if __name__ == "__main__":
    main()
