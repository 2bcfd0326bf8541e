"""CPU oracle of the three schedulers the reference wires to the base pipeline -- TEST INFRASTRUCTURE, NOT PRODUCT.

predict.py:74-96 builds DDIMScheduler, DDPMScheduler and EulerDiscreteScheduler (diffusers==0.16.0, environment.yml:14)
from the SD-1.4 ``scheduler/scheduler_config.json`` with the betas of base/configs/sample.yaml:23-25
(linear 1e-4 .. 2e-2); VideoGenPipeline calls ``scale_model_input`` and ``step`` (pipeline_videogen.py:666-683).

* DDIM: restated from the reference's vendored copy, vsr/diffusion/scheduling_ddim.py:345-394, and PINNED against it
  (tests/golden/make_golden_ddim.py -> tests/test_oracle.py::test_ddim_step_matches_reference_scheduler).
* DDPM and EulerDiscrete: diffusers 0.16.0 is a third-party dependency that is neither vendored nor installed here,
  so these two are restated from the published 0.16.0 sources (scheduling_ddpm.py ``step`` / ``_get_variance``,
  scheduling_euler_discrete.py ``set_timesteps`` / ``scale_model_input`` / ``step``) -- **parity unpinned**; the tests
  anchor them on algebraic properties instead (exact recovery of x0 when eps is the true noise, the deterministic
  limit, sigma bookkeeping).  SD-1.4 scheduler config values used: steps_offset 1 (DDIM only reads it), clip_sample
  False, set_alpha_to_one False, prediction_type epsilon, variance_type fixed_small.

All arithmetic in float64; callers cast.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch


def alphas_cumprod(num_train: int = 1000, beta_start: float = 1e-4, beta_end: float = 2e-2) -> torch.Tensor:
    betas = torch.linspace(beta_start, beta_end, num_train, dtype=torch.float32)     # schedulers build betas in fp32
    return torch.cumprod(1.0 - betas, dim=0).double()


class DDIM:
    """eta = 0 (vsr/diffusion/scheduling_ddim.py:345-394); stock 0.16 timesteps (arange(n) * ratio)[::-1] + offset."""

    def __init__(self, steps: int = 50, num_train: int = 1000, steps_offset: int = 1):
        self.acp = alphas_cumprod(num_train)
        self.ratio = num_train // steps
        self.timesteps = ((np.arange(steps) * self.ratio).round()[::-1] + steps_offset).astype(np.int64).tolist()
        self.init_noise_sigma = 1.0

    def scale(self, i: int) -> float:
        return 1.0

    def step(self, eps, i, sample, noise=None):
        t = self.timesteps[i]
        prev = t - self.ratio
        a_t = self.acp[t]
        a_p = self.acp[prev] if prev >= 0 else self.acp[0]
        x0 = (sample.double() - (1 - a_t).sqrt() * eps.double()) / a_t.sqrt()
        return a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps.double()


class DDPM:
    """diffusers 0.16.0 DDPMScheduler.step, variance_type fixed_small, clip_sample False [3p]."""

    def __init__(self, steps: int = 50, num_train: int = 1000):
        self.acp = alphas_cumprod(num_train)
        self.ratio = num_train // steps
        self.timesteps = (np.arange(steps) * self.ratio).round()[::-1].astype(np.int64).tolist()
        self.init_noise_sigma = 1.0

    def scale(self, i: int) -> float:
        return 1.0

    def coefficients(self, i: int):
        t = self.timesteps[i]
        prev = t - self.ratio
        a_t = self.acp[t]
        a_p = self.acp[prev] if prev >= 0 else torch.tensor(1.0, dtype=torch.float64)
        b_t, b_p = 1 - a_t, 1 - a_p
        cur_a = a_t / a_p
        cur_b = 1 - cur_a
        c_x0 = a_p.sqrt() * cur_b / b_t
        c_x = cur_a.sqrt() * b_p / b_t
        var = (b_p / b_t * cur_b).clamp(min=1e-20)
        sigma = var.sqrt() if t > 0 else torch.tensor(0.0, dtype=torch.float64)
        return a_t, c_x0, c_x, sigma

    def step(self, eps, i, sample, noise=None):
        a_t, c_x0, c_x, sigma = self.coefficients(i)
        x0 = (sample.double() - (1 - a_t).sqrt() * eps.double()) / a_t.sqrt()
        out = c_x0 * x0 + c_x * sample.double()
        if float(sigma) > 0:
            out = out + sigma * noise.double()
        return out


class EulerDiscrete:
    """diffusers 0.16.0 EulerDiscreteScheduler, interpolation_type linear, s_churn 0 [3p]."""

    def __init__(self, steps: int = 50, num_train: int = 1000):
        acp = alphas_cumprod(num_train).numpy()
        self.timesteps = np.linspace(0, num_train - 1, steps, dtype=float)[::-1].copy().tolist()
        sig = ((1 - acp) / acp) ** 0.5
        sig = np.interp(np.array(self.timesteps), np.arange(0, len(sig)), sig)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32).astype(np.float64)   # stored fp32 by diffusers
        self.init_noise_sigma = float(self.sigmas.max())

    def scale(self, i: int) -> float:
        return float(1.0 / (self.sigmas[i] ** 2 + 1.0) ** 0.5)

    def step(self, eps, i, sample, noise=None):
        sigma = self.sigmas[i]
        x0 = sample.double() - sigma * eps.double()
        derivative = (sample.double() - x0) / sigma
        return sample.double() + derivative * (self.sigmas[i + 1] - sigma)


def make(name: str, steps: int = 50):
    return {"ddim": DDIM, "ddpm": DDPM, "eulerdiscrete": EulerDiscrete}[name](steps)


@torch.no_grad()
def cfg_loop(forward: Callable, latents: torch.Tensor, text_uncond_cond: torch.Tensor, scheduler, guidance: float = 7.5,
             noises: Optional[list] = None) -> torch.Tensor:
    """VideoGenPipeline.__call__'s loop (pipeline_videogen.py:664-689): ``forward(model_in, t, text) -> eps [2, ...]``;
    ``noises[i]`` = the DDPM variance noise of step i (same tensor the product is given)."""
    latents = latents.float() * scheduler.init_noise_sigma
    for i, t in enumerate(scheduler.timesteps):
        model_in = torch.cat([latents] * 2) * scheduler.scale(i)
        eps = forward(model_in.float(), t, text_uncond_cond)
        u, c = eps.chunk(2)
        eps = u + guidance * (c - u)
        latents = scheduler.step(eps, i, latents, None if noises is None else noises[i]).float()
    return latents
