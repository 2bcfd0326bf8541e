"""The caller of the denoiser: the classifier-free-guidance loop of ``VideoGenPipeline.__call__``
(base/pipelines/pipeline_videogen.py:664-689) with a DDIM scheduler, restated over the B200 module.

Per step: ``cat([latents]*2)`` -> ``unet(...)`` -> ``u + g (t - u)`` -> ``scheduler.step``.  The guidance combine and the
DDIM (eta = 0) update run as one small kernel (``lavie_cfg_ddim_step``).  Scheduler constants follow the reference's
sampling config (base/configs/sample.yaml:23-25: linear betas 1e-4..2e-2) and diffusers-0.16 DDIM timesteps
``(arange(n) * (1000 // n))[::-1] + 1`` (mirror: vsr/diffusion/scheduling_ddim.py:259-265, 345-394).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from . import ops

F32 = torch.float32


class DDIMSchedule:
    def __init__(self, num_inference_steps: int = 50, num_train_timesteps: int = 1000, beta_start: float = 1e-4,
                 beta_end: float = 2e-2, steps_offset: int = 1):
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.ratio = num_train_timesteps // num_inference_steps
        self.timesteps: List[int] = ((torch.arange(num_inference_steps) * self.ratio).flip(0) + steps_offset).tolist()
        self.init_noise_sigma = 1.0

    def alphas(self, t: int):
        prev = t - self.ratio
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.alphas_cumprod[0])
        return a_t, a_prev


class CFGDenoiser:
    """One CFG denoising step on device-resident state.  ``text`` = cat([negative, positive]) [2, L, ctx] exactly as
    ``_encode_prompt`` returns it (pipeline_videogen.py:418)."""

    def __init__(self, unet, guidance_scale: float = 7.5, schedule: Optional[DDIMSchedule] = None):
        self.unet = unet
        self.guidance_scale = float(guidance_scale)
        self.schedule = schedule or DDIMSchedule()
        self._model_in = None

    def step(self, latents: torch.Tensor, t: int, text: torch.Tensor) -> torch.Tensor:
        """latents fp32 [1,C,F,H,W] on the device -> next latents (new tensor)."""
        if self._model_in is None or self._model_in.shape[1:] != latents.shape[1:]:
            self._model_in = torch.empty((2,) + tuple(latents.shape[1:]), dtype=F32, device=latents.device)
        self._model_in[0].copy_(latents[0])          # torch.cat([latents] * 2), pipeline_videogen.py:666
        self._model_in[1].copy_(latents[0])
        noise = self.unet(self._model_in, t, encoder_hidden_states=text).sample
        a_t, a_prev = self.schedule.alphas(t)
        return ops.cfg_ddim_step(noise[0:1], noise[1:2], self.guidance_scale, a_t, a_prev, latents)

    def loop(self, latents: torch.Tensor, text: torch.Tensor,
             callback: Optional[Callable[[int, int, torch.Tensor], None]] = None) -> torch.Tensor:
        latents = latents.to(device=self.unet.device, dtype=F32).contiguous() * self.schedule.init_noise_sigma
        text = text.to(self.unet.device)
        for i, t in enumerate(self.schedule.timesteps):
            latents = self.step(latents, t, text)
            if callback is not None:
                callback(i, t, latents)
        return latents
