"""The caller of the denoiser: the classifier-free-guidance loop of ``VideoGenPipeline.__call__``
(base/pipelines/pipeline_videogen.py:664-689) with the three schedulers the reference wires up (predict.py:74-96:
DDIM, DDPM, EulerDiscrete), restated over the B200 module.

Per step: ``cat([latents]*2)`` -> ``scheduler.scale_model_input`` -> ``unet(...)`` -> ``u + g (t - u)`` ->
``scheduler.step``.  Guidance combine + scheduler update run as ONE small kernel: every epsilon-prediction update is
``a * latents + b * eps (+ c * noise)`` (``lavie_cfg_linear_step``; DDIM keeps its dedicated ``lavie_cfg_ddim_step``), the
coefficients are computed here in float64 per step.  EulerDiscrete's ``scale_model_input`` (x / sqrt(sigma^2 + 1)) is
folded into conv_in through ``unet(..., input_scale=...)``.  Scheduler constants follow the reference's sampling config
(base/configs/sample.yaml:23-25: linear betas 1e-4..2e-2) and diffusers 0.16.0 (DDIM mirror in the reference:
vsr/diffusion/scheduling_ddim.py:259-265, 345-394; DDPM / EulerDiscrete from the published 0.16.0 sources).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import numpy as np
import torch

from . import ops

F32 = torch.float32


def _alphas_cumprod(num_train_timesteps: int, beta_start: float, beta_end: float) -> torch.Tensor:
    betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    return torch.cumprod(1.0 - betas, dim=0)


class DDIMSchedule:
    """DDIMScheduler, eta = 0, steps_offset 1, set_alpha_to_one False, clip_sample False."""
    name = "ddim"
    needs_noise = False

    def __init__(self, num_inference_steps: int = 50, num_train_timesteps: int = 1000, beta_start: float = 1e-4,
                 beta_end: float = 2e-2, steps_offset: int = 1):
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end)
        self.ratio = num_train_timesteps // num_inference_steps
        self.timesteps: List[int] = ((torch.arange(num_inference_steps) * self.ratio).flip(0) + steps_offset).tolist()
        self.init_noise_sigma = 1.0

    def alphas(self, t: int):
        prev = t - self.ratio
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.alphas_cumprod[0])
        return a_t, a_prev

    def input_scale(self, i: int) -> float:
        return 1.0

    def coefficients(self, i: int):
        """(a, b, c_noise) of latents' = a latents + b eps + c noise  (x0 = (x - sqrt(1-a_t) eps) / sqrt(a_t))."""
        a_t, a_p = self.alphas(self.timesteps[i])
        return (math.sqrt(a_p / a_t), math.sqrt(1 - a_p) - math.sqrt(a_p) * math.sqrt(1 - a_t) / math.sqrt(a_t), 0.0)


class DDPMSchedule:
    """DDPMScheduler (diffusers 0.16.0): ancestral sampling, variance_type fixed_small, clip_sample False."""
    name = "ddpm"
    needs_noise = True

    def __init__(self, num_inference_steps: int = 50, num_train_timesteps: int = 1000, beta_start: float = 1e-4,
                 beta_end: float = 2e-2):
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end).double()
        self.ratio = num_train_timesteps // num_inference_steps
        self.timesteps: List[int] = (torch.arange(num_inference_steps) * self.ratio).flip(0).tolist()
        self.init_noise_sigma = 1.0

    def input_scale(self, i: int) -> float:
        return 1.0

    def coefficients(self, i: int):
        t = self.timesteps[i]
        prev = t - self.ratio
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else 1.0
        b_t, b_p = 1.0 - a_t, 1.0 - a_p
        cur_a = a_t / a_p
        cur_b = 1.0 - cur_a
        c_x0 = math.sqrt(a_p) * cur_b / b_t            # weight of the predicted x0
        c_x = math.sqrt(cur_a) * b_p / b_t             # weight of the current sample
        sigma = math.sqrt(max(b_p / b_t * cur_b, 1e-20)) if t > 0 else 0.0
        # x0 = (x - sqrt(b_t) eps) / sqrt(a_t)
        return (c_x0 / math.sqrt(a_t) + c_x, -c_x0 * math.sqrt(b_t) / math.sqrt(a_t), sigma)


class EulerDiscreteSchedule:
    """EulerDiscreteScheduler (diffusers 0.16.0), linear sigma interpolation, s_churn = 0 (deterministic)."""
    name = "eulerdiscrete"
    needs_noise = False

    def __init__(self, num_inference_steps: int = 50, num_train_timesteps: int = 1000, beta_start: float = 1e-4,
                 beta_end: float = 2e-2):
        acp = _alphas_cumprod(num_train_timesteps, beta_start, beta_end).double().numpy()
        ts = np.linspace(0, num_train_timesteps - 1, num_inference_steps, dtype=float)[::-1].copy()
        sig = ((1 - acp) / acp) ** 0.5
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32).astype(np.float64)
        self.timesteps: List[float] = ts.tolist()        # fractional timesteps go to the time embedding as they are
        self.init_noise_sigma = float(self.sigmas.max())

    def input_scale(self, i: int) -> float:
        return float(1.0 / math.sqrt(self.sigmas[i] ** 2 + 1.0))

    def coefficients(self, i: int):
        # x0 = x - sigma eps; derivative = (x - x0) / sigma = eps; x' = x + eps (sigma_next - sigma)
        return (1.0, float(self.sigmas[i + 1] - self.sigmas[i]), 0.0)


def make_schedule(name: str, num_inference_steps: int = 50, **kw):
    """'ddim' | 'ddpm' | 'eulerdiscrete' -- the keys of predict.py:74-96's scheduler table."""
    table = {"ddim": DDIMSchedule, "ddpm": DDPMSchedule, "eulerdiscrete": EulerDiscreteSchedule}
    if name not in table:
        raise ValueError(f"unknown sample_method {name!r}; choose from {sorted(table)}")
    return table[name](num_inference_steps, **kw)


class CFGDenoiser:
    """One CFG denoising step on device-resident state.  ``text`` = cat([negative, positive]) [2, L, ctx] exactly as
    ``_encode_prompt`` returns it (pipeline_videogen.py:418)."""

    def __init__(self, unet, guidance_scale: float = 7.5, schedule=None):
        self.unet = unet
        self.guidance_scale = float(guidance_scale)
        self.schedule = schedule or DDIMSchedule()
        self._model_in = None
        self._index = {t: i for i, t in enumerate(self.schedule.timesteps)}

    def step(self, latents: torch.Tensor, t, text: torch.Tensor, noise: Optional[torch.Tensor] = None,
             generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """latents fp32 [1,C,F,H,W] on the device -> next latents (new tensor).  ``noise``: the variance noise of a
        stochastic scheduler (DDPM); drawn from ``generator`` on the latents' device when omitted, like diffusers'
        ``randn_tensor``."""
        if self._model_in is None or self._model_in.shape[1:] != latents.shape[1:]:
            self._model_in = torch.empty((2,) + tuple(latents.shape[1:]), dtype=F32, device=latents.device)
        self._model_in[0].copy_(latents[0])          # torch.cat([latents] * 2), pipeline_videogen.py:666
        self._model_in[1].copy_(latents[0])
        sched = self.schedule
        i = self._index[t]
        noise_pred = self.unet(self._model_in, t, encoder_hidden_states=text, input_scale=sched.input_scale(i)).sample
        if isinstance(sched, DDIMSchedule):
            a_t, a_prev = sched.alphas(t)
            return ops.cfg_ddim_step(noise_pred[0:1], noise_pred[1:2], self.guidance_scale, a_t, a_prev, latents)
        a, b, c = sched.coefficients(i)
        if c != 0.0 and noise is None:
            noise = torch.randn(latents.shape, generator=generator, device=latents.device, dtype=F32)
        return ops.cfg_linear_step(noise_pred[0:1], noise_pred[1:2], self.guidance_scale, a, b, latents,
                                   noise=noise if c != 0.0 else None, c_noise=c)

    def loop(self, latents: torch.Tensor, text: torch.Tensor,
             callback: Optional[Callable[[int, int, torch.Tensor], None]] = None,
             noises: Optional[list] = None, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        latents = latents.to(device=self.unet.device, dtype=F32).contiguous() * self.schedule.init_noise_sigma
        text = text.to(self.unet.device)
        for i, t in enumerate(self.schedule.timesteps):
            noise = None if noises is None else noises[i].to(device=latents.device, dtype=F32).contiguous()
            latents = self.step(latents, t, text, noise=noise, generator=generator)
            if callback is not None:
                callback(i, t, latents)
        return latents


class InterpolationSampler:
    """The caller of the frame-interpolation denoiser (interpolation/sample.py:138-166 ``auto_inpainting_copy_no_mask``
    -> interpolation/diffusion/gaussian_diffusion.py ``ddim_sample_loop`` with eta 0, clip_denoised False, use_concat):
    respaced DDIM (respace.py:9-63: ``count`` timesteps with a fractional stride) over cat([x_t, key-frame latents], 1)
    with ``forward_with_cfg`` (cfg_scale 4.0, text = [prompt, negative prompt]).  The reference's ``p_mean_variance``
    runs the UNet TWICE per step by accident (a failed ``.sample`` attribute access re-invokes the model,
    gaussian_diffusion.py:284-291); the result is identical, so one evaluation is made here."""

    def __init__(self, unet, cfg_scale: float = 4.0, num_steps: int = 50, num_train: int = 1000):
        self.unet = unet
        self.cfg_scale = float(cfg_scale)
        betas = torch.linspace(1e-4, 2e-2, num_train, dtype=torch.float64)
        acp = torch.cumprod(1.0 - betas, dim=0)
        stride = 1 if num_steps <= 1 else (num_train - 1) / (num_steps - 1)
        self.timesteps = [round(i * stride) for i in range(num_steps)]            # ascending; sampled in reverse
        self.ab = acp[self.timesteps]
        self.ab_prev = torch.cat([torch.ones(1, dtype=torch.float64), self.ab[:-1]])

    def coefficients(self, i: int):
        """x0 = x / sqrt(ab) - sqrt(1/ab - 1) eps;  x' = sqrt(ab_prev) x0 + sqrt(1 - ab_prev) eps."""
        ab, abp = float(self.ab[i]), float(self.ab_prev[i])
        return (math.sqrt(abp) / math.sqrt(ab), math.sqrt(1.0 - abp) - math.sqrt(abp) * math.sqrt(1.0 / ab - 1.0))

    @torch.no_grad()
    def loop(self, z: torch.Tensor, x_start: torch.Tensor, text: torch.Tensor) -> torch.Tensor:
        """z, x_start: [2, 4, F, H, W] (both CFG halves, as the reference passes them); text [2, L, ctx]."""
        dev = self.unet.device
        x = z.to(device=dev, dtype=F32).contiguous()
        cond = x_start.to(device=dev, dtype=F32).contiguous()
        text = text.to(dev)
        for i in reversed(range(len(self.timesteps))):
            # forward_with_cfg (text ordered [cond, uncond]) on cat([x_t, x_start], 1); the guided eps comes back for
            # both halves and the DDIM update (eta 0) is applied to the whole batch, as gaussian_diffusion.py:587-642 does
            g = self.unet.forward_with_cfg(torch.cat([x, cond], dim=1), self.timesteps[i], encoder_hidden_states=text,
                                           cfg_scale=self.cfg_scale).contiguous()
            a, b = self.coefficients(i)
            x = ops.cfg_linear_step(g, g, 0.0, a, b, x)
        return x
