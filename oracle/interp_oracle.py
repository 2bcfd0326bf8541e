"""CPU oracle of the LaVie frame-INTERPOLATION denoiser -- groundwork for the next round.

TEST INFRASTRUCTURE ONLY, like oracle/unet3d_oracle.py: nothing under lavie_b200/ may import this module.  There is
no B200 path for this model yet; what exists is (a) the parameter table (lavie_b200.config.INTERP_CONFIG: 798 keys,
909 131 524 parameters, checked against the reference with strict loading) and (b) this fp32 restatement, pinned
against golden vectors the UNMODIFIED reference produced in the build container
(tests/golden/make_golden_interp.py -> tests/golden/interp_*.pt, tests/test_oracle_interp.py).

The interpolation UNet (interpolation/models/unet.py:320-475) is the base UNet3D with
  * 8 input channels (noisy latent ++ masked key-frame latent, unet.py:501-507),
  * a different transformer block (interpolation/models/attention.py:566-608):
        SparseCausal self-attention -> text cross-attention -> GEGLU feed-forward -> temporal attention
    (the base block runs self, cross, temporal, feed-forward),
  * SparseCausalAttention (attention.py:611-664): the queries of frame f attend to the keys / values of frame 0
    concatenated with those of frame max(f - 1, 0) -- 2 x H*W keys per query,
  * a plain temporal attention: softmax(q k^T / sqrt(d)) v over the frames of one pixel, no rotary embedding and no
    relative-position bias (the state_dict has neither key).
ResNet blocks, down / up sampling, time embedding and the 5-D GroupNorms are the base model's, so everything else is
shared with oracle/unet3d_oracle.py through its ``block`` hook.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import unet3d_oracle as B

SD = B.SD


def sparse_causal_attention(sd: SD, p: str, x: torch.Tensor, frames: int) -> torch.Tensor:
    """SparseCausalAttention.forward (interpolation/models/attention.py:611-664); x is [(B F), HW, C]."""
    bf, hw, c = x.shape
    b = bf // frames
    q = B._split_heads(F.linear(x, sd[f"{p}.to_q.weight"]))
    k = F.linear(x, sd[f"{p}.to_k.weight"]).reshape(b, frames, hw, c)
    v = F.linear(x, sd[f"{p}.to_v.weight"]).reshape(b, frames, hw, c)
    former = (torch.arange(frames) - 1).clamp_min(0)                       # frame f looks at frame f-1 (frame 0 at itself)
    first = torch.zeros(frames, dtype=torch.long)
    k = torch.cat([k[:, first], k[:, former]], dim=2).reshape(bf, 2 * hw, c)   # [first frame | former frame] keys
    v = torch.cat([v[:, first], v[:, former]], dim=2).reshape(bf, 2 * hw, c)
    k, v = B._split_heads(k), B._split_heads(v)
    probs = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5), dim=-1)
    o = B._merge_heads(torch.matmul(probs, v))
    return F.linear(o, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"])


def transformer_block(sd: SD, p: str, x: torch.Tensor, text: torch.Tensor, frames: int) -> torch.Tensor:
    """BasicTransformerBlock.forward of the interpolation model (attention.py:566-608); x is [(B F), HW, C]."""
    C = x.shape[-1]
    ln = lambda t, n: F.layer_norm(t, (C,), sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"], B.LN_EPS)
    x = sparse_causal_attention(sd, f"{p}.attn1", ln(x, "norm1"), frames) + x
    x = B.attention(sd, f"{p}.attn2", ln(x, "norm2"), text) + x
    x = B.geglu_ff(sd, f"{p}.ff", ln(x, "norm3")) + x
    bf, d, _ = x.shape
    b = bf // frames
    xt = x.reshape(b, frames, d, C).permute(0, 2, 1, 3).reshape(b * d, frames, C)      # (b f) d c -> (b d) f c
    xt = B.attention(sd, f"{p}.attn_temp", ln(xt, "norm_temp")) + xt                    # plain attention over frames
    return xt.reshape(b, d, frames, C).permute(0, 2, 1, 3).reshape(bf, d, C)


@torch.no_grad()
def unet_forward(sd: SD, sample: torch.Tensor, timestep, text: torch.Tensor) -> torch.Tensor:
    """UNet3DConditionModel.forward of the interpolation model (interpolation/models/unet.py:320-475):
    sample [B, 8, F, H, W] -> noise prediction [B, 4, F, H, W]."""
    return B.unet_forward(sd, sample, timestep, text, block=transformer_block)


# ----------------------------------------------------------------------------------------------------------------
# The caller: classifier-free guidance + respaced DDIM with channel-concat conditioning
# (interpolation/sample.py:138-174 `auto_inpainting_copy_no_mask`, interpolation/diffusion/gaussian_diffusion.py)
# ----------------------------------------------------------------------------------------------------------------
def space_timesteps(num_timesteps: int, count: int):
    """respace.py:9-63 for a single section ("50"): `count` indices from 0 .. num_timesteps-1 with a fractional stride,
    rounded with Python's round()."""
    stride = 1 if count <= 1 else (num_timesteps - 1) / (count - 1)
    return [round(i * stride) for i in range(count)]


def ddim_schedule(num_steps: int = 50, num_train: int = 1000):
    """create_diffusion(str(num_steps)) (diffusion/__init__.py:10-47): linear betas 1e-4 .. 2e-2 in float64
    (gaussian_diffusion.py:98-113), alphas_cumprod of the retained timesteps (the SpacedDiffusion betas reproduce exactly
    these products, respace.py:76-89).  Returns (timesteps kept, alpha_bar, alpha_bar_prev) in ascending order."""
    betas = torch.linspace(1e-4, 2e-2, num_train, dtype=torch.float64)
    acp = torch.cumprod(1.0 - betas, dim=0)
    use = space_timesteps(num_train, num_steps)
    ab = acp[use]
    ab_prev = torch.cat([torch.ones(1, dtype=torch.float64), ab[:-1]])
    return use, ab, ab_prev


def forward_with_cfg(sd: SD, x: torch.Tensor, t: int, text: torch.Tensor, cfg_scale: float = 4.0) -> torch.Tensor:
    """UNet3DConditionModel.forward_with_cfg (interpolation/models/unet.py:453-474): the first half of the batch is run
    with [prompt, negative prompt]; eps = uncond + s (cond - uncond), returned for both halves."""
    half = x[: len(x) // 2]
    eps = unet_forward(sd, torch.cat([half, half], dim=0), t, text)
    cond, uncond = eps.chunk(2, dim=0)
    g = uncond + cfg_scale * (cond - uncond)
    return torch.cat([g, g], dim=0)


@torch.no_grad()
def ddim_loop(sd: SD, z: torch.Tensor, x_start: torch.Tensor, text: torch.Tensor, num_steps: int = 50,
              cfg_scale: float = 4.0, model=None) -> torch.Tensor:
    """ddim_sample_loop(model.forward_with_cfg, ..., eta=0, clip_denoised=False, use_concat=True, copy_no_mask=True)
    (gaussian_diffusion.py:282-288, 362-395, 587-642, 723-778): every step feeds cat([x_t, x_start], dim=1) -- the noisy
    latent next to the copied key-frame latent -- to the UNet.  z, x_start: [2, 4, F, H, W] (both CFG halves);
    text = [prompt, negative prompt].  ``model(x8, t, text)`` may replace the oracle's guided forward."""
    use, ab, ab_prev = ddim_schedule(num_steps)
    fwd = model if model is not None else (lambda x8, t, e: forward_with_cfg(sd, x8, t, e, cfg_scale))
    x = z.double()
    for i in reversed(range(len(use))):
        eps = fwd(torch.cat([x.float(), x_start.float()], dim=1), use[i], text).double()
        x0 = x / ab[i].sqrt() - (1.0 / ab[i] - 1.0).sqrt() * eps        # _predict_xstart_from_eps
        x = ab_prev[i].sqrt() * x0 + (1.0 - ab_prev[i]).sqrt() * eps     # eta = 0: no noise term
        x = x.float().double()                                           # the reference keeps x in fp32 between steps
    return x.float()
