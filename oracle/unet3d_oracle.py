"""CPU oracle for the LaVie base denoiser step -- TEST INFRASTRUCTURE, NOT PRODUCT.

A plain-PyTorch fp32 restatement of the reference's per-step denoiser
(`UNet3DConditionModel.forward`) and of the CFG + DDIM caller loop, written as
pure functions over a ``state_dict``.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this file;
the product path (``lavie_b200``) never does and fails loudly when its CUDA
library is missing.

Parity pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against the reference ITSELF: ``tests/golden/make_golden.py``
imports the unmodified reference from /root/reference (with stand-ins for the two
un-vendored third-party packages, diffusers==0.16.0 and rotary_embedding_torch),
runs it on seeded inputs/weights and commits the outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks this file against those vectors (max-abs error
~1e-6 fp32).  Third-party arithmetic restated here from published behaviour:
diffusers 0.16.0 ``get_timestep_embedding``/``TimestepEmbedding``/``GEGLU``
(in-tree mirrors: base/models/utils.py:74-94, vsr/models/diffusers_attention.py:
734-822) and rotary_embedding_torch ``rotate_queries_or_keys`` (no mirror).

Tensors keep the reference's layouts ([B,C,F,H,W] feature maps, [(B F),HW,C]
tokens) on purpose: the CUDA path is channels-last throughout, so a layout slip
there cannot cancel against the same slip here.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# ----------------------------------------------------------------------------
# Architecture constants of the SD-1.4-derived base config.
# ----------------------------------------------------------------------------
BLOCK_OUT = (320, 640, 1280, 1280)
DOWN_HAS_ATTN = (True, True, True, False)
UP_HAS_ATTN = (False, True, True, True)
LAYERS_PER_BLOCK = 2
HEADS = 8
GROUPS = 32
RESNET_EPS = 1e-5          # norm_eps -> ResnetBlock3D(eps=resnet_eps), unet.py:199
TRANSFORMER_GN_EPS = 1e-6  # attention.py:324
LN_EPS = 1e-5              # nn.LayerNorm default, attention.py:444
ROT_DIM = 32


# ----------------------------------------------------------------------------
# Leaves
# ----------------------------------------------------------------------------
def timestep_embedding(timesteps: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """diffusers 0.16 Timesteps(320, flip_sin_to_cos=True, freq_shift=0): [cos | sin],
    w_i = 10000^(-i/160).  Call site unet.py:153,428; mirror base/models/utils.py:74-94."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def time_mlp(sd: SD, t_emb: torch.Tensor) -> torch.Tensor:
    """TimestepEmbedding: Linear -> SiLU -> Linear (unet.py:156,434)."""
    h = F.linear(t_emb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
    return F.linear(F.silu(h), sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])


def inflated_conv(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """InflatedConv3d = the same 2-D conv on every frame (resnet.py:13-21)."""
    B, C, Fr, H, W = x.shape
    k = w.shape[-1]
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W), w, b, stride=stride, padding=k // 2)
    return y.reshape(B, Fr, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)


def resnet_block(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """ResnetBlock3D.forward (resnet.py:177-207).  GroupNorm sees the 5-D tensor, so
    its statistics span channels-in-group x FRAMES x H x W per batch item."""
    h = F.group_norm(x, GROUPS, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], RESNET_EPS)
    h = inflated_conv(F.silu(h), sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"])
    t = F.linear(F.silu(emb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"])
    h = h + t[:, :, None, None, None]
    h = F.group_norm(h, GROUPS, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], RESNET_EPS)
    h = inflated_conv(F.silu(h), sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"])
    if f"{p}.conv_shortcut.weight" in sd:
        x = inflated_conv(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"])
    return x + h


def _split_heads(t: torch.Tensor) -> torch.Tensor:
    n, s, c = t.shape
    return t.reshape(n, s, HEADS, c // HEADS).permute(0, 2, 1, 3)      # [n, heads, s, d]


def _merge_heads(t: torch.Tensor) -> torch.Tensor:
    n, h, s, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(n, s, h * d)


def attention(sd: SD, p: str, x: torch.Tensor, ctx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """CrossAttention.forward/_attention (attention.py:146-239): softmax(q k^T / sqrt(d)) v,
    q/k/v without bias, to_out[0] with bias.  ctx=None -> self-attention."""
    ctx = x if ctx is None else ctx
    q = _split_heads(F.linear(x, sd[f"{p}.to_q.weight"]))
    k = _split_heads(F.linear(ctx, sd[f"{p}.to_k.weight"]))
    v = _split_heads(F.linear(ctx, sd[f"{p}.to_v.weight"]))
    scale = q.shape[-1] ** -0.5
    probs = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * scale, dim=-1)
    o = _merge_heads(torch.matmul(probs, v))
    return F.linear(o, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"])


def rel_pos_bucket(rel: torch.Tensor, num_buckets: int = 32, max_distance: int = 32) -> torch.Tensor:
    """T5-style bucket index (attention.py:680-698)."""
    n = -rel
    half = num_buckets // 2
    ret = (n < 0).long() * half
    n = n.abs()
    max_exact = half // 2
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact)
                         * (half - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, half - 1))
    return ret + torch.where(n < max_exact, n, large)


def rel_pos_bias(table: torch.Tensor, n: int) -> torch.Tensor:
    """RelativePositionBias.forward (attention.py:700-707): [heads, n, n], rel = k_pos - q_pos."""
    pos = torch.arange(n)
    rel = pos[None, :] - pos[:, None]
    return table[rel_pos_bucket(rel, table.shape[0], 32)].permute(2, 0, 1)


def rope(t: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """rotary_embedding_torch.rotate_queries_or_keys on [..., F, d]: rotates dims [0, 2*len(freqs))
    as adjacent pairs (x0,x1)->(x0 cos - x1 sin, x1 cos + x0 sin), angle = position * freqs[i]."""
    n = t.shape[-2]
    rot = 2 * freqs.shape[0]
    ang = (torch.arange(n, dtype=freqs.dtype)[:, None] * freqs[None, :]).repeat_interleave(2, dim=-1)
    a, rest = t[..., :rot], t[..., rot:]
    pairs = a.reshape(*a.shape[:-1], rot // 2, 2)
    rotated = torch.stack((-pairs[..., 1], pairs[..., 0]), dim=-1).reshape(a.shape)
    return torch.cat((a * ang.cos() + rotated * ang.sin(), rest), dim=-1)


def temporal_attention(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """TemporalAttention.forward/_attention (attention.py:580-667) on [(B HW), F, C]:
    q scaled BEFORE RoPE (:640), RoPE on q and k (:644-646), + per-head bias (:650)."""
    q = _split_heads(F.linear(x, sd[f"{p}.to_q.weight"]))
    k = _split_heads(F.linear(x, sd[f"{p}.to_k.weight"]))
    v = _split_heads(F.linear(x, sd[f"{p}.to_v.weight"]))
    q = q * (q.shape[-1] ** -0.5)
    freqs = sd[f"{p}.rotary_emb.freqs"]
    q, k = rope(q, freqs), rope(k, freqs)
    scores = torch.matmul(q, k.transpose(-1, -2))
    scores = scores + rel_pos_bias(sd[f"{p}.time_rel_pos_bias.relative_attention_bias.weight"], x.shape[1])
    probs = torch.softmax(scores - scores.amax(dim=-1, keepdim=True), dim=-1)
    o = _merge_heads(torch.matmul(probs, v))
    return F.linear(o, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"])


def geglu_ff(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """diffusers FeedForward(activation_fn='geglu') (mirror vsr/models/diffusers_attention.py:734-822):
    Linear(C->8C) -> h * gelu_erf(gate) -> Linear(4C->C)."""
    hg = F.linear(x, sd[f"{p}.net.0.proj.weight"], sd[f"{p}.net.0.proj.bias"])
    h, gate = hg.chunk(2, dim=-1)
    return F.linear(h * F.gelu(gate), sd[f"{p}.net.2.weight"], sd[f"{p}.net.2.bias"])


def transformer_block(sd: SD, p: str, x: torch.Tensor, text: torch.Tensor, frames: int) -> torch.Tensor:
    """BasicTransformerBlock.forward, eval branch (attention.py:511-560); x is [(B F), HW, C]."""
    C = x.shape[-1]
    ln = lambda t, n: F.layer_norm(t, (C,), sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"], LN_EPS)
    x = attention(sd, f"{p}.attn1", ln(x, "norm1")) + x
    x = attention(sd, f"{p}.attn2", ln(x, "norm2"), text) + x
    bf, d, _ = x.shape
    b = bf // frames
    xt = x.reshape(b, frames, d, C).permute(0, 2, 1, 3).reshape(b * d, frames, C)      # (b f) d c -> (b d) f c
    xt = temporal_attention(sd, f"{p}.attn_temp", ln(xt, "norm_temp")) + xt
    x = xt.reshape(b, d, frames, C).permute(0, 2, 1, 3).reshape(bf, d, C)
    return geglu_ff(sd, f"{p}.ff", ln(x, "norm3")) + x


def transformer3d(sd: SD, p: str, x: torch.Tensor, text: torch.Tensor, block=None) -> torch.Tensor:
    """Transformer3DModel.forward (attention.py:358-407): per-FRAME GroupNorm (4-D input,
    eps 1e-6) -> 1x1 conv -> tokens -> block -> 1x1 conv -> + residual."""
    B, C, Fr, H, W = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W)
    text_f = text[:, None].expand(B, Fr, *text.shape[1:]).reshape(B * Fr, *text.shape[1:])
    h = F.group_norm(xf, GROUPS, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], TRANSFORMER_GN_EPS)
    h = F.conv2d(h, sd[f"{p}.proj_in.weight"], sd[f"{p}.proj_in.bias"])
    h = h.permute(0, 2, 3, 1).reshape(B * Fr, H * W, C)
    h = (block or transformer_block)(sd, f"{p}.transformer_blocks.0", h, text_f, Fr)
    h = h.reshape(B * Fr, H, W, C).permute(0, 3, 1, 2)
    h = F.conv2d(h, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"]) + xf
    return h.reshape(B, Fr, C, H, W).permute(0, 2, 1, 3, 4)


def upsample(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """Upsample3D (resnet.py:44-76): nearest x2 in H and W only, then 3x3 conv."""
    x = F.interpolate(x, scale_factor=[1.0, 2.0, 2.0], mode="nearest")
    return inflated_conv(x, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"])


# ----------------------------------------------------------------------------
# The denoiser step
# ----------------------------------------------------------------------------
@torch.no_grad()
def unet_forward(sd: SD, sample: torch.Tensor, timestep, text: torch.Tensor,
                 taps: Optional[dict] = None, block=None) -> torch.Tensor:
    """UNet3DConditionModel.forward (unet.py:366-512).  ``taps`` (optional dict) receives
    a few intermediate activations for debugging the CUDA path.  ``block`` replaces the
    transformer block (oracle/interp_oracle.py: the interpolation UNet shares everything else)."""
    sample = sample.float()
    text = text.float()
    B = sample.shape[0]
    if not torch.is_tensor(timestep):
        timestep = torch.tensor([timestep])
    timestep = timestep.reshape(-1).expand(B)
    emb = time_mlp(sd, timestep_embedding(timestep, BLOCK_OUT[0]))

    def tap(name, t):
        if taps is not None:
            taps[name] = t.clone()

    tap("emb", emb)
    x = inflated_conv(sample, sd["conv_in.weight"], sd["conv_in.bias"])
    tap("conv_in", x)
    skips = [x]
    for i, has_attn in enumerate(DOWN_HAS_ATTN):                      # unet_blocks.py:320-362, 417-441
        for j in range(LAYERS_PER_BLOCK):
            x = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", x, emb)
            if i == 0 and j == 0:
                tap("down0_res0", x)
            if has_attn:
                x = transformer3d(sd, f"down_blocks.{i}.attentions.{j}", x, text, block)
                if i == 0 and j == 0:
                    tap("down0_attn0", x)
            skips.append(x)
        if i != len(BLOCK_OUT) - 1:
            pd = f"down_blocks.{i}.downsamplers.0.conv"
            x = inflated_conv(x, sd[f"{pd}.weight"], sd[f"{pd}.bias"], stride=2)   # resnet.py:102-110
            skips.append(x)
    tap("down_out", x)
    x = resnet_block(sd, "mid_block.resnets.0", x, emb)              # unet_blocks.py:226-232
    x = transformer3d(sd, "mid_block.attentions.0", x, text, block)
    x = resnet_block(sd, "mid_block.resnets.1", x, emb)
    tap("mid", x)
    for i, has_attn in enumerate(UP_HAS_ATTN):                        # unet_blocks.py:524-574, 625-648
        for j in range(LAYERS_PER_BLOCK + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", x, emb)
            if has_attn:
                x = transformer3d(sd, f"up_blocks.{i}.attentions.{j}", x, text, block)
        if i != len(BLOCK_OUT) - 1:
            x = upsample(sd, f"up_blocks.{i}.upsamplers.0", x)
    tap("up_out", x)
    x = F.group_norm(x, GROUPS, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], RESNET_EPS)
    return inflated_conv(F.silu(x), sd["conv_out.weight"], sd["conv_out.bias"])


# ----------------------------------------------------------------------------
# The caller: CFG combine + DDIM update (pipeline_videogen.py:664-689)
# ----------------------------------------------------------------------------
def ddim_schedule(num_steps: int = 50, num_train: int = 1000, beta_start: float = 1e-4,
                  beta_end: float = 2e-2, steps_offset: int = 1):
    """Linear betas (base/configs/sample.yaml:23-25) and stock diffusers-0.16 DDIM timesteps
    (commented block vsr/diffusion/scheduling_ddim.py:259-265): (arange(n)*ratio)[::-1] + offset."""
    betas = torch.linspace(beta_start, beta_end, num_train, dtype=torch.float32)
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    ratio = num_train // num_steps
    timesteps = (torch.arange(num_steps) * ratio).flip(0) + steps_offset
    return alphas_cumprod, timesteps, ratio


def ddim_step(noise_pred, t: int, latents, alphas_cumprod, ratio: int):
    """eta=0, epsilon prediction, clip_sample=False, set_alpha_to_one=False
    (vsr/diffusion/scheduling_ddim.py:345-394)."""
    prev_t = t - ratio
    a_t = alphas_cumprod[t]
    a_prev = alphas_cumprod[prev_t] if prev_t >= 0 else alphas_cumprod[0]
    x0 = (latents - (1 - a_t).sqrt() * noise_pred) / a_t.sqrt()
    return a_prev.sqrt() * x0 + (1 - a_prev).sqrt() * noise_pred


@torch.no_grad()
def cfg_ddim_loop(sd: SD, latents: torch.Tensor, text_uncond_cond: torch.Tensor,
                  guidance_scale: float = 7.5, num_steps: int = 50, unet=None) -> torch.Tensor:
    """The denoising loop of VideoGenPipeline.__call__ (pipeline_videogen.py:664-689) with the
    DDIM scheduler.  ``unet`` may replace the oracle forward (same signature) so tests can drive
    the product module through the identical loop."""
    fwd = unet if unet is not None else (lambda x, t, e: unet_forward(sd, x, t, e))
    acp, timesteps, ratio = ddim_schedule(num_steps)
    for t in timesteps.tolist():
        model_in = torch.cat([latents] * 2)
        noise = fwd(model_in, t, text_uncond_cond)
        n_uncond, n_text = noise.chunk(2)
        noise = n_uncond + guidance_scale * (n_text - n_uncond)
        latents = ddim_step(noise, t, latents, acp, ratio)
    return latents
