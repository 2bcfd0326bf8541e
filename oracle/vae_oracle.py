"""CPU oracle for the VAE decoder -- TEST INFRASTRUCTURE, NOT PRODUCT.

**PARITY UNPINNED.**  The reference decodes with ``diffusers.AutoencoderKL`` (``diffusers==0.16.0``, third party, not vendored
and not installable offline).  Its in-tree mirror vsr/models/autoencoder_kl.py:179-192 shows the call order
(``post_quant_conv`` -> ``Decoder``) but imports the ``Decoder`` itself from diffusers.  This file restates the published
diffusers-0.16 ``Decoder`` of the Stable Diffusion VAE (block_out_channels (128, 256, 512, 512), layers_per_block 2, 32
groups, eps 1e-6): conv_in -> UNetMidBlock2D (ResnetBlock2D, single-head AttentionBlock, ResnetBlock2D) -> four
UpDecoderBlock2D (three ResnetBlock2D each, nearest-x2 Upsample2D + 3x3 conv on the first three) -> GroupNorm -> SiLU ->
conv_out, and the reference's own caller ``decode_latents`` (base/pipelines/pipeline_videogen.py:422-429).  No reference
run or golden vector of the real class exists here; the GPU path is tested against THIS restatement only.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

GROUPS = 32
EPS = 1e-6
UP_CHANNELS = (512, 512, 256, 128)


def resnet2d(sd, p, x):
    """diffusers ResnetBlock2D with temb=None, output_scale_factor 1."""
    h = F.group_norm(x, GROUPS, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], EPS)
    h = F.conv2d(F.silu(h), sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    h = F.group_norm(h, GROUPS, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], EPS)
    h = F.conv2d(F.silu(h), sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"])
    return x + h


def attention_block(sd, p, x):
    """diffusers 0.16 AttentionBlock, one head: GroupNorm -> q, k, v Linear -> softmax(q k^T / sqrt(C)) v -> proj -> + x."""
    N, C, H, W = x.shape
    h = F.group_norm(x, GROUPS, sd[f"{p}.group_norm.weight"], sd[f"{p}.group_norm.bias"], EPS)
    h = h.reshape(N, C, H * W).transpose(1, 2)
    q = F.linear(h, sd[f"{p}.query.weight"], sd[f"{p}.query.bias"])
    k = F.linear(h, sd[f"{p}.key.weight"], sd[f"{p}.key.bias"])
    v = F.linear(h, sd[f"{p}.value.weight"], sd[f"{p}.value.bias"])
    probs = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1)
    o = F.linear(probs @ v, sd[f"{p}.proj_attn.weight"], sd[f"{p}.proj_attn.bias"])
    return x + o.transpose(1, 2).reshape(N, C, H, W)


@torch.no_grad()
def decode(sd, z: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL._decode (mirror vsr/models/autoencoder_kl.py:179-192): z [N,4,h,w] -> image [N,3,8h,8w]."""
    sd = {k: v.float() for k, v in sd.items()}
    x = F.conv2d(z.float(), sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = F.conv2d(x, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    x = resnet2d(sd, "decoder.mid_block.resnets.0", x)
    x = attention_block(sd, "decoder.mid_block.attentions.0", x)
    x = resnet2d(sd, "decoder.mid_block.resnets.1", x)
    for i in range(len(UP_CHANNELS)):
        for j in range(3):
            x = resnet2d(sd, f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i != len(UP_CHANNELS) - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"],
                         sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    x = F.group_norm(x, GROUPS, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], EPS)
    return F.conv2d(F.silu(x), sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


@torch.no_grad()
def decode_latents(sd, latents: torch.Tensor) -> torch.Tensor:
    """VideoGenPipeline.decode_latents (base/pipelines/pipeline_videogen.py:422-429): [B,4,F,h,w] -> uint8 [B,F,H,W,3]."""
    B, C, Fr, h, w = latents.shape
    z = (1 / 0.18215 * latents).permute(0, 2, 1, 3, 4).reshape(B * Fr, C, h, w)
    video = decode(sd, z)
    video = video.reshape(B, Fr, 3, video.shape[-2], video.shape[-1]).permute(0, 1, 3, 4, 2)
    return ((video / 2 + 0.5) * 255).add_(0.5).clamp_(0, 255).to(dtype=torch.uint8).contiguous()
