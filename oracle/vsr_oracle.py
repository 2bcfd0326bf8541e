"""CPU oracle for the LaVie video super-resolution denoiser -- TEST INFRASTRUCTURE, NOT PRODUCT.

A plain-PyTorch fp32 restatement of ``UNet3DVSRModel.forward`` (vsr/models/unet.py:408-590 with
vsr/configs/unet_3d_config.json) as pure functions over a ``state_dict``.  Everything the VSR UNet shares with the base
denoiser (3x3 InflatedConv3d, ResnetBlock3D, attention, RoPE + relative-bias temporal attention, GEGLU) is imported from
``oracle/unet3d_oracle.py``; this file restates what is new:

* 7 input channels = noisy latent (4) | low-resolution RGB frames (3), concatenated on C (unet.py:446);
* the noise-level class embedding added to the time embedding (unet.py:180, 494-507);
* ``ResnetBlock3DCNN`` (resnet.py:220-316): GroupNorm(eps 1e-6) -> SiLU -> Conv3d (k,1,1) over FRAMES -> (+ time
  embedding) -> GroupNorm -> SiLU -> Conv3d (3,1,1) -> + input;
* ``Transformer3DModel`` of the VSR tree (attention.py:314-438): a (3,1,1) ResnetBlock3DCNN without time embedding in
  front, Linear proj_in / proj_out, and a block (attention.py:556-593) whose first attention reads the TEXT when
  ``only_cross_attention`` (the three high-resolution levels), order attn1 -> attn2 -> temporal -> feed-forward;
* ``TemporalModule3D`` (temporal_module.py:65-178, attention_block_types ("", ""), video_condition False) behind every
  block: (5,1,1) ResnetBlock3DCNN with time embedding -> spatial ResnetBlock3D (eps 1e-6) -> 1x1 shift_conv -> + input.

Parity pinning: ``tests/golden/make_golden_vsr.py`` runs the UNMODIFIED reference (imported from /root/reference/vsr with
the stand-ins of tests/golden/shims) on seeded inputs and weights; ``tests/test_oracle_vsr.py`` checks this file against
those vectors.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import unet3d_oracle as O

SD = O.SD
BLOCK_OUT = (256, 512, 512, 1024)
DOWN_HAS_ATTN = (False, True, True, True)
UP_HAS_ATTN = (True, True, True, False)
ONLY_CROSS = (True, True, True, False)          # per resolution level (unet.py:134-139)
LAYERS_PER_BLOCK = 2
GROUPS = 32
RESNET_EPS = 1e-5            # norm_eps -> the blocks' ResnetBlock3D (unet.py:221)
CNN_EPS = 1e-6               # default eps of ResnetBlock3DCNN and of TemporalModule3D's ResnetBlock3D (resnet.py:134,231)


def frame_conv(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.Conv3d with a (k,1,1) kernel and (k//2,0,0) zero padding (resnet.py:253-254,269): a 1-D conv over frames."""
    k = w.shape[2]
    return F.conv3d(x, w, b, padding=(k // 2, 0, 0))


def resnet_block_cnn(sd: SD, p: str, x: torch.Tensor, emb: Optional[torch.Tensor]) -> torch.Tensor:
    """ResnetBlock3DCNN.forward (resnet.py:284-316), in == out channels (no shortcut conv), output_scale_factor 1."""
    h = F.group_norm(x, GROUPS, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], CNN_EPS)
    h = frame_conv(F.silu(h), sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"])
    if emb is not None:
        t = F.linear(F.silu(emb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"])
        h = h + t[:, :, None, None, None]
    h = F.group_norm(h, GROUPS, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], CNN_EPS)
    h = frame_conv(F.silu(h), sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"])
    return x + h


def resnet_block(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor, eps: float) -> torch.Tensor:
    """ResnetBlock3D.forward (vsr/models/resnet.py:189-217) -- the base oracle's block with an explicit eps."""
    h = F.group_norm(x, GROUPS, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], eps)
    h = O.inflated_conv(F.silu(h), sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"])
    t = F.linear(F.silu(emb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"])
    h = h + t[:, :, None, None, None]
    h = F.group_norm(h, GROUPS, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], eps)
    h = O.inflated_conv(F.silu(h), sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"])
    if f"{p}.conv_shortcut.weight" in sd:
        x = O.inflated_conv(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"])
    return x + h


def temporal_module(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """TemporalModule3D.forward (temporal_module.py:151-178) without attention layers / video condition."""
    h = resnet_block_cnn(sd, f"{p}.resblocks_3d_t", x, emb)
    h = resnet_block(sd, f"{p}.resblocks_3d_s", h, emb, CNN_EPS)
    return x + O.inflated_conv(h, sd[f"{p}.shift_conv.weight"], sd[f"{p}.shift_conv.bias"])


def transformer_block(sd: SD, p: str, x: torch.Tensor, text: torch.Tensor, frames: int, only_cross: bool) -> torch.Tensor:
    """BasicTransformerBlock.forward (vsr/models/attention.py:556-593); x is [(B F), HW, C], text [(B F), L, 1024]."""
    C = x.shape[-1]
    ln = lambda t, n: F.layer_norm(t, (C,), sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"], O.LN_EPS)
    x = O.attention(sd, f"{p}.attn1", ln(x, "norm1"), text if only_cross else None) + x
    x = O.attention(sd, f"{p}.attn2", ln(x, "norm2"), text) + x
    bf, d, _ = x.shape
    b = bf // frames
    xt = x.reshape(b, frames, d, C).permute(0, 2, 1, 3).reshape(b * d, frames, C)      # (b f) d c -> (b d) f c
    xt = O.temporal_attention(sd, f"{p}.attn_temporal", ln(xt, "norm_temporal")) + xt
    x = xt.reshape(b, d, frames, C).permute(0, 2, 1, 3).reshape(bf, d, C)
    return O.geglu_ff(sd, f"{p}.ff", ln(x, "norm3")) + x


def transformer3d(sd: SD, p: str, x: torch.Tensor, text: torch.Tensor, only_cross: bool) -> torch.Tensor:
    """Transformer3DModel.forward (vsr/models/attention.py:386-438): temporal ResNet on the 5-D tensor, then per-frame
    GroupNorm (eps 1e-6) -> Linear -> block -> Linear -> + residual (the residual is the temporal ResNet's OUTPUT)."""
    x = resnet_block_cnn(sd, f"{p}.resblock_temporal", x, None)
    B, C, Fr, H, W = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W)
    text_f = text[:, None].expand(B, Fr, *text.shape[1:]).reshape(B * Fr, *text.shape[1:])
    h = F.group_norm(xf, GROUPS, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], O.TRANSFORMER_GN_EPS)
    h = h.permute(0, 2, 3, 1).reshape(B * Fr, H * W, C)
    h = F.linear(h, sd[f"{p}.proj_in.weight"], sd[f"{p}.proj_in.bias"])
    h = transformer_block(sd, f"{p}.transformer_blocks.0", h, text_f, Fr, only_cross)
    h = F.linear(h, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"])
    h = h.reshape(B * Fr, H, W, C).permute(0, 3, 1, 2) + xf
    return h.reshape(B, Fr, C, H, W).permute(0, 2, 1, 3, 4)


@torch.no_grad()
def unet_forward(sd: SD, sample: torch.Tensor, timestep, low_res: torch.Tensor, text: torch.Tensor, class_labels,
                 taps: Optional[dict] = None) -> torch.Tensor:
    """UNet3DVSRModel.forward (vsr/models/unet.py:408-590).  sample [B,4,F,H,W], low_res [B,3,F,H,W], text [B,L,1024],
    class_labels int64 [B] (the noise level; the reference's default scalar 20 does not index per batch item)."""
    sample, low_res, text = sample.float(), low_res.float(), text.float()
    B = sample.shape[0]
    if not torch.is_tensor(timestep):
        timestep = torch.tensor([timestep])
    timestep = timestep.reshape(-1).expand(B)
    emb = O.time_mlp(sd, O.timestep_embedding(timestep, BLOCK_OUT[0]))
    emb = emb + sd["class_embedding.weight"][torch.as_tensor(class_labels).reshape(-1).expand(B)]     # unet.py:494-507

    def tap(name, t):
        if taps is not None:
            taps[name] = t.clone()

    tap("emb", emb)
    x = O.inflated_conv(torch.cat([sample, low_res], dim=1), sd["conv_in.weight"], sd["conv_in.bias"])
    tap("conv_in", x)
    skips = [x]
    for i, has_attn in enumerate(DOWN_HAS_ATTN):
        for j in range(LAYERS_PER_BLOCK):
            x = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", x, emb, RESNET_EPS)
            if has_attn:
                x = transformer3d(sd, f"down_blocks.{i}.attentions.{j}", x, text, ONLY_CROSS[i])
            skips.append(x)
        if i != len(BLOCK_OUT) - 1:
            pd = f"down_blocks.{i}.downsamplers.0.conv"
            x = O.inflated_conv(x, sd[f"{pd}.weight"], sd[f"{pd}.bias"], stride=2)
            skips.append(x)
        x = temporal_module(sd, f"down_temporal_blocks.{i}", x, emb)          # unet.py:529-536 (skips are taken before)
        if i == 0:
            tap("down0", x)
    x = resnet_block(sd, "mid_block.resnets.0", x, emb, RESNET_EPS)
    x = transformer3d(sd, "mid_block.attentions.0", x, text, ONLY_CROSS[-1])
    x = resnet_block(sd, "mid_block.resnets.1", x, emb, RESNET_EPS)
    x = temporal_module(sd, "mid_temporal_block", x, emb)
    tap("mid", x)
    for i, has_attn in enumerate(UP_HAS_ATTN):
        for j in range(LAYERS_PER_BLOCK + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", x, emb, RESNET_EPS)
            if has_attn:
                x = transformer3d(sd, f"up_blocks.{i}.attentions.{j}", x, text, ONLY_CROSS[len(BLOCK_OUT) - 1 - i])
        if i != len(BLOCK_OUT) - 1:
            x = O.upsample(sd, f"up_blocks.{i}.upsamplers.0", x)
        x = temporal_module(sd, f"up_temporal_blocks.{i}", x, emb)
    tap("up_out", x)
    x = F.group_norm(x, GROUPS, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], RESNET_EPS)
    return O.inflated_conv(F.silu(x), sd["conv_out.weight"], sd["conv_out.bias"])
