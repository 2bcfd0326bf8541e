"""Architecture description of the LaVie base denoiser and its parameter table.

The reference builds its UNet from the Stable-Diffusion-1.4 ``unet/config.json``
(/root/reference/base/models/unet.py:540-570, constructor defaults :104-140).  That
file is not shipped with the reference; the values below are the ones SURVEY.md
section 8a records.  ``param_spec`` derives the complete state_dict key -> shape
table (830 keys for the base config) so that the host module, the synthetic
weight generator and the loader all agree without instantiating anything.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field, asdict
from typing import Dict, Optional, Tuple


@dataclass(frozen=True)
class UNetConfig:
    sample_size: int = 64
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = (
        "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D")
    up_block_types: Tuple[str, ...] = (
        "UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D")
    layers_per_block: int = 2
    attention_head_dim: int = 8          # NOTE: used as the HEAD COUNT (unet_blocks.py:289-291)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    act_fn: str = "silu"
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    center_input_sample: bool = False
    use_linear_projection: bool = False
    rotary_dim: int = 32                 # RotaryEmbedding(32), unet.py:185
    rel_pos_buckets: int = 32            # RelativePositionBias(num_buckets=32, max_distance=32), attention.py:577
    rel_pos_max_distance: int = 32
    # "base" = the T2V denoiser (this round's hot path).  "interp" = the frame-interpolation UNet (SURVEY 8f N1,
    # interpolation/models/unet.py:476-557): 8 input channels, SparseCausal self-attention, plain temporal attention
    # (no rotary embedding, no relative-position bias keys in the state_dict).  Only the oracle uses "interp" so far.
    variant: str = "base"
    # "vsr" = the x4 video super-resolution UNet (vsr/models/unet.py:102-646 with
    # vsr/configs/unet_3d_config.json): 7 input channels (4 latent + 3 low-res RGB), noise-level class embedding, Linear
    # proj_in / proj_out, text-only "self" attention on the three high-resolution levels (only_cross_attention), a
    # temporal (3,1,1) ResNet in front of every transformer and a TemporalModule3D behind every block.
    only_cross_attention: Tuple[bool, ...] = (False, False, False, False)
    num_class_embeds: Optional[int] = None        # None (not 0): the reference builds the table whenever it is not None
    temporal_modules: bool = False
    max_noise_level: int = 350
    _diffusers_version: str = "0.16.0"

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @property
    def heads(self) -> int:
        return self.attention_head_dim

    def to_dict(self):
        return asdict(self)


BASE_CONFIG = UNetConfig()
INTERP_CONFIG = UNetConfig(in_channels=8, variant="interp")     # copy_no_mask / use_concat: interpolation/models/unet.py:501-507
VSR_CONFIG = UNetConfig(                                         # vsr/configs/unet_3d_config.json
    sample_size=128, in_channels=7, out_channels=4, block_out_channels=(256, 512, 512, 1024),
    down_block_types=("DownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D"),
    up_block_types=("CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "UpBlock3D"),
    cross_attention_dim=1024, variant="vsr", only_cross_attention=(True, True, True, False), num_class_embeds=1000,
    use_linear_projection=True, temporal_modules=True)


def _resnet(spec, p, cin, cout, temb):
    spec[f"{p}.norm1.weight"] = (cin,)
    spec[f"{p}.norm1.bias"] = (cin,)
    spec[f"{p}.conv1.weight"] = (cout, cin, 3, 3)
    spec[f"{p}.conv1.bias"] = (cout,)
    spec[f"{p}.time_emb_proj.weight"] = (cout, temb)
    spec[f"{p}.time_emb_proj.bias"] = (cout,)
    spec[f"{p}.norm2.weight"] = (cout,)
    spec[f"{p}.norm2.bias"] = (cout,)
    spec[f"{p}.conv2.weight"] = (cout, cout, 3, 3)
    spec[f"{p}.conv2.bias"] = (cout,)
    if cin != cout:
        spec[f"{p}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        spec[f"{p}.conv_shortcut.bias"] = (cout,)


def _attn(spec, p, c, ctx):
    spec[f"{p}.to_q.weight"] = (c, c)
    spec[f"{p}.to_k.weight"] = (c, ctx)
    spec[f"{p}.to_v.weight"] = (c, ctx)
    spec[f"{p}.to_out.0.weight"] = (c, c)
    spec[f"{p}.to_out.0.bias"] = (c,)


def _resnet_cnn(spec, p, c, k, temb):
    """ResnetBlock3DCNN (vsr/models/resnet.py:220-316): temporal (k,1,1) conv, then a (3,1,1) conv; same width."""
    spec[f"{p}.norm1.weight"] = (c,)
    spec[f"{p}.norm1.bias"] = (c,)
    spec[f"{p}.conv1.weight"] = (c, c, k, 1, 1)
    spec[f"{p}.conv1.bias"] = (c,)
    if temb:
        spec[f"{p}.time_emb_proj.weight"] = (c, temb)
        spec[f"{p}.time_emb_proj.bias"] = (c,)
    spec[f"{p}.norm2.weight"] = (c,)
    spec[f"{p}.norm2.bias"] = (c,)
    spec[f"{p}.conv2.weight"] = (c, c, 3, 1, 1)
    spec[f"{p}.conv2.bias"] = (c,)


def _temporal_module(spec, p, c, temb):
    """TemporalModule3D without attention layers (vsr/models/temporal_module.py:65-178, attention_block_types ("", ""))."""
    _resnet_cnn(spec, f"{p}.resblocks_3d_t", c, 5, temb)
    _resnet(spec, f"{p}.resblocks_3d_s", c, c, temb)
    spec[f"{p}.shift_conv.weight"] = (c, c, 1, 1)
    spec[f"{p}.shift_conv.bias"] = (c,)


def _transformer_vsr(spec, p, c, cfg: UNetConfig, only_cross: bool):
    _resnet_cnn(spec, f"{p}.resblock_temporal", c, 3, 0)
    spec[f"{p}.norm.weight"] = (c,)
    spec[f"{p}.norm.bias"] = (c,)
    spec[f"{p}.proj_in.weight"] = (c, c)
    spec[f"{p}.proj_in.bias"] = (c,)
    b = f"{p}.transformer_blocks.0"
    _attn(spec, f"{b}.attn1", c, cfg.cross_attention_dim if only_cross else c)
    spec[f"{b}.norm1.weight"] = (c,)
    spec[f"{b}.norm1.bias"] = (c,)
    _attn(spec, f"{b}.attn2", c, cfg.cross_attention_dim)
    spec[f"{b}.norm2.weight"] = (c,)
    spec[f"{b}.norm2.bias"] = (c,)
    _attn(spec, f"{b}.attn_temporal", c, c)
    spec[f"{b}.attn_temporal.time_rel_pos_bias.relative_attention_bias.weight"] = (cfg.rel_pos_buckets, cfg.heads)
    spec[f"{b}.attn_temporal.rotary_emb.freqs"] = (cfg.rotary_dim // 2,)
    spec[f"{b}.norm_temporal.weight"] = (c,)
    spec[f"{b}.norm_temporal.bias"] = (c,)
    spec[f"{b}.ff.net.0.proj.weight"] = (8 * c, c)
    spec[f"{b}.ff.net.0.proj.bias"] = (8 * c,)
    spec[f"{b}.ff.net.2.weight"] = (c, 4 * c)
    spec[f"{b}.ff.net.2.bias"] = (c,)
    spec[f"{b}.norm3.weight"] = (c,)
    spec[f"{b}.norm3.bias"] = (c,)
    spec[f"{p}.proj_out.weight"] = (c, c)
    spec[f"{p}.proj_out.bias"] = (c,)


def _transformer(spec, p, c, cfg: UNetConfig, only_cross: bool = False):
    if cfg.variant == "vsr":
        return _transformer_vsr(spec, p, c, cfg, only_cross)
    spec[f"{p}.norm.weight"] = (c,)
    spec[f"{p}.norm.bias"] = (c,)
    spec[f"{p}.proj_in.weight"] = (c, c, 1, 1)
    spec[f"{p}.proj_in.bias"] = (c,)
    b = f"{p}.transformer_blocks.0"
    _attn(spec, f"{b}.attn1", c, c)
    spec[f"{b}.norm1.weight"] = (c,)
    spec[f"{b}.norm1.bias"] = (c,)
    _attn(spec, f"{b}.attn2", c, cfg.cross_attention_dim)
    spec[f"{b}.norm2.weight"] = (c,)
    spec[f"{b}.norm2.bias"] = (c,)
    _attn(spec, f"{b}.attn_temp", c, c)
    if cfg.variant == "base":
        spec[f"{b}.attn_temp.time_rel_pos_bias.relative_attention_bias.weight"] = (cfg.rel_pos_buckets, cfg.heads)
        spec[f"{b}.attn_temp.rotary_emb.freqs"] = (cfg.rotary_dim // 2,)
    spec[f"{b}.norm_temp.weight"] = (c,)
    spec[f"{b}.norm_temp.bias"] = (c,)
    spec[f"{b}.ff.net.0.proj.weight"] = (8 * c, c)
    spec[f"{b}.ff.net.0.proj.bias"] = (8 * c,)
    spec[f"{b}.ff.net.2.weight"] = (c, 4 * c)
    spec[f"{b}.ff.net.2.bias"] = (c,)
    spec[f"{b}.norm3.weight"] = (c,)
    spec[f"{b}.norm3.bias"] = (c,)
    spec[f"{p}.proj_out.weight"] = (c, c, 1, 1)
    spec[f"{p}.proj_out.bias"] = (c,)


def param_spec(cfg: UNetConfig = BASE_CONFIG) -> "OrderedDict[str, Tuple[int, ...]]":
    """state_dict key -> shape, mirroring the module tree the reference builds in
    unet.py:147-290 / unet_blocks.py (block constructors)."""
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    boc = cfg.block_out_channels
    temb = cfg.time_embed_dim
    spec["conv_in.weight"] = (boc[0], cfg.in_channels, 3, 3)
    spec["conv_in.bias"] = (boc[0],)
    spec["time_embedding.linear_1.weight"] = (temb, boc[0])
    spec["time_embedding.linear_1.bias"] = (temb,)
    spec["time_embedding.linear_2.weight"] = (temb, temb)
    spec["time_embedding.linear_2.bias"] = (temb,)
    vsr = cfg.variant == "vsr"
    oca = cfg.only_cross_attention
    if cfg.num_class_embeds:
        spec["class_embedding.weight"] = (cfg.num_class_embeds, temb)         # vsr/models/unet.py:180
    if vsr:
        spec["temporal_rotary_emb.freqs"] = (cfg.rotary_dim // 2,)
    # down path (unet.py:187-218)
    out_c = boc[0]
    for i, kind in enumerate(cfg.down_block_types):
        in_c, out_c = out_c, boc[i]
        for j in range(cfg.layers_per_block):
            _resnet(spec, f"down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c, temb)
            if kind == "CrossAttnDownBlock3D":
                _transformer(spec, f"down_blocks.{i}.attentions.{j}", out_c, cfg, oca[i])
        if i != len(boc) - 1:
            spec[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (out_c, out_c, 3, 3)
            spec[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (out_c,)
        if cfg.temporal_modules:
            _temporal_module(spec, f"down_temporal_blocks.{i}", out_c, temb)
    # mid (unet.py:221-238)
    c = boc[-1]
    _resnet(spec, "mid_block.resnets.0", c, c, temb)
    _transformer(spec, "mid_block.attentions.0", c, cfg, oca[-1])
    _resnet(spec, "mid_block.resnets.1", c, c, temb)
    if cfg.temporal_modules:
        _temporal_module(spec, "mid_temporal_block", c, temb)
    # up path (unet.py:246-285; channel bookkeeping of unet_blocks.py:476-486, 598-600)
    rev = list(reversed(boc))
    out_c = rev[0]
    for i, kind in enumerate(cfg.up_block_types):
        prev_c, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, len(boc) - 1)]
        n = cfg.layers_per_block + 1
        for j in range(n):
            skip_c = in_c if j == n - 1 else out_c
            res_in = prev_c if j == 0 else out_c
            _resnet(spec, f"up_blocks.{i}.resnets.{j}", res_in + skip_c, out_c, temb)
            if kind == "CrossAttnUpBlock3D":
                _transformer(spec, f"up_blocks.{i}.attentions.{j}", out_c, cfg, oca[len(boc) - 1 - i])
        if i != len(boc) - 1:
            spec[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (out_c, out_c, 3, 3)
            spec[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (out_c,)
        if cfg.temporal_modules:
            _temporal_module(spec, f"up_temporal_blocks.{i}", out_c, temb)
    spec["conv_norm_out.weight"] = (boc[0],)
    spec["conv_norm_out.bias"] = (boc[0],)
    spec["conv_out.weight"] = (cfg.out_channels, boc[0], 3, 3)
    spec["conv_out.bias"] = (cfg.out_channels,)
    return spec
