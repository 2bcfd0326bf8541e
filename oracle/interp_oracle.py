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
