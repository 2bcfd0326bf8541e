"""CPU oracle for the CLIP text encoder -- TEST INFRASTRUCTURE, NOT PRODUCT.

The reference calls ``transformers.CLIPTextModel`` (third party, not vendored in /root/reference; call sites
base/pipelines/pipeline_videogen.py:337-348, 395-406: ``self.text_encoder(ids)[0]``).  This file restates the published
forward of ``CLIPTextTransformer`` as pure functions over the model's ``state_dict``: token + position embedding; per layer
pre-LN causal multi-head self-attention (q scaled by d^-1/2, additive causal mask) and a pre-LN MLP with quick-GELU
(x * sigmoid(1.702 x)) or erf GELU; final LayerNorm.  Pinned against ``transformers`` itself (the library the reference
imports; version in tests/golden/make_golden_clip.py) on seeded weights: tests/test_oracle_clip.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


@torch.no_grad()
def clip_text_forward(sd, input_ids: torch.Tensor, heads: int, act: str = "quick_gelu", eps: float = 1e-5) -> torch.Tensor:
    """-> last_hidden_state [B, L, C] (fp32)."""
    tok = sd["text_model.embeddings.token_embedding.weight"].float()
    pos = sd["text_model.embeddings.position_embedding.weight"].float()
    B, L = input_ids.shape
    C = tok.shape[1]
    d = C // heads
    x = tok[input_ids] + pos[:L][None]
    mask = torch.full((L, L), float("-inf")).triu(1)
    i = 0
    while f"text_model.encoder.layers.{i}.layer_norm1.weight" in sd:
        p = f"text_model.encoder.layers.{i}"
        lin = lambda t, n: F.linear(t, sd[f"{p}.{n}.weight"].float(), sd[f"{p}.{n}.bias"].float())
        n = F.layer_norm(x, (C,), sd[f"{p}.layer_norm1.weight"].float(), sd[f"{p}.layer_norm1.bias"].float(), eps)
        q = (lin(n, "self_attn.q_proj") * d ** -0.5).reshape(B, L, heads, d).transpose(1, 2)
        k = lin(n, "self_attn.k_proj").reshape(B, L, heads, d).transpose(1, 2)
        v = lin(n, "self_attn.v_proj").reshape(B, L, heads, d).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(-1, -2) + mask, dim=-1) @ v
        x = x + lin(a.transpose(1, 2).reshape(B, L, C), "self_attn.out_proj")
        n = F.layer_norm(x, (C,), sd[f"{p}.layer_norm2.weight"].float(), sd[f"{p}.layer_norm2.bias"].float(), eps)
        h = lin(n, "mlp.fc1")
        h = h * torch.sigmoid(1.702 * h) if act == "quick_gelu" else F.gelu(h)
        x = x + lin(h, "mlp.fc2")
        i += 1
    return F.layer_norm(x, (C,), sd["text_model.final_layer_norm.weight"].float(),
                        sd["text_model.final_layer_norm.bias"].float(), eps)
