"""CLIP text encoder on the B200 path: the ``text_encoder`` the pipelines call once per prompt
(base/pipelines/pipeline_videogen.py:337-348 and :395-406, ``self.text_encoder(text_input_ids)[0]``).

The reference takes it from ``transformers`` (``CLIPTextModel``, un-vendored third party): this module keeps that class's
``state_dict`` layout (``text_model.embeddings.*``, ``text_model.encoder.layers.N.*``, ``text_model.final_layer_norm.*``) and
call convention (``encoder(input_ids)[0]`` = ``last_hidden_state`` [B, L, C]) and runs the forward as C-ABI launches:
token + position embedding, then per layer LayerNorm -> fused q|k|v GEMM (+bias) -> causal attention over the <= 77 tokens
-> out_proj GEMM (+bias, +residual) -> LayerNorm -> fc1 GEMM (+bias) -> quick-GELU / GELU -> fc2 GEMM (+bias, +residual),
and the final LayerNorm.  Activations are bf16 with fp32 accumulation, like the denoiser.  No CPU fallback."""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Tuple

import torch
from torch import nn

from . import ops
from .synthetic import _gen

BF16 = torch.bfloat16
F32 = torch.float32


@dataclass(frozen=True)
class CLIPTextConfig:
    """transformers.CLIPTextConfig fields the forward depends on.  Defaults = openai/clip-vit-large-patch14 text tower,
    the text_encoder of Stable Diffusion 1.4 that LaVie's base and interpolation pipelines load."""
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    max_position_embeddings: int = 77
    hidden_act: str = "quick_gelu"
    layer_norm_eps: float = 1e-5


SD14_TEXT = CLIPTextConfig()
# text tower of stabilityai/stable-diffusion-x4-upscaler (OpenCLIP ViT-H/14), the VSR pipeline's text_encoder
X4_UPSCALER_TEXT = CLIPTextConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=23, num_attention_heads=16,
                                  hidden_act="gelu")


def clip_param_spec(cfg: CLIPTextConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    C, I = cfg.hidden_size, cfg.intermediate_size
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    spec["text_model.embeddings.token_embedding.weight"] = (cfg.vocab_size, C)
    spec["text_model.embeddings.position_embedding.weight"] = (cfg.max_position_embeddings, C)
    for i in range(cfg.num_hidden_layers):
        p = f"text_model.encoder.layers.{i}"
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            spec[f"{p}.self_attn.{n}.weight"] = (C, C)
            spec[f"{p}.self_attn.{n}.bias"] = (C,)
        spec[f"{p}.layer_norm1.weight"] = (C,)
        spec[f"{p}.layer_norm1.bias"] = (C,)
        spec[f"{p}.mlp.fc1.weight"] = (I, C)
        spec[f"{p}.mlp.fc1.bias"] = (I,)
        spec[f"{p}.mlp.fc2.weight"] = (C, I)
        spec[f"{p}.mlp.fc2.bias"] = (C,)
        spec[f"{p}.layer_norm2.weight"] = (C,)
        spec[f"{p}.layer_norm2.bias"] = (C,)
    spec["text_model.final_layer_norm.weight"] = (C,)
    spec["text_model.final_layer_norm.bias"] = (C,)
    return spec


def clip_synthetic_state_dict(cfg: CLIPTextConfig = SD14_TEXT, seed: int = 0):
    """Deterministic random-init weights (same keyed generator as lavie_b200.synthetic): Linear U(+-1/sqrt(fan_in)),
    embeddings N(0, 0.02), LayerNorm affine perturbed away from (1, 0)."""
    sd = OrderedDict()
    spec = clip_param_spec(cfg)
    for key, shape in spec.items():
        g = _gen(seed, "clip:" + key)
        if "embedding" in key:
            t = 0.02 * torch.randn(shape, generator=g)
        elif "layer_norm" in key:
            t = 0.1 * torch.randn(shape, generator=g) + (1.0 if key.endswith("weight") else 0.0)
        else:
            wshape = spec[key[: -len("bias")] + "weight"] if key.endswith("bias") else shape
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) / math.sqrt(wshape[1])
        sd[key] = t
    return sd


class _Out(tuple):
    """(last_hidden_state,) with attribute access, like transformers' BaseModelOutputWithPooling for what the pipelines
    read (``[0]`` / ``.last_hidden_state``)."""
    @property
    def last_hidden_state(self):
        return self[0]


class CLIPTextEncoder(nn.Module):
    """Drop-in for ``transformers.CLIPTextModel`` as the LaVie pipelines use it: ``encoder(input_ids)[0]``."""

    def __init__(self, config: CLIPTextConfig = SD14_TEXT):
        super().__init__()
        self.cfg = config
        self.config = SimpleNamespace(**config.__dict__)
        d = config.hidden_size // config.num_attention_heads
        if d not in (64, 128) or config.max_position_embeddings > 128 or config.hidden_act not in ("quick_gelu", "gelu"):
            raise NotImplementedError("CLIP text towers with 64- or 128-wide heads, <= 128 positions, quick_gelu / gelu")
        for key, shape in clip_param_spec(config).items():
            prefix, leaf = key.rsplit(".", 1)
            cur = self
            parts = prefix.split(".")
            for p in parts:
                nxt = cur._modules.get(p)
                if nxt is None:
                    nxt = nn.Module()
                    cur.add_module(p, nxt)
                cur = nxt
            cur.register_parameter(leaf, nn.Parameter(torch.empty(shape), requires_grad=False))
        self._packed = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self):
        self._packed = None

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._invalidate()
        return out

    @property
    def device(self):
        return self.text_model.final_layer_norm.weight.device

    @property
    def dtype(self):
        return self.text_model.final_layer_norm.weight.dtype

    def _pack(self):
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("lavie_b200.CLIPTextEncoder runs on CUDA (sm_100a) only; move it with .to('cuda')")
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        f32 = lambda k: sd[k].to(device=dev, dtype=F32).contiguous()
        b16 = lambda t: t.to(device=dev, dtype=BF16).contiguous()
        P: Dict[str, object] = {"tok": f32("text_model.embeddings.token_embedding.weight"),
                                "pos": f32("text_model.embeddings.position_embedding.weight"), "layers": []}
        for i in range(self.cfg.num_hidden_layers):
            p = f"text_model.encoder.layers.{i}"
            P["layers"].append({
                "ln1": (f32(f"{p}.layer_norm1.weight"), f32(f"{p}.layer_norm1.bias")),
                "ln2": (f32(f"{p}.layer_norm2.weight"), f32(f"{p}.layer_norm2.bias")),
                "w_qkv": b16(torch.cat([sd[f"{p}.self_attn.{n}_proj.weight"] for n in "qkv"], 0)),
                "b_qkv": torch.cat([sd[f"{p}.self_attn.{n}_proj.bias"] for n in "qkv"], 0).to(device=dev, dtype=F32).contiguous(),
                "w_o": b16(sd[f"{p}.self_attn.out_proj.weight"]), "b_o": f32(f"{p}.self_attn.out_proj.bias"),
                "w1": b16(sd[f"{p}.mlp.fc1.weight"]), "b1": f32(f"{p}.mlp.fc1.bias"),
                "w2": b16(sd[f"{p}.mlp.fc2.weight"]), "b2": f32(f"{p}.mlp.fc2.bias")})
        P["ln_f"] = (f32("text_model.final_layer_norm.weight"), f32("text_model.final_layer_norm.bias"))
        self._packed = P
        return P

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask=None, **unused):
        """``attention_mask`` is accepted like transformers does; the LaVie pipelines pass ``None`` (the SD text encoder
        config has no ``use_attention_mask``, pipeline_videogen.py:337-340), and only the causal mask applies."""
        if attention_mask is not None:
            raise NotImplementedError("padding masks are not part of the LaVie text-encoding path")
        if input_ids.dim() != 2 or input_ids.shape[1] > self.cfg.max_position_embeddings:
            raise ValueError(f"input_ids must be [B, L <= {self.cfg.max_position_embeddings}]")
        P = self._packed or self._pack()
        cfg = self.cfg
        B, L = input_ids.shape
        H = cfg.num_attention_heads
        d = cfg.hidden_size // H
        eps = cfg.layer_norm_eps
        ids = input_ids.to(device=self.device, dtype=torch.int64).reshape(-1).contiguous()
        x = ops.clip_embed(ids, P["tok"], P["pos"], L)
        for lyr in P["layers"]:
            n = ops.layernorm(x, lyr["ln1"][0], lyr["ln1"][1], eps)
            qkv = ops.gemm(n, lyr["w_qkv"], bias=lyr["b_qkv"])
            a = ops.causal_attention_small(qkv, B, L, H, d)
            x = ops.gemm(a, lyr["w_o"], bias=lyr["b_o"], residual=x)
            n = ops.layernorm(x, lyr["ln2"][0], lyr["ln2"][1], eps)
            h = ops.activation_(ops.gemm(n, lyr["w1"], bias=lyr["b1"]), cfg.hidden_act)
            x = ops.gemm(h, lyr["w2"], bias=lyr["b2"], residual=x)
        x = ops.layernorm(x, P["ln_f"][0], P["ln_f"][1], eps)
        return _Out((x.float().reshape(B, L, cfg.hidden_size),))
