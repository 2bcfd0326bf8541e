"""SURVEY 8f row N4 on the B200: the CLIP text encoder (lavie_b200.clip.CLIPTextEncoder) against vectors produced by
transformers.CLIPTextModel and against the pinned oracle; the VAE decoder (lavie_b200.vae.VAEDecoder) against the oracle
restatement of diffusers' decoder (parity unpinned: no diffusers offline, see oracle/vae_oracle.py)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
# The VAE decoder stacks 31 convolutions, 30 GroupNorms and an attention block between the latent and the pixels; with bf16
# storage after every one of them the ORACLE ITSELF (weights and every intermediate rounded to bf16 on the CPU) is 2.3e-2
# from its fp32 run on these inputs, so the stated tolerance for this path is 4e-2 (the images are quantised to 8 bits
# right after: 4e-2 of the signal is < 0.3 grey levels rms).
VAE_TOL = 4e-2
DEV = "cuda"


def test_causal_attention_small_kernel():
    from lavie_b200 import ops
    for B, L, H, d in [(2, 77, 12, 64), (3, 20, 16, 64), (1, 128, 4, 128), (2, 1, 2, 64)]:
        qkv = torch.randn(B * L, 3 * H * d, generator=torch.Generator().manual_seed(L)).to(torch.bfloat16).to(DEV)
        out = ops.causal_attention_small(qkv, B, L, H, d)
        q, k, v = [t.float().reshape(B, L, H, d).transpose(1, 2) for t in qkv.reshape(B * L, 3, H * d).unbind(1)]
        ref = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B * L, H * d)
        assert rel_l2(out.float(), ref) < 5e-3


def test_activation_softmax_embed_uint8_kernels():
    from lavie_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = (3 * torch.randn(1000, 64, generator=g)).to(torch.bfloat16).to(DEV)
    xf = x.float()
    assert rel_l2(ops.activation_(x.clone(), "quick_gelu").float(), xf * torch.sigmoid(1.702 * xf)) < 4e-3
    assert rel_l2(ops.activation_(x.clone(), "gelu").float(), F.gelu(xf)) < 4e-3
    for rows, n in [(300, 2560), (64, 64), (10, 4096), (5, 24)]:
        s = (2 * torch.randn(rows, n, generator=g)).to(torch.bfloat16).to(DEV)
        ref = torch.softmax(0.7 * s.float(), dim=-1)
        assert rel_l2(ops.softmax_rows_(s.clone(), 0.7).float(), ref) < 4e-3
    tok, pos = torch.randn(500, 768, generator=g).to(DEV), torch.randn(77, 768, generator=g).to(DEV)
    ids = torch.randint(0, 500, (3 * 77,), generator=g).to(DEV)
    ref = tok[ids] + pos.repeat(3, 1)
    assert rel_l2(ops.clip_embed(ids, tok, pos, 77).float(), ref) < 4e-3
    y = (1.2 * torch.randn(4096, 32, generator=g)).to(torch.bfloat16).to(DEV)
    want = ((y[:, :3].float() / 2 + 0.5) * 255).add_(0.5).clamp_(0, 255).to(torch.uint8)
    assert torch.equal(ops.image_to_uint8(y), want)


@pytest.mark.parametrize("name", ["clip_sd14_b2", "clip_vith3_b3"])
def test_clip_matches_transformers_golden(name):
    from lavie_b200.clip import CLIPTextConfig, CLIPTextEncoder, clip_synthetic_state_dict
    g = load_golden(name)
    cfg = CLIPTextConfig(**g["cfg"])
    enc = CLIPTextEncoder(cfg)
    enc.load_state_dict(clip_synthetic_state_dict(cfg, seed=g["weights_seed"]), strict=True)
    enc = enc.to(DEV).eval()
    out = enc(g["ids"].to(DEV))
    assert out[0].shape == g["out"].shape and out.last_hidden_state is out[0]
    assert rel_l2(out[0].cpu(), g["out"]) < BF16_TOL
    assert torch.equal(enc(g["ids"].to(DEV))[0], out[0])


def test_clip_short_prompt_vs_oracle():
    """L < 77 and a batch of one, against the pinned oracle."""
    from lavie_b200.clip import CLIPTextConfig, CLIPTextEncoder, clip_synthetic_state_dict
    from oracle import clip_oracle as C
    cfg = CLIPTextConfig(num_hidden_layers=4)
    sd = clip_synthetic_state_dict(cfg, seed=1)
    enc = CLIPTextEncoder(cfg)
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(DEV).eval()
    ids = torch.randint(0, cfg.vocab_size, (1, 13), generator=torch.Generator().manual_seed(3))
    ref = C.clip_text_forward(sd, ids, cfg.num_attention_heads, cfg.hidden_act)
    assert rel_l2(enc(ids.to(DEV))[0].cpu(), ref) < BF16_TOL


@pytest.fixture(scope="module")
def vae():
    from lavie_b200.vae import VAEDecoder, vae_synthetic_state_dict
    sd = vae_synthetic_state_dict(seed=0)
    m = VAEDecoder()
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


@pytest.mark.parametrize("N,h,w", [(3, 8, 8), (2, 16, 24), (1, 40, 64)])
def test_vae_decode_vs_oracle(vae, N, h, w):
    from oracle import vae_oracle as V
    m, sd = vae
    z = torch.randn(N, 4, h, w, generator=torch.Generator().manual_seed(h * w))
    ref = V.decode(sd, z)
    out = m.decode(z.to(DEV)).sample
    assert out.shape == ref.shape == (N, 3, 8 * h, 8 * w) and out.dtype == torch.float32
    assert rel_l2(out.cpu(), ref) < VAE_TOL


def test_vae_decode_latents_uint8(vae):
    """decode_latents (pipeline_videogen.py:422-429): uint8 frames [B,F,H,W,3]; bf16 rounding may move a value across an
    integer boundary, so compare as images: mean absolute difference below one grey level, no outliers beyond 12."""
    from oracle import vae_oracle as V
    m, sd = vae
    lat = 0.18215 * torch.randn(1, 4, 3, 16, 16, generator=torch.Generator().manual_seed(9))
    ref = V.decode_latents(sd, lat)
    out = m.decode_latents(lat.to(DEV))
    assert out.shape == ref.shape == (1, 3, 128, 128, 3) and out.dtype == torch.uint8 and out.device.type == "cpu"
    diff = (out.int() - ref.int()).abs()
    assert float(diff.float().mean()) < 1.0 and int(diff.max()) <= 12


def test_vae_loads_a_full_autoencoder_checkpoint_layout(vae):
    """Encoder-side keys are ignored and the post-0.16 attention parameter names are accepted."""
    from lavie_b200.vae import VAEDecoder, load_vae_state_dict
    m, sd = vae
    full = dict(sd)
    full["encoder.conv_in.weight"] = torch.zeros(128, 3, 3, 3)
    full["quant_conv.weight"] = torch.zeros(8, 8, 1, 1)
    a = "decoder.mid_block.attentions.0"
    for new, old in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn")):
        full[f"{a}.{new}.weight"] = full.pop(f"{a}.{old}.weight")
        full[f"{a}.{new}.bias"] = full.pop(f"{a}.{old}.bias")
    m2 = VAEDecoder()
    load_vae_state_dict(m2, full, strict=True)
    m2 = m2.to(DEV).eval()
    z = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(2)).to(DEV)
    assert torch.equal(m2.decode(z).sample, m.decode(z).sample)
