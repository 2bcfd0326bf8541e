"""SURVEY 8f row N4 (CLIP text encode): the oracle restatement of transformers' CLIPTextTransformer is pinned against
vectors produced by transformers.CLIPTextModel itself (tests/golden/make_golden_clip.py).  CPU only."""
import pytest
import torch

from conftest import load_golden, rel_l2


@pytest.mark.parametrize("name", ["clip_sd14_b2", "clip_vith3_b3"])
def test_clip_oracle_matches_transformers_golden(name):
    from lavie_b200.clip import CLIPTextConfig, clip_synthetic_state_dict
    from oracle import clip_oracle as C
    g = load_golden(name)
    cfg = CLIPTextConfig(**g["cfg"])
    sd = clip_synthetic_state_dict(cfg, seed=g["weights_seed"])
    out = C.clip_text_forward(sd, g["ids"], cfg.num_attention_heads, cfg.hidden_act, cfg.layer_norm_eps)
    assert out.shape == g["out"].shape
    assert rel_l2(out, g["out"]) < 2e-5


def test_clip_module_keeps_the_transformers_state_dict_layout():
    from lavie_b200.clip import CLIPTextEncoder, SD14_TEXT, clip_param_spec, clip_synthetic_state_dict
    m = CLIPTextEncoder()
    spec = clip_param_spec(SD14_TEXT)
    assert len(spec) == 196 and set(m.state_dict().keys()) == set(spec.keys())
    m.load_state_dict(clip_synthetic_state_dict(SD14_TEXT), strict=True)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 77, dtype=torch.long))                  # no CPU fallback
    # causality: the oracle's token i never sees token j > i
    from oracle import clip_oracle as C
    from lavie_b200.clip import CLIPTextConfig
    cfg = CLIPTextConfig(num_hidden_layers=2)
    sd = clip_synthetic_state_dict(cfg)
    ids = torch.randint(0, 49408, (1, 12), generator=torch.Generator().manual_seed(0))
    a = C.clip_text_forward(sd, ids, 12)
    ids2 = ids.clone(); ids2[0, 8] = 5
    b = C.clip_text_forward(sd, ids2, 12)
    assert float((a[0, :8] - b[0, :8]).abs().max()) == 0.0 and float((a[0, 8:] - b[0, 8:]).abs().max()) > 1e-3


def test_vae_decoder_module_and_oracle_shapes():
    """Host-side checks of the VAE decoder (its oracle is parity-unpinned, see oracle/vae_oracle.py): the module holds the
    decode-side keys of diffusers' AutoencoderKL, the oracle maps [N,4,h,w] -> [N,3,8h,8w], decode_latents returns the
    pipelines' uint8 layout, and there is no CPU fallback."""
    from lavie_b200.vae import VAEDecoder, vae_decoder_param_spec, vae_synthetic_state_dict
    from oracle import vae_oracle as V
    spec = vae_decoder_param_spec()
    assert len(spec) == 140 and spec["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"] == (256, 512, 1, 1)
    assert sum(__import__("math").prod(s) for s in spec.values()) == 49_490_199
    m = VAEDecoder()
    sd = vae_synthetic_state_dict()
    m.load_state_dict(sd, strict=True)
    z = torch.randn(1, 4, 4, 4, generator=torch.Generator().manual_seed(0))
    assert V.decode(sd, z).shape == (1, 3, 32, 32)
    v = V.decode_latents(sd, z.reshape(1, 4, 1, 4, 4))
    assert v.shape == (1, 1, 32, 32, 3) and v.dtype == torch.uint8
    with pytest.raises(RuntimeError):
        m.decode(z)
