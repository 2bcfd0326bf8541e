"""SURVEY 8f row N2 (video super-resolution UNet): the 1158-key parameter table and the CPU oracle are pinned against
golden vectors produced by the unmodified reference (tests/golden/make_golden_vsr.py).  CPU only."""
import math

import pytest
import torch

from conftest import load_golden, rel_l2


def test_vsr_param_table():
    from lavie_b200.config import VSR_CONFIG, param_spec
    spec = param_spec(VSR_CONFIG)
    assert len(spec) == 1158
    assert sum(math.prod(s) for s in spec.values()) == 691_036_692
    assert spec["conv_in.weight"] == (256, 7, 3, 3)
    assert spec["class_embedding.weight"] == (1000, 1024)
    assert spec["down_temporal_blocks.0.resblocks_3d_t.conv1.weight"] == (256, 256, 5, 1, 1)
    assert spec["down_blocks.1.attentions.0.resblock_temporal.conv1.weight"] == (512, 512, 3, 1, 1)
    # text-only first attention on the three high-resolution levels, true self-attention at 1/8 resolution
    assert spec["down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_k.weight"] == (512, 1024)
    assert spec["mid_block.attentions.0.transformer_blocks.0.attn1.to_k.weight"] == (1024, 1024)
    assert spec["down_blocks.3.attentions.0.transformer_blocks.0.attn1.to_k.weight"] == (1024, 1024)


@pytest.mark.parametrize("name", ["vsr_b2_f4_16x16", "vsr_b1_f3_24x8"])
def test_vsr_oracle_matches_reference_golden(name):
    from lavie_b200.config import VSR_CONFIG
    from lavie_b200.synthetic import synthetic_state_dict
    from oracle import vsr_oracle as V
    g = load_golden(name)
    sd = synthetic_state_dict(VSR_CONFIG, seed=g["weights_seed"])
    taps = {}
    out = V.unet_forward(sd, g["sample"], g["timestep"], g["low_res"], g["text"], g["class_labels"], taps=taps)
    assert out.shape == g["out"].shape
    assert rel_l2(taps["down0"], g["down0"]) < 2e-5
    assert rel_l2(taps["mid"], g["mid"]) < 2e-5
    assert rel_l2(out, g["out"]) < 2e-5


def test_vsr_new_paths_are_visible():
    """The reference zero-initialises shift_conv and attn_temporal.to_out (temporal_module.py:147, attention.py:520), which
    would hide the temporal modules; the synthetic weights do not.  Dropping either must move the output."""
    from lavie_b200.config import VSR_CONFIG
    from lavie_b200.synthetic import synthetic_state_dict
    from oracle import vsr_oracle as V
    g = load_golden("vsr_b1_f3_24x8")
    sd = synthetic_state_dict(VSR_CONFIG, seed=0)
    args = (g["sample"], g["timestep"], g["low_res"], g["text"], g["class_labels"])
    base = V.unet_forward(sd, *args)
    for pat in ("shift_conv.weight", "resblock_temporal.conv2.weight", "class_embedding.weight"):
        sd2 = {k: (torch.zeros_like(v) if k.endswith(pat) else v) for k, v in sd.items()}
        assert rel_l2(V.unet_forward(sd2, *args), base) > 1e-3, pat
    # frames interact only through the frame convs, the 5-D GroupNorms and the temporal attention: a change of the LAST
    # frame's low-res input must reach frame 0
    lr = g["low_res"].clone()
    lr[:, :, -1] += 1.0
    moved = V.unet_forward(sd, g["sample"], g["timestep"], lr, g["text"], g["class_labels"])
    assert float((moved[:, :, 0] - base[:, :, 0]).abs().max()) > 1e-4
