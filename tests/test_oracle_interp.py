"""Groundwork for SURVEY 8f row N1 (frame-interpolation UNet): the parameter table and the CPU oracle are pinned
against golden vectors produced by the unmodified reference (tests/golden/make_golden_interp.py).  CPU only."""
import math

import pytest
import torch

from conftest import load_golden, rel_l2


def test_interp_param_table():
    from lavie_b200.config import BASE_CONFIG, INTERP_CONFIG, param_spec
    spec = param_spec(INTERP_CONFIG)
    assert len(spec) == 798
    assert sum(math.prod(s) for s in spec.values()) == 909_131_524
    assert spec["conv_in.weight"] == (320, 8, 3, 3)
    assert not any("rotary_emb" in k or "time_rel_pos_bias" in k for k in spec)
    base = param_spec(BASE_CONFIG)
    assert set(spec) < set(base) and len(base) - len(spec) == 32           # 16 blocks x (rotary freqs + rel-pos table)


@pytest.mark.parametrize("name", ["interp_b2_f7_8x8", "interp_b1_f5_16x8"])
def test_interp_oracle_matches_reference_golden(name):
    from lavie_b200.config import INTERP_CONFIG
    from lavie_b200.synthetic import synthetic_state_dict
    from oracle import interp_oracle as O
    g = load_golden(name)
    sd = synthetic_state_dict(INTERP_CONFIG, seed=g["weights_seed"])
    out = O.unet_forward(sd, g["sample"], g["timestep"], g["text"])
    assert out.shape == g["out"].shape
    assert rel_l2(out, g["out"]) < 2e-5


def test_sparse_causal_attention_sees_first_and_former_frame():
    """Changing frame 2 must move the self-attention output of frames 2 (residual path aside: its own queries) and 3
    (former frame), and of no later frame; changing frame 0 moves every frame."""
    from lavie_b200.config import INTERP_CONFIG
    from lavie_b200.synthetic import synthetic_state_dict
    from oracle import interp_oracle as O
    sd = synthetic_state_dict(INTERP_CONFIG, seed=0)
    p = "down_blocks.0.attentions.0.transformer_blocks.0.attn1"
    frames, hw, c = 6, 4, 320
    x = torch.randn(frames, hw, c, generator=torch.Generator().manual_seed(1))
    y = O.sparse_causal_attention(sd, p, x, frames)
    x2 = x.clone(); x2[2] += 1.0
    d = (O.sparse_causal_attention(sd, p, x2, frames) - y).abs().amax(dim=(1, 2))
    assert d[2] > 1e-4 and d[3] > 1e-4 and float(d[[0, 1, 4, 5]].max()) == 0.0
    x0 = x.clone(); x0[0] += 1.0
    d0 = (O.sparse_causal_attention(sd, p, x0, frames) - y).abs().amax(dim=(1, 2))
    assert bool((d0 > 1e-4).all())


def test_interp_ddim_loop_matches_reference_golden():
    """Respaced DDIM + CFG 4.0 with channel-concat conditioning (interpolation/sample.py:138-166 through the reference's
    own diffusion/ package and forward_with_cfg) against the oracle's restatement of the loop."""
    from lavie_b200.config import INTERP_CONFIG
    from lavie_b200.synthetic import synthetic_state_dict
    from oracle import interp_oracle as O
    g = load_golden("interp_loop_f5_8x8")
    sd = synthetic_state_dict(INTERP_CONFIG, seed=g["weights_seed"])
    out = O.ddim_loop(sd, g["z"], g["x_start"], g["text"], num_steps=g["steps"], cfg_scale=4.0)
    assert out.shape == g["out"].shape
    assert rel_l2(out, g["out"]) < 1e-4
    assert O.space_timesteps(1000, 50)[:3] == [0, 20, 41] and O.space_timesteps(1000, 50)[-1] == 999
