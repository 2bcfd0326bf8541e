"""Whole-denoiser parity on the B200, through the public module API and the C-ABI library underneath.

Tolerance (BASELINE.json north_star): noise-prediction relative L2 <= 2e-2 for the bf16 path against the
reference's fp32 forward (golden vectors produced by the reference itself, and the pinned CPU oracle)."""
import pytest
import torch

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def unet(synthetic_sd):
    from lavie_b200 import UNet3DConditionModel
    m = UNet3DConditionModel()
    m.load_state_dict(synthetic_sd, strict=True)
    return m.to("cuda").eval()


@pytest.mark.parametrize("name", ["b2_f16_8x8", "b1_f3_16x24", "b2_f2_8x8_tvec"])
def test_matches_reference_golden(unet, name):
    g = load_golden(name)
    out = unet(g["sample"].cuda(), g["timestep"], encoder_hidden_states=g["text"].cuda()).sample
    assert out.shape == g["out"].shape and out.dtype == torch.float32
    err = rel_l2(out.cpu(), g["out"])
    print(f"{name}: rel-L2 vs reference fp32 = {err:.3e}")
    assert err <= BF16_TOL


def test_matches_oracle_on_tma_geometry(unet, synthetic_sd):
    """A geometry where every level takes the TMA implicit-GEMM conv path (W = 64, 32, 16, 8)."""
    from lavie_b200.synthetic import synthetic_inputs
    from oracle import unet3d_oracle as O
    sample, t, text = synthetic_inputs(2, 4, 8, 64, seed=3)
    taps_o, taps_g = {}, {}
    ref = O.unet_forward(synthetic_sd, sample, t, text, taps=taps_o)
    out = unet(sample.cuda(), t, encoder_hidden_states=text.cuda(), taps=taps_g).sample
    for k in ("emb", "conv_in", "down0_res0", "down0_attn0", "mid"):
        print(f"tap {k}: rel-L2 = {rel_l2(taps_g[k].cpu(), taps_o[k]):.3e}")
    err = rel_l2(out.cpu(), ref)
    print(f"8x64 x4 frames: rel-L2 vs oracle = {err:.3e}")
    assert err <= BF16_TOL
    # the CUDA-graph path must give the same numbers as the eager launch sequence
    out_g = unet(sample.cuda(), t, encoder_hidden_states=text.cuda()).sample
    assert rel_l2(out_g.cpu(), out.cpu()) < 1e-3


def test_matches_oracle_on_ragged_geometry(unet, synthetic_sd):
    """Nothing divides nicely: 3 batch items, 5 frames, a 24x40 latent (levels 24x40, 12x20, 6x10, 3x5), 33 text
    tokens.  Exercises the M / N tails of the GEMM tiles, partial key tiles of both attentions, F < 16 in the temporal
    kernel (padded frames) and the image-edge handling of the im2col-mode TMA on odd maps."""
    from lavie_b200.synthetic import synthetic_inputs
    from oracle import unet3d_oracle as O
    sample, t, text = synthetic_inputs(3, 5, 24, 40, seed=11, text_len=33)
    ref = O.unet_forward(synthetic_sd, sample, t, text)
    out = unet(sample.cuda(), t, encoder_hidden_states=text.cuda()).sample
    err = rel_l2(out.cpu(), ref)
    print(f"3 x 5 frames x 24x40, 33 tokens: rel-L2 vs oracle = {err:.3e}")
    assert out.shape == ref.shape and err <= BF16_TOL


@pytest.mark.parametrize("frames,h,w", [(2, 8, 8), (16, 16, 24)])
def test_fifty_step_ddim_latent(unet, synthetic_sd, frames, h, w):
    """BASELINE north star: the FINAL latent of the 50-step DDIM + CFG 7.5 loop (pipeline_videogen.py:664-689) against
    the oracle's loop on the same start noise (the oracle's per-step update is pinned to the reference's own
    DDIMScheduler.step, tests/test_oracle.py::test_ddim_step_matches_reference_scheduler).  bf16 noise compounds over 50
    feed-back steps through a random-init network; stated tolerance 5e-2 (measured 7.4e-3 at 2 frames x 8x8 on B200).
    The second case is a full 16-frame video on a 16x24 latent (all four levels, temporal attention over 16 frames)."""
    from lavie_b200.pipeline import CFGDenoiser, DDIMSchedule
    from lavie_b200.synthetic import synthetic_inputs
    from oracle import unet3d_oracle as O
    torch.set_num_threads(max(1, __import__("os").cpu_count() or 1))
    sample, _, text = synthetic_inputs(2, frames, h, w, seed=21)
    lat0 = sample[:1]
    ref = O.cfg_ddim_loop(synthetic_sd, lat0, text, 7.5, 50)
    den = CFGDenoiser(unet, 7.5, DDIMSchedule(50))
    out = den.loop(lat0.cuda(), text.cuda()).cpu()
    err = rel_l2(out, ref)
    print(f"50-step DDIM latent: rel-L2 vs oracle loop = {err:.3e}")
    assert torch.isfinite(out).all() and err <= 5e-2


def test_full_size_matches_reference(unet, synthetic_sd):
    """The HEADLINE geometry of BASELINE.json (config 1/2: batch 2 x [4,16,40,64], 77 tokens) against the reference's
    fp32 CPU forward on the same inputs: the unmodified reference model from baseline/_ref when it travelled with the
    snapshot, else the pinned CPU oracle.  320-wide tiles over 320 row tiles, tail windows, the 20-tile 2560-key
    attention and the 148-chunk GroupNorm statistics all run together only here.  ~20 s of CPU on the GPU box."""
    import os
    from lavie_b200.synthetic import synthetic_inputs
    from oracle import reference_loader as R
    from oracle import unet3d_oracle as O
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sample, t, text = synthetic_inputs(2, 16, 40, 64, seed=0)
    if R.available("base"):
        ref_model = R.load_reference_unet("base", synthetic_sd)
        with torch.no_grad():
            ref = ref_model(sample, t, encoder_hidden_states=text).sample
        who = "unmodified reference (baseline/_ref)"
    else:
        ref = O.unet_forward(synthetic_sd, sample, t, text)
        who = "CPU oracle"
    out = unet(sample.cuda(), t, encoder_hidden_states=text.cuda()).sample.cpu()
    err = rel_l2(out, ref)
    print(f"[2,4,16,40,64]: rel-L2 vs {who} = {err:.3e}")
    assert out.shape == ref.shape and err <= BF16_TOL


def test_full_size_properties(unet):
    """BASELINE config 2 geometry [2,4,16,40,64]: too big for the CPU oracle inside a test, so check the
    size-independent properties: finite, deterministic, and the two CFG halves do not interact
    (every op of the UNet is per batch item)."""
    from lavie_b200.synthetic import synthetic_inputs
    sample, t, text = synthetic_inputs(2, 16, 40, 64, seed=0)
    s, e = sample.cuda(), text.cuda()
    both = unet(s, t, encoder_hidden_states=e).sample
    again = unet(s, t, encoder_hidden_states=e).sample
    assert torch.isfinite(both).all()
    assert rel_l2(again, both) < 1e-3
    # Alone, the cond half runs with half the rows, so tile shapes / split-K factors (and with them fp32 summation
    # order) differ; the bf16 network amplifies such rounding-level changes to ~1e-2 (same size as bf16-vs-fp32), so the
    # bound is the parity tolerance.  A real cross-talk bug (wrong batch stride, shared statistics) gives O(1).
    lone = unet(s[1:], t, encoder_hidden_states=e[1:]).sample
    err = rel_l2(lone, both[1:])
    print(f"CFG halves independent: rel-L2(alone, batched) = {err:.3e}")
    assert err < 2e-2
    assert 0.05 < float(both.std()) < 5.0


def test_api_contract(unet):
    from lavie_b200.synthetic import synthetic_inputs
    sample, t, text = synthetic_inputs(1, 2, 8, 8, seed=5)
    s, e = sample.cuda(), text.cuda()
    a = unet(s, t, e).sample
    b = unet(s, torch.tensor(t), encoder_hidden_states=e, return_dict=False)[0]          # 0-d tensor, tuple return
    c = unet(s, torch.tensor([float(t)], device="cuda"), encoder_hidden_states=e).sample  # [1] float tensor on device
    assert torch.equal(a, b) and torch.equal(a, c)
    h = unet(s.half(), t, encoder_hidden_states=e.half()).sample
    assert h.dtype == torch.float16
    with pytest.raises(ValueError):
        unet(s[:, :, :, :7], t, encoder_hidden_states=e)
    assert unet.config.in_channels == 4 and unet.config.sample_size == 64
    unet.set_use_memory_efficient_attention_xformers(True)
    unet.set_attention_slice(1)


def test_text_projections_are_cached_per_prompt_tensor(synthetic_sd):
    """The K/V projections of the text run once per prompt tensor, outside the captured step: the same tensor object,
    unmodified, skips the GEMM (one launch less); an in-place edit (version counter) or another tensor recomputes."""
    from lavie_b200 import UNet3DConditionModel
    from lavie_b200.synthetic import synthetic_inputs
    unet = UNet3DConditionModel()
    unet.load_state_dict(synthetic_sd, strict=True)
    unet = unet.to("cuda").eval()
    sample, t, text = synthetic_inputs(2, 2, 8, 8, seed=5)
    sample, text = sample.to("cuda"), text.to("cuda")
    a = unet(sample, t, encoder_hidden_states=text).sample
    n_first = unet.launches_per_step()
    b = unet(sample, t, encoder_hidden_states=text).sample
    assert unet.launches_per_step() == n_first - 1 and torch.equal(a, b)
    c = unet(sample, t, encoder_hidden_states=text.clone()).sample          # another object, same values
    assert unet.launches_per_step() == n_first and torch.equal(a, c)
    text2 = text.clone()
    unet(sample, t, encoder_hidden_states=text2)
    text2.mul_(0.5)                                                          # in-place edit of the cached prompt
    d = unet(sample, t, encoder_hidden_states=text2).sample
    assert unet.launches_per_step() == n_first and not torch.equal(a, d)
    ref = unet(sample, t, encoder_hidden_states=(text * 0.5)).sample
    assert torch.equal(d, ref)
