"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the host module keeps the
reference's state_dict contract, weight repacking is lossless, and the product path refuses to run without CUDA."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lavie_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lavie_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from lavie_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lavie_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.lavie_abi_version() == 1
    # pure host queries (no GPU work)
    assert lib.lavie_conv3x3_supported(40, 64, 320) == 1
    assert lib.lavie_conv3x3_supported(16, 24, 320) == 1 and lib.lavie_conv3x3_supported(16, 24, 4) == 0
    assert lib.lavie_groupnorm_chunks(2, 16 * 40 * 64) == 148 and lib.lavie_groupnorm_chunks(32, 40) == 3


def test_gemm_planner_decisions():
    """Host-only: tile width / split-K / tail-window choices for the shapes of the full-size step (148 SMs = 74 CTA
    pairs).  These pin the reasoning of DESIGN.md section 4: wide tiles where L2 feed bounds the conv, split-K where a
    level has fewer tiles than CTA pairs, a tail window where the last wave would be nearly empty."""
    import ctypes
    from lavie_b200 import _lib
    lib = _lib.load()

    def plan(M, N, K, conv=0, geglu=0, ws=128 << 20):
        bn, s, tail, ts = (ctypes.c_int() for _ in range(4))
        assert lib.lavie_gemm_plan(M, N, K, conv, geglu, ws, bn, s, tail, ts) == 0
        return bn.value, s.value, tail.value, ts.value

    assert plan(81920, 320, 2880, conv=1) == (320, 1, 0, 1)          # 40x64 resnet conv: one 320-wide tile per row tile
    assert plan(81920, 2560, 320, geglu=1)[0] == 256                  # GEGLU pairs value / gate inside a 256-wide tile
    bn, s, tail, ts = plan(5120, 1280, 11520, conv=1)                 # 10x16 level: 80 tiles on 74 pairs
    assert bn == 320 and s == 1 and tail == 2 and ts > 1
    bn, s, tail, ts = plan(1280, 1280, 11520, conv=1)                 # 5x8 level: 20 tiles -> split-K fills the machine
    assert s >= 2 and tail == 0
    assert plan(1280, 1280, 11520, conv=1, ws=0)[1] == 1              # no workspace, no split-K
    assert lib.lavie_gemm_plan(0, 320, 320, 0, 0, 0, None, None, None, None) != 0


def test_epilogue_struct_matches_header():
    import ctypes
    from lavie_b200._lib import Epilogue
    text = open(os.path.join(ROOT, "include", "lavie_b200.h")).read()
    body = text[text.index("typedef struct {"):text.index("} lavie_epilogue;")]
    fields = re.findall(r"\b(\w+);", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert [f for f, _ in Epilogue._fields_] == fields
    assert ctypes.sizeof(Epilogue) == 48      # 4 pointers + 4 ints, natural alignment


def test_param_table_is_the_reference_contract():
    from lavie_b200.config import param_spec
    spec = param_spec()
    assert len(spec) == 830
    n = 0
    for shape in spec.values():
        k = 1
        for s in shape:
            k *= s
        n += k
    assert n == 909_124_356
    assert spec["conv_in.weight"] == (320, 4, 3, 3)
    assert spec["up_blocks.0.resnets.0.conv1.weight"] == (1280, 2560, 3, 3)
    assert spec["up_blocks.3.resnets.2.conv_shortcut.weight"] == (320, 640, 1, 1)
    assert spec["mid_block.attentions.0.transformer_blocks.0.attn2.to_k.weight"] == (1280, 768)
    assert spec["down_blocks.0.attentions.0.transformer_blocks.0.ff.net.0.proj.weight"] == (2560, 320)
    assert "down_blocks.3.downsamplers.0.conv.weight" not in spec and "up_blocks.3.upsamplers.0.conv.weight" not in spec


def test_module_state_dict_roundtrip(synthetic_sd):
    from lavie_b200 import UNet3DConditionModel
    m = UNet3DConditionModel()
    assert set(m.state_dict().keys()) == set(synthetic_sd.keys())
    missing, unexpected = m.load_state_dict(synthetic_sd, strict=True)
    assert not missing and not unexpected
    back = m.state_dict()
    for k in ("conv_in.weight", "mid_block.attentions.0.transformer_blocks.0.attn_temp.rotary_emb.freqs",
              "up_blocks.2.attentions.1.transformer_blocks.0.ff.net.2.bias"):
        assert torch.equal(back[k], synthetic_sd[k])
    # PEFT-style probing finds real nn.Linear projections (fine_tuning.py:296-308)
    lin = [n for n, mod in m.named_modules() if isinstance(mod, torch.nn.Linear) and n.endswith(("to_q", "to_out.0"))]
    assert len(lin) == 16 * 3 * 2
    assert m.config.in_channels == 4 and m.config["sample_size"] == 64 and "_diffusers_version" in m.config
    # no CUDA here: the product path must fail loudly, not fall back to PyTorch/CPU math
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            m(torch.zeros(1, 4, 2, 8, 8), 1, encoder_hidden_states=torch.zeros(1, 77, 768))


def test_packing_helpers():
    from lavie_b200 import packing as P
    w = torch.randn(16, 8, 3, 3)
    pk = P.pack_conv3x3(w).float()
    assert pk.shape == (16, 72)
    assert torch.equal(pk[:, 3 * 8:4 * 8], w[:, :, 1, 0].to(torch.bfloat16).float())   # tap (kh=1,kw=0) slab
    q = torch.randn(8 * 40, 320)
    qp = P.pad_heads(q, 8)
    assert qp.shape == (8 * 48, 320) and P.head_pitch(40) == 48 and P.head_pitch(80) == 80
    assert torch.equal(qp.reshape(8, 48, 320)[:, :40], q.reshape(8, 40, 320)) and qp.reshape(8, 48, 320)[:, 40:].abs().max() == 0
    w8 = torch.randn(2560, 320)
    b8 = torch.randn(2560)
    wi, bi = P.interleave_geglu(w8, b8)
    assert torch.equal(wi[0:128], w8[0:128]) and torch.equal(wi[128:256], w8[1280:1408]) and torch.equal(wi[256:384], w8[128:256])
    assert torch.equal(bi[128:256], b8[1280:1408])
    # rel-pos bias and RoPE tables agree with the oracle's restatement of the reference
    from oracle import unet3d_oracle as O
    emb = torch.randn(32, 8)
    assert torch.equal(P.rel_pos_bias_table(emb, 16), O.rel_pos_bias(emb, 16))
    freqs = 1.0 / (10000.0 ** (torch.arange(0, 32, 2).float() / 32))
    tab = P.rope_table(freqs, 16)
    x = torch.randn(16, 40)
    y = O.rope(x, freqs)
    x0, x1 = x[:, 0:32:2], x[:, 1:32:2]
    assert torch.allclose(y[:, 0:32:2], x0 * tab[..., 0] - x1 * tab[..., 1], atol=1e-6)
    assert torch.allclose(y[:, 1:32:2], x1 * tab[..., 0] + x0 * tab[..., 1], atol=1e-6)


def test_ddim_schedule_matches_oracle():
    from lavie_b200.pipeline import DDIMSchedule
    from oracle import unet3d_oracle as O
    s = DDIMSchedule(50)
    acp, ts, ratio = O.ddim_schedule(50)
    assert s.timesteps == ts.tolist() and s.ratio == ratio
    assert torch.equal(s.alphas_cumprod, acp)
    assert s.alphas(1) == (float(acp[1]), float(acp[0]))


def test_vsr_module_keeps_the_reference_state_dict_layout():
    """UNet3DVSRModel holds exactly the 1158 keys of the reference VSR UNet (shapes from lavie_b200.config, proven against
    the reference in tests/golden/make_golden_vsr.py) and refuses to run without CUDA."""
    import pytest
    import torch
    from lavie_b200.config import VSR_CONFIG, param_spec
    from lavie_b200.vsr import UNet3DVSRModel, pack_frame_conv
    m = UNet3DVSRModel()
    spec = param_spec(VSR_CONFIG)
    sd = m.state_dict()
    assert list(sd.keys()) == list(spec.keys()) or set(sd.keys()) == set(spec.keys())
    assert all(tuple(sd[k].shape) == tuple(v) for k, v in spec.items())
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 2, 8, 8), 1, torch.zeros(1, 3, 2, 8, 8), encoder_hidden_states=torch.zeros(1, 4, 1024))
    # frame-conv weight packing: K ordered (tap, cin)
    w = torch.arange(2 * 3 * 5, dtype=torch.float32).reshape(2, 3, 5, 1, 1)
    p = pack_frame_conv(w)
    assert p.shape == (2, 15) and float(p[1, 2 * 3 + 1]) == float(w[1, 1, 2, 0, 0])


def test_upsample_conv_weight_packing_is_the_subpixel_identity():
    """packing.pack_upsample_conv3x3: a 3x3 conv on the nearest-2x upsampled map == four 2x2 phase convs on the source map
    with summed taps (what lavie_upsample_conv3x3_bf16 computes), borders included.  CPU, fp32."""
    import torch
    import torch.nn.functional as F
    from lavie_b200.packing import pack_upsample_conv3x3
    g = torch.Generator().manual_seed(0)
    co, ci, H, W = 5, 3, 4, 6
    w = torch.randn(co, ci, 3, 3, generator=g)
    x = torch.randn(2, ci, H, W, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    wp = pack_upsample_conv3x3(w, dtype=None).reshape(2, 2, co, 2, 2, ci)        # [py, px, n, a, b, c]
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for py in range(2):
        for px in range(2):
            acc = torch.zeros(2, co, H, W)
            for a in range(2):
                for b in range(2):                                              # source pixel (y + py - 1 + a, x + px - 1 + b)
                    acc += torch.einsum("nchw,oc->nohw", xp[:, :, py + a:py + a + H, px + b:px + b + W], wp[py, px, :, a, b, :])
            out[:, :, py::2, px::2] = acc
    assert float((out - ref).abs().max()) < 1e-5
