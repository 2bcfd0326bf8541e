"""Generate the golden vectors that pin ``oracle/interp_oracle.py`` to the reference.

Run ONLY inside the build container (needs /root/reference, read-only):

    python tests/golden/make_golden_interp.py

Imports the UNMODIFIED interpolation model from /root/reference/interpolation/models with the same stand-ins as
make_golden.py (tests/golden/shims/), builds it with the configuration ``from_pretrained_2d`` would produce for
``copy_no_mask`` (in_channels 8, use_first_frame True; interpolation/models/unet.py:487-507), loads the deterministic
synthetic weights with ``load_state_dict(strict=True)`` (which proves the 798-key table of
``lavie_b200.config.param_spec(INTERP_CONFIG)``), runs the forward on CPU in fp32 and stores inputs and output.
"""
import os
import sys
import time
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference/interpolation")

import torch  # noqa: E402

from lavie_b200.config import INTERP_CONFIG, param_spec  # noqa: E402
from lavie_b200.synthetic import synthetic_state_dict  # noqa: E402
from models.unet import UNet3DConditionModel  # noqa: E402  (the reference's interpolation UNet)

CASES = {
    # name: (batch, frames, height, width, timestep, text tokens)
    "interp_b2_f7_8x8": (2, 7, 8, 8, 500, 77),
    "interp_b1_f5_16x8": (1, 5, 16, 8, 37, 20),
}


def main():
    torch.set_num_threads(os.cpu_count())
    cfg = INTERP_CONFIG.to_dict()
    cfg["use_first_frame"] = True
    ref = UNet3DConditionModel.from_config(cfg).eval()
    sd = synthetic_state_dict(INTERP_CONFIG, seed=0)
    assert set(ref.state_dict().keys()) == set(param_spec(INTERP_CONFIG).keys())
    ref.load_state_dict(sd, strict=True)
    for name, (b, f, h, w, t, ntok) in CASES.items():
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
        sample = torch.randn(b, 8, f, h, w, generator=g)
        text = torch.randn(b, ntok, 768, generator=g)
        t0 = time.time()
        with torch.no_grad():
            out = ref(sample, t, encoder_hidden_states=text).sample
        torch.save({"sample": sample, "timestep": t, "text": text, "out": out, "weights_seed": 0,
                    "shape": (b, f, h, w)}, os.path.join(HERE, f"{name}.pt"))
        print(f"{name}: out {tuple(out.shape)} std {out.std():.4f} in {time.time() - t0:.1f}s")
    # the caller: respaced DDIM + classifier-free guidance with channel-concat conditioning, exactly as
    # interpolation/sample.py:138-166 drives it (copy_no_mask, use_concat, eta 0, clip_denoised False, cfg_scale 4.0)
    from diffusion import create_diffusion  # noqa: E402  (the reference's IDDPM code)
    steps, f, h, w, ntok = 6, 5, 8, 8, 12
    g = torch.Generator().manual_seed(zlib.crc32(b"interp_loop"))
    z = torch.randn(1, 4, f, h, w, generator=g)
    copied = torch.randn(1, 4, f, h, w, generator=g)
    text = torch.randn(2, ntok, 768, generator=g)                  # [prompt, negative prompt]
    z2, x_start = torch.cat([z] * 2), torch.cat([copied] * 2)
    diffusion = create_diffusion(str(steps))
    t0 = time.time()
    samples = diffusion.ddim_sample_loop(ref.forward_with_cfg, z2.shape, z2, clip_denoised=False,
                                         model_kwargs=dict(encoder_hidden_states=text, class_labels=None), progress=False,
                                         device="cpu", mask=None, x_start=x_start, use_concat=True, copy_no_mask=True)
    torch.save({"z": z2, "x_start": x_start, "text": text, "steps": steps, "out": samples, "weights_seed": 0},
               os.path.join(HERE, "interp_loop_f5_8x8.pt"))
    print(f"interp_loop_f5_8x8: {steps} DDIM steps, out std {samples.std():.4f} in {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
