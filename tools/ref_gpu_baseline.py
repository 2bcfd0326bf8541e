"""Library baseline (VERDICT r1 missing #8): the UNMODIFIED reference UNet3DConditionModel (baseline/_ref + the stand-ins
for its two un-vendored dependencies) run end to end on the same B200 through stock PyTorch (cuDNN / cuBLAS, the
reference's own baddbmm + softmax attention), at the headline shape [2,4,16,40,64] with the same synthetic weights.
Not the product path and not a bench.py arm: it answers "what does the reference itself do on this GPU".

    python tools/ref_gpu_baseline.py [--variant base|interp] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lavie_b200.config import BASE_CONFIG, INTERP_CONFIG  # noqa: E402
from lavie_b200.synthetic import synthetic_state_dict  # noqa: E402


def load_reference_unet(variant, state_dict):
    """The unmodified reference model from baseline/_ref (copied there by __graft_entry__.build()) with the stand-ins for
    its two un-vendored dependencies -- a local copy of what bench.py's reference arm does, so that this tool does not
    import the test-only oracle package."""
    import importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tree = "base" if variant == "base" else "interpolation"
    for pth in (os.path.join(root, "tests", "golden", "shims"), os.path.join(root, "baseline", "_ref", tree)):
        sys.path.insert(0, pth)
    mod = importlib.import_module("models.unet")
    cfg = (BASE_CONFIG if variant == "base" else INTERP_CONFIG).to_dict()
    if variant != "base":
        cfg["use_first_frame"] = True
    ref = mod.UNet3DConditionModel.from_config(cfg).eval()
    ref.load_state_dict(state_dict, strict=True)
    return ref


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="base")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--frames", type=int, default=0)
    args = ap.parse_args()
    cfg = BASE_CONFIG if args.variant == "base" else INTERP_CONFIG
    sd = synthetic_state_dict(cfg, seed=0)
    ref = load_reference_unet(args.variant, sd)
    frames = args.frames or (16 if args.variant == "base" else 61)
    g = torch.Generator().manual_seed(3)
    sample = torch.randn(2, cfg.in_channels, frames, 40, 64, generator=g)
    text = torch.randn(2, 77, 768, generator=g)
    res = {"what": f"unmodified reference {args.variant} UNet, stock PyTorch {torch.__version__} on "
                   f"{torch.cuda.get_device_name(0)}", "shape": list(sample.shape), "steps": args.steps}
    out32 = None
    for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16), ("fp32", torch.float32)):
        try:
            m = ref.to(device="cuda", dtype=dt)
            x, e = sample.to("cuda", dt), text.to("cuda", dt)
            with torch.no_grad():
                for _ in range(2):
                    out = m(x, 500, encoder_hidden_states=e).sample
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    out = m(x, 500, encoder_hidden_states=e).sample
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            res[name] = {"ms_per_step": ms, "steps_per_s": 1000.0 / ms,
                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
            if dt == torch.float32:
                out32 = out.float().cpu()
            else:
                res[name]["_out"] = out.float().cpu()
        except Exception as ex:  # e.g. out of memory for the materialised score matrices
            res[name] = {"error": repr(ex)[:200]}
        torch.cuda.empty_cache()
    for name in ("fp16", "bf16"):
        o = res.get(name, {}).pop("_out", None)
        if o is not None and out32 is not None:
            res[name]["rel_l2_vs_its_own_fp32"] = float((o.double() - out32.double()).norm() / out32.double().norm())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
