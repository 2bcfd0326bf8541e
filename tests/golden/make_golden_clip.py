"""Golden vectors that pin ``oracle/clip_oracle.py`` to ``transformers.CLIPTextModel`` -- the third-party class the
reference instantiates as its text encoder (base/pipelines/pipeline_videogen.py:103, sample.py loads it with
``CLIPTextModel.from_pretrained(sd_path, subfolder="text_encoder")``).  transformers is installed in the build image
(version printed below); run once:

    python tests/golden/make_golden_clip.py

Two cases on the deterministic synthetic weights of ``lavie_b200.clip.clip_synthetic_state_dict``: the full SD-1.4 text tower
(ViT-L/14 text: 12 layers, 768 wide, 12 heads, quick_gelu) and a 3-layer cut of the x4-upscaler's tower (1024 wide, 16
heads, erf GELU).  Stored: input ids and last_hidden_state.
"""
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import torch  # noqa: E402
import transformers  # noqa: E402
from transformers import CLIPTextConfig as HFConfig, CLIPTextModel  # noqa: E402

from lavie_b200.clip import CLIPTextConfig, SD14_TEXT, clip_param_spec, clip_synthetic_state_dict  # noqa: E402

CASES = {
    "clip_sd14_b2": (SD14_TEXT, 2, 77),
    "clip_vith3_b3": (CLIPTextConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=3,
                                     num_attention_heads=16, hidden_act="gelu"), 3, 77),
}


def main():
    print("transformers", transformers.__version__)
    for name, (cfg, b, L) in CASES.items():
        hf = CLIPTextModel(HFConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                                    intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_hidden_layers,
                                    num_attention_heads=cfg.num_attention_heads,
                                    max_position_embeddings=cfg.max_position_embeddings, hidden_act=cfg.hidden_act,
                                    layer_norm_eps=cfg.layer_norm_eps)).eval()
        sd = clip_synthetic_state_dict(cfg, seed=0)
        hf_keys = {k for k in hf.state_dict().keys() if "position_ids" not in k}
        assert hf_keys == set(clip_param_spec(cfg).keys()), hf_keys ^ set(clip_param_spec(cfg).keys())
        hf.load_state_dict(sd, strict=False)
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
        ids = torch.randint(0, cfg.vocab_size, (b, L), generator=g)
        ids[:, 0] = 49406
        ids[:, 20:] = 49407                                        # BOS ... EOS padding, like the tokenizer emits
        with torch.no_grad():
            out = hf(ids)[0]
        torch.save({"ids": ids, "out": out, "weights_seed": 0, "cfg": cfg.__dict__,
                    "transformers": transformers.__version__}, os.path.join(HERE, f"{name}.pt"))
        print(name, tuple(out.shape), float(out.std()))


if __name__ == "__main__":
    main()
