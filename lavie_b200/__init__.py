"""lavie_b200: B200-native (sm_100a) implementation of LaVie's per-step denoiser.

Public surface = the reference's module API for this path:

    from lavie_b200 import UNet3DConditionModel
    unet = UNet3DConditionModel().to("cuda"); unet.load_state_dict(torch.load("lavie_base.pt"))
    noise = unet(latents, t, encoder_hidden_states=prompt_embeds).sample
"""
from .config import BASE_CONFIG, UNetConfig, param_spec  # noqa: F401
from .unet import UNet3DConditionModel, UNet3DConditionOutput  # noqa: F401

__all__ = ["UNet3DConditionModel", "UNet3DConditionOutput", "UNetConfig", "BASE_CONFIG", "param_spec"]
