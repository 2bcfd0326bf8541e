"""Name-only stand-in for diffusers.models.attention_processor (imported by vsr/models, unused by the VSR UNet)."""


class Attention:
    pass
