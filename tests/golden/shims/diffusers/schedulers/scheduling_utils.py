"""Stand-in for diffusers.schedulers.scheduling_utils (names only; the vendored reference scheduler
vsr/diffusion/scheduling_ddim.py needs them at import time)."""
from enum import Enum


class KarrasDiffusionSchedulers(Enum):
    DDIMScheduler = 1
    DDPMScheduler = 2
    EulerDiscreteScheduler = 7


class SchedulerMixin:
    config_name = "scheduler_config.json"
