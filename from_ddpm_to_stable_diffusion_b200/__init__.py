"""B200-native (sm_100a) implementation of the tiny-Stable-Diffusion DDPM hot path.

Mirrors the reference's Python surface (06_tiny_stable_diffusion/diffusion.py, utils.py):
``Diffusion``, ``TrainerDDPM``, ``SamplerDDPM``, ``extract``.
"""
__all__ = ["Diffusion", "TrainerDDPM", "SamplerDDPM", "extract", "EMA", "CosineWarmupScheduler", "denormalize",
           "normalize_u8", "image_grid_u8", "train_step", "generate_grid", "means", "stds", "GraphedTrainStep", "training_state", "load_training_state"]


def __getattr__(name):
    if name == "Diffusion":
        from .diffusion import Diffusion
        return Diffusion
    if name in ("TrainerDDPM", "SamplerDDPM", "extract"):
        from . import utils
        return getattr(utils, name)
    if name in ("EMA", "CosineWarmupScheduler", "denormalize", "normalize_u8", "image_grid_u8", "train_step",
                "generate_grid", "means", "stds", "GraphedTrainStep", "training_state", "load_training_state"):
        from . import training
        return getattr(training, name)
    raise AttributeError(name)
