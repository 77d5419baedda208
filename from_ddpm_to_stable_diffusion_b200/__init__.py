"""B200-native (sm_100a) implementation of the tiny-Stable-Diffusion DDPM hot path.

Mirrors the reference's Python surface (06_tiny_stable_diffusion/diffusion.py, utils.py):
``Diffusion``, ``TrainerDDPM``, ``SamplerDDPM``, ``extract``.
"""
__all__ = ["Diffusion", "TrainerDDPM", "SamplerDDPM", "extract"]


def __getattr__(name):
    if name == "Diffusion":
        from .diffusion import Diffusion
        return Diffusion
    if name in ("TrainerDDPM", "SamplerDDPM", "extract"):
        from . import utils
        return getattr(utils, name)
    raise AttributeError(name)
