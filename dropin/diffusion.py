"""Drop-in replacement for 06_tiny_stable_diffusion/diffusion.py: `from diffusion import Diffusion`
(02_train_direct.py:8) resolves to the B200-native module.  Put this directory first on sys.path
(or copy the two files next to the training script)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from from_ddpm_to_stable_diffusion_b200.diffusion import Diffusion  # noqa: E402,F401
