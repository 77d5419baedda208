"""Drop-in replacement for the DDPM part of 06_tiny_stable_diffusion/utils.py:
`from utils import SamplerDDPM, TrainerDDPM, ...` (02_train_direct.py:7) resolves to the B200-native
classes.  The reference's data / LR glue (animal_faces_loader, denormalize, EMA, CosineWarmupScheduler,
utils.py:10-93) is out of scope for this path and is re-exported unchanged from the reference's own file
when it has been kept beside this one as `utils_reference.py`."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

try:  # the maintainer's renamed copy of the original utils.py, if present
    from utils_reference import (CosineWarmupScheduler, EMA, animal_faces_loader, denormalize, means, stds)  # noqa: F401
except ImportError:  # pragma: no cover
    pass

from from_ddpm_to_stable_diffusion_b200.utils import SamplerDDPM, TrainerDDPM, extract  # noqa: E402,F401
from from_ddpm_to_stable_diffusion_b200.optim import FusedClipAdamW  # noqa: E402,F401
