"""Drop-in replacement for the DDPM part of 06_tiny_stable_diffusion/utils.py:
`from utils import SamplerDDPM, TrainerDDPM, ...` (02_train_direct.py:7) resolves to the B200-native
classes.  `denormalize`, `EMA` and `CosineWarmupScheduler` (utils.py:14-18, 42-93) resolve to the B200-native versions too;
only `animal_faces_loader` (CPU DataLoader, utils.py:21-29) is re-exported from the reference's own file when it has been
kept beside this one as `utils_reference.py`."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

try:  # the data loader stays the reference's (CPU workers, PIL resize): the maintainer's renamed copy of utils.py
    from utils_reference import animal_faces_loader  # noqa: F401
except ImportError:  # pragma: no cover
    pass

from from_ddpm_to_stable_diffusion_b200.utils import SamplerDDPM, TrainerDDPM, extract  # noqa: E402,F401
from from_ddpm_to_stable_diffusion_b200.optim import FusedClipAdamW  # noqa: E402,F401
from from_ddpm_to_stable_diffusion_b200.training import (CosineWarmupScheduler, EMA, GraphedTrainStep, denormalize,  # noqa: E402,F401
                                                          generate_grid, image_grid_u8, load_training_state, means,
                                                          normalize_u8, stds, train_step, training_state)
