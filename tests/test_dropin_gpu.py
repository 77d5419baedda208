"""The drop-in boundary the way the reference uses it: bare ``from diffusion import ...`` / ``from utils import ...``
from the script's working directory (06_tiny_stable_diffusion/02_train_direct.py:7-8)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_reference_script_body_runs_on_dropin(cuda):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ROOT, "dropin")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "dropin_script.py")],
                       cwd=os.path.join(ROOT, "dropin"), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_dropin_modules_export_the_reference_names():
    """CPU: the two stub modules resolve and expose the names 02_train_direct.py:7-8 imports."""
    code = ("import diffusion, utils\n"
            "assert diffusion.Diffusion.__module__ == 'from_ddpm_to_stable_diffusion_b200.diffusion'\n"
            "for n in ('SamplerDDPM', 'TrainerDDPM', 'CosineWarmupScheduler', 'denormalize', 'extract', 'EMA'):\n"
            "    assert hasattr(utils, n), n\n"
            "m = diffusion.Diffusion(channel_img=3, channel_multy=[1, 2, 2, 2], num_class=3)\n"
            "assert len(m.state_dict()) == 425\n"
            "t = utils.TrainerDDPM(m, 0.0015, 0.0195, 1000)\n"
            "assert t.sqrt_alphas_bar.dtype.is_floating_point and t.sqrt_alphas_bar.shape[0] == 1000\n"
            "print('OK')\n")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ROOT, "dropin")
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.join(ROOT, "dropin"), env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
