"""Runs the reference's training / sampling glue (06_tiny_stable_diffusion/02_train_direct.py:7-8, 13-28, 52-74)
against the drop-in modules, imported by BARE module name from the working directory exactly like the reference's
scripts do.  Launched by tests/test_dropin_gpu.py with cwd = dropin/."""
import sys

import numpy as np
import torch
from utils import SamplerDDPM, TrainerDDPM, CosineWarmupScheduler, denormalize  # noqa: E402  (02_train_direct.py:7)
from diffusion import Diffusion  # noqa: E402  (02_train_direct.py:8)

assert Diffusion.__module__.startswith("from_ddpm_to_stable_diffusion_b200"), Diffusion.__module__
config = dict(img_channel=3, channel_multy=[1, 2, 2, 2], channel_base=128, num_class=3, dropout=0.1, beta_1=0.0015,
              beta_T=0.0195, T=1000, w=1.8, lr=2e-6, grad_clip=1.0, train_rand=0.05, img_size=32, epoch=70)
device = torch.device("cuda:0")
torch.manual_seed(0)
# 02_train_direct.py:33-38,52-57
diffusion = Diffusion(channel_img=config['img_channel'], channel_base=config['channel_base'],
                      num_class=config['num_class'], channel_multy=config['channel_multy'],
                      dropout=config['dropout']).to(device)
optimizer = torch.optim.AdamW(diffusion.parameters(), lr=config['lr'], weight_decay=1e-5)
scheduler = CosineWarmupScheduler(optimizer=optimizer, warmup_epochs=config['epoch'] // 7, max_lr=1e-4,
                                  total_epochs=config['epoch'])
trainer = TrainerDDPM(diffusion, config['beta_1'], config['beta_T'], config['T']).to(device)
g = torch.Generator().manual_seed(3)
losses = []
for it in range(3):  # 02_train_direct.py:64-74, verbatim step body
    images = torch.randn(4, 3, config['img_size'], config['img_size'], generator=g)
    labels = torch.randint(0, config['num_class'], (4,), generator=g)
    optimizer.zero_grad()
    bs = images.shape[0]
    x_0 = images.to(device)
    labels = labels.to(device) + 1
    if np.random.rand() < config['train_rand']:
        labels = torch.zeros_like(labels).to(device)
    loss = trainer(x_0, labels).sum() / bs ** 2.
    loss.backward()
    torch.nn.utils.clip_grad_norm_(diffusion.parameters(), config['grad_clip'])
    optimizer.step()
    losses.append(loss.item())
scheduler.step()
assert all(np.isfinite(losses)), losses
assert len(diffusion.state_dict()) == 425
# 02_train_direct.py:13-23: generate()
nrow = 2
diffusion.eval()
sampler = SamplerDDPM(diffusion, config['beta_1'], config['beta_T'], 8, w=config['w']).to(device)  # short schedule
values = torch.arange(1, config['num_class'] + 1)
labels = values.repeat_interleave(nrow).to(device)
x_i = torch.randn(size=[config['num_class'] * nrow, config['img_channel'], config['img_size'], config['img_size']],
                  device=device)
with torch.no_grad():
    x0 = sampler(x_i, labels)
img = denormalize(x0)
assert x0.shape == x_i.shape and float(x0.abs().max()) <= 1.0 and torch.isfinite(img).all()
print("DROPIN_OK", losses)
sys.exit(0)
