"""Generate tests/golden/callers.pt from the UNMODIFIED reference classes (imported from /root/reference, never
copied) and torchvision itself.  Run in the build container only:  python oracle/make_golden_callers.py"""
import os
import sys

import torch
from torchvision import transforms
from torchvision.utils import make_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/06_tiny_stable_diffusion"
sys.path.insert(0, REF)
import utils as ref_utils  # noqa: E402  (the reference's module)

OUT = os.path.join(ROOT, "tests", "golden", "callers.pt")


def main():
    g = torch.Generator().manual_seed(77)
    fx = {}
    # --- loader transform after the resize: ToTensor + Normalize (utils.py:21-25)
    u8 = torch.randint(0, 256, (5, 16, 12, 3), generator=g, dtype=torch.uint8)
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=ref_utils.means, std=ref_utils.stds)])
    fx["u8"] = u8
    fx["normalized"] = torch.stack([tf(img.numpy()) for img in u8])
    # --- generate(): save_image(denormalize(x), nrow, padding) up to the PNG encoder (02_train_direct.py:24-27)
    grids = []
    for (n, h, w, nrow, pad) in [(21, 8, 8, 7, 0), (5, 6, 10, 3, 2), (1, 4, 4, 8, 2)]:
        x = torch.randn(n, 3, h, w, generator=g) * 1.3
        grid = make_grid(ref_utils.denormalize(x), nrow=nrow, padding=pad)
        u = grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8)  # torchvision save_image
        grids.append({"x": x, "nrow": nrow, "padding": pad, "grid_u8": u})
    fx["grids"] = grids
    # --- EMA (utils.py:42-72) on a toy module
    torch.manual_seed(5)
    lin = torch.nn.Linear(7, 5)
    ema = ref_utils.EMA(lin, 0.999)
    w0 = {k: v.clone() for k, v in lin.state_dict().items()}
    steps = []
    for i in range(3):
        with torch.no_grad():
            for p in lin.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
        ema.update()
        steps.append({"params": {k: v.clone() for k, v in lin.state_dict().items()},
                      "shadow": {k: v.clone() for k, v in ema.shadow.items()}})
    fx["ema"] = {"decay": 0.999, "init": w0, "steps": steps}
    # --- CosineWarmupScheduler stepped once per epoch (02_train_direct.py:52-56, 83; tiny_sd_direct.yml)
    scheds = []
    for (base, mx, epochs) in [(2.0e-6, 1.0e-4, 70), (1.0e-5, 3.0e-4, 21)]:
        opt = torch.optim.AdamW(lin.parameters(), lr=base, weight_decay=1e-5)
        sch = ref_utils.CosineWarmupScheduler(optimizer=opt, warmup_epochs=epochs // 7, max_lr=mx, total_epochs=epochs)
        lrs = []
        for _ in range(epochs):
            lrs.append(opt.param_groups[0]["lr"])
            opt.step()
            sch.step()
        scheds.append({"base_lr": base, "max_lr": mx, "epochs": epochs, "warmup": epochs // 7, "lrs": lrs})
    fx["lr"] = scheds
    torch.save(fx, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
