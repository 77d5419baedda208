"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference, never
copied) on seeded inputs.  Run in the build container only:  python oracle/make_golden.py

The weights are not stored (124 MB): they are regenerated on any machine by
oracle.ref_unet.init_state_dict(seed, ...), and each fixture carries a digest of them.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/06_tiny_stable_diffusion"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import ref_unet as R  # noqa: E402
import diffusion as ref_diffusion  # noqa: E402  (the reference's module)
import utils as ref_utils  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
MULTY = [1, 2, 2, 2]
B1, BT, T = 0.0015, 0.0195, 1000


def ref_model(channel_img, num_class, seed):
    sd = R.init_state_dict(seed, channel_img, MULTY, 128, num_class)
    m = ref_diffusion.Diffusion(channel_img=channel_img, channel_multy=MULTY, channel_base=128, num_class=num_class,
                                dropout=0.1)
    m.load_state_dict(sd)
    return m.eval(), sd


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())

    # ---- state_dict contract
    m, sd = ref_model(3, 3, 0)
    keys = [(k, list(v.shape)) for k, v in m.state_dict().items()]
    json.dump(keys, open(os.path.join(OUT, "state_dict_keys.json"), "w"))

    # ---- forward, 3x64x64 (tiny_sd_direct.yml config)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 64, 64, generator=g)
    t = torch.tensor([7, 940])
    y = torch.tensor([2, 0])
    with torch.no_grad():
        eps = m(x, t, y)
    torch.save({"cfg": dict(channel_img=3, num_class=3, seed=0), "digest": R.state_dict_digest(sd), "x": x, "t": t,
                "y": y, "eps": eps}, os.path.join(OUT, "fwd_3x64.pt"))

    # ---- forward, latent 4x16x16 (03_train_with_vae.py:36-37: channel_img=4, default num_class=10)
    ml, sdl = ref_model(4, 10, 1)
    xl = torch.randn(4, 4, 16, 16, generator=g)
    tl = torch.tensor([0, 1, 500, 999])
    yl = torch.tensor([0, 10, 3, 7])
    with torch.no_grad():
        epsl = ml(xl, tl, yl)
    torch.save({"cfg": dict(channel_img=4, num_class=10, seed=1), "digest": R.state_dict_digest(sdl), "x": xl, "t": tl,
                "y": yl, "eps": epsl}, os.path.join(OUT, "fwd_4x16.pt"))

    # ---- trainer: the reference draws t and noise itself; replaying the same RNG calls recovers them
    trainer = ref_utils.TrainerDDPM(m, B1, BT, T)
    x0 = torch.randn(2, 3, 32, 32, generator=g)
    lab = torch.tensor([1, 3])
    torch.manual_seed(4321)
    m.zero_grad()
    loss = trainer(x0, lab)  # eval mode => dropout off, deterministic
    (loss.sum() / 2 ** 2.).backward()
    torch.manual_seed(4321)
    t_rep = torch.randint(T, size=(2,))
    noise_rep = torch.randn_like(x0)
    grads = {k: p.grad.clone() for k, p in m.named_parameters()}
    keep = ["tail.2.weight", "tail.2.bias", "time_embedding.mlp.0.bias", "label_embedding.0.weight",
            "encoders.1.0.linear_time.1.bias", "bottleneck.1.atten_1.1.out_proj.bias", "decoders.7.0.conv_1.0.weight",
            "encoders.0.0.weight", "decoders.3.2.conv.bias", "bottleneck.1.atten_2.v_proj.weight"]
    torch.save({"x0": x0, "labels": lab, "t": t_rep, "noise": noise_rep, "loss": loss.detach(),
                "grad_norms": {k: float(v.norm()) for k, v in grads.items()},
                "grads": {k: grads[k] for k in keep}}, os.path.join(OUT, "trainer_3x32.pt"))

    # ---- sampler: the reference's full loop on a 4-step schedule (t=0 branch + final clip included), CFG w=1.8
    w = 1.8
    sampler4 = ref_utils.SamplerDDPM(m, B1, BT, 4, w=w)
    xT = torch.randn(2, 3, 32, 32, generator=g)
    labs = torch.tensor([3, 1])
    torch.manual_seed(99)
    with torch.no_grad():
        x0_ref = sampler4(xT, labs)
    torch.manual_seed(99)
    zs = [torch.randn_like(xT) for _ in range(3)]  # steps 3, 2, 1 draw noise; step 0 does not
    # single steps of the T=1000 schedule through the reference's own p_mean_variance
    sampler = ref_utils.SamplerDDPM(m, B1, BT, T, w=w)
    singles = []
    for ts in (999, 400, 1, 0):
        xt = torch.randn(2, 3, 32, 32, generator=g) * 2.0
        z = torch.randn(2, 3, 32, 32, generator=g)
        tt = xt.new_ones([2], dtype=torch.long) * ts
        with torch.no_grad():
            mean, var = sampler.p_mean_variance(x_t=xt, t=tt, labels=labs)
        nz = z if ts > 0 else 0
        singles.append({"t": ts, "x_t": xt, "z": z, "x_prev": mean + torch.sqrt(var) * nz})
    torch.save({"w": w, "T4": {"x_T": xT, "labels": labs, "zs": zs, "x_0": x0_ref}, "singles": singles, "labels": labs},
               os.path.join(OUT, "sampler_3x32.pt"))

    # ---- schedule tables (fp64 buffers of both reference classes)
    tab = {k: v.clone() for k, v in trainer.named_buffers()}
    tab.update({k: v.clone() for k, v in sampler.named_buffers()})
    tt = torch.tensor([0, 1, 500, 999])
    tab["extract_sqrt_alphas_bar"] = ref_utils.extract(trainer.sqrt_alphas_bar, tt, (4, 3, 8, 8))
    torch.save(tab, os.path.join(OUT, "schedule.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
