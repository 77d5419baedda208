"""ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY.  Recipe that stages the UNMODIFIED reference modules of the hot path
under oracle/_ref/ so that bench.py's reference arm (and its GPU-eager context leg) can run the reference's own code
on the GPU box, where /root/reference does not exist.

  /root/reference/06_tiny_stable_diffusion/diffusion.py  ->  oracle/_ref/tiny_sd/diffusion.py
  /root/reference/06_tiny_stable_diffusion/utils.py      ->  oracle/_ref/tiny_sd/utils.py
  /root/reference/03_variational_autoencoder/models.py   ->  oracle/_ref/vae/models.py      (VQ-VAE codec)

oracle/_ref/ is an OUTPUT directory: git-ignored (no reference source enters the history), not gpurun-ignored (it
travels to the GPU box like the built .so).  The reference is pure Python: "building" it is this byte-for-byte staging;
a manifest with the SHA-256 of every staged file is written next to them.  Run by __graft_entry__.build() whenever
/root/reference is present; on the GPU box the staged copy is used as is.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TSD_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = [("06_tiny_stable_diffusion/diffusion.py", "tiny_sd/diffusion.py"),
         ("06_tiny_stable_diffusion/utils.py", "tiny_sd/utils.py"),
         ("03_variational_autoencoder/models.py", "vae/models.py")]


def stage():
    """Returns the manifest dict, or None when the reference tree is not available here."""
    if not os.path.isdir(REF):
        return None
    manifest = {}
    for src, dst in FILES:
        s, d = os.path.join(REF, src), os.path.join(OUT, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[dst] = {"source": src, "sha256": hashlib.sha256(open(d, "rb").read()).hexdigest()}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


def available():
    return all(os.path.exists(os.path.join(OUT, dst)) for _, dst in FILES)


def load_tiny_sd():
    """Imports the staged reference modules (`diffusion`, `utils`) under private names; returns (diffusion, utils)."""
    import importlib.util
    mods = []
    for name in ("diffusion", "utils"):
        spec = importlib.util.spec_from_file_location(f"_tsd_reference_{name}", os.path.join(OUT, "tiny_sd", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


def load_vae():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_tsd_reference_vae_models", os.path.join(OUT, "vae", "models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    m = stage()
    print("staged" if m else "reference tree not found", json.dumps(m, indent=1) if m else "")
