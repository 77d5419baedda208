"""Compile the CUDA sources in this directory into libtinysd_b200.so (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting
.so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libtinysd_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _headers_digest():
    h = hashlib.sha1()
    for f in sorted(os.listdir(HERE)):
        if f.endswith((".cuh", ".h")):
            h.update(open(os.path.join(HERE, f), "rb").read())
    inc = os.path.join(os.path.dirname(PKG), "include", "tinysd_b200.h")
    h.update(open(inc, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile_one(src, hdig, verbose):
    path = os.path.join(HERE, src)
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    stamp = obj + ".stamp"
    dig = hashlib.sha1(open(path, "rb").read() + hdig.encode()).hexdigest()
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(dig)
    return obj, r.stderr if verbose else ""


def build(verbose=False, force=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    hdig = _headers_digest()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(s, hdig, verbose), srcs))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose and log:
        print(log)
    return OUT


if __name__ == "__main__":
    out = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print("built", out)
