"""ctypes binding of libtinysd_b200.so (the C ABI declared in include/tinysd_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSD_LIB") or os.path.join(_HERE, "libtinysd_b200.so")  # TSD_LIB: an alternative build (A/B runs)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m from_ddpm_to_stable_diffusion_b200.csrc.build` "
                "(or __graft_entry__.build()); there is no CPU/PyTorch fallback for this path")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.tsd_last_error.restype = ctypes.c_char_p
    return _lib


def _as_arg(a):
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, torch.Tensor):
        return ctypes.c_void_p(a.data_ptr())
    if isinstance(a, bool):
        return ctypes.c_int(int(a))
    if isinstance(a, int):
        return ctypes.c_int(a)
    if isinstance(a, float):
        return ctypes.c_float(a)
    return a  # already a ctypes value


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


PROFILE = bool(int(os.environ.get("TSD_PROFILE", "0")))
NVTX = bool(int(os.environ.get("TSD_NVTX", "0")))  # one NVTX range per C-ABI call (entry-point name) for nsys / ncu --nvtx
_prof = []


def call(name, *args):
    """Call `int name(void* stream, ...)` on the current torch CUDA stream."""
    fn = getattr(lib(), name)
    if PROFILE:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if NVTX:
        torch.cuda.nvtx.range_push(name)
    rc = fn(stream_ptr(), *[_as_arg(a) for a in args])
    if NVTX:
        torch.cuda.nvtx.range_pop()
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib().tsd_last_error().decode()}")
    if PROFILE:
        e1.record()
        _prof.append((name, tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool)), e0, e1))


def profile_report(reset=True):
    """Per (entry point, integer arguments) device time, from CUDA events (TSD_PROFILE=1 only)."""
    torch.cuda.synchronize()
    agg = {}
    for name, key, e0, e1 in _prof:
        t = e0.elapsed_time(e1)
        n, tot = agg.get((name, key), (0, 0.0))
        agg[(name, key)] = (n + 1, tot + t)
    if reset:
        _prof.clear()
    return agg


def call_nostream(name, *args):
    fn = getattr(lib(), name)
    rc = fn(*[_as_arg(a) for a in args])
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib().tsd_last_error().decode()}")


def i64(v):
    return ctypes.c_int64(int(v))


def u64(v):
    return ctypes.c_uint64(int(v) & 0xFFFFFFFFFFFFFFFF)


def f32(v):
    return ctypes.c_float(float(v))
