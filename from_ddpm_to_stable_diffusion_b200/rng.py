"""Position of a Philox random stream, kept in device memory.

The reference draws its random numbers from torch's global generator (``torch.randint`` / ``randn_like`` at
06_tiny_stable_diffusion/utils.py:112-113,163; ``nn.Dropout`` at diffusion.py:97): every call and every sample sees
fresh numbers.  Here the numbers come from counter-based Philox inside the kernels, keyed by
``(seed, call counter, GLOBAL sample index, element)``:

  * the call counter lives on the device and is advanced by a one-thread kernel, so a captured CUDA graph (the training
    iteration) draws new numbers on every replay and two ``SamplerDDPM`` calls never reuse a noise sequence;
  * the sample index is global (``sample0`` = index of this rank's first sample), so a batch sharded over N ranks draws
    exactly what a single rank would -- results are independent of N.

``seed`` defaults to a value derived from ``torch.initial_seed()`` at construction, so ``torch.manual_seed`` controls the
stream like it does in the reference.
"""
import torch

from . import ops


def default_seed(salt):
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + salt) & 0xFFFFFFFFFFFFFFFF


class DeviceRng:
    def __init__(self, salt, seed=None):
        self.seed = default_seed(salt) if seed is None else int(seed)
        self._buf = None           # int64 [2] on the device: (calls, sample0)
        self._calls = 0            # host mirror, exact while no graph replay has advanced the device counter
        self._sample0 = 0
        self._replayed = False
        self.frozen = False

    def tensor(self, device):
        if self._buf is None or self._buf.device != device:
            self._buf = torch.tensor([self._calls, self._sample0], dtype=torch.int64, device=device)
        return self._buf

    def advance(self, device, force=False):
        """calls += 1 on the device (stream-ordered; capturable).  Returns the device tensor.  While ``frozen`` the
        counter is left alone (gradient accumulation: every micro-batch of a step belongs to the same call; the owner
        of the step advances it once with ``force=True``)."""
        buf = self.tensor(device)
        if self.frozen and not force:
            return buf
        ops.counter_add_u64(buf, 1)
        if torch.cuda.is_current_stream_capturing():
            self._replayed = True  # from now on the device counter runs ahead of the host mirror
        else:
            self._calls += 1
        return buf

    def set_sample0(self, sample0):
        """Global index of this rank's first sample (data-parallel shards, micro-batches)."""
        if int(sample0) == self._sample0:
            return
        self._sample0 = int(sample0)
        if self._buf is not None:
            self._buf[1:2].fill_(self._sample0)

    @property
    def calls(self):
        if self._buf is not None and self._replayed:
            self._calls = int(self._buf[0].item())
        return self._calls

    def state_dict(self):
        return {"seed": self.seed, "calls": self.calls, "sample0": self._sample0}

    def load_state_dict(self, sd):
        self.seed = int(sd["seed"])
        self._calls = int(sd["calls"])
        self._sample0 = int(sd.get("sample0", 0))
        if self._buf is not None:
            self._buf.copy_(torch.tensor([self._calls, self._sample0], dtype=torch.int64))
