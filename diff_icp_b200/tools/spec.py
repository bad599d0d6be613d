"""Tensor "specs" (dtype + device), same contract as the reference's tools/spec.py:24-61.

The B200 build computes on CUDA in fp32 only; `defspec` is the CUDA spec when a device is visible.  On a
machine without a GPU the package can still be imported (for host-side logic and CPU unit tests of the
orchestration), but every compute entry point raises: there is no CPU fallback.
"""

import io
import pickle

import torch

cpuspec = {"device": "cpu", "dtype": torch.float32}
gpuspec = {"device": "cuda", "dtype": torch.float32}
use_cuda = torch.cuda.is_available()
defspec = gpuspec if use_cuda else cpuspec


def getspec(*T):
    """Common (device, dtype) of the given tensors (None entries ignored); ValueError if they differ
    (reference: tools/spec.py:39-43)."""
    found = {(t.device, t.dtype) for t in T if t is not None}
    if len(found) != 1:
        raise ValueError("the different input tensors to this function should be on the same device and use the same dtype !")
    dev, dt = next(iter(found))
    return {"device": dev, "dtype": dt}


class CPU_Unpickler(pickle.Unpickler):
    """Unpickle on the CPU objects whose tensors lived on a GPU when pickled (reference: tools/spec.py:57-61)."""

    def find_class(self, module, name):
        if module == "torch.storage" and name == "_load_from_bytes":
            return lambda b: torch.load(io.BytesIO(b), map_location="cpu")
        return super().find_class(module, name)
