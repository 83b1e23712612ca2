"""Device plumbing: one libhpose handle per GPU; torch supplies device memory and streams only."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

from . import _lib


class Context:
    """Owns an ``hp_handle`` (one per GPU)."""

    def __init__(self, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise _lib.HposeError(-2, "no CUDA device visible; hpose_b200 has no CPU fallback")
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        h = C.c_void_p()
        _lib.check(_lib.lib().hp_create(self.device, C.byref(h)))
        self.handle = h

    def stream_ptr(self) -> C.c_void_p:
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.torch_device).cuda_stream)

    def launch_count(self) -> int:
        return int(_lib.lib().hp_launch_count(self.handle))

    def set_impl(self, impl: int):
        _lib.check(_lib.lib().hp_set_impl(self.handle, int(impl)))

    def fma_peak_tflops(self, packed: bool = False) -> float:
        out = C.c_double()
        _lib.check(_lib.lib().hp_fma_peak(self.handle, 1 if packed else 0, C.byref(out)))
        return float(out.value)

    def close(self):
        if self.handle:
            _lib.lib().hp_destroy(self.handle)
            self.handle = None


_contexts: Dict[int, Context] = {}


def default_context(device: Optional[int] = None) -> Context:
    import torch
    if device is None:
        if not torch.cuda.is_available():
            raise _lib.HposeError(-2, "no CUDA device visible; hpose_b200 has no CPU fallback")
        device = torch.cuda.current_device()
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
