"""Shared policy of the operator modules: when does a module run its hand-written kernel?

* no-grad mode (`torch.no_grad()` / eval drivers): always the sm_100a kernel through the C ABI.
* autograd mode: the differentiable operators (time encoding, GRU, temporal attention, MergeLayer,
  restarters) build their graph with torch ops on the GPU so that `loss.backward()` of the
  reference's training loops works; index/no-grad operators (neighbor finder, argmax-by-timestamp,
  message store, memory get/set) run the kernels in both modes.  Native backward kernels are the
  next scope row.
There is no CPU path: modules raise on CPU tensors when the kernel route is taken.
"""
import torch

from www2023tiger_b200 import ops
from www2023tiger_b200._lib import TigerLibraryError, raise_on_err_flags


def use_kernel(*tensors) -> bool:
    """True when no autograd graph is being recorded for this call."""
    if not torch.is_grad_enabled():
        return True
    return False


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise TigerLibraryError(f'{what}: CUDA tensors required (the TIGER B200 path has no CPU fallback)')


class ErrFlags:
    """Device error word shared by the kernels of one model (include/tiger_b200.h TIGER_ERR_*)."""

    def __init__(self):
        self.t = None

    def get(self, device) -> torch.Tensor:
        if self.t is None or self.t.device != device:
            self.t = torch.zeros(1, dtype=torch.int32, device=device)
        return self.t

    def check(self):
        if self.t is not None:
            v = int(self.t.item()) & 0xffffffff
            if v:
                self.t.zero_()
                raise_on_err_flags(v)


def f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()
