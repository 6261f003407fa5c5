"""Shared policy of the operator modules: when does a module run its hand-written kernel?

* no-grad mode (`torch.no_grad()` / eval drivers): the sm_100a inference kernel through the C ABI.
* training: `TIGER.contrast_and_mutual_learning` is one autograd node whose forward and backward are the
  hand-written training step of www2023tiger_b200/train.py (default operator variants).
* the operator modules called on their own under autograd (non-default variants, tests) build a torch-op graph;
  index/no-grad operators (neighbor finder, argmax-by-timestamp, message store, memory get/set) run the kernels
  in every mode.
There is no CPU path: modules raise on CPU tensors when the kernel route is taken.
"""
import torch

from www2023tiger_b200 import ops
from www2023tiger_b200._lib import TigerLibraryError, raise_on_err_flags


def use_kernel(module=None) -> bool:
    """True when the inference kernel of an operator may serve this call: no autograd graph is being recorded, and
    the module is not a training-mode module with active dropout (the reference applies dropout whenever
    `module.training`, also under `torch.no_grad()` - `TIGER.restart` is called from the training loop in train()
    mode, train_self_supervised.py:158-163 - and the inference kernels have no dropout)."""
    if torch.is_grad_enabled():
        return False
    if module is not None and module.training:
        p = getattr(module, 'dropout', 0.0)
        p = getattr(p, 'p', p)
        if isinstance(p, (int, float)) and p > 0:
            return False
    return True


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
