"""Index utilities (reference tiger/model/utils.py:10-27)."""
from typing import Tuple

import numpy as np
import torch
from torch import Tensor

from www2023tiger_b200 import ops


def _cuda_device():
    if not torch.cuda.is_available():
        raise ops._lib.TigerLibraryError('select_latest_nids / anonymized_reindex need a CUDA device')
    return torch.device('cuda', torch.cuda.current_device())


def select_latest_nids(nids: Tensor, ts: Tensor) -> Tuple[Tensor, Tensor]:
    """Distinct ids (ascending) and, per id, the position of its maximum timestamp; ties go to the
    lowest position (torch_scatter.scatter_max CPU rule, which the reference's collator relies on)."""
    home = nids.device
    dev = home if nids.is_cuda else _cuda_device()
    d_ids = nids.to(dev, torch.int64).contiguous()
    d_ts = ts.to(dev).contiguous()
    if d_ts.dtype not in (torch.float32, torch.float64):
        d_ts = d_ts.double()
    scratch = ops.SelectScratch(int(d_ids.max()) + 1, dev) if d_ids.numel() > 2048 else None
    _, uniq, index, count = ops.select_latest(d_ids, d_ts, scratch)
    n = int(count) if d_ids.numel() else 0
    return uniq[:n].to(home), index[:n].to(home)


def anonymized_reindex(hist_nids: np.ndarray) -> np.ndarray:
    """Per row: rank of each id by its last occurrence (most recent distinct id = 1); 0 stays 0."""
    if hist_nids.size == 0:
        return np.zeros_like(hist_nids)
    t = torch.as_tensor(np.ascontiguousarray(hist_nids), dtype=torch.int64).to(_cuda_device())
    return ops.anonymized_reindex(t).cpu().numpy().astype(hist_nids.dtype)
