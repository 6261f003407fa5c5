"""Stand-in for torch_scatter.scatter_max used ONLY by make_golden.py to drive the
unmodified reference (torch_scatter is not installed in the build container).
Implements torch_scatter's CPU reducer semantics: sequential scan, strict '>' update,
so the lowest position wins ties; empty groups get arg == len(src)."""
import torch


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    n = int(index.max()) + 1 if len(index) else 0
    best = torch.full((n,), float('-inf'), dtype=src.dtype)
    arg = torch.full((n,), len(src), dtype=torch.long)
    for pos in range(len(src)):
        g = int(index[pos])
        if src[pos] > best[g]:
            best[g] = src[pos]
            arg[g] = pos
    return best, arg
