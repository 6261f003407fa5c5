"""Shared helpers of the GPU parity tests (the CUDA path is always reached through the C ABI)."""
import numpy as np
import torch

from oracle import tiger_oracle as O
from www2023tiger_b200 import ops
from www2023tiger_b200.engine import TigerEngine

DEV = 'cuda'


def dev(x, dtype=None):
    t = torch.as_tensor(x)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV).contiguous()


def device_csr(src, dst, ts, eids, n_nodes):
    return ops.csr_build(dev(src, torch.int64), dev(dst, torch.int64), dev(ts, torch.float64),
                         dev(eids, torch.int64), n_nodes)


def engine_from(W, csr, *, N, dim, efeats, nfeats, K, H, B, msg_src, upd_src, restarter=None, lazy_restart=False,
                want_targets=True, hist_len=40):
    return TigerEngine({k: torch.as_tensor(v) for k, v in W.items()}, csr, n_nodes=N, dim=dim,
                       efeats=None if efeats is None else dev(efeats, torch.float32),
                       nfeats=None if nfeats is None else dev(nfeats, torch.float32),
                       n_neighbors=K, n_head=H, batch_size=B, msg_src=msg_src, upd_src=upd_src,
                       restarter=restarter, hist_len=hist_len, lazy_restart=lazy_restart,
                       want_restarter_targets=want_targets)


def oracle_from(W, src, dst, ts, eids, *, N, dim, efeats, nfeats, K, H, msg_src, upd_src, restarter='static',
                hist_len=40):
    graph = O.OracleGraph(src, dst, ts, eids, n_nodes=N)
    model = O.OracleTIGER(W, graph, N, dim, efeats, nfeats, n_neighbors=K, n_head=H, msg_src=msg_src,
                          upd_src=upd_src, restarter=restarter, hist_len=hist_len)
    return graph, model


def cpu(t):
    return t.detach().cpu().numpy()
