"""Link-prediction evaluation and warm-up loops (reference tiger/eval_utils.py:15-68,102-129).

Same call signatures and the same AP/AUC protocol (mean over chunks of 200 events, non-finite scores
dropped with a warning); the per-batch restart bookkeeping keeps the reference's Python-set
semantics because the `uptodate_nodes` set is part of the public signature."""
import math
import warnings
from typing import Optional, Tuple

import numpy as np
import torch

from .utils import BackgroundThreadGenerator


def _batches(dl, device):
    for src, dst, neg, ts, eids, _, cg in BackgroundThreadGenerator(dl):
        yield (src.long().to(device), dst.long().to(device), neg.long().to(device), ts.float().to(device),
               eids.long().to(device), cg.to(device))


def _lazy_restart(model, cg, ts, uptodate_nodes: set, device):
    """eval_utils.py:37-42: restart every involved node not yet seen, at the batch's earliest time."""
    fresh = set(cg.np_computation_graph_nodes.tolist()) - uptodate_nodes
    ids = torch.tensor(sorted(fresh), dtype=torch.long, device=device)
    model.restart(ids, torch.full((len(ids),), ts.min().item(), device=device))
    uptodate_nodes.update(fresh)


def eval_edge_prediction(model, dl, device: torch.device, restart_mode: bool,
                         uptodate_nodes: Optional[set] = None, mean_over_n_samples: int = 200
                         ) -> Tuple[float, float]:
    from sklearn.metrics import average_precision_score, roc_auc_score
    model.eval()
    uptodate_nodes = set() if uptodate_nodes is None else uptodate_nodes
    pos_all, neg_all = [], []
    with torch.no_grad():
        for src, dst, neg, ts, eids, cg in _batches(dl, device):
            if restart_mode:
                _lazy_restart(model, cg, ts, uptodate_nodes, device)
            _, _, pos_scores, neg_scores, *_ = model.contrast_learning(src, dst, neg, ts, eids, cg)
            pos_all.append(pos_scores.sigmoid())
            neg_all.append(neg_scores.sigmoid())
    pos = torch.cat(pos_all).cpu().numpy()
    neg = torch.cat(neg_all).cpu().numpy()
    aps, aucs = [], []
    for lo in range(0, len(pos), mean_over_n_samples):
        hi = min(lo + mean_over_n_samples, len(pos))
        score = np.concatenate([pos[lo:hi], neg[lo:hi]])
        label = np.concatenate([np.ones(hi - lo), np.zeros(hi - lo)])
        ok = np.isfinite(score)
        if not ok.all():
            warnings.warn(f'Encounter invalid values: {score[~ok]}')
            score, label = score[ok], label[ok]
        aps.append(average_precision_score(label, score))
        aucs.append(roc_auc_score(label, score))
    return float(np.mean(aps)), float(np.mean(aucs))


def warmup(model, dl, device: torch.device, uptodate_nodes: Optional[set] = None) -> set:
    """Run the stream through the model in restart mode without scoring (only valid with a restarter)."""
    model.eval()
    uptodate_nodes = set() if uptodate_nodes is None else uptodate_nodes
    with torch.no_grad():
        for src, dst, neg, ts, eids, cg in _batches(dl, device):
            _lazy_restart(model, cg, ts, uptodate_nodes, device)
            model.contrast_learning(src, dst, neg, ts, eids, cg)
    return uptodate_nodes
