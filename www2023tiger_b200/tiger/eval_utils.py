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


def average_precision_score(label: np.ndarray, score: np.ndarray) -> float:
    """sklearn.metrics.average_precision_score for binary labels: AP = sum_n (R_n - R_{n-1}) P_n over the distinct
    score thresholds, descending.  The reference calls sklearn once per 200-event chunk (eval_utils.py:55-62), where its
    input validation costs ~2 ms per call - more than the whole device path of the batch; same numbers
    (tests/test_dropin_cpu.py compares against sklearn, ties included)."""
    order = np.argsort(-score, kind='mergesort')
    y, sc = label[order].astype(np.float64), score[order]
    last = np.r_[np.nonzero(np.diff(sc))[0], len(sc) - 1]           # last index of every group of equal scores
    tps = np.cumsum(y)[last]
    precision = tps / (last + 1.0)
    recall = tps / tps[-1] if tps[-1] > 0 else np.full_like(tps, np.nan)
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def roc_auc_score(label: np.ndarray, score: np.ndarray) -> float:
    """sklearn.metrics.roc_auc_score for binary labels (trapezoidal ROC area = Mann-Whitney U with average ranks
    for ties)."""
    order = np.argsort(score, kind='mergesort')
    sc = score[order]
    starts = np.r_[0, np.nonzero(np.diff(sc))[0] + 1]
    ends = np.r_[starts[1:], len(sc)]
    rank_of_group = (starts + ends + 1) / 2.0                       # average 1-based rank of each tie group
    ranks = np.empty(len(sc))
    ranks[order] = np.repeat(rank_of_group, ends - starts)
    pos = label > 0
    n_pos, n_neg = int(pos.sum()), int((~pos).sum())
    if n_pos == 0 or n_neg == 0:
        raise ValueError('Only one class present in y_true. ROC AUC score is not defined in that case.')
    return float((ranks[pos].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


def eval_edge_prediction(model, dl, device: torch.device, restart_mode: bool,
                         uptodate_nodes: Optional[set] = None, mean_over_n_samples: int = 200
                         ) -> Tuple[float, float]:
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
