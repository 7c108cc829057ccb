"""CPU oracle for verification / identification scoring -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in plain numpy (float64 / exact integers):
  * nn.CosineSimilarity(dim=1, eps=1e-6) as used at /root/reference utils/modules.py:150-151 (third-party arithmetic:
    PyTorch, requirements.txt pins torch==2.5.1, this image has 2.11.0; ATen normalises each vector by
    max(|x|, eps) and sums the products),
  * sklearn.metrics.roc_curve as called at utils/modules.py:54 (third-party: scikit-learn, NOT pinned by the
    reference's requirements.txt; this image has 1.9.0, whose published algorithm -- metrics/_ranking.py roc_curve +
    confusion_matrix_at_thresholds -- is restated here: stable descending sort, one point per distinct score,
    tps = cumsum(y == 1), fps = 1 + index - tps, drop_intermediate by second differences, a leading (inf, 0, 0) point),
  * the reference's own host arithmetic on that curve: get_tpr (utils/modules.py:40-47), EER / AUC / score
    (utils/modules.py:56-61), rank-1 identification (utils/modules.py:81-87).

Parity pin: tests/golden/make_golden_scoring.py runs torch's CosineSimilarity, scikit-learn's roc_curve / auc and the
reference's own get_tpr / calculate_scores / calculate_identification_acc (function bodies executed from
/root/reference/utils/modules.py) on seeded inputs and commits tests/golden/scoring_*.npz;
tests/test_scoring.py checks every function below against them (exact for counts, thresholds and rates).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import numpy as np


def pair_cosine(x1, x2, eps=1e-6):
    """utils/modules.py:150-151 -- cosine of matching rows, each norm clamped from below by eps."""
    x1 = np.asarray(x1, np.float64)
    x2 = np.asarray(x2, np.float64)
    n1 = np.maximum(np.sqrt((x1 * x1).sum(1, keepdims=True)), eps)
    n2 = np.maximum(np.sqrt((x2 * x2).sum(1, keepdims=True)), eps)
    return ((x1 / n1) * (x2 / n2)).sum(1)


def roc_counts(y_true, y_score, drop_intermediate=True):
    """Integer stage of roc_curve (scikit-learn 1.9.0 metrics/_ranking.py): (thresholds, fps, tps) without the
    leading (inf, 0, 0) point.  Scores are compared as fp32 (utils/modules.py:152 turns an fp32 tensor into a list)."""
    score = np.asarray(y_score, np.float32).ravel()
    if np.isnan(score).any():
        raise ValueError("Input y_score contains NaN.")
    pos = (np.asarray(y_true).ravel() == 1)
    order = np.argsort(score, kind="stable")[::-1]
    score, pos = score[order], pos[order]
    ends = np.r_[np.nonzero(np.diff(score))[0], score.size - 1].astype(np.int64) if score.size else np.zeros(0, np.int64)
    tps = np.cumsum(pos.astype(np.int64))[ends]
    fps = 1 + ends - tps
    thr = score[ends]
    if drop_intermediate and fps.size > 2:
        keep = np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True]
        thr, fps, tps = thr[keep], fps[keep], tps[keep]
    return thr, fps, tps


def roc_curve(y_true, y_score, drop_intermediate=True):
    """utils/modules.py:54 -- (fpr, tpr, thresholds), float64, as scikit-learn returns them."""
    thr, fps, tps = roc_counts(y_true, y_score, drop_intermediate)
    fps = np.r_[0.0, fps.astype(np.float64)]
    tps = np.r_[0.0, tps.astype(np.float64)]
    thr = np.r_[np.inf, thr.astype(np.float64)]
    fpr = fps / fps[-1] if fps[-1] > 0 else np.full(fps.shape, np.nan)
    tpr = tps / tps[-1] if tps[-1] > 0 else np.full(tps.shape, np.nan)
    return fpr, tpr, thr


def get_tpr(fprs, tprs):
    """utils/modules.py:40-47 -- the first point (in the order given) whose FPR is nearest to each target."""
    out = []
    for target in (1e-5, 1e-4, 1e-3):
        gap = np.abs(np.asarray(fprs) - target)
        out.append(tprs[int(np.flatnonzero(gap == gap.min())[0])] * 100)
    return out


def calculate_scores(y_score, y_true):
    """utils/modules.py:52-61 -- AUC, EER, TPR@FPR row and their sum, on the flipped curve."""
    fpr, tpr, _ = roc_curve(y_true, y_score)
    fprs, tprs = fpr[::-1], tpr[::-1]
    eer = fprs[np.nanargmin(np.abs((1 - tprs) - fprs))]
    # sklearn.metrics.auc on a decreasing x: minus the trapezoid sum
    auc = -np.trapezoid(tprs, fprs) if np.any(np.diff(fprs) < 0) else np.trapezoid(tprs, fprs)
    row = get_tpr(fprs, tprs)
    return {"auc": float(auc), "eer": float(eer), "tpr_at_fpr": [float(v) for v in row], "score": float(sum(row))}


def identification(y_score, total_sub):
    """utils/modules.py:81-87 -- (argmax per subject, accuracy in %)."""
    s = np.asarray(y_score)
    each = s.size // total_sub
    best = np.argmax(s[: total_sub * each].reshape(total_sub, each), axis=1)
    return best, float((best == np.arange(total_sub)).sum() / total_sub * 100)
