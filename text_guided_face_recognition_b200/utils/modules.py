"""Verification / identification scoring of the reference's utils/modules.py on the B200.

Mirrors, with the reference's names and argument meaning:
    get_tpr(fprs, tprs)                          utils/modules.py:40-47
    calculate_scores(y_score, y_true, args)      utils/modules.py:52-72   (AUC, EER, TPR@FPR=1e-5/1e-4/1e-3)
    calculate_identification_acc(y_score, args)  utils/modules.py:76-88
and the device work inside `test` (utils/modules.py:150-166):
    pair_scores(out1, out2)                      nn.CosineSimilarity(dim=1, eps=1e-6)
    roc_curve(y_true, y_score)                   sklearn.metrics.roc_curve as called at utils/modules.py:54

The cosine, the descending sort, the per-threshold counts, the collinear-point drop and the per-subject argmax run
in libtgfr_b200.so (csrc/scoring.cu).  What stays on the host is float64 arithmetic on the few ROC points that
survive drop_intermediate (the reference does the same with numpy).  There is no CPU fallback: inputs are moved to the
current CUDA device, and a missing library or device raises.
"""
from __future__ import annotations

import os

import numpy as np
import torch

try:
    from .. import ops
except (ImportError, ValueError):          # imported as a top-level `utils` package
    from text_guided_face_recognition_b200 import ops

__all__ = ["pair_scores", "roc_curve", "get_tpr", "calculate_scores", "calculate_identification_acc", "score_pairs"]

_FPR_TARGETS = (1e-5, 1e-4, 1e-3)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("text_guided_face_recognition_b200.utils.modules runs on CUDA (sm_100a) only; no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as_cuda(x, dtype):
    if isinstance(x, torch.Tensor):
        t = x.detach()
        if t.device.type != "cuda":
            t = t.to(_device())
        return t.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=_device())


def pair_scores(out1, out2, eps=1e-6):
    """cosine_sim(out1, out2) of utils/modules.py:150-151: [N, D] x 2 -> [N] CUDA fp32."""
    return ops.pair_cosine(out1, out2, eps)


def roc_curve(y_true, y_score, drop_intermediate=True):
    """(fpr, tpr, thresholds) exactly as sklearn.metrics.roc_curve(y_true, y_score) returns them (float64 numpy).

    y_score: list / numpy / tensor of scores (compared as fp32, which is what the reference's scores are);
    y_true: 0/1 (or -1/1) labels, positive = 1."""
    scores = _as_cuda(y_score, torch.float32)
    labels = _as_cuda(y_true, torch.int64)
    thr, fps, tps, _ = ops.roc_counts(scores, labels, drop_intermediate)
    host = torch.stack((fps, tps)).cpu().numpy().astype(np.float64)
    fps = np.concatenate(([0.0], host[0]))
    tps = np.concatenate(([0.0], host[1]))
    thresholds = np.concatenate(([np.inf], thr.cpu().numpy().astype(np.float64)))
    with np.errstate(invalid="ignore", divide="ignore"):
        fpr = fps / fps[-1] if fps[-1] > 0 else np.full(fps.shape, np.nan)
        tpr = tps / tps[-1] if tps[-1] > 0 else np.full(tps.shape, np.nan)
    return fpr, tpr, thresholds


def get_tpr(fprs, tprs):
    """TPR (in %) at the ROC point whose FPR is nearest to 1e-5, 1e-4, 1e-3 (first such point of the given order)."""
    fprs, tprs = np.asarray(fprs), np.asarray(tprs)
    return [tprs[int(np.argmin(np.abs(fprs - target)))] * 100 for target in _FPR_TARGETS]


def _area(x, y):
    """sklearn.metrics.auc for a monotone x (the reference passes the flipped, i.e. decreasing, fpr)."""
    dx = np.diff(x)
    if np.any(dx < 0) and not np.all(dx <= 0):
        raise ValueError("x is neither increasing nor decreasing")
    direction = -1.0 if np.any(dx < 0) else 1.0
    trap = getattr(np, "trapezoid", None) or np.trapz          # NumPy >= 2.0 renamed trapz
    return float(direction * trap(y, x))


def calculate_scores(y_score, y_true, args=None):
    """Prints the reference's summary line and returns the numbers as a dict (the reference returns None)."""
    fpr, tpr, _ = roc_curve(y_true, y_score)
    fprs, tprs = fpr[::-1], tpr[::-1]
    eer = fprs[np.nanargmin(np.absolute((1 - tprs) - fprs))]
    auc = _area(fprs, tprs)
    row = get_tpr(fprs, tprs)
    total = row[0] + row[1] + row[2]
    print("AUC {:.4f} | EER {:.4f} | TPR@FPR=1e-5 {:.4f} | TPR@FPR=1e-4 {:.4f} | TPR@FPR=1e-3 {:.4f} | score {:.4f}".format(
        auc, eer, row[0], row[1], row[2], total))
    if args is not None and getattr(args, "is_roc", False):
        filename = os.path.join(".", args.roc_file + ".npy")
        print("saving npy file in :", filename)
        with open(filename, "wb") as f:
            np.save(f, np.asarray(_host(y_true)))
            np.save(f, np.asarray(_host(y_score)))
    return {"auc": auc, "eer": float(eer), "tpr_at_fpr": [float(v) for v in row], "score": float(total)}


def _host(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x


def calculate_identification_acc(y_score, args):
    """Rank-1 identification: subject s is correct when the maximum of its `pair_each_sub` scores is at position s."""
    total_sub = int(args.test_sub)
    path = getattr(args, "checkpoints_path", None)
    if path:
        with open(os.path.join(path, "ident_file"), "wb") as f:
            np.save(f, np.asarray(_host(y_score)))
    scores = _as_cuda(y_score, torch.float32).view(-1)
    pair_each_sub = scores.numel() // total_sub
    print("total subjects: ", total_sub)
    best = ops.row_argmax(scores[: total_sub * pair_each_sub].view(total_sub, pair_each_sub))
    acc = int((best == torch.arange(total_sub, device=best.device)).sum().item())
    print("identification accuracy (%)", (acc / total_sub) * 100)
    return acc / total_sub * 100


def score_pairs(batches, args=None, is_ident=False):
    """The scoring part of `test` (utils/modules.py:150-166): batches yields (out1 [n, D], out2 [n, D], pair_label [n]);
    scores stay on the device until the ROC."""
    preds, labels = [], []
    for out1, out2, pair_label in batches:
        preds.append(pair_scores(out1, out2))
        labels.append(_as_cuda(pair_label, torch.int64).view(-1))
    preds, labels = torch.cat(preds), torch.cat(labels)
    result = {}
    if is_ident:
        result["identification_acc"] = calculate_identification_acc(preds, args)
    result.update(calculate_scores(preds, labels, args))
    return result
