"""Generate tests/golden/scoring_*.npz from the real thing: torch's CosineSimilarity, scikit-learn's roc_curve / auc,
and the reference's own get_tpr / calculate_scores / calculate_identification_acc, whose function bodies are executed
straight from /root/reference/utils/modules.py (the module itself cannot be imported here: it pulls in easydict, nltk
and torchvision transforms through utils.prepare).  Run in the build container:

    python tests/golden/make_golden_scoring.py
"""
import ast
import contextlib
import io
import os
import re
import tempfile
import types

import numpy as np
import sklearn
import torch
from sklearn import metrics

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/utils/modules.py"


def reference_functions():
    tree = ast.parse(open(REF).read())
    wanted = {"get_tpr", "calculate_scores", "calculate_identification_acc"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    ns = {"np": np, "metrics": metrics, "os": os}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


def printed(fn, *a):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


def summary_numbers(line):
    return np.array([float(v) for v in re.findall(r"(?:AUC|EER|1e-5|1e-4|1e-3|score) (-?[0-9.]+|nan)", line)])


def main():
    ref = reference_functions()
    args = types.SimpleNamespace(is_roc=False)
    rng = np.random.RandomState(100)
    g = torch.Generator().manual_seed(100)
    cases = {}

    # embeddings -> cosine -> ROC (the whole path of utils/modules.py:150-166)
    n, d = 600, 40
    ident = torch.randn(n, d, generator=g)
    label = (torch.rand(n, generator=g) < 0.4).long()
    x1 = ident + 0.7 * torch.randn(n, d, generator=g)
    x2 = torch.where(label[:, None] == 1, ident, torch.randn(n, d, generator=g)) + 0.7 * torch.randn(n, d, generator=g)
    x1[5] = 0.0                                   # eps clamp
    x2[9] *= 1e-9
    score = torch.nn.CosineSimilarity(dim=1, eps=1e-6)(x1, x2)
    cases["embed"] = dict(x1=x1.numpy(), x2=x2.numpy(), labels=label.numpy(), scores=score.numpy())

    # heavy ties (scores on a coarse grid), signed zeros
    n = 2000
    label = (rng.rand(n) < 0.3).astype(np.int64)
    s = np.round((rng.randn(n) * 0.25 + 0.35 * label) * 20) / 20
    s = s.astype(np.float32)
    s[::97] = -0.0
    s[1::97] = 0.0
    cases["ties"] = dict(labels=label, scores=s)

    # 6000 pairs with well separated classes (TPR@FPR targets are resolved), fp32 scores in [-1, 1]
    n = 6000
    label = (np.arange(n) % 10 == 0).astype(np.int64)
    s = np.clip(rng.randn(n) * 0.12 + np.where(label == 1, 0.62, 0.05), -1, 1).astype(np.float32)
    cases["pairs6000"] = dict(labels=label, scores=s)

    # degenerate: one score for everybody; two samples
    cases["constant"] = dict(labels=np.array([0, 1, 1, 0, 1], np.int64), scores=np.full(5, 0.25, np.float32))
    cases["two"] = dict(labels=np.array([1, 0], np.int64), scores=np.array([0.9, -0.1], np.float32))

    for name, c in cases.items():
        y_score = c["scores"].tolist()            # the reference accumulates Python lists (utils/modules.py:152-153)
        y_true = c["labels"].tolist()
        fpr, tpr, thr = metrics.roc_curve(y_true, y_score)
        fpr_all, tpr_all, thr_all = metrics.roc_curve(y_true, y_score, drop_intermediate=False)
        line = printed(ref["calculate_scores"], y_score, y_true, args)
        c.update(fpr=fpr, tpr=tpr, thr=thr, fpr_all=fpr_all, tpr_all=tpr_all, thr_all=thr_all,
                 summary=summary_numbers(line), summary_line=np.array(line.strip()),
                 get_tpr=np.array(ref["get_tpr"](np.flipud(fpr), np.flipud(tpr)), np.float64),
                 auc=np.float64(metrics.auc(np.flipud(fpr), np.flipud(tpr))))
        np.savez_compressed(os.path.join(HERE, f"scoring_{name}.npz"), **c)
        print(name, line.strip(), "points", fpr.size, "of", fpr_all.size)

    # identification (utils/modules.py:76-88): 50 subjects x 12 scores, with ties and a duplicated maximum
    total_sub, each = 50, 12
    s = rng.rand(total_sub, each).astype(np.float32)
    for k in range(0, total_sub, 2):
        s[k, k % each] = 2.0
    s[4, 7] = 2.0
    s[4, 4 % each] = 2.0                         # two maxima: the first one wins
    s[7] = 0.5                                   # all equal -> index 0
    with tempfile.TemporaryDirectory() as tmp:
        a = types.SimpleNamespace(checkpoints_path=tmp, test_sub=total_sub)
        line = printed(ref["calculate_identification_acc"], s.ravel().tolist(), a)
    acc = float(re.search(r"accuracy \(%\) ([0-9.]+)", line).group(1))
    np.savez_compressed(os.path.join(HERE, "scoring_ident.npz"), scores=s.ravel(), total_sub=np.int64(total_sub),
                        argmax=np.argmax(s, axis=1), acc=np.float64(acc))
    print("ident", line.strip().replace("\n", " | "))
    print("versions: torch", torch.__version__, "sklearn", sklearn.__version__, "numpy", np.__version__)


if __name__ == "__main__":
    main()
