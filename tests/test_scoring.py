"""Verification / identification scoring (reference utils/modules.py:40-88,150-166).

CPU: the numpy oracle against fixtures produced by torch's CosineSimilarity, scikit-learn's roc_curve / auc and the
reference's own calculate_scores / get_tpr / calculate_identification_acc (tests/golden/make_golden_scoring.py).
GPU: the CUDA path (utils/modules.py mirror -> ops -> C ABI, csrc/scoring.cu) against the fixtures and the oracle.
Bar: ROC counts, thresholds, rates and argmax decisions BIT-EXACT; cosine scores within 2e-6 absolute (fp32)."""
import contextlib
import io
import os
import types

import numpy as np
import pytest
import torch

from oracle import scoring_oracle as SO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["embed", "ties", "pairs6000", "constant", "two"]


def load(name):
    return np.load(os.path.join(GOLDEN, f"scoring_{name}.npz"))


# ------------------------------------------------------------------ oracle vs the real thing (CPU)
@pytest.mark.parametrize("name", CASES)
def test_oracle_roc_is_sklearn_roc(name):
    g = load(name)
    for drop, suffix in ((True, ""), (False, "_all")):
        fpr, tpr, thr = SO.roc_curve(g["labels"], g["scores"], drop_intermediate=drop)
        np.testing.assert_array_equal(fpr, g["fpr" + suffix])
        np.testing.assert_array_equal(tpr, g["tpr" + suffix])
        np.testing.assert_array_equal(thr, g["thr" + suffix])


@pytest.mark.parametrize("name", CASES)
def test_oracle_summary_is_the_reference_summary(name):
    g = load(name)
    out = SO.calculate_scores(g["scores"], g["labels"])
    got = np.array([out["auc"], out["eer"], *out["tpr_at_fpr"], out["score"]])
    np.testing.assert_allclose(got, g["summary"], atol=5.1e-5, rtol=0)       # the reference prints 4 decimals
    np.testing.assert_array_equal(np.array(out["tpr_at_fpr"]), g["get_tpr"])
    assert out["auc"] == float(g["auc"])


def test_oracle_cosine_is_torch_cosine():
    g = load("embed")
    np.testing.assert_allclose(SO.pair_cosine(g["x1"], g["x2"]), g["scores"], atol=1e-6, rtol=0)
    assert SO.pair_cosine(g["x1"], g["x2"])[5] == 0.0                          # zero vector: eps clamp, not NaN


def test_oracle_identification_is_the_reference():
    g = load("ident")
    best, acc = SO.identification(g["scores"], int(g["total_sub"]))
    np.testing.assert_array_equal(best, g["argmax"])
    assert acc == float(g["acc"])


def test_oracle_rejects_nan_scores():
    with pytest.raises(ValueError):
        SO.roc_counts([0, 1], [0.5, float("nan")])


def test_host_side_of_the_mirror_is_sklearn_arithmetic():
    """utils/modules.py mirror, host half (no GPU): get_tpr picks the reference's point, the trapezoid AUC on the flipped
    (decreasing) fpr equals sklearn.metrics.auc bit for bit."""
    from sklearn import metrics as skm
    from text_guided_face_recognition_b200.utils import modules
    for name in CASES:
        g = load(name)
        fprs, tprs = np.flipud(g["fpr"]), np.flipud(g["tpr"])
        np.testing.assert_array_equal(np.array(modules.get_tpr(fprs, tprs)), g["get_tpr"])
        assert modules._area(fprs, tprs) == skm.auc(fprs, tprs) == float(g["auc"])
    with pytest.raises(ValueError):
        modules._area(np.array([0.0, 1.0, 0.5]), np.array([0.0, 1.0, 1.0]))


def test_product_refuses_cpu():
    from text_guided_face_recognition_b200 import _lib, ops
    with pytest.raises((_lib.TgfrError, RuntimeError)):
        ops.pair_cosine(torch.randn(4, 8), torch.randn(4, 8))
    with pytest.raises((_lib.TgfrError, RuntimeError)):
        ops.row_argmax(torch.randn(4, 8))


# ------------------------------------------------------------------ CUDA path (B200)
def cuda_roc(labels, scores, drop=True):
    from text_guided_face_recognition_b200.utils import modules
    return modules.roc_curve(labels, scores, drop_intermediate=drop)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_roc_matches_sklearn_fixture(name):
    g = load(name)
    for drop, suffix in ((True, ""), (False, "_all")):
        fpr, tpr, thr = cuda_roc(g["labels"], g["scores"], drop)
        np.testing.assert_array_equal(fpr, g["fpr" + suffix])
        np.testing.assert_array_equal(tpr, g["tpr" + suffix])
        np.testing.assert_array_equal(thr, g["thr" + suffix])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_summary_matches_reference_line(name):
    from text_guided_face_recognition_b200.utils import modules
    g = load(name)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = modules.calculate_scores(torch.from_numpy(g["scores"]).cuda(), torch.from_numpy(g["labels"]).cuda(), None)
    assert buf.getvalue().strip() == str(g["summary_line"])                    # the very line the reference prints
    ref = SO.calculate_scores(g["scores"], g["labels"])
    assert out["auc"] == ref["auc"] and out["eer"] == ref["eer"] and out["tpr_at_fpr"] == ref["tpr_at_fpr"]


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 31, 4095, 4096, 4097, 60000, (1 << 20) + 12345])
@pytest.mark.parametrize("grid", [0, 64])
def test_gpu_roc_counts_match_oracle(n, grid):
    """Sizes around the 4096-key tile, BASELINE config 5's 60 000 pairs, and > 2^20; grid = score quantisation (ties)."""
    from text_guided_face_recognition_b200 import ops
    rs = np.random.RandomState(n + grid)
    labels = (rs.rand(n) < 0.1).astype(np.int64)
    scores = np.clip(rs.randn(n) * 0.2 + 0.5 * labels, -1, 1).astype(np.float32)
    if grid:
        scores = (np.round(scores * grid) / grid).astype(np.float32)
    for drop in (True, False):
        thr, fps, tps, distinct = ops.roc_counts(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(), drop)
        rthr, rfps, rtps = SO.roc_counts(labels, scores, drop)
        np.testing.assert_array_equal(fps.cpu().numpy(), rfps)
        np.testing.assert_array_equal(tps.cpu().numpy(), rtps)
        np.testing.assert_array_equal(thr.cpu().numpy(), rthr)
        assert distinct == np.unique(scores).size
        # size-independent properties: the last point counts everybody; both counts are non-decreasing
        assert int(fps[-1] + tps[-1]) == n and int(tps[-1]) == int(labels.sum())
        assert bool((fps[1:] >= fps[:-1]).all()) and bool((tps[1:] >= tps[:-1]).all())


@pytest.mark.gpu
def test_gpu_roc_edge_cases():
    from text_guided_face_recognition_b200 import ops
    # infinities, signed zeros, denormals, every label equal
    scores = np.array([np.inf, -np.inf, 0.0, -0.0, 1e-42, -1e-42, 1.0, -1.0, 1.0], np.float32)
    for labels in (np.array([1, 0, 1, 0, 1, 0, 1, 0, 0]), np.ones(9, np.int64), np.zeros(9, np.int64),
                   np.array([1, -1, 1, -1, 1, -1, 1, -1, -1])):
        for drop in (True, False):
            thr, fps, tps, _ = ops.roc_counts(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(), drop)
            rthr, rfps, rtps = SO.roc_counts(labels, scores, drop)
            np.testing.assert_array_equal(thr.cpu().numpy(), rthr)
            np.testing.assert_array_equal(fps.cpu().numpy(), rfps)
            np.testing.assert_array_equal(tps.cpu().numpy(), rtps)
    with pytest.raises(ValueError):
        ops.roc_counts(torch.tensor([0.1, float("nan")]).cuda(), torch.tensor([0, 1]).cuda())
    thr, fps, tps, distinct = ops.roc_counts(torch.zeros(0).cuda(), torch.zeros(0, dtype=torch.int64).cuda())
    assert thr.numel() == 0 and distinct == 0


@pytest.mark.gpu
def test_gpu_pair_cosine():
    from text_guided_face_recognition_b200 import ops
    g = load("embed")
    x1, x2 = torch.from_numpy(g["x1"]).cuda(), torch.from_numpy(g["x2"]).cuda()
    got = ops.pair_cosine(x1, x2).cpu().numpy()
    np.testing.assert_allclose(got, SO.pair_cosine(g["x1"], g["x2"]), atol=2e-6, rtol=0)
    np.testing.assert_allclose(got, g["scores"], atol=2e-6, rtol=0)            # torch's own CosineSimilarity
    assert got[5] == 0.0
    # config-5 width (640 = 512 face + 128 text), strided / unaligned / odd-width inputs through the generic kernel
    rs = np.random.RandomState(5)
    for n, d in ((1000, 640), (257, 1024), (33, 2048), (65, 37), (3, 1)):
        a, b = rs.randn(n, d).astype(np.float32), rs.randn(n, d).astype(np.float32)
        ref = SO.pair_cosine(a, b)
        ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        np.testing.assert_allclose(ops.pair_cosine(ta, tb).cpu().numpy(), ref, atol=2e-6, rtol=0)
        np.testing.assert_allclose(ops.pair_cosine(ta.t().contiguous().t(), tb).cpu().numpy(), ref, atol=2e-6, rtol=0)
        wide = torch.zeros(n, d + 3, device="cuda")
        wide[:, 1:d + 1] = ta
        np.testing.assert_allclose(ops.pair_cosine(wide[:, 1:d + 1], tb).cpu().numpy(), ref, atol=2e-6, rtol=0)


@pytest.mark.gpu
def test_gpu_identification():
    from text_guided_face_recognition_b200 import ops
    from text_guided_face_recognition_b200.utils import modules
    g = load("ident")
    args = types.SimpleNamespace(test_sub=int(g["total_sub"]), checkpoints_path=None)
    with contextlib.redirect_stdout(io.StringIO()):
        acc = modules.calculate_identification_acc(g["scores"].tolist(), args)
    assert acc == float(g["acc"])
    s = torch.from_numpy(g["scores"]).cuda().view(int(g["total_sub"]), -1)
    np.testing.assert_array_equal(ops.row_argmax(s).cpu().numpy(), g["argmax"])
    rs = np.random.RandomState(11)
    for rows, cols in ((6000, 10), (7, 1), (100, 1000), (3, 4097)):
        x = np.round(rs.rand(rows, cols) * 50).astype(np.float32)              # plenty of duplicated maxima
        x[rows // 2, cols // 2] = np.nan                                       # np.argmax: NaN is the maximum
        np.testing.assert_array_equal(ops.row_argmax(torch.from_numpy(x).cuda()).cpu().numpy(), np.argmax(x, axis=1))


@pytest.mark.gpu
def test_gpu_score_pairs_end_to_end():
    """Embeddings in, the reference's summary out: cosine on the device feeds the ROC without leaving it."""
    from text_guided_face_recognition_b200.utils import modules
    g = load("embed")
    x1, x2, lab = (torch.from_numpy(g[k]).cuda() for k in ("x1", "x2", "labels"))
    batches = [(x1[i:i + 128], x2[i:i + 128], lab[i:i + 128]) for i in range(0, x1.size(0), 128)]
    with contextlib.redirect_stdout(io.StringIO()):
        out = modules.score_pairs(batches)
    scores = modules.pair_scores(x1, x2).cpu().numpy()
    ref = SO.calculate_scores(scores, g["labels"])                              # same fp32 scores -> identical decisions
    assert out["auc"] == ref["auc"] and out["eer"] == ref["eer"] and out["tpr_at_fpr"] == ref["tpr_at_fpr"]
