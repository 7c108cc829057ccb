"""tgfr_matmul_split (csrc/gemm_tc.cu gemm_tc_pair): the hi / lo split tensor-core contraction behind IMIM's 1x1
convolutions, nn.Linear layers and torch.bmm calls (reference models/models.py:380-405, models/fusion_nets.py:97-115),
against float64 numpy on the shapes IMIM uses: NT / NN / TN, a strided batch of 196 x 196 x 256 samples, split-K over
K = B * 196 rows.  Tolerance: 2e-6 of the result's scale per entry (fp32 class), and -- because the tensor core's
accumulator truncates -- a check that column sums over thousands of rows carry no systematic bias beyond 2e-5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, mode):
    a, b = a.astype(np.float64), b.astype(np.float64)
    if mode == 0:
        return a @ np.swapaxes(b, -1, -2)
    if mode == 1:
        return a @ b
    return np.swapaxes(a, -1, -2) @ b


def _shapes(mode, M, N, K, batch):
    sa = (K, M) if mode == 2 else (M, K)
    sb = (N, K) if mode == 0 else (K, N)
    return ((batch,) + sa, (batch,) + sb) if batch else (sa, sb)


@pytest.mark.parametrize("mode,M,N,K,batch,splits", [
    (0, 25088, 768, 256, 0, 1),     # q | k | v projections at B = 128
    (0, 1000, 128, 256, 0, 1),      # ragged row tile
    (0, 196, 196, 256, 5, 1),       # attention scores per sample
    (1, 196, 256, 196, 5, 1),       # response = attention . value  (K = 196: partial K step)
    (1, 4000, 128, 256, 0, 1),      # d input of a 1x1 convolution
    (2, 196, 256, 196, 5, 1),       # d value = attention^T d response
    (2, 256, 128, 25088, 0, 37),    # weight gradient: K = B * 196 rows, split-K
    (2, 768, 256, 5000, 0, 8),
    (0, 37, 50, 24, 0, 1),          # smaller than one tile / one K step
])
@pytest.mark.parametrize("persistent", ["0", "1"])
def test_matmul_split_vs_float64(mode, M, N, K, batch, splits, persistent, monkeypatch):
    """persistent = 1: the one-CTA-per-SM kernel with double-buffered TMEM accumulators (TGFR_GEMM_PERSIST=1)."""
    from text_guided_face_recognition_b200 import ops
    monkeypatch.setenv("TGFR_GEMM_PERSIST", persistent)
    rs = np.random.RandomState(M + 7 * N + 13 * K + mode)
    sa, sb = _shapes(mode, M, N, K, batch)
    a = (rs.randn(*sa) * np.exp(rs.randn(*sa))).astype(np.float32)      # heavy-tailed: exercises the lo parts
    b = rs.randn(*sb).astype(np.float32) * 1e-3
    got = ops.matmul_split(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), mode=mode, splits=splits).cpu().numpy()
    ref = _ref(a, b, mode)
    scale = np.sqrt(np.mean(ref ** 2))
    bound = 1e-6 * (_ref(np.abs(a), np.abs(b), mode) + scale)            # fp32 class: 2^-22 per product, fp32 accumulation
    assert np.all(np.abs(got - ref) < bound)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-6


def test_matmul_split_bias_relu_alpha():
    from text_guided_face_recognition_b200 import ops
    rs = np.random.RandomState(5)
    a, b, bias = rs.randn(700, 256).astype(np.float32), rs.randn(128, 256).astype(np.float32), rs.randn(128).astype(np.float32)
    got = ops.matmul_split(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), mode=0, alpha=0.25,
                           bias=torch.from_numpy(bias).cuda(), relu=True).cpu().numpy()
    ref = np.maximum(0.25 * _ref(a, b, 0) + bias, 0.0)
    assert np.max(np.abs(got - ref)) < 2e-5


@pytest.mark.parametrize("mode", [0, 1])
def test_matmul_split_column_sums_unbiased(mode):
    """colsum over 25 088 rows of a product (IMIM's bias gradients): a truncating accumulator shows up here first."""
    from text_guided_face_recognition_b200 import ops
    rs = np.random.RandomState(11 + mode)
    M, N, K = 25088, 256, 256
    sa, sb = _shapes(mode, M, N, K, 0)
    a, b = rs.randn(*sa).astype(np.float32), (rs.randn(*sb) / 16).astype(np.float32)
    got = ops.matmul_split(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), mode=mode).cpu().numpy().astype(np.float64)
    ref = _ref(a, b, mode)
    err = (got - ref).sum(axis=0)
    print("mean entry error / rms entry:", (got - ref).mean() / np.sqrt(np.mean(ref ** 2)))
    assert np.max(np.abs(err)) < 2e-5 * np.sqrt(M) * np.sqrt(np.mean(ref ** 2))
