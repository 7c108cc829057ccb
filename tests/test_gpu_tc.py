"""GPU self-tests of the tcgen05 / TMA building blocks (csrc/tc.cuh): every UMMA operand layout the
fused kernels rely on (K-major and MN-major, TMA-written and thread-written with the hand-computed
128B swizzle) and the TMA reduce-add store, each checked against a plain matmul."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from text_guided_face_recognition_b200 import _lib as L
    L.ensure_device(torch.device("cuda", 0))
    return L


@pytest.mark.parametrize("a_mn,b_mn,manual_a,N,K", [
    (0, 0, 0, 128, 256),     # GEMM-1: scores  = regions(K-major) x words(K-major)
    (0, 0, 0, 64, 64),
    (1, 1, 0, 256, 128),     # GEMM-2: context = E^T(MN-major) x regions(MN-major)
    (1, 1, 1, 256, 128),     # ... with E written by the threads (hand swizzle)
    (1, 1, 1, 256, 64),
    (0, 1, 0, 256, 128),     # dC GEMM: dS(K-major) x words(MN-major)
    (1, 0, 0, 128, 64),
])
def test_umma_layouts(a_mn, b_mn, manual_a, N, K):
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(1234 + 7 * N + K + a_mn + 2 * b_mn)
    A = torch.randn(128, K, generator=g).half()
    B = torch.randn(N, K, generator=g).half()
    ref = A.float() @ B.float().t()
    a_dev = (A.t().contiguous() if a_mn else A.contiguous()).cuda()
    b_dev = (B.t().contiguous() if b_mn else B.contiguous()).cuda()
    out = torch.full((128, N), float("nan"), device="cuda")
    L.check(lib.tgfr_debug_umma(a_dev.data_ptr(), b_dev.data_ptr(), out.data_ptr(), N, K, a_mn, b_mn, manual_a,
                                torch.cuda.current_stream().cuda_stream), "tgfr_debug_umma")
    torch.cuda.synchronize()
    got = out.cpu()
    err = (got - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


def test_tma_reduce_add():
    L = _lib()
    lib = L.load()
    rows, cols = 128, 64
    init = torch.arange(rows * cols, dtype=torch.float32).reshape(rows, cols) * 0.5
    out = init.clone().cuda()
    L.check(lib.tgfr_debug_tma_reduce(out.data_ptr(), rows, cols, torch.cuda.current_stream().cuda_stream),
            "tgfr_debug_tma_reduce")
    torch.cuda.synchronize()
    r = torch.arange(rows, dtype=torch.float32)[:, None]
    c = torch.arange(cols, dtype=torch.float32)[None, :]
    assert torch.equal(out.cpu(), init + 1000.0 * r + c)
