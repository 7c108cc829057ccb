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
    (0, 1, 2, 64, 128),      # A operand staged in TMEM by tcgen05.st (dS / E16 in the backward pass)
    (0, 1, 2, 256, 128),
    (0, 0, 2, 128, 256),
    (0, 2, 0, 128, 256),     # B as planes [K/8][N][8] (one bulk copy), K-major without swizzle: dE~ = C x V^T
    (0, 3, 2, 256, 128),     # the same storage read MN-major, A in TMEM: d ctx += Ek x V
    (0, 3, 0, 256, 128),
])
def test_umma_layouts(a_mn, b_mn, manual_a, N, K):
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(1234 + 7 * N + K + a_mn + 2 * b_mn)
    A = torch.randn(128, K, generator=g).half()
    B = torch.randn(N, K, generator=g).half()
    ref = A.float() @ B.float().t()
    a_dev = (A.t().contiguous() if a_mn else A.contiguous()).cuda()
    if b_mn == 2:
        b_dev = B.reshape(N, K // 8, 8).permute(1, 0, 2).contiguous().cuda()
    elif b_mn == 3:
        b_dev = B.t().reshape(K, N // 8, 8).permute(1, 0, 2).contiguous().cuda()
    else:
        b_dev = (B.t().contiguous() if b_mn else B.contiguous()).cuda()
    out = torch.full((128, N), float("nan"), device="cuda")
    L.check(lib.tgfr_debug_umma(a_dev.data_ptr(), b_dev.data_ptr(), out.data_ptr(), N, K, a_mn, b_mn, manual_a,
                                torch.cuda.current_stream().cuda_stream), "tgfr_debug_umma")
    torch.cuda.synchronize()
    got = out.cpu()
    err = (got - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.parametrize("N,K", [(128, 64), (128, 256), (256, 128), (64, 64)])
def test_umma_cta_pair(N, K):
    """cta_group::2: one M = 256 MMA spans a cluster of two CTAs (each: its 128 rows of A, half of B, its half of D)."""
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(77 + N + K)
    A = torch.randn(256, K, generator=g).half()
    B = torch.randn(N, K, generator=g).half()
    ref = A.float() @ B.float().t()
    a_dev, b_dev = A.cuda(), B.cuda()
    out = torch.full((256, N), float("nan"), device="cuda")
    L.check(lib.tgfr_debug_umma_2cta(a_dev.data_ptr(), b_dev.data_ptr(), out.data_ptr(), N, K,
                                     torch.cuda.current_stream().cuda_stream), "tgfr_debug_umma_2cta")
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().max().item()
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


# ------------------------------------------------------------------------------------------------
# fused tensor-core word-region kernels (TGFR_PREC_TC) vs the fp64 oracle and the fp32 SIMT path
# fp16 operands / fp32 accumulation: |sim| ~ 30 is reproduced to ~1e-3 absolute, the losses to
# ~1e-5 relative (contract: 1e-4), see DESIGN.md "numerics".
# ------------------------------------------------------------------------------------------------
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth  # noqa: E402
from oracle import fcam_oracle as O  # noqa: E402

TC_SIM_ATOL = 5e-3
TC_LOSS_RTOL = 1e-4


@pytest.mark.parametrize("B,T,R,D,flavour,ragged", [
    (8, 22, 196, 256, "BERT", False),
    (16, 18, 196, 256, "LSTM", True),
    (6, 7, 16, 64, "LSTM", True),
    (5, 30, 49, 128, "BERT", False),
    (128, 22, 196, 256, "BERT", False),       # BASELINE config 2
])
def test_tc_forward_vs_oracle(B, T, R, D, flavour, ragged):
    from text_guided_face_recognition_b200 import _lib, ops
    ctx, words, cap = synth.wordregion_inputs(B, T, R, D, flavour, seed=100, ragged=ragged)
    feats = torch.from_numpy(ctx).cuda()
    wd = torch.from_numpy(words).cuda()
    capt = None if cap is None else torch.from_numpy(cap).cuda()
    sim, attn = ops.wordregion_sim(feats, wd, capt, 4.0, 5.0, 10.0, precision=_lib.PREC_TC)
    sim32, attn32 = ops.wordregion_sim(feats, wd, capt, 4.0, 5.0, 10.0, precision=_lib.PREC_FP32)
    torch.cuda.synchronize()
    ref, ref_att = O.wordregion_sim(ctx, words, cap, 4.0, 5.0, 10.0)
    got = sim.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.max(np.abs(got - ref)) < TC_SIM_ATOL, np.max(np.abs(got - ref))
    assert np.max(np.abs(sim32.cpu().numpy() - ref)) < 2e-4
    l0, l1 = ops.pair_ce(sim)
    r0, r1 = O.pair_ce(ref)
    assert abs(l0.item() - r0) < TC_LOSS_RTOL * r0 and abs(l1.item() - r1) < TC_LOSS_RTOL * r1
    # the diagonal attention maps are emitted by the tensor-core forward itself (epi-1 already holds their numerators):
    # fp16-operand scores, so TF32 class -- 2e-3 of a map's largest entry; TGFR_ATTN_MAPS=fp32 selects the exact kernel
    a_tc, a_32 = attn.cpu().numpy(), attn32.cpu().numpy()
    assert np.isfinite(a_tc).all()
    for i, a in enumerate(ref_att):
        assert np.max(np.abs(a_32[i, : a.shape[0]] - a)) < 1e-5
        assert np.max(np.abs(a_tc[i, : a.shape[0]] - a)) < 2e-3 * np.max(a), (i, np.max(np.abs(a_tc[i, : a.shape[0]] - a)), np.max(a))
        assert np.max(np.abs(a_tc[i, : a.shape[0]].sum(axis=1) - 1.0)) < 1e-5         # rows are distributions over the regions
        assert not a_tc[i, a.shape[0]:].any()                                          # zero beyond the caption's length


def test_tc_attention_maps_fp32_switch(monkeypatch):
    """TGFR_ATTN_MAPS=fp32: the maps come from the exact fp32 kernel, bit-identical to the fp32 mode's."""
    from text_guided_face_recognition_b200 import _lib, ops
    monkeypatch.setenv("TGFR_ATTN_MAPS", "fp32")
    ctx, words, cap = synth.wordregion_inputs(16, 18, 196, 256, "LSTM", seed=100, ragged=True)
    feats, wd, capt = torch.from_numpy(ctx).cuda(), torch.from_numpy(words).cuda(), torch.from_numpy(cap).cuda()
    _, attn = ops.wordregion_sim(feats, wd, capt, 4.0, 5.0, 10.0, precision=_lib.PREC_TC)
    _, attn32 = ops.wordregion_sim(feats, wd, capt, 4.0, 5.0, 10.0, precision=_lib.PREC_FP32)
    assert torch.allclose(attn, attn32, rtol=1e-5, atol=1e-8)


TC_GRAD_RTOL = 1e-3      # contract for the TF32-class path; emulation of the fp16-operand pipeline gives ~4e-4


@pytest.fixture(params=["save", "wu", "recompute"])
def bwd_mode(request, monkeypatch):
    """What the forward leaves for the backward: its fp16 word-softmax / attention records (default), only the Wu
    tiles (TGFR_WORDREGION_SAVE=wu: the backward recomputes the scores), or nothing (full recomputation)."""
    monkeypatch.setenv("TGFR_WORDREGION_SAVE", {"save": "1", "wu": "wu", "recompute": "0"}[request.param])
    return request.param


@pytest.mark.parametrize("B,T,R,D,flavour,ragged", [
    (8, 22, 196, 256, "BERT", False),
    (16, 18, 196, 256, "LSTM", True),
    (6, 7, 16, 64, "LSTM", True),
    (5, 30, 49, 128, "BERT", False),
    (7, 12, 130, 192, "LSTM", True),
    (32, 22, 196, 256, "BERT", False),
])
def test_tc_backward_vs_oracle(B, T, R, D, flavour, ragged, bwd_mode):
    from text_guided_face_recognition_b200 import _lib, ops
    ctx, words, cap = synth.wordregion_inputs(B, T, R, D, flavour, seed=100, ragged=ragged)
    feats = torch.from_numpy(ctx).cuda().requires_grad_(True)
    wd = torch.from_numpy(words).cuda()                     # text side detached, as in the reference's scripts
    capt = None if cap is None else torch.from_numpy(cap).cuda()
    sim, _ = ops.wordregion_sim(feats, wd, capt, 4.0, 5.0, 10.0, precision=_lib.PREC_TC, want_attn=False)
    l0, l1 = ops.pair_ce(sim)
    (l0 + l1).backward()
    torch.cuda.synchronize()
    got = feats.grad.cpu().numpy()
    assert np.isfinite(got).all()
    ref, _ = O.words_loss_grads(ctx, words, None, cap, 4.0, 5.0, 10.0)
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err < TC_GRAD_RTOL, err
    worst = max(np.linalg.norm(got[b] - ref[b]) / np.linalg.norm(ref[b]) for b in range(B))
    assert worst < 2 * TC_GRAD_RTOL, worst


@pytest.mark.parametrize("B,T,R,D,flavour,ragged", [
    (8, 22, 196, 256, "BERT", False),
    (16, 18, 196, 256, "LSTM", True),
    (6, 7, 16, 64, "LSTM", True),
    (5, 30, 49, 128, "BERT", False),
    (7, 12, 130, 192, "LSTM", True),
    (32, 22, 196, 256, "BERT", False),
])
def test_tc_backward_words_vs_oracle(B, T, R, D, flavour, ragged, bwd_mode):
    """Text-side gradient on the tensor cores (DQ pass), alone and together with the face-side gradient."""
    from text_guided_face_recognition_b200 import _lib, ops
    ctx, words, cap = synth.wordregion_inputs(B, T, R, D, flavour, seed=100, ragged=ragged)
    capt = None if cap is None else torch.from_numpy(cap).cuda()
    rc, rw = O.words_loss_grads(ctx, words, None, cap, 4.0, 5.0, 10.0)
    for both in (False, True):
        feats = torch.from_numpy(ctx).cuda().requires_grad_(both)
        wd = torch.from_numpy(words).cuda().requires_grad_(True)
        sim, _ = ops.wordregion_sim(feats, wd, capt, 4.0, 5.0, 10.0, precision=_lib.PREC_TC, want_attn=False)
        l0, l1 = ops.pair_ce(sim)
        (l0 + l1).backward()
        torch.cuda.synchronize()
        got = wd.grad.cpu().numpy()
        assert np.isfinite(got).all()
        err = np.linalg.norm(got - rw) / np.linalg.norm(rw)
        assert err < TC_GRAD_RTOL, err
        if cap is not None:                                   # words beyond a caption's length get exactly zero
            for i in range(B):
                assert not got[i, int(cap[i]):].any()
        if both:
            errc = np.linalg.norm(feats.grad.cpu().numpy() - rc) / np.linalg.norm(rc)
            assert errc < TC_GRAD_RTOL, errc


# ------------------------------------------------------------------------------------------------
# gradient parity on the exact instances bench.py times: BASELINE config 2
# (B=128, T=22 -> Tp=24, R=196, D=256; 148 persistent CTAs, CTA pairs, TMA reduce-add ordering) and the
# config-4 caption length (T=30 -> Tp=32).  The fp64 oracle runs once per session (~2 min at B=128).
# ------------------------------------------------------------------------------------------------
_ORACLE_CACHE = {}


def _oracle_grads(B, T, R, D, seed):
    key = (B, T, R, D, seed)
    if key not in _ORACLE_CACHE:
        ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=seed)
        _ORACLE_CACHE[key] = (ctx, words) + tuple(O.words_loss_grads(ctx, words, None, None, 4.0, 5.0, 10.0))
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize("B,T", [(128, 22), (32, 30), (48, 30)])
@pytest.mark.parametrize("grads", ["ctx", "both"])
def test_tc_grads_benchmarked_instances_vs_oracle(B, T, grads, bwd_mode):
    from text_guided_face_recognition_b200 import _lib, ops
    R, D = 196, 256
    ctx, words, rc, rw = _oracle_grads(B, T, R, D, 100)
    feats = torch.from_numpy(ctx).cuda().requires_grad_(True)
    wd = torch.from_numpy(words).cuda().requires_grad_(grads == "both")
    sim, _ = ops.wordregion_sim(feats, wd, None, 4.0, 5.0, 10.0, precision=_lib.PREC_TC, want_attn=False)
    l0, l1 = ops.pair_ce(sim)
    (l0 + l1).backward()
    torch.cuda.synchronize()
    got = feats.grad.cpu().numpy()
    assert np.isfinite(got).all()
    err = np.linalg.norm(got - rc) / np.linalg.norm(rc)
    assert err < TC_GRAD_RTOL, err
    worst = max(np.linalg.norm(got[b] - rc[b]) / np.linalg.norm(rc[b]) for b in range(B))
    assert worst < 2 * TC_GRAD_RTOL, worst
    if grads == "both":
        gw = wd.grad.cpu().numpy()
        errw = np.linalg.norm(gw - rw) / np.linalg.norm(rw)
        assert errw < TC_GRAD_RTOL, errw


@pytest.mark.parametrize("save", ["1", "wu", "0"])
def test_tc_operand_range_guard(save, monkeypatch):
    """The fp16 operand copies carry no per-tensor scale (the reference's inputs are unit-norm rows, models/models.py:119,
    403).  Inside the range fp16 holds well the result must not depend on how the magnitude is split between the two
    operands (ctx x 64, words / 16 and a rescaled gamma give the same sim); outside it (largest entry above 65 504 or below
    2^-9: x1e6, x1e-4) the tensor-core forward must fail LOUDLY -- every sim NaN -- while the fp32 path stays exact."""
    from text_guided_face_recognition_b200 import _lib, ops
    monkeypatch.setenv("TGFR_WORDREGION_SAVE", save)
    B, T, R, D = 8, 22, 196, 256
    ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=100)
    f, w = torch.from_numpy(ctx).cuda(), torch.from_numpy(words).cuda()

    def sim_of(ff, ww, prec, grad=False):
        ff = ff.clone().requires_grad_(grad)
        s, _ = ops.wordregion_sim(ff, ww, None, 4.0, 5.0, 10.0, precision=prec, want_attn=False)
        return s

    base = sim_of(f, w, _lib.PREC_TC, grad=True)
    assert torch.isfinite(base).all()
    # <c * 64, q / 64> = <c, q>: power-of-two rescaling is exact in fp16 while everything stays normal
    moved = sim_of(f * 64.0, w / 64.0, _lib.PREC_TC, grad=True)
    assert torch.isfinite(moved).all()
    assert float((moved - base).detach().abs().max()) < 5e-3 * float(base.detach().abs().max())
    for scale in (1e6, 1e-4):
        for ff, ww in ((f * scale, w), (f, w * scale)):
            bad = sim_of(ff, ww, _lib.PREC_TC, grad=True)
            assert torch.isnan(bad).all(), scale
            ok32 = sim_of(ff, ww, _lib.PREC_FP32)
            assert torch.isfinite(ok32).all()
