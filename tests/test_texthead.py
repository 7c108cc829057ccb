"""TextHeading (reference models/models.py:170-232).

CPU: the fp64 oracle against the fixtures generated from the reference (tests/golden/make_golden_texthead.py).
GPU: the CUDA path (models/text_heading.py -> C ABI) against the fixtures and, at config-2 size, the oracle.
fp32 arithmetic: outputs within 2e-5, gradients within 1e-4 relative."""
import os
import types

import numpy as np
import pytest
import torch

import synth
from oracle import texthead_oracle as TO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def upstream(g):
    rs = np.random.RandomState(int(g["seed"]) + 17)
    B, wn, F = int(g["B"]), int(g["words_num"]), int(g["F"])
    return rs.randn(B, wn - 2, F).astype(np.float32), rs.randn(B, F).astype(np.float32)


@pytest.mark.parametrize("name", ["texthead_small", "texthead_bert24"])
def test_oracle_matches_reference_fixture(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, wn, F = int(g["B"]), int(g["words_num"]), int(g["F"])
    tokens, ws, bs = synth.texthead_inputs(B, wn, F, seed=int(g["seed"]))
    words, sent = TO.forward(tokens, ws, bs, wn)
    assert np.max(np.abs(words - g["words"])) < 2e-6 and np.max(np.abs(sent - g["sent"])) < 2e-6
    np.testing.assert_allclose(np.linalg.norm(words, axis=2), 1.0, atol=1e-12)
    gw, gs = upstream(g)
    dws, dbs = TO.backward(tokens, ws, bs, wn, gw, gs)
    for k in range(3):
        assert rel(dbs[k], g[f"db{k}"]) < 2e-5
        if f"dw{k}" in g.files:
            assert rel(dws[k], g[f"dw{k}"]) < 2e-5
        else:
            assert rel(dws[k][:4], g[f"dw{k}_head"]) < 2e-5
            assert abs(np.linalg.norm(dws[k]) - float(g[f"dw{k}_norm"])) < 2e-5 * float(g[f"dw{k}_norm"])


def test_oracle_last_word_is_detached():
    """models.py:206 copies the last word through torch.cuda.FloatTensor: no gradient flows through it."""
    tokens, ws, bs = synth.texthead_inputs(2, 10, 8, seed=5)
    gw = np.zeros((2, 8, 8), np.float32)
    gw[:, -1] = 1.0
    dws, dbs = TO.backward(tokens, ws, bs, 10, gw, None)
    assert all(not d.any() for d in dws) and all(not d.any() for d in dbs)


def _head(B, wn, F, seed):
    from text_guided_face_recognition_b200.models.text_heading import TextHeading
    tokens, ws, bs = synth.texthead_inputs(B, wn, F, seed=seed)
    th = TextHeading(types.SimpleNamespace(aux_feat_dim_per_granularity=F, bert_words_num=wn)).cuda()
    with torch.no_grad():
        for conv, w, b in zip(th.bwm.convs1, ws, bs):
            conv.weight.copy_(torch.from_numpy(w).unsqueeze(1))
            conv.bias.copy_(torch.from_numpy(b))
    return th, tokens, ws, bs


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["texthead_small", "texthead_bert24"])
def test_gpu_texthead_vs_reference_fixture(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, wn, F = int(g["B"]), int(g["words_num"]), int(g["F"])
    th, tokens, ws, bs = _head(B, wn, F, int(g["seed"]))
    words, sent = th(torch.from_numpy(tokens).cuda(), None)
    assert tuple(words.shape) == (B, F, wn - 2) and words.stride(1) == 1          # the reference's [B, F, T] view of [B, T, F]
    w = words.transpose(1, 2).detach().cpu().numpy()
    assert np.max(np.abs(w - g["words"])) < 2e-5 and np.max(np.abs(sent.detach().cpu().numpy() - g["sent"])) < 2e-5
    gw, gs = upstream(g)
    ((words.transpose(1, 2) * torch.from_numpy(gw).cuda()).sum() + (sent * torch.from_numpy(gs).cuda()).sum()).backward()
    for k, conv in enumerate(th.bwm.convs1):
        dw = conv.weight.grad.squeeze(1).cpu().numpy()
        assert rel(conv.bias.grad.cpu().numpy(), g[f"db{k}"]) < 1e-4
        if f"dw{k}" in g.files:
            assert rel(dw, g[f"dw{k}"]) < 1e-4
        else:
            assert rel(dw[:4], g[f"dw{k}_head"]) < 1e-4
            assert abs(np.linalg.norm(dw.astype(np.float64)) - float(g[f"dw{k}_norm"])) < 1e-4 * float(g[f"dw{k}_norm"])


@pytest.mark.gpu
def test_gpu_texthead_config2_size_vs_oracle_and_feeds_words_loss():
    """B = 128, bert_words_num = 24, F = 256 (configs[1]): parity with the oracle, and the [B, F, T] output goes
    straight into words_loss (same memory layout the reference's TextHeading produces)."""
    from text_guided_face_recognition_b200.models import losses
    B, wn, F = 128, 24, 256
    th, tokens, ws, bs = _head(B, wn, F, 7)
    words, sent = th(torch.from_numpy(tokens).cuda(), None)
    rw, rs_ = TO.forward(tokens, ws, bs, wn)
    assert np.max(np.abs(words.transpose(1, 2).detach().cpu().numpy() - rw)) < 2e-5
    assert np.max(np.abs(sent.detach().cpu().numpy() - rs_)) < 2e-5
    rng = np.random.RandomState(3)
    gw, gs = rng.randn(B, wn - 2, F).astype(np.float32), rng.randn(B, F).astype(np.float32)
    ((words.transpose(1, 2) * torch.from_numpy(gw).cuda()).sum() + (sent * torch.from_numpy(gs).cuda()).sum()).backward()
    dws, dbs = TO.backward(tokens, ws, bs, wn, gw, gs)
    for k, conv in enumerate(th.bwm.convs1):
        assert rel(conv.weight.grad.squeeze(1).cpu().numpy(), dws[k]) < 1e-4
        assert rel(conv.bias.grad.cpu().numpy(), dbs[k]) < 1e-4
    # downstream: the words feed the word-region loss with gradients reaching the head's weights
    ctx, _, _ = synth.wordregion_inputs(B, wn - 2, 196, F, "BERT", seed=1)
    feats = torch.from_numpy(ctx).cuda().view(B, 14, 14, F).permute(0, 3, 1, 2)
    ns = types.SimpleNamespace
    args = ns(en_type="BERT", bert_words_num=wn, CUDA=True, device="cuda",
              TRAIN=ns(SMOOTH=ns(GAMMA1=4.0, GAMMA2=5.0, GAMMA3=10.0)))
    for conv in th.bwm.convs1:
        conv.weight.grad = None
    words, _ = th(torch.from_numpy(tokens).cuda(), None)
    l0, l1, _ = losses.words_loss(feats, words, torch.arange(B).cuda(), None, None, B, args)
    (l0 + l1).backward()
    assert all(torch.isfinite(c.weight.grad).all() and c.weight.grad.abs().sum() > 0 for c in th.bwm.convs1)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["tc", "mixed", "fp32"])
def test_gpu_texthead_modes_at_config2_size(monkeypatch, mode):
    """Every arithmetic mode of the products (TGFR_TEXTHEAD_PRECISION) meets the same bar at B = 128: outputs within
    1e-5, weight / bias gradients within 1e-4 of the fp64 oracle.  In the all-tensor-core mode this needs the exact
    re-evaluation of the near-tied activations: without it about two arg-max decisions in 720 k flip (dW off by 3e-3)."""
    monkeypatch.setenv("TGFR_TEXTHEAD_PRECISION", mode)
    B, wn, F = 128, 24, 256
    th, tokens, ws, bs = _head(B, wn, F, 7)
    words, sent = th(torch.from_numpy(tokens).cuda(), None)
    rw, rs_ = TO.forward(tokens, ws, bs, wn)
    assert np.max(np.abs(words.transpose(1, 2).detach().cpu().numpy() - rw)) < 1e-5
    assert np.max(np.abs(sent.detach().cpu().numpy() - rs_)) < 1e-5
    rng = np.random.RandomState(3)
    gw, gs = rng.randn(B, wn - 2, F).astype(np.float32), rng.randn(B, F).astype(np.float32)
    ((words.transpose(1, 2) * torch.from_numpy(gw).cuda()).sum() + (sent * torch.from_numpy(gs).cuda()).sum()).backward()
    dws, dbs = TO.backward(tokens, ws, bs, wn, gw, gs)
    for k, conv in enumerate(th.bwm.convs1):
        assert rel(conv.weight.grad.squeeze(1).cpu().numpy(), dws[k]) < 1e-4
        assert rel(conv.bias.grad.cpu().numpy(), dbs[k]) < 1e-4
