"""Pin oracle/fcam_oracle.py against fixtures produced by the reference's own PyTorch code
(tests/golden/make_golden.py).  The reference computed in fp32; the oracle in fp64, so the
tolerances below are fp32 round-off of the *reference*, not slack in the oracle."""
import os

import numpy as np
import pytest

import synth
from oracle import fcam_oracle as O

RTOL_LOSS = 2e-6
RTOL_GRAD = 2e-5      # ||delta|| / ||ref||


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def test_func_attention(golden_dir):
    g = load(golden_dir, "attention_small")
    wc, attn = O.func_attention(g["query"], g["context"], float(g["gamma1"]))
    assert rel(wc, g["wc"]) < 1e-6
    assert rel(attn, g["attn"]) < 1e-6
    dq, dc = O.func_attention_bwd(g["query"], g["context"], float(g["gamma1"]), g["gw"], g["ga"])
    assert rel(dq, g["dquery"]) < RTOL_GRAD
    assert rel(dc, g["dcontext"]) < RTOL_GRAD


@pytest.mark.parametrize("name", ["wordregion_bert_small", "wordregion_lstm_ragged"])
def test_wordregion_small(golden_dir, name):
    g = load(golden_dir, name)
    g1, g2, g3 = (float(v) for v in g["gammas"])
    cap = g["cap_lens"] if g["cap_lens"].size else None
    l0, l1, attn, _ = O.words_loss(g["ctx"], g["words"], None, cap, g1, g2, g3)
    assert abs(l0 - float(g["loss0"])) < RTOL_LOSS * abs(float(g["loss0"]))
    assert abs(l1 - float(g["loss1"])) < RTOL_LOSS * abs(float(g["loss1"]))
    for i, a in enumerate(attn):
        assert rel(a, g["att"][i, : a.shape[0]]) < 2e-6
    dctx, dwords = O.words_loss_grads(g["ctx"], g["words"], None, cap, g1, g2, g3,
                                      float(g["w0"]), float(g["w1"]))
    assert rel(dctx, g["dctx"]) < RTOL_GRAD
    assert rel(dwords, g["dwords"]) < RTOL_GRAD


@pytest.mark.parametrize("name", ["wordregion_bert_mid", "wordregion_config1"])
def test_wordregion_full_width(golden_dir, name):
    g = load(golden_dir, name)
    B, T, D = int(g["B"]), int(g["T"]), int(g["D"])
    R = int(g["ih"]) * int(g["iw"])
    ctx, words, cap = synth.wordregion_inputs(B, T, R, D, str(g["flavour"]), seed=100, ragged=False)
    g1, g2, g3 = (float(v) for v in g["gammas"])
    l0, l1, attn, _ = O.words_loss(ctx, words, None, cap, g1, g2, g3)
    assert abs(l0 - float(g["loss0"])) < 5e-6 * abs(float(g["loss0"]))
    assert abs(l1 - float(g["loss1"])) < 5e-6 * abs(float(g["loss1"]))
    assert rel(np.stack(attn), g["att"]) < 5e-6
    dctx, dwords = O.words_loss_grads(ctx, words, None, cap, g1, g2, g3)
    assert rel(dwords, g["dwords"]) < RTOL_GRAD
    assert rel(dctx[:2], g["dctx_head"]) < RTOL_GRAD
    proj = np.random.RandomState(7).randn(D).astype(np.float32)
    assert rel(dctx @ proj.astype(np.float64), g["dctx_proj"]) < 5e-5
    assert abs(np.linalg.norm(dctx) - float(g["dctx_norm"])) < 1e-5 * float(g["dctx_norm"])


@pytest.mark.parametrize("name", ["sentence_plain", "sentence_collisions"])
def test_sentence_losses(golden_dir, name):
    g = load(golden_dir, name)
    img, txt, cid = g["img"], g["txt"], g["class_ids"]
    l0, l1, sc = O.sent_loss(img, txt, None, cid, float(g["gamma3"]))
    assert abs(l0 - float(g["sent_loss0"])) < RTOL_LOSS * abs(l0)
    assert abs(l1 - float(g["sent_loss1"])) < RTOL_LOSS * abs(l1)
    if name.endswith("collisions"):
        assert np.isinf(sc).any(), "fixture must exercise the class-id mask"
    dx, dy = O.sent_loss_grads(img, txt, None, cid, float(g["gamma3"]), 1.0, 0.5)
    assert rel(dx, g["sent_dimg"]) < RTOL_GRAD
    assert rel(dy, g["sent_dtxt"]) < RTOL_GRAD
    assert abs(O.global_loss(img, txt) - float(g["global_loss"])) < RTOL_LOSS * float(g["global_loss"])
    dx, dy = O.sent_loss_grads(img, txt, None, None, 10.0, 1.0, 1.0)
    assert rel(dx, g["global_dimg"]) < RTOL_GRAD
    assert rel(dy, g["global_dtxt"]) < RTOL_GRAD
    a, b = img * 3.0, txt * 2.0
    assert abs(O.clip_loss(b, a, 1.0) - float(g["clip_loss"])) < RTOL_LOSS * float(g["clip_loss"])
    sc = O.cosine_scores(a, b, 1.0, normalise=False)
    gs = O.pair_ce_bwd(sc, None, 0.5, 0.5)
    dx, dy = O.cosine_scores_bwd(a, b, 1.0, gs, normalise=False)
    assert rel(dx, g["clip_dimg"]) < RTOL_GRAD
    assert rel(dy, g["clip_dtxt"]) < RTOL_GRAD


@pytest.mark.parametrize("name", ["arc_small", "arc_small_easy"])
def test_arc_margin_small(golden_dir, name):
    g = load(golden_dir, name)
    s, m, easy = float(g["s"]), float(g["m"]), bool(g["easy"])
    logits = O.arc_margin(g["x"], g["weight"], g["label"], s, m, easy)
    assert np.max(np.abs(logits - g["logits"])) < 2e-5
    assert np.array_equal(logits.argmax(1), g["argmax"])
    loss = O.focal_loss(logits, g["label"], float(g["gamma"]))
    assert abs(loss - float(g["loss"])) < RTOL_LOSS * abs(loss)
    gl = O.focal_loss_bwd(logits, g["label"], float(g["gamma"]))
    dx, dw = O.arc_margin_bwd(g["x"], g["weight"], g["label"], gl, s, m, easy)
    assert rel(dx, g["dx"]) < RTOL_GRAD
    assert rel(dw, g["dweight"]) < RTOL_GRAD


def test_arc_margin_mid(golden_dir):
    g = load(golden_dir, "arc_mid")
    B, Din, C = int(g["B"]), int(g["Din"]), int(g["C"])
    x, w, label = synth.margin_inputs(B, Din, C, seed=100)
    assert np.array_equal(label, g["label"])
    s, m, easy = float(g["s"]), float(g["m"]), bool(g["easy"])
    logits = O.arc_margin(x, w, label, s, m, easy)
    assert np.max(np.abs(logits[:8] - g["logits_head"])) < 2e-5
    assert np.array_equal(logits.argmax(1), g["argmax"])
    loss = O.focal_loss(logits, label, 2.0)
    assert abs(loss - float(g["loss"])) < RTOL_LOSS * abs(loss)
    dx, dw = O.arc_margin_bwd(x, w, label, O.focal_loss_bwd(logits, label, 2.0), s, m, easy)
    assert rel(dx, g["dx"]) < RTOL_GRAD
    assert rel(dw[:64], g["dweight_head"]) < RTOL_GRAD
    assert abs(np.linalg.norm(dw) - float(g["dweight_norm"])) < 1e-5 * float(g["dweight_norm"])


@pytest.mark.parametrize("name", ["mag_small_easy", "mag_small_hard"])
def test_mag_head(golden_dir, name):
    g = load(golden_dir, name)
    kw = dict(l_a=float(g["l_a"]), u_a=float(g["u_a"]), l_margin=float(g["l_margin"]),
              u_margin=float(g["u_margin"]), scale=float(g["scale"]), easy_margin=bool(g["easy"]))
    (cos, cos_m), xn = O.mag_linear(g["x"], g["weight"], **kw)
    assert np.max(np.abs(cos - g["cos"])) < 5e-5
    assert np.max(np.abs(cos_m - g["cos_m"])) < 5e-5
    assert rel(xn, g["x_norm"]) < 1e-6
    loss, loss_g, one_hot = O.mag_loss((cos, cos_m), g["label"], xn, kw["u_a"])
    assert abs(loss - float(g["loss"])) < RTOL_LOSS * abs(loss)
    assert abs(loss_g - float(g["loss_g"])) < RTOL_LOSS * abs(loss_g)
    assert np.array_equal(one_hot, g["one_hot"])
    dx, dw = O.mag_head_grads(g["x"], g["weight"], g["label"], g_loss=1.0, g_lossg=float(g["lam_g"]), **kw)
    assert rel(dx, g["dx"]) < RTOL_GRAD
    assert rel(dw, g["dweight"]) < RTOL_GRAD


def test_ref_port_matches_golden(golden_dir):
    """oracle/ref_port.py (the CPU-baseline arm) reproduces the real reference's numbers."""
    import torch
    from oracle import ref_port as P
    g = load(golden_dir, "wordregion_lstm_ragged")
    B, T, ih, iw = int(g["B"]), int(g["T"]), int(g["ih"]), int(g["iw"])
    c = torch.from_numpy(g["ctx"]).clone().requires_grad_(True)
    w = torch.from_numpy(g["words"]).clone().requires_grad_(True)
    g1, g2, g3 = (float(v) for v in g["gammas"])
    l0, l1, att = P.words_loss_port(c.view(B, ih, iw, -1).permute(0, 3, 1, 2), w.transpose(1, 2), torch.arange(B),
                                    [int(v) for v in g["cap_lens"]], g1, g2, g3)
    assert abs(l0.item() - float(g["loss0"])) < 1e-6 * abs(float(g["loss0"]))
    assert abs(l1.item() - float(g["loss1"])) < 1e-6 * abs(float(g["loss1"]))
    (float(g["w0"]) * l0 + float(g["w1"]) * l1).backward()
    assert rel(c.grad.numpy(), g["dctx"]) < 1e-6
    assert rel(w.grad.numpy(), g["dwords"]) < 1e-6
    s = load(golden_dir, "sentence_collisions")
    a = torch.from_numpy(s["img"]).clone().requires_grad_(True)
    b = torch.from_numpy(s["txt"]).clone().requires_grad_(True)
    s0, s1 = P.sent_loss_port(a, b, torch.arange(a.shape[0]), s["class_ids"], float(s["gamma3"]))
    assert abs(s0.item() - float(s["sent_loss0"])) < 1e-6 and abs(s1.item() - float(s["sent_loss1"])) < 1e-6
    h = load(golden_dir, "arc_small")
    x = torch.from_numpy(h["x"])
    out, loss = P.arc_focal_port(x, torch.from_numpy(h["weight"]), torch.from_numpy(h["label"]), float(h["s"]),
                                 float(h["m"]), float(h["gamma"]), bool(h["easy"]))
    assert np.max(np.abs(out.numpy() - h["logits"])) < 1e-5 and abs(loss.item() - float(h["loss"])) < 1e-5


# ------------------------------------------------------------------------------------------------
# round-2 fixtures (tests/golden/make_golden_r2.py)
# ------------------------------------------------------------------------------------------------
def test_cosine_similarity(golden_dir):
    g = load(golden_dir, "cosine_small")
    out = O.cosine_similarity(g["x1"], g["x2"])
    assert np.max(np.abs(out - g["out"])) < 1e-6
    d1, d2 = O.cosine_similarity_bwd(g["x1"], g["x2"], g["g"])
    assert rel(d1, g["dx1"]) < RTOL_GRAD and rel(d2, g["dx2"]) < RTOL_GRAD


def test_mag_head_config3(golden_dir):
    """MagLinear(512, 10177) + MagLoss, B = 512 (BASELINE configs[2]) as the reference computes it."""
    g = load(golden_dir, "mag_config3")
    B, Din, C = int(g["B"]), int(g["Din"]), int(g["C"])
    x, w, label = synth.margin_inputs(B, Din, C, seed=100, mag=True)
    x = x * 4.0
    kw = dict(l_a=float(g["l_a"]), u_a=float(g["u_a"]), l_margin=float(g["l_margin"]),
              u_margin=float(g["u_margin"]), scale=float(g["scale"]), easy_margin=True)
    (cos, cos_m), xn = O.mag_linear(x, w, **kw)
    assert np.max(np.abs(cos[:8] - g["cos_head"])) < 5e-5 and np.max(np.abs(cos_m[:8] - g["cos_m_head"])) < 5e-5
    loss, loss_g, one_hot = O.mag_loss((cos, cos_m), label, xn, kw["u_a"])
    assert abs(loss - float(g["loss"])) < RTOL_LOSS * abs(loss)
    assert abs(loss_g - float(g["loss_g"])) < RTOL_LOSS * abs(loss_g)
    assert one_hot.sum() == float(g["one_hot_sum"])
    dx, dw = O.mag_head_grads(x, w, label, g_loss=1.0, g_lossg=float(g["lam_g"]), **kw)
    assert rel(dx, g["dx"]) < RTOL_GRAD
    assert rel(dw[:, :64], g["dweight_head"]) < RTOL_GRAD
    assert abs(np.linalg.norm(dw) - float(g["dweight_norm"])) < 1e-5 * float(g["dweight_norm"])


def test_train_block_composition(golden_dir):
    """The reference's Train.train loss block (src/train_encoders_bert.py:267-323) equals the oracle's composition
    words_loss + sent_loss + lambda_id (focal(arc(sent)) + focal(arc(img))) + lambda_clip global_loss."""
    import sys
    sys.path.insert(0, golden_dir)
    from make_golden_r2 import train_block_inputs
    g = load(golden_dir, "train_block_bert")
    B, T, D, C = int(g["B"]), int(g["T"]), int(g["D"]), int(g["C"])
    ctx, words, img, txt, cid, w_img, w_txt = train_block_inputs(B, T, D, C)
    w0, w1, _, _ = O.words_loss(ctx, words, None, None, 4.0, 5.0, 10.0)
    s0, s1, _ = O.sent_loss(img, txt, None, cid, 10.0)
    t_logits = O.arc_margin(txt, w_txt, cid, 35.0, 0.5, False)
    i_logits = O.arc_margin(img, w_img, cid, 30.0, 0.5, False)
    tid, iid = O.focal_loss(t_logits, cid, 2.0), O.focal_loss(i_logits, cid, 2.0)
    cl = O.global_loss(img, txt)
    total = w0 + w1 + s0 + s1 + 100 * (tid + iid) + 2.0 * cl
    assert abs(total - float(g["total_loss"])) < 5e-6 * abs(total)
    assert abs((w0 + w1) - float(g["w_total"])) < 5e-6 * (w0 + w1)
    dctx, _ = O.words_loss_grads(ctx, words, None, None, 4.0, 5.0, 10.0)
    got = g["d_words_features"].transpose(0, 2, 3, 1).reshape(B, 196, D)
    assert rel(got, dctx) < RTOL_GRAD
    dimg_s, _ = O.sent_loss_grads(img, txt, None, cid, 10.0)
    dimg_i, dw_i = O.arc_margin_bwd(img, w_img, cid, O.focal_loss_bwd(i_logits, cid, 2.0, 100.0), 30.0, 0.5, False)
    _, dw_t = O.arc_margin_bwd(txt, w_txt, cid, O.focal_loss_bwd(t_logits, cid, 2.0, 100.0), 35.0, 0.5, False)
    assert rel(dw_i, g["d_image_cls"]) < RTOL_GRAD and rel(dw_t, g["d_text_cls"]) < RTOL_GRAD
    dimg_c = O.global_loss_grads(img, txt)[0] if hasattr(O, "global_loss_grads") else None
    if dimg_c is not None:
        assert rel(dimg_s + dimg_i + 2.0 * dimg_c, g["d_img_features"]) < RTOL_GRAD
