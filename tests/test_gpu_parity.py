"""GPU parity tests: the CUDA path (through the reference-shaped Python API, i.e. through the C ABI)
against (a) the golden fixtures produced by the reference's own PyTorch code and (b) the fp64
oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): losses <= 1e-4 relative, gradients <= 1e-3 relative
(||delta|| / ||ref||) for the fp32 path; argmax decisions bit-exact.  The fp32 SIMT path is in
practice ~1e-6 / ~1e-5; the asserts below use tighter bounds than the contract where it holds.
"""
import os
import types

import numpy as np
import pytest
import torch

import synth
from oracle import fcam_oracle as O

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_RTOL = 1e-3
FP32_LOSS_RTOL = 2e-5     # what the fp32 path actually has to hold
FP32_GRAD_RTOL = 1e-4


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def make_args(flavour, T, g=(4.0, 5.0, 10.0)):
    ns = types.SimpleNamespace
    return ns(en_type=flavour, bert_words_num=T + 2, CUDA=True, device="cuda",
              TRAIN=ns(SMOOTH=ns(GAMMA1=g[0], GAMMA2=g[1], GAMMA3=g[2])))


def ref_layout(ctx, words, ih, iw, channels_last=True):
    """canonical numpy -> leaf tensors + reference-shaped views."""
    B, R, D = ctx.shape
    if channels_last:
        c = torch.from_numpy(ctx).cuda().requires_grad_(True)            # memory [B,ih,iw,D]
        c_ref = c.view(B, ih, iw, D).permute(0, 3, 1, 2)
        w = torch.from_numpy(words).cuda().requires_grad_(True)          # memory [B,T,D]
        w_ref = w.transpose(1, 2)
        def grads():
            return c.grad.cpu().numpy(), w.grad.cpu().numpy()
    else:                                                                # logically contiguous (after a gather)
        c = torch.from_numpy(np.ascontiguousarray(ctx.transpose(0, 2, 1))).cuda().requires_grad_(True)
        c_ref = c.view(B, D, ih, iw)
        w = torch.from_numpy(np.ascontiguousarray(words.transpose(0, 2, 1))).cuda().requires_grad_(True)
        w_ref = w
        def grads():
            return (c.grad.cpu().numpy().transpose(0, 2, 1), w.grad.cpu().numpy().transpose(0, 2, 1))
    return c_ref, w_ref, grads


@pytest.fixture(params=["fp32", "tc"])
def prec(request, monkeypatch):
    """Runs a word-region test once per arithmetic mode: fp32 = SIMT kernels (tight bounds), tc = tcgen05
    kernels with fp16 operands / fp32 accumulation (the contract's 1e-4 / 1e-3).  Shapes the tensor-core
    kernel does not take (D % 64 != 0) run the fp32 kernels in both modes."""
    monkeypatch.setenv("TGFR_WORDREGION_PRECISION", request.param)
    tc = request.param == "tc"
    return types.SimpleNamespace(name=request.param, loss=LOSS_RTOL if tc else FP32_LOSS_RTOL,
                                 grad=GRAD_RTOL if tc else FP32_GRAD_RTOL, sim_atol=5e-3 if tc else 2e-4,
                                 att=2e-3 if tc else 1e-5)     # tc: the maps come out of the fp16-operand forward


@pytest.fixture(params=["fp32", "tc"])
def hprec(request, monkeypatch):
    """Margin-head tests run once per arithmetic mode of the cos-theta / gradient contractions: fp32 SIMT (tight
    bounds, bit-exact argmax) and tcgen05 with fp16 operands / fp32 accumulation (contract bounds; |logit| <= s, so
    the absolute logit bound is s * 2e-4; argmax compared where the top-2 gap exceeds that resolution)."""
    monkeypatch.setenv("TGFR_HEAD_PRECISION", request.param)
    tc = request.param == "tc"
    return types.SimpleNamespace(name=request.param, tc=tc, loss=LOSS_RTOL if tc else FP32_LOSS_RTOL,
                                 grad=GRAD_RTOL if tc else FP32_GRAD_RTOL, logit=1e-2 if tc else 1e-4)


def argmax_matches(got, ref, gap):
    """argmax equality on the rows whose reference top-2 gap is above `gap` (all rows when gap == 0)."""
    top2 = np.sort(ref, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > gap
    return clear.mean() > 0.5 and np.array_equal(got.argmax(1)[clear], ref.argmax(1)[clear])


@pytest.fixture(scope="module")
def api():
    from text_guided_face_recognition_b200.models import attention, losses, magface, metrics
    return types.SimpleNamespace(attention=attention, losses=losses, metrics=metrics, magface=magface)


# ------------------------------------------------------------------------------------------------
# word-region loss
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,channels_last", [("wordregion_bert_small", True), ("wordregion_bert_small", False),
                                                ("wordregion_lstm_ragged", True), ("wordregion_lstm_ragged", False)])
def test_words_loss_small_vs_golden(api, golden_dir, name, channels_last, prec):
    g = load(golden_dir, name)
    B, T, ih, iw = int(g["B"]), int(g["T"]), int(g["ih"]), int(g["iw"])
    flavour = str(g["flavour"])
    args = make_args(flavour, T, tuple(float(v) for v in g["gammas"]))
    c_ref, w_ref, grads = ref_layout(g["ctx"], g["words"], ih, iw, channels_last)
    cap = torch.from_numpy(g["cap_lens"]).cuda() if g["cap_lens"].size else None
    l0, l1, att = api.losses.words_loss(c_ref, w_ref, torch.arange(B).cuda(), cap, np.arange(B), B, args)
    assert abs(l0.item() - float(g["loss0"])) < prec.loss * abs(float(g["loss0"]))
    assert abs(l1.item() - float(g["loss1"])) < prec.loss * abs(float(g["loss1"]))
    assert len(att) == B
    for i, a in enumerate(att):
        n = int(g["cap_lens"][i]) if g["cap_lens"].size else T
        assert tuple(a.shape) == (1, n, ih, iw)
        assert rel(a.cpu().numpy().reshape(n, -1), g["att"][i, :n]) < prec.att
    (float(g["w0"]) * l0 + float(g["w1"]) * l1).backward()
    dctx, dwords = grads()
    assert rel(dctx, g["dctx"]) < prec.grad
    assert rel(dwords, g["dwords"]) < prec.grad


@pytest.mark.parametrize("name", ["wordregion_bert_mid", "wordregion_config1"])
def test_words_loss_full_width_vs_golden(api, golden_dir, name, prec):
    g = load(golden_dir, name)
    B, T, D, ih, iw = int(g["B"]), int(g["T"]), int(g["D"]), int(g["ih"]), int(g["iw"])
    flavour = str(g["flavour"])
    ctx, words, cap = synth.wordregion_inputs(B, T, ih * iw, D, flavour, seed=100)
    args = make_args(flavour, T)
    c_ref, w_ref, grads = ref_layout(ctx, words, ih, iw)
    capt = None if cap is None else torch.from_numpy(cap).cuda()
    l0, l1, att = api.losses.words_loss(c_ref, w_ref, torch.arange(B).cuda(), capt, np.arange(B), B, args)
    assert abs(l0.item() - float(g["loss0"])) < prec.loss * abs(float(g["loss0"]))
    assert abs(l1.item() - float(g["loss1"])) < prec.loss * abs(float(g["loss1"]))
    got = np.stack([a.cpu().numpy().reshape(T, -1) for a in att])
    assert rel(got, g["att"]) < prec.att
    (l0 + l1).backward()
    dctx, dwords = grads()
    assert rel(dwords, g["dwords"]) < prec.grad
    assert rel(dctx[:2], g["dctx_head"]) < prec.grad
    proj = np.random.RandomState(7).randn(D).astype(np.float32)
    assert rel(dctx @ proj, g["dctx_proj"]) < 5 * prec.grad
    assert abs(np.linalg.norm(dctx.astype(np.float64)) - float(g["dctx_norm"])) < prec.grad * float(g["dctx_norm"])


def test_words_loss_only_context_grad(api, prec):
    """The reference's training scripts detach the text side (utils/dataset_utils.py:42-46)."""
    B, T, R, D = 8, 6, 16, 32
    ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=3)
    args = make_args("BERT", T)
    c = torch.from_numpy(ctx).cuda().requires_grad_(True)
    w = torch.from_numpy(words).cuda()
    l0, l1, _ = api.losses.words_loss(c.view(B, 4, 4, D).permute(0, 3, 1, 2), w.transpose(1, 2),
                                      torch.arange(B).cuda(), None, None, B, args)
    (l0 + l1).backward()
    dctx, _ = O.words_loss_grads(ctx, words, None, None, 4.0, 5.0, 10.0)
    assert rel(c.grad.cpu().numpy(), dctx) < prec.grad
    l0n, l1n, att = api.losses.words_loss(c.view(B, 4, 4, D).permute(0, 3, 1, 2), w.transpose(1, 2),
                                          None, None, None, B, args)
    assert l0n is None and l1n is None and len(att) == B


def test_words_loss_config2_size_vs_oracle(api, prec):
    """BASELINE config 2 (B=128, T=22, R=196, D=256): similarity matrix and losses at full size."""
    B, T, R, D = 128, 22, 196, 256
    ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=100)
    from text_guided_face_recognition_b200 import ops
    feats = torch.from_numpy(ctx).cuda()
    wd = torch.from_numpy(words).cuda()
    sim, attn = ops.wordregion_sim(feats, wd, None, 4.0, 5.0, 10.0)
    ref, ref_att = O.wordregion_sim(ctx, words, None, 4.0, 5.0, 10.0)
    assert np.max(np.abs(sim.cpu().numpy() - ref)) < prec.sim_atol  # |sim| ~ 30
    assert rel(attn.cpu().numpy(), np.stack(ref_att)) < prec.att
    l0, l1 = ops.pair_ce(sim)
    r0, r1 = O.pair_ce(ref)
    assert abs(l0.item() - r0) < prec.loss * r0 and abs(l1.item() - r1) < prec.loss * r1
    # rows of both softmaxes sum to one => the gradient of (loss0+loss1) w.r.t. sim sums to zero
    s = sim.detach().requires_grad_(True)
    a, b = ops.pair_ce(s)
    (a + b).backward()
    assert abs(s.grad.sum().item()) < 1e-4
    assert rel(s.grad.cpu().numpy(), O.pair_ce_bwd(ref)) < (1e-4 if prec.name == "fp32" else 5e-3)


def test_words_loss_grads_b32_vs_oracle(api, prec):
    B, T, R, D = 32, 22, 196, 256
    ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=11)
    args = make_args("BERT", T)
    c_ref, w_ref, grads = ref_layout(ctx, words, 14, 14)
    l0, l1, _ = api.losses.words_loss(c_ref, w_ref, torch.arange(B).cuda(), None, None, B, args)
    (l0 + 2.0 * l1).backward()
    dctx, dwords = grads()
    rc, rw = O.words_loss_grads(ctx, words, None, None, 4.0, 5.0, 10.0, 1.0, 2.0)
    assert rel(dctx, rc) < prec.grad
    assert rel(dwords, rw) < prec.grad


def test_func_attention_vs_golden(api, golden_dir):
    g = load(golden_dir, "attention_small")
    q = torch.from_numpy(g["query"]).cuda().requires_grad_(True)
    c = torch.from_numpy(g["context"]).cuda().requires_grad_(True)
    wc, attn = api.attention.func_attention(q, c, float(g["gamma1"]))
    assert tuple(wc.shape) == g["wc"].shape and tuple(attn.shape) == g["attn"].shape
    assert rel(wc.detach().cpu().numpy(), g["wc"]) < 1e-5
    assert rel(attn.detach().cpu().numpy(), g["attn"]) < 1e-5
    ((wc * torch.from_numpy(g["gw"]).cuda()).sum() + (attn * torch.from_numpy(g["ga"]).cuda()).sum()).backward()
    assert rel(q.grad.cpu().numpy(), g["dquery"]) < FP32_GRAD_RTOL
    assert rel(c.grad.cpu().numpy(), g["dcontext"]) < FP32_GRAD_RTOL


# ------------------------------------------------------------------------------------------------
# sentence / global / CLIP losses
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["sentence_plain", "sentence_collisions"])
def test_sentence_losses_vs_golden(api, golden_dir, name):
    g = load(golden_dir, name)
    B = g["img"].shape[0]
    args = make_args("BERT", 22)
    labels = torch.arange(B).cuda()

    def leaves(a, b):
        return (torch.from_numpy(a).cuda().requires_grad_(True), torch.from_numpy(b).cuda().requires_grad_(True))

    a, b = leaves(g["img"], g["txt"])
    l0, l1 = api.losses.sent_loss(a, b, labels, g["class_ids"], B, args)
    assert abs(l0.item() - float(g["sent_loss0"])) < FP32_LOSS_RTOL * float(g["sent_loss0"])
    assert abs(l1.item() - float(g["sent_loss1"])) < FP32_LOSS_RTOL * float(g["sent_loss1"])
    (l0 + 0.5 * l1).backward()
    assert rel(a.grad.cpu().numpy(), g["sent_dimg"]) < FP32_GRAD_RTOL
    assert rel(b.grad.cpu().numpy(), g["sent_dtxt"]) < FP32_GRAD_RTOL
    assert api.losses.sent_loss(a, b, None, g["class_ids"], B, args) == (None, None)

    a, b = leaves(g["img"], g["txt"])
    gl = api.losses.global_loss(a, b)
    assert gl.dim() == 0
    assert abs(gl.item() - float(g["global_loss"])) < FP32_LOSS_RTOL * float(g["global_loss"])
    gl.backward()
    assert rel(a.grad.cpu().numpy(), g["global_dimg"]) < FP32_GRAD_RTOL
    assert rel(b.grad.cpu().numpy(), g["global_dtxt"]) < FP32_GRAD_RTOL

    a, b = leaves(g["img"] * 3.0, g["txt"] * 2.0)
    cl = api.losses.ClipLoss()(b, a, args, 1)
    assert abs(cl.item() - float(g["clip_loss"])) < FP32_LOSS_RTOL * float(g["clip_loss"])
    cl.backward()
    assert rel(a.grad.cpu().numpy(), g["clip_dimg"]) < FP32_GRAD_RTOL
    assert rel(b.grad.cpu().numpy(), g["clip_dtxt"]) < FP32_GRAD_RTOL


def test_sentence_loss_b1024_vs_oracle(api):
    """Config-4 width of the B x B block (global batch 1024) with class collisions."""
    B, D = 1024, 256
    img, txt, cid = synth.sentence_inputs(B, D, seed=5, collisions=True)
    args = make_args("BERT", 30)
    a = torch.from_numpy(img).cuda().requires_grad_(True)
    b = torch.from_numpy(txt).cuda().requires_grad_(True)
    l0, l1 = api.losses.sent_loss(a, b, torch.arange(B).cuda(), cid, B, args)
    r0, r1, _ = O.sent_loss(img, txt, None, cid, 10.0)
    assert abs(l0.item() - r0) < FP32_LOSS_RTOL * r0 and abs(l1.item() - r1) < FP32_LOSS_RTOL * r1
    (l0 + l1).backward()
    dx, dy = O.sent_loss_grads(img, txt, None, cid, 10.0)
    assert rel(a.grad.cpu().numpy(), dx) < FP32_GRAD_RTOL
    assert rel(b.grad.cpu().numpy(), dy) < FP32_GRAD_RTOL


# ------------------------------------------------------------------------------------------------
# margin heads
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["arc_small", "arc_small_easy"])
def test_arc_margin_small_vs_golden(api, golden_dir, name, hprec):
    g = load(golden_dir, name)
    B, Din = g["x"].shape
    C = g["weight"].shape[0]
    head = api.metrics.ArcMarginProduct(Din, C, s=float(g["s"]), m=float(g["m"]), easy_margin=bool(g["easy"])).cuda()
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(g["weight"]))
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    label = torch.from_numpy(g["label"]).cuda()
    logits = head(x, label)
    assert tuple(logits.shape) == (B, C)
    assert np.max(np.abs(logits.detach().cpu().numpy() - g["logits"])) < (hprec.logit if hprec.tc else 5e-5)
    assert argmax_matches(logits.detach().cpu().numpy(), g["logits"], 2 * hprec.logit if hprec.tc else 0.0)
    loss = api.losses.FocalLoss(gamma=float(g["gamma"]))(logits, label)
    assert abs(loss.item() - float(g["loss"])) < hprec.loss * float(g["loss"])
    loss.backward()
    assert rel(x.grad.cpu().numpy(), g["dx"]) < hprec.grad
    assert rel(head.weight.grad.cpu().numpy(), g["dweight"]) < hprec.grad


def test_arc_margin_mid_vs_golden(api, golden_dir, hprec):
    g = load(golden_dir, "arc_mid")
    B, Din, C = int(g["B"]), int(g["Din"]), int(g["C"])
    xn, wn, label = synth.margin_inputs(B, Din, C, seed=100)
    head = api.metrics.ArcMarginProduct(Din, C, s=float(g["s"]), m=float(g["m"])).cuda()
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(wn))
    x = torch.from_numpy(xn).cuda().requires_grad_(True)
    lab = torch.from_numpy(label).cuda()
    logits = head(x, lab)
    if not hprec.tc:
        assert np.array_equal(logits.argmax(1).cpu().numpy(), g["argmax"])
    else:
        ref_full = O.arc_margin(xn, wn, label, float(g["s"]), float(g["m"]), False)
        assert argmax_matches(logits.detach().cpu().numpy(), ref_full, 2 * hprec.logit)
    assert np.max(np.abs(logits.detach().cpu().numpy()[:8] - g["logits_head"])) < (hprec.logit if hprec.tc else 5e-5)
    loss = api.losses.FocalLoss(gamma=2)(logits, lab)
    assert abs(loss.item() - float(g["loss"])) < hprec.loss * float(g["loss"])
    loss.backward()
    assert rel(x.grad.cpu().numpy(), g["dx"]) < hprec.grad
    assert rel(head.weight.grad.cpu().numpy()[:64], g["dweight_head"]) < hprec.grad
    assert abs(np.linalg.norm(head.weight.grad.cpu().numpy().astype(np.float64)) - float(g["dweight_norm"])) \
        < hprec.grad * float(g["dweight_norm"])


def test_arc_margin_config3_size_vs_oracle(api, hprec):
    """BASELINE config 3: ArcMarginProduct(512, 10177, s=30, m=0.5) + FocalLoss(2), B=512."""
    B, Din, C = 512, 512, 10177
    xn, wn, label = synth.margin_inputs(B, Din, C, seed=100)
    head = api.metrics.ArcMarginProduct(Din, C, s=30.0, m=0.5).cuda()
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(wn))
    x = torch.from_numpy(xn).cuda().requires_grad_(True)
    lab = torch.from_numpy(label).cuda()
    logits = head(x, lab)
    ref = O.arc_margin(xn, wn, label, 30.0, 0.5, False)
    got = logits.detach().cpu().numpy()
    assert np.max(np.abs(got - ref)) < hprec.logit
    # decisions: identical argmax wherever the fp64 top-2 gap is above the mode's resolution
    assert argmax_matches(got, ref, 2 * hprec.logit if hprec.tc else 1e-3)
    loss = api.losses.FocalLoss(gamma=2)(logits, lab)
    rl = O.focal_loss(ref, label, 2.0)
    assert abs(loss.item() - rl) < hprec.loss * rl
    loss.backward()
    dx, dw = O.arc_margin_bwd(xn, wn, label, O.focal_loss_bwd(ref, label, 2.0), 30.0, 0.5, False)
    assert rel(x.grad.cpu().numpy(), dx) < hprec.grad
    assert rel(head.weight.grad.cpu().numpy(), dw) < hprec.grad


@pytest.mark.parametrize("name", ["arc_small", "arc_small_easy"])
def test_arc_fused_loss_small_vs_golden(api, golden_dir, name):
    """ArcMarginProduct.fused_loss (no logits: margin / softmax statistics / softmax gradient in the GEMM epilogues)
    against the reference's logits -> FocalLoss pair."""
    g = load(golden_dir, name)
    B, Din = g["x"].shape
    C = g["weight"].shape[0]
    head = api.metrics.ArcMarginProduct(Din, C, s=float(g["s"]), m=float(g["m"]), easy_margin=bool(g["easy"])).cuda()
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(g["weight"]))
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    label = torch.from_numpy(g["label"]).cuda()
    loss = head.fused_loss(x, label, gamma=float(g["gamma"]))
    assert abs(loss.item() - float(g["loss"])) < LOSS_RTOL * float(g["loss"])
    loss.backward()
    assert rel(x.grad.cpu().numpy(), g["dx"]) < GRAD_RTOL
    assert rel(head.weight.grad.cpu().numpy(), g["dweight"]) < GRAD_RTOL


@pytest.mark.parametrize("B,Din,C,gamma,easy", [(512, 512, 10177, 2.0, False), (96, 256, 4500, 0.0, False),
                                               (37, 72, 1003, 2.0, True)])
def test_arc_fused_loss_vs_oracle(api, B, Din, C, gamma, easy):
    """Config-3 size (and ragged sizes: rows / classes / features that do not fill the 128-wide tiles); gamma = 0 is
    plain cross entropy; an upstream gradient other than 1 (the reference scales by lambda_id = 100)."""
    xn, wn, label = synth.margin_inputs(B, Din, C, seed=100)
    head = api.metrics.ArcMarginProduct(Din, C, s=30.0, m=0.5, easy_margin=easy).cuda()
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(wn))
    x = torch.from_numpy(xn).cuda().requires_grad_(True)
    lab = torch.from_numpy(label).cuda()
    loss = head.fused_loss(x, lab, gamma=gamma)
    ref = O.arc_margin(xn, wn, label, 30.0, 0.5, easy)
    rl = O.focal_loss(ref, label, gamma)
    assert abs(loss.item() - rl) < LOSS_RTOL * rl
    (100.0 * loss).backward()
    dx, dw = O.arc_margin_bwd(xn, wn, label, O.focal_loss_bwd(ref, label, gamma, 100.0), 30.0, 0.5, easy)
    assert rel(x.grad.cpu().numpy(), dx) < GRAD_RTOL
    assert rel(head.weight.grad.cpu().numpy(), dw) < GRAD_RTOL


def test_arc_fused_class_shards_on_one_gpu():
    """The class-sharded fused head (distributed.ShardedArcMarginProduct.loss) emulated on one GPU: two shards of
    1003 classes, their (row max, sum-exp, target logit) merged by hand.  Labels 502..511 belong to shard 1 but lie
    inside the zero padding of shard 0's last 128-wide tile: shard 0 must not see a label column there."""
    from text_guided_face_recognition_b200 import ops
    B, Din, C = 64, 256, 1003
    xn, wn, label = synth.margin_inputs(B, Din, C, seed=7)
    label[:4] = [505, 511, 502, 501]
    cuts = [0, 502, C]
    x = torch.from_numpy(xn).cuda()
    lab = torch.from_numpy(label).cuda()
    stats = []
    for k in range(2):                                   # pass 1: every shard's own statistics
        w = torch.from_numpy(wn[cuts[k]:cuts[k + 1]]).cuda()
        ops.arc_fused_focal(x, w, lab, 30.0, 0.5, False, 2.0, cuts[k],
                            merge=lambda mx, sm, tg: (stats.append((mx.clone(), sm.clone(), tg.clone())), (mx, sm, tg))[1])

    def merged(mx, sm, tg, other):
        omx, osm, otg = other
        g = torch.maximum(mx, omx)
        return g, (sm * torch.exp(mx - g) + osm * torch.exp(omx - g)).contiguous(), (tg + otg).contiguous()
    ref = O.arc_margin(xn, wn, label, 30.0, 0.5, False)
    rl = O.focal_loss(ref, label, 2.0)
    dx_ref, dw_ref = O.arc_margin_bwd(xn, wn, label, O.focal_loss_bwd(ref, label, 2.0), 30.0, 0.5, False)
    dx = torch.zeros_like(x)
    for k in range(2):                                   # pass 2: the global loss and this shard's gradients
        xk = x.clone().requires_grad_(True)
        w = torch.from_numpy(wn[cuts[k]:cuts[k + 1]]).cuda().requires_grad_(True)
        loss = ops.arc_fused_focal(xk, w, lab, 30.0, 0.5, False, 2.0, cuts[k],
                                   merge=lambda mx, sm, tg, o=stats[1 - k]: merged(mx, sm, tg, o))
        assert abs(loss.item() - rl) < LOSS_RTOL * rl, (k, loss.item(), rl)
        loss.backward()
        assert rel(w.grad.cpu().numpy(), dw_ref[cuts[k]:cuts[k + 1]]) < GRAD_RTOL
        dx += xk.grad
    assert rel(dx.cpu().numpy(), dx_ref) < GRAD_RTOL


@pytest.mark.parametrize("name", ["mag_small_easy", "mag_small_hard"])
def test_mag_head_vs_golden(api, golden_dir, name, hprec):
    g = load(golden_dir, name)
    B, Din = g["x"].shape
    C = g["weight"].shape[1]
    l_a, u_a, l_m, u_m = float(g["l_a"]), float(g["u_a"]), float(g["l_margin"]), float(g["u_margin"])
    head = api.magface.MagLinear(Din, C, scale=float(g["scale"]), easy_margin=bool(g["easy"])).cuda()
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(g["weight"]))
    crit = api.magface.MagLoss(l_a, u_a, l_m, u_m, float(g["scale"]))
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    lab = torch.from_numpy(g["label"]).cuda()
    logits, x_norm = head(x, lambda xn: (u_m - l_m) / (u_a - l_a) * (xn - l_a) + l_m, l_a, u_a)
    assert np.max(np.abs(logits[0].detach().cpu().numpy() - g["cos"])) < (2 * hprec.logit if hprec.tc else 1e-4)
    assert np.max(np.abs(logits[1].detach().cpu().numpy() - g["cos_m"])) < (4 * hprec.logit if hprec.tc else 1e-4)
    assert rel(x_norm.detach().cpu().numpy(), g["x_norm"]) < 1e-6
    loss, loss_g, one_hot = crit(logits, lab, x_norm)
    assert abs(loss.item() - float(g["loss"])) < hprec.loss * float(g["loss"])
    assert abs(loss_g.item() - float(g["loss_g"])) < FP32_LOSS_RTOL * float(g["loss_g"])
    assert np.array_equal(one_hot.cpu().numpy(), g["one_hot"])
    (loss + float(g["lam_g"]) * loss_g).backward()
    assert rel(x.grad.cpu().numpy(), g["dx"]) < hprec.grad
    assert rel(head.weight.grad.cpu().numpy(), g["dweight"]) < hprec.grad


def test_other_heads_run(api, hprec):
    """AddMargin / Sphere heads ride the same cosine-logits kernel (API completeness)."""
    x = torch.randn(8, 32, device="cuda", requires_grad=True)
    lab = torch.randint(0, 20, (8,), device="cuda")
    add = api.metrics.AddMarginProduct(32, 20).cuda()
    out = add(x, lab)
    ref = O.arc_margin(x.detach().cpu().numpy(), add.weight.detach().cpu().numpy(), lab.cpu().numpy(), 30.0, 0.0, True)
    ref[np.arange(8), lab.cpu().numpy()] -= 30.0 * 0.40
    assert np.max(np.abs(out.detach().cpu().numpy() - ref)) < hprec.logit
    out.sum().backward()
    assert torch.isfinite(x.grad).all()
    sp = api.metrics.SphereProduct(32, 20).cuda()
    assert tuple(sp(x, lab).shape) == (8, 20)


def test_fcam_losses_overlapped_matches_sequential_calls():
    """fcam.fcam_losses (sentence chain on a side stream) gives the same losses and gradients as the two
    reference-shaped calls issued one after the other, eagerly and as a replayed CUDA graph."""
    from text_guided_face_recognition_b200 import fcam
    from text_guided_face_recognition_b200.graphs import GraphedStep
    from text_guided_face_recognition_b200.models import losses
    B, T, R, D = 16, 18, 196, 256
    # BERT flavour: no caption lengths (the LSTM flavour reads them back on the host, as the reference does, and a
    # host read cannot be captured); device-resident class ids: nothing to copy during capture
    ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=3)
    img, txt, cid = synth.sentence_inputs(B, D, seed=3, collisions=True)
    cid = torch.from_numpy(cid).cuda()
    args = make_args("BERT", T)
    labels = torch.arange(B, device="cuda")
    capt = None
    leaves = [torch.from_numpy(a).cuda().requires_grad_(True) for a in (ctx, words, img, txt)]

    def views():
        c, w, a, b = leaves
        return c.view(B, 14, 14, D).permute(0, 3, 1, 2), w.transpose(1, 2), a, b

    def sequential():
        for t in leaves:
            t.grad = None
        c, w, a, b = views()
        s0, s1 = losses.sent_loss(a, b, labels, cid, B, args)
        w0, w1, _ = losses.words_loss(c, w, labels, capt, cid, B, args)
        tot = w0 + w1 + s0 + s1
        tot.backward()
        return tot

    def overlapped():
        for t in leaves:
            t.grad = None
        c, w, a, b = views()
        w0, w1, _, s0, s1 = fcam.fcam_losses(c, w, a, b, labels, capt, cid, B, args)
        tot = w0 + w1 + s0 + s1
        tot.backward()
        return tot

    ref = sequential().item()
    ref_grads = [t.grad.clone() for t in leaves]
    got = overlapped().item()
    torch.cuda.synchronize()
    assert got == ref
    for g, t in zip(ref_grads, leaves):
        # the tensor-core d ctx is accumulated with reduce-adds in a run-dependent order
        assert torch.allclose(t.grad, g, rtol=1e-4, atol=1e-7)
    step = GraphedStep(overlapped)
    for _ in range(2):
        out = step()
    torch.cuda.synchronize()
    assert abs(out.item() - ref) <= 1e-6 * abs(ref)
    for g, t in zip(ref_grads, leaves):
        assert torch.allclose(t.grad, g, rtol=1e-4, atol=1e-7)


# ------------------------------------------------------------------------------------------------
# round 2: exported cosine_similarity, MagFace at BASELINE config 3 (rows a9 / a10)
# ------------------------------------------------------------------------------------------------
def test_cosine_similarity_vs_golden(api, golden_dir):
    """models/losses.py:12-16 through tgfr_cosine_rows_fwd/bwd: zero rows, rows below the eps clamp, gradients."""
    g = load(golden_dir, "cosine_small")
    a = torch.from_numpy(g["x1"]).cuda().requires_grad_(True)
    b = torch.from_numpy(g["x2"]).cuda().requires_grad_(True)
    out = api.losses.cosine_similarity(a, b)
    assert tuple(out.shape) == g["out"].shape
    assert np.max(np.abs(out.detach().cpu().numpy() - g["out"])) < 1e-6
    out.backward(torch.from_numpy(g["g"]).cuda())
    assert rel(a.grad.cpu().numpy(), g["dx1"]) < FP32_GRAD_RTOL
    assert rel(b.grad.cpu().numpy(), g["dx2"]) < FP32_GRAD_RTOL
    # the call shape of words_loss (losses.py:99-104): [B*T, D] views of transposed tensors, dim=1 reduced
    rs = np.random.RandomState(3)
    w = rs.randn(6, 32, 5).astype(np.float32)
    c = rs.randn(6, 32, 5).astype(np.float32)
    wt = torch.from_numpy(w).cuda().transpose(1, 2).contiguous().view(30, 32)
    ct = torch.from_numpy(c).cuda().transpose(1, 2).contiguous().view(30, 32)
    got = api.losses.cosine_similarity(wt, ct).cpu().numpy()
    ref = O.cosine_similarity(w.transpose(0, 2, 1).reshape(30, 32), c.transpose(0, 2, 1).reshape(30, 32))
    assert np.max(np.abs(got - ref)) < 1e-6
    # another reduced dimension and the trailing squeeze
    x3 = torch.from_numpy(w).cuda()
    got3 = api.losses.cosine_similarity(x3, torch.from_numpy(c).cuda(), dim=2)
    assert np.max(np.abs(got3.cpu().numpy() - O.cosine_similarity(w, c, dim=2))) < 1e-6
    one = api.losses.cosine_similarity(x3[:1, :, :1], torch.from_numpy(c).cuda()[:1, :, :1])
    assert one.dim() == 0


def test_mag_head_config3_vs_golden_and_oracle(api, golden_dir, hprec):
    """MagLinear(512, 10177, scale=64) + MagLoss(10, 110, 0.45, 0.8), B = 512: the reference's own numbers
    (tests/golden/mag_config3.npz) and the fp64 oracle for the full logits."""
    g = load(golden_dir, "mag_config3")
    B, Din, C = int(g["B"]), int(g["Din"]), int(g["C"])
    l_a, u_a, l_m, u_m, scale = (float(g[k]) for k in ("l_a", "u_a", "l_margin", "u_margin", "scale"))
    xn, wn, label = synth.margin_inputs(B, Din, C, seed=100, mag=True)
    xn = xn * 4.0
    head = api.magface.MagLinear(Din, C, scale=scale, easy_margin=True).cuda()
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(wn))
    crit = api.magface.MagLoss(l_a, u_a, l_m, u_m, scale)
    x = torch.from_numpy(xn).cuda().requires_grad_(True)
    lab = torch.from_numpy(label).cuda()
    logits, x_norm = head(x, lambda v: (u_m - l_m) / (u_a - l_a) * (v - l_a) + l_m, l_a, u_a)
    (rc, rm), rxn = O.mag_linear(xn, wn, l_a, u_a, l_m, u_m, scale, True)
    # |logit| <= scale = 64: the tc bound is scale * 2e-4 * sqrt-ish; fp32 SIMT is exact to round-off
    tol = 4 * hprec.logit if hprec.tc else 2e-4
    got_c, got_m = logits[0].detach().cpu().numpy(), logits[1].detach().cpu().numpy()
    assert np.max(np.abs(got_c - rc)) < tol
    # easy_margin switches at cos = 0 (magface.py:98): within the arithmetic's resolution of the threshold either
    # branch is a correct rounding of the input, so cos_theta_m is compared away from it (everywhere in fp32 mode
    # except the few elements closer to 0 than fp32 round-off)
    clear = np.abs(rc) > (scale * 1e-3 if hprec.tc else scale * 1e-6)
    assert clear.mean() > 0.97 and np.max(np.abs(got_m - rm)[clear]) < 2 * tol
    assert np.max(np.abs(got_c[:8] - g["cos_head"])) < tol
    # decisions: argmax of the cosine logits (cos_theta_m is discontinuous at cos = 0, its row maximum sits AT the
    # threshold by construction, so it carries no decision)
    assert argmax_matches(got_c, rc, 4 * tol if hprec.tc else 1e-3)
    assert rel(x_norm.detach().cpu().numpy(), g["x_norm"]) < 1e-6
    loss, loss_g, one_hot = crit(logits, lab, x_norm)
    assert abs(loss.item() - float(g["loss"])) < hprec.loss * float(g["loss"])
    assert abs(loss_g.item() - float(g["loss_g"])) < FP32_LOSS_RTOL * float(g["loss_g"])
    oh = one_hot.cpu().numpy()
    assert oh.sum() == float(g["one_hot_sum"]) and np.array_equal(oh.argmax(1), label) and oh.max() == 1.0
    (loss + float(g["lam_g"]) * loss_g).backward()
    assert rel(x.grad.cpu().numpy(), g["dx"]) < hprec.grad
    dw = head.weight.grad.cpu().numpy()
    assert rel(dw[:, :64], g["dweight_head"]) < hprec.grad
    assert abs(np.linalg.norm(dw.astype(np.float64)) - float(g["dweight_norm"])) < hprec.grad * float(g["dweight_norm"])
    _, dw_ref = O.mag_head_grads(xn, wn, label, l_a, u_a, l_m, u_m, scale, True, 1.0, float(g["lam_g"]))
    assert rel(dw, dw_ref) < hprec.grad
