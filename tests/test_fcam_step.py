"""The assembled FCAM training step (BASELINE configs[3], src/train_encoders_bert.py:254-326) at world size 1:
TextHeading (no_grad) + ImageHeading + words / sent / global losses + two ArcFace heads + backward, against the fp64
oracle composed from the same pieces (oracle/texthead_oracle, fusion_oracle.imim_forward in training mode,
fcam_oracle losses and heads)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import synth  # noqa: E402
from make_golden_imim_r2 import imim_inputs  # noqa: E402

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("prec", ["fp32", "tc"])
def test_fcam_train_step_vs_oracle(prec, monkeypatch):
    monkeypatch.setenv("TGFR_WORDREGION_PRECISION", prec)
    monkeypatch.setenv("TGFR_HEAD_PRECISION", prec)
    from oracle import fcam_oracle as O
    from oracle import fusion_oracle as FO
    from oracle import texthead_oracle as TO
    from text_guided_face_recognition_b200.fcam import FcamTrainStep
    B, bwn, C = 8, 32, 300                                   # 32 BERT tokens -> T = 30 words (config 4)
    ns = types.SimpleNamespace
    args = ns(aux_feat_dim_per_granularity=256, bert_words_num=bwn, en_type="BERT",
              TRAIN=ns(SMOOTH=ns(GAMMA1=4.0, GAMMA2=5.0, GAMMA3=10.0)))
    step = FcamTrainStep(args, C, torch.device("cuda", 0))
    x, _, xg, _, _, _ = imim_inputs(B, 5)
    tok, _, _ = synth.texthead_inputs(B, bwn, 256, seed=5)
    cid = np.random.RandomState(5).randint(0, C // 2, size=B).astype(np.int64)
    cid[1] = cid[0]                                          # a class collision: the sent_loss mask is exercised
    captured = {}
    orig_fwd = step.image_head.forward

    def fwd(g, l):
        gi, wf = orig_fwd(g, l)
        wf.retain_grad()
        gi.retain_grad()
        captured["wf"], captured["gi"] = wf, gi
        return gi, wf
    step.image_head.forward = fwd
    total = step(torch.from_numpy(xg).cuda(), torch.from_numpy(x).cuda(), torch.from_numpy(tok).cuda(),
                 torch.from_numpy(cid).cuda())
    torch.cuda.synchronize()
    # ---- oracle composition
    sd = {k: v.detach().cpu().numpy() for k, v in step.image_head.imim.state_dict().items()}
    wf_ref = FO.imim_forward(sd, x, training=True)                                         # [B,256,14,14]
    pg = step.image_head.project_global.projection
    z = xg.astype(np.float64) @ pg.weight.detach().cpu().numpy().astype(np.float64).T + pg.bias.detach().cpu().numpy()
    gi_ref = z / np.linalg.norm(z, axis=1, keepdims=True)
    convs = step.text_head.bwm.convs1
    ws = [c.weight.detach().cpu().numpy()[:, 0] for c in convs]
    bs = [c.bias.detach().cpu().numpy() for c in convs]
    words_ref, sent_ref = TO.forward(tok, ws, bs, bwn)                                     # [B,T,256], [B,256]
    ctx_ref = wf_ref.transpose(0, 2, 3, 1).reshape(B, 196, 256)
    assert np.max(np.abs(captured["wf"].detach().cpu().numpy() - wf_ref)) < 5e-5
    w0, w1, _, _ = O.words_loss(ctx_ref, words_ref, None, None, 4.0, 5.0, 10.0)
    s0, s1, _ = O.sent_loss(gi_ref, sent_ref, None, cid, 10.0)
    cl = O.global_loss(gi_ref, sent_ref)
    wi = step.image_cls.weight.detach().cpu().numpy()
    wt = step.text_cls.weight.detach().cpu().numpy()
    li = O.arc_margin(gi_ref, wi, cid, 30.0, 0.5, False)
    lt = O.arc_margin(sent_ref, wt, cid, 35.0, 0.5, False)
    iid, tid = O.focal_loss(li, cid, 2.0), O.focal_loss(lt, cid, 2.0)
    ref_total = w0 + w1 + s0 + s1 + 100.0 * (tid + iid) + 2.0 * cl
    ltol, gtol = (5e-5, 5e-4) if prec == "fp32" else (2e-4, 2e-3)
    assert abs(total.item() - ref_total) < ltol * abs(ref_total), (total.item(), ref_total)
    # gradients at the head's outputs (what flows into ImageHeading's backward) and of the classifier weights
    dctx, _ = O.words_loss_grads(ctx_ref, words_ref, None, None, 4.0, 5.0, 10.0)
    got_dwf = captured["wf"].grad.permute(0, 2, 3, 1).reshape(B, 196, 256).cpu().numpy()
    assert rel(got_dwf, dctx) < gtol
    dimg_s, _ = O.sent_loss_grads(gi_ref, sent_ref, None, cid, 10.0)
    dimg_c, _ = O.global_loss_grads(gi_ref, sent_ref)
    dimg_i, dwi = O.arc_margin_bwd(gi_ref, wi, cid, O.focal_loss_bwd(li, cid, 2.0, 100.0), 30.0, 0.5, False)
    _, dwt = O.arc_margin_bwd(sent_ref, wt, cid, O.focal_loss_bwd(lt, cid, 2.0, 100.0), 35.0, 0.5, False)
    assert rel(captured["gi"].grad.cpu().numpy(), dimg_s + 2.0 * dimg_c + dimg_i) < gtol
    assert rel(step.image_cls.weight.grad.cpu().numpy(), dwi) < gtol
    assert rel(step.text_cls.weight.grad.cpu().numpy(), dwt) < gtol
    # every trainable parameter of the image head received a gradient; the text head (detached) none
    for name, p in step.image_head.named_parameters():
        assert (p.grad is not None) == (not name.endswith(("fc.weight", "fc.bias"))), name
    assert all(p.grad is None for p in step.text_head.parameters())
