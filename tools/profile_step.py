"""One eager pass over every kernel of the path, for ncu (no graphs, no timing):
configs[1] loss step (words_loss + sent_loss fwd + bwd, face-side gradient), configs[2] margin head step through the
dense-logits API and through fused_loss, and a TextHeading step.  Usage (see profiles/README or DESIGN.md section 6):

    ncu --set full --clock-control none --import-source on -o gpurun_out/prof_all python tools/profile_step.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import synth  # noqa: E402
from text_guided_face_recognition_b200 import fcam  # noqa: E402
from text_guided_face_recognition_b200.models import losses, metrics  # noqa: E402
from text_guided_face_recognition_b200.models.text_heading import TextHeading  # noqa: E402

B, T, R, D = 128, 22, 196, 256
ns = types.SimpleNamespace
args = ns(en_type="BERT", bert_words_num=T + 2, CUDA=True, device="cuda",
          TRAIN=ns(SMOOTH=ns(GAMMA1=4.0, GAMMA2=5.0, GAMMA3=10.0)))
ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=100)
img, txt, cid = synth.sentence_inputs(B, D, seed=100)
c = torch.from_numpy(ctx).cuda().requires_grad_(True)
w = torch.from_numpy(words).cuda()
a = torch.from_numpy(img).cuda().requires_grad_(True)
b = torch.from_numpy(txt).cuda().requires_grad_(True)
labels = torch.arange(B, device="cuda")
cid = torch.arange(B, device="cuda")
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(passes):
    w0, w1, _, s0, s1 = fcam.fcam_losses(c.view(B, 14, 14, D).permute(0, 3, 1, 2), w.transpose(1, 2), a, b, labels, None,
                                         cid, B, args)
    (w0 + w1 + s0 + s1).backward()

    hB, Din, C = 512, 512, 10177
    xn, wn, lab = synth.margin_inputs(hB, Din, C, seed=100)
    head = metrics.ArcMarginProduct(Din, C, s=30.0, m=0.5).cuda()
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(wn))
    x = torch.from_numpy(xn).cuda().requires_grad_(True)
    labt = torch.from_numpy(lab).cuda()
    losses.FocalLoss(gamma=2)(head(x, labt), labt).backward()
    head.fused_loss(x, labt, gamma=2.0).backward()

    tok, tw, tb = synth.texthead_inputs(B, T + 2, D, seed=100)
    th = TextHeading(ns(aux_feat_dim_per_granularity=D, bert_words_num=T + 2)).cuda()
    wo, so = th(torch.from_numpy(tok).cuda(), None)
    (wo.sum() + so.sum()).backward()
    # verification scoring at configs[4]'s pair count: cosine, exact ROC, identification argmax
    from text_guided_face_recognition_b200 import ops
    NP, DF = 60000, 640
    o1, o2 = torch.randn(NP, DF, device="cuda"), torch.randn(NP, DF, device="cuda")
    sc = ops.pair_cosine(o1, o2)
    ops.roc_counts(sc, (torch.arange(NP, device="cuda") % 10 == 0).long())
    ops.row_argmax(sc.view(6000, 10))
    # round 2: MagFace head at configs[2] size, the exported cosine_similarity, the image head (IMIM + global projection) and
    # the FCFM fusion net in training mode (fwd + bwd each)
    from text_guided_face_recognition_b200.models import magface
    from text_guided_face_recognition_b200.models.fusion_nets import Working
    from text_guided_face_recognition_b200.models.image_heading import ImageHeading
    mxn, mwn, mlab = synth.margin_inputs(hB, Din, C, seed=100, mag=True)
    mhead = magface.MagLinear(Din, C, scale=64.0, easy_margin=True).cuda()
    with torch.no_grad():
        mhead.weight.copy_(torch.from_numpy(mwn))
    mx = torch.from_numpy(mxn * 4.0).cuda().requires_grad_(True)
    lg, xnorm = mhead(mx, lambda v: 0.0035 * (v - 10.0) + 0.45, 10.0, 110.0)
    ml, mg, _ = magface.MagLoss(10.0, 110.0, 0.45, 0.8, 64.0)(lg, torch.from_numpy(mlab).cuda(), xnorm)
    (ml + 35.0 * mg).backward()
    losses.cosine_similarity(x, x.detach().roll(1, 0)).sum().backward()
    ih = ImageHeading(ns(aux_feat_dim_per_granularity=D)).cuda().train()
    gi, li = ih(torch.randn(B, 512, device="cuda"), torch.randn(B, 256, 14, 14, device="cuda"))
    (gi.sum() + (li * li.detach().roll(1, 0)).sum()).backward()
    fus = Working(channel_dim=256).cuda().train()
    fo = fus(li.detach(), w.transpose(1, 2), gi.detach(), b.detach())
    fo.square().sum().backward()
    # the verification path of the fusion net: eval forward with the 3x3 convolution as an implicit tensor-core GEMM
    with torch.no_grad():
        fus.eval()
        nb = 4096
        fus(torch.randn(nb, 14, 14, 256, device="cuda").permute(0, 3, 1, 2), torch.randn(nb, T, 256, device="cuda").transpose(1, 2),
            torch.randn(nb, 256, device="cuda"), torch.randn(nb, 256, device="cuda"))
torch.cuda.synchronize()
print("profile_step ok")
