"""Generate the golden fixtures in this directory from the UNMODIFIED reference code.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's ``models.attention``, ``models.losses``,
``models.metrics`` and ``models.magface`` (shims: an empty ``termcolor`` module;
``torch.zeros(device='cuda')`` redirected to CPU for ArcMarginProduct, which
hard-codes the device at models/metrics.py:53), runs them forward + autograd in
fp32 on the seeded inputs of ``tests/synth.py`` and stores inputs, outputs and
gradients as ``*.npz``.  The reference has no golden vectors of its own
(SURVEY.md section 4); these files are what pins ``oracle/fcam_oracle.py``.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import synth  # noqa: E402

REF = "/root/reference"


def import_reference():
    tc = types.ModuleType("termcolor")
    tc.cprint = print
    sys.modules.setdefault("termcolor", tc)
    sys.path.insert(0, REF)
    from models import attention, losses, metrics, magface  # type: ignore
    sys.path.pop(0)
    # ArcMarginProduct hard-codes device='cuda' (metrics.py:53): redirect to CPU.
    real_zeros = torch.zeros

    class _TorchProxy:
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def zeros(*a, **kw):
            kw.pop("device", None)
            return real_zeros(*a, **kw)

    metrics.torch = _TorchProxy()
    return attention, losses, metrics, magface


def ns(**kw):
    return types.SimpleNamespace(**kw)


def make_args(flavour, T, g1=4.0, g2=5.0, g3=10.0):
    return ns(en_type=flavour, bert_words_num=T + 2, CUDA=False, device="cpu",
              TRAIN=ns(SMOOTH=ns(GAMMA1=g1, GAMMA2=g2, GAMMA3=g3)))


def to_ref_layout(ctx, words, ih, iw):
    """canonical [B,R,D]/[B,T,D] -> reference [B,D,ih,iw]/[B,D,T] views (a0 layouts)."""
    B, R, D = ctx.shape
    c = torch.from_numpy(ctx).clone().requires_grad_(True)
    w = torch.from_numpy(words).clone().requires_grad_(True)
    return c, w, c.view(B, ih, iw, D).permute(0, 3, 1, 2), w.transpose(1, 2)


def wordregion_case(losses, name, B, T, ih, iw, D, flavour, ragged, store_inputs=True,
                    g=(4.0, 5.0, 10.0), w0=1.0, w1=1.0):
    R = ih * iw
    ctx, words, cap_lens = synth.wordregion_inputs(B, T, R, D, flavour, seed=100, ragged=ragged)
    args = make_args(flavour, T, *g)
    c, w, c_ref, w_ref = to_ref_layout(ctx, words, ih, iw)
    labels = torch.arange(B)
    cl = None if cap_lens is None else torch.from_numpy(cap_lens)
    l0, l1, att = losses.words_loss(c_ref, w_ref, labels, cl, np.arange(B), B, args)
    (w0 * l0 + w1 * l1).backward()
    out = dict(B=B, T=T, ih=ih, iw=iw, D=D, flavour=flavour, ragged=int(ragged), gammas=np.array(g),
               w0=w0, w1=w1, loss0=l0.item(), loss1=l1.item(),
               cap_lens=np.zeros(0, np.int64) if cap_lens is None else cap_lens)
    Tmax = max(a.shape[1] for a in att)
    att_np = np.zeros((B, Tmax, R), np.float32)
    for i, a in enumerate(att):
        att_np[i, : a.shape[1]] = a.detach().numpy().reshape(a.shape[1], R)
    out["att"] = att_np
    dctx, dwords = c.grad.numpy(), w.grad.numpy()
    if store_inputs:
        out.update(ctx=ctx, words=words, dctx=dctx, dwords=dwords)
    else:
        rs = np.random.RandomState(7)
        proj = rs.randn(D).astype(np.float32)
        out.update(dwords=dwords, dctx_head=dctx[:2].copy(), dctx_proj=dctx @ proj,
                   dctx_norm=float(np.linalg.norm(dctx.astype(np.float64))))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss0", out["loss0"], "loss1", out["loss1"])


def attention_case(attention, name, B, T, ih, iw, D):
    rs = np.random.RandomState(5)
    q = rs.randn(B, D, T).astype(np.float32)
    c = rs.randn(B, D, ih, iw).astype(np.float32) * 0.3
    qt = torch.from_numpy(q).requires_grad_(True)
    ct = torch.from_numpy(c).requires_grad_(True)
    wc, attn = attention.func_attention(qt, ct, 4.0)
    gw = rs.randn(*wc.shape).astype(np.float32)
    ga = rs.randn(*attn.shape).astype(np.float32)
    ((wc * torch.from_numpy(gw)).sum() + (attn * torch.from_numpy(ga)).sum()).backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), query=q, context=c, gamma1=4.0,
                        wc=wc.detach().numpy(), attn=attn.detach().numpy(), gw=gw, ga=ga,
                        dquery=qt.grad.numpy(), dcontext=ct.grad.numpy())
    print(name, "ok")


def sentence_case(losses, name, B, D, collisions):
    img, txt, class_ids = synth.sentence_inputs(B, D, seed=100, collisions=collisions)
    args = make_args("BERT", 22)
    out = dict(img=img, txt=txt, class_ids=class_ids, gamma3=10.0)
    # sent_loss (with the class-id mask)
    a = torch.from_numpy(img).clone().requires_grad_(True)
    b = torch.from_numpy(txt).clone().requires_grad_(True)
    l0, l1 = losses.sent_loss(a, b, torch.arange(B), class_ids, B, args)
    (l0 + 0.5 * l1).backward()
    out.update(sent_loss0=l0.item(), sent_loss1=l1.item(), sent_dimg=a.grad.numpy(), sent_dtxt=b.grad.numpy())
    # global_loss
    a = torch.from_numpy(img).clone().requires_grad_(True)
    b = torch.from_numpy(txt).clone().requires_grad_(True)
    gl = losses.global_loss(a, b)
    gl.backward()
    out.update(global_loss=gl.item(), global_dimg=a.grad.numpy(), global_dtxt=b.grad.numpy())
    # ClipLoss (unnormalised logits; use scaled copies so that it differs from global_loss)
    a = torch.from_numpy(img * 3.0).clone().requires_grad_(True)
    b = torch.from_numpy(txt * 2.0).clone().requires_grad_(True)
    cl = losses.ClipLoss()(b, a, args, 1)
    cl.backward()
    out.update(clip_loss=cl.item(), clip_dimg=a.grad.numpy(), clip_dtxt=b.grad.numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, out["sent_loss0"], out["sent_loss1"], out["global_loss"], out["clip_loss"])


def arc_case(metrics, losses, name, B, Din, C, s, m, easy, store=True):
    x, w, label = synth.margin_inputs(B, Din, C, seed=100)
    head = metrics.ArcMarginProduct(Din, C, s=s, m=m, easy_margin=easy)
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(w))
    xt = torch.from_numpy(x).clone().requires_grad_(True)
    logits = head(xt, torch.from_numpy(label))
    loss = losses.FocalLoss(gamma=2)(logits, torch.from_numpy(label))
    loss.backward()
    out = dict(s=s, m=m, easy=int(easy), label=label, loss=loss.item(), gamma=2.0,
               argmax=logits.argmax(1).numpy())
    if store:
        out.update(x=x, weight=w, logits=logits.detach().numpy(), dx=xt.grad.numpy(),
                   dweight=head.weight.grad.numpy())
    else:
        out.update(B=B, Din=Din, C=C, dx=xt.grad.numpy(),
                   logits_head=logits.detach().numpy()[:8].copy(),
                   dweight_norm=float(np.linalg.norm(head.weight.grad.numpy().astype(np.float64))),
                   dweight_head=head.weight.grad.numpy()[:64].copy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", out["loss"])


def mag_case(magface, name, B, Din, C, easy):
    l_a, u_a, l_m, u_m, scale = 10.0, 110.0, 0.45, 0.8, 64.0
    x, w, label = synth.margin_inputs(B, Din, C, seed=100, mag=True)
    x = x * 4.0                                   # spread |x| across the [l_a, u_a] clamp
    head = magface.MagLinear(Din, C, scale=scale, easy_margin=easy)
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(w))
    crit = magface.MagLoss(l_a, u_a, l_m, u_m, scale)

    def margin(xn):
        return (u_m - l_m) / (u_a - l_a) * (xn - l_a) + l_m

    xt = torch.from_numpy(x).clone().requires_grad_(True)
    logits, x_norm = head(xt, margin, l_a, u_a)
    loss, loss_g, one_hot = crit(logits, torch.from_numpy(label), x_norm)
    lam_g = 35.0
    (loss + lam_g * loss_g).backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, weight=w, label=label, easy=int(easy),
                        l_a=l_a, u_a=u_a, l_margin=l_m, u_margin=u_m, scale=scale, lam_g=lam_g,
                        cos=logits[0].detach().numpy(), cos_m=logits[1].detach().numpy(),
                        x_norm=x_norm.detach().numpy(), loss=loss.item(), loss_g=loss_g.item(),
                        one_hot=one_hot.numpy(), dx=xt.grad.numpy(), dweight=head.weight.grad.numpy())
    print(name, "loss", loss.item(), "loss_g", loss_g.item())


def main():
    torch.manual_seed(100)
    torch.set_num_threads(8)
    attention, losses, metrics, magface = import_reference()
    attention_case(attention, "attention_small", B=3, T=5, ih=3, iw=4, D=16)
    wordregion_case(losses, "wordregion_bert_small", B=4, T=5, ih=3, iw=3, D=16, flavour="BERT", ragged=False)
    wordregion_case(losses, "wordregion_lstm_ragged", B=6, T=7, ih=4, iw=4, D=32, flavour="LSTM", ragged=True,
                    w0=1.0, w1=0.5)
    wordregion_case(losses, "wordregion_bert_mid", B=8, T=22, ih=14, iw=14, D=256, flavour="BERT", ragged=False,
                    store_inputs=False)
    wordregion_case(losses, "wordregion_config1", B=16, T=18, ih=14, iw=14, D=256, flavour="LSTM", ragged=False,
                    store_inputs=False)
    sentence_case(losses, "sentence_plain", B=16, D=256, collisions=False)
    sentence_case(losses, "sentence_collisions", B=12, D=64, collisions=True)
    arc_case(metrics, losses, "arc_small", B=8, Din=32, C=50, s=30.0, m=0.5, easy=False)
    arc_case(metrics, losses, "arc_small_easy", B=8, Din=32, C=50, s=35.0, m=0.5, easy=True)
    arc_case(metrics, losses, "arc_mid", B=64, Din=256, C=4500, s=30.0, m=0.5, easy=False, store=False)
    mag_case(magface, "mag_small_easy", B=8, Din=32, C=50, easy=True)
    mag_case(magface, "mag_small_hard", B=8, Din=32, C=50, easy=False)


if __name__ == "__main__":
    main()
