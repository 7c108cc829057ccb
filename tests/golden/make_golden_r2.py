"""Round-2 golden fixtures, generated from the UNMODIFIED reference code (build container only):

    python tests/golden/make_golden_r2.py

  cosine_small.npz       models/losses.py:12-16 cosine_similarity forward + autograd (zero rows, rows below eps)
  mag_config3.npz        MagLinear(512, 10177) + MagLoss at B = 512 (BASELINE configs[2]), compact: loss, loss_g,
                         dx, leading columns of the logits / d weight, norms
  train_block_bert.npz   the loss block of Train.train (src/train_encoders_bert.py:267-323): the script reads THOSE
                         SOURCE LINES from the reference file at run time, dedents them and executes them against the
                         reference's own words_loss / sent_loss / global_loss / ArcMarginProduct / FocalLoss on
                         synthetic tensors; total_loss and the gradients that reach image_head's outputs and the two
                         classifier weights are stored.  tests/test_dropin_loop.py runs the same block on the mirror.

Nothing of the reference is copied into the repository: only numbers are stored.
"""
from __future__ import annotations

import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import synth  # noqa: E402
from make_golden import REF, import_reference, make_args  # noqa: E402


def cosine_case(losses, name):
    rs = np.random.RandomState(5)
    a = rs.randn(37, 64).astype(np.float32)
    b = rs.randn(37, 64).astype(np.float32)
    a[3] = 0.0                       # zero row: 0 / eps
    a[5] *= 1e-6                     # |x1||x2| below eps: the clamp is active
    b[5] *= 1e-6
    g = rs.randn(37).astype(np.float32)
    ta, tb = torch.from_numpy(a).requires_grad_(True), torch.from_numpy(b).requires_grad_(True)
    out = losses.cosine_similarity(ta, tb)
    out.backward(torch.from_numpy(g))
    # a [B*T, D] pair viewed through strides, as words_loss calls it (losses.py:99-104)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x1=a, x2=b, g=g, out=out.detach().numpy(),
                        dx1=ta.grad.numpy(), dx2=tb.grad.numpy())
    print(name, float(out.sum()))


def mag_config3_case(magface, name):
    B, Din, C = 512, 512, 10177
    l_a, u_a, l_m, u_m, scale = 10.0, 110.0, 0.45, 0.8, 64.0
    x, w, label = synth.margin_inputs(B, Din, C, seed=100, mag=True)
    x = x * 4.0
    head = magface.MagLinear(Din, C, scale=scale, easy_margin=True)
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(w))
    crit = magface.MagLoss(l_a, u_a, l_m, u_m, scale)
    xt = torch.from_numpy(x).clone().requires_grad_(True)
    logits, x_norm = head(xt, lambda xn: (u_m - l_m) / (u_a - l_a) * (xn - l_a) + l_m, l_a, u_a)
    loss, loss_g, one_hot = crit(logits, torch.from_numpy(label), x_norm)
    lam_g = 35.0
    (loss + lam_g * loss_g).backward()
    dw = head.weight.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), B=B, Din=Din, C=C, l_a=l_a, u_a=u_a, l_margin=l_m,
                        u_margin=u_m, scale=scale, lam_g=lam_g, loss=loss.item(), loss_g=loss_g.item(),
                        cos_head=logits[0][:8].detach().numpy(), cos_m_head=logits[1][:8].detach().numpy(),
                        x_norm=x_norm.detach().numpy(), dx=xt.grad.numpy(), dweight_head=dw[:, :64].copy(),
                        dweight_norm=float(np.linalg.norm(dw.astype(np.float64))),
                        argmax=logits[1].argmax(1).numpy(), one_hot_sum=float(one_hot.sum()))
    print(name, "loss", loss.item(), "loss_g", loss_g.item())


TRAIN_BLOCK_FILE = os.path.join(REF, "src", "train_encoders_bert.py")
TRAIN_BLOCK_LINES = (267, 323)      # optimizer.zero_grad() ... total_loss.backward(); the optimiser calls are stubbed


def train_block_inputs(B=12, T=22, D=256, C=300, seed=100):
    """Synthetic stand-ins for what the loop body sees after image_head / prepare_train_data_for_Bert."""
    ctx, words, _ = synth.wordregion_inputs(B, T, 196, D, "BERT", seed=seed)
    img, txt, cid = synth.sentence_inputs(B, D, seed=seed, collisions=True)
    cid = (cid % C).astype(np.int64)
    _, w_img, _ = synth.margin_inputs(B, D, C, seed=seed)
    _, w_txt, _ = synth.margin_inputs(B, D, C, seed=seed + 1)
    return ctx, words, img, txt, cid, w_img, w_txt


def train_block_case(losses, metrics, name):
    B, T, D, C = 12, 22, 256, 300
    ctx, words, img, txt, cid, w_img, w_txt = train_block_inputs(B, T, D, C)
    src = open(TRAIN_BLOCK_FILE).read().splitlines()
    lo, hi = TRAIN_BLOCK_LINES
    block = textwrap.dedent("\n".join(src[lo - 1:hi]))
    assert "words_loss(words_features, words_emb, labels," in block and "total_loss.backward()" in block, block
    args = make_args("BERT", T)
    for k, v in dict(is_DAMSM=True, is_WRA=False, is_ident_loss=True, is_CLIP=True, is_CMP=False, lambda_clip=2.0,
                     lambda_id=100, model_type="arcface", batch_size=B).items():
        setattr(args, k, v)
    image_cls = metrics.ArcMarginProduct(D, C, s=30, m=0.5, easy_margin=False)
    text_cls = metrics.ArcMarginProduct(D, C, s=35, m=0.5, easy_margin=False)
    with torch.no_grad():
        image_cls.weight.copy_(torch.from_numpy(w_img))
        text_cls.weight.copy_(torch.from_numpy(w_txt))
    noop = types.SimpleNamespace(zero_grad=lambda: None, step=lambda: None)
    self_ = types.SimpleNamespace(args=args, text_cls=text_cls, image_cls=image_cls,
                                  ident_loss=losses.FocalLoss(gamma=2), optimizer=noop, optimizer_head=noop,
                                  optimizer_cls=noop)
    words_features = torch.from_numpy(ctx).view(B, 14, 14, D).permute(0, 3, 1, 2).requires_grad_(True)   # IMIM layout
    img_features = torch.from_numpy(img).clone().requires_grad_(True)
    words_emb = torch.from_numpy(words).transpose(1, 2)            # detached text side (utils/dataset_utils.py:42-46)
    sent_emb = torch.from_numpy(txt)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self                 # the block calls class_ids.cuda(); this box has no GPU
    try:
        env = dict(self=self_, words_loss=losses.words_loss, sent_loss=losses.sent_loss, global_loss=losses.global_loss,
                   words_features=words_features, img_features=img_features, words_emb=words_emb, sent_emb=sent_emb,
                   labels=torch.arange(B), cap_lens=None, class_ids=torch.from_numpy(cid), batch_size=B, torch=torch,
                   total_damsm_loss=0, w_total_loss=0, s_total_loss=0, total_cl_loss=0, total_cmp_loss=0)
        exec(compile(block, TRAIN_BLOCK_FILE, "exec"), env)
    finally:
        torch.Tensor.cuda = real_cuda
    np.savez_compressed(os.path.join(HERE, name + ".npz"), B=B, T=T, D=D, C=C,
                        total_loss=float(env["total_loss"].item()), damsm=float(env["total_damsm_loss"]),
                        w_total=float(env["w_total_loss"]), s_total=float(env["s_total_loss"]),
                        cl=float(env["total_cl_loss"].item()), d_words_features=words_features.grad.numpy(),
                        d_img_features=img_features.grad.numpy(), d_image_cls=image_cls.weight.grad.numpy(),
                        d_text_cls=text_cls.weight.grad.numpy())
    print(name, "total_loss", env["total_loss"].item())


def main():
    torch.manual_seed(100)
    torch.set_num_threads(8)
    attention, losses, metrics, magface = import_reference()
    cosine_case(losses, "cosine_small")
    mag_config3_case(magface, "mag_config3")
    train_block_case(losses, metrics, "train_block_bert")


if __name__ == "__main__":
    main()
