"""CPU port of the reference's PyTorch implementation of the hot path -- TEST/BENCH INFRASTRUCTURE.

This is the "cpu_baseline" / `bench.py --impl reference` arm: the reference is pure Python and
lives in /root/reference, which does not exist on the GPU box, so its CPU path is restated here
op for op (same per-caption loop, the same bmm / softmax / transpose-copy sequence and autograd
graph as models/losses.py:61-135 + models/attention.py:10-43, models/losses.py:19-57,
models/metrics.py:42-60, models/losses.py:313-325) so that it has the reference's cost profile.
`tests/test_oracle_golden.py::test_ref_port_matches_golden` pins it against the fixtures produced
by the real reference; DESIGN.md records its step time next to the real reference's in this
container.  Never imported by the product package.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def attention_port(query, context, gamma1):
    """models/attention.py:10-43 -- query [B,D,T], context [B,D,ih,iw]."""
    B, T = query.size(0), query.size(2)
    ih, iw = context.size(2), context.size(3)
    R = ih * iw
    ctx = context.view(B, -1, R)
    ctx_t = ctx.transpose(1, 2).contiguous()
    scores = torch.bmm(ctx_t, query)                               # [B,R,T]
    scores = torch.softmax(scores.view(B * R, T), dim=-1).view(B, R, T)
    scores = scores.transpose(1, 2).contiguous().view(B * T, R)
    attn = torch.softmax(scores * gamma1, dim=-1).view(B, T, R)
    attn_t = attn.transpose(1, 2).contiguous()
    return torch.bmm(ctx, attn_t), attn.view(B, -1, ih, iw)


def _cos(x1, x2, eps=1e-8):
    num = (x1 * x2).sum(1)
    return (num / (x1.norm(2, 1) * x2.norm(2, 1)).clamp(min=eps)).squeeze()


def words_loss_port(img_features, words_emb, labels, cap_lens, g1, g2, g3):
    """models/losses.py:61-135 (per-caption Python loop, B-fold repeat of the caption)."""
    B = img_features.size(0)
    lens = cap_lens if isinstance(cap_lens, (list, tuple)) else [words_emb.size(2)] * B
    cols, att = [], []
    for i in range(B):
        n = int(lens[i])
        word = words_emb[i, :, :n].unsqueeze(0).contiguous().repeat(B, 1, 1)
        wctx, attn = attention_port(word, img_features, g1)
        att.append(attn[i].unsqueeze(0).contiguous())
        wflat = word.transpose(1, 2).contiguous().view(B * n, -1)
        cflat = wctx.transpose(1, 2).contiguous().view(B * n, -1)
        row = _cos(wflat, cflat).view(B, n)
        row = row.mul(g2).exp()                                  # (reference does this in place)
        cols.append(torch.log(row.sum(dim=1, keepdim=True)))
    sim = torch.cat(cols, 1) * g3
    return F.cross_entropy(sim, labels), F.cross_entropy(sim.t(), labels), att


def sent_loss_port(cnn_code, rnn_code, labels, class_ids, g3, eps=1e-8):
    """models/losses.py:19-57 including the host-side numpy mask loop."""
    import numpy as np
    B = cnn_code.size(0)
    mask = None
    if class_ids is not None:
        rows = []
        for i in range(B):
            m = (class_ids == class_ids[i]).astype(bool)
            m[i] = 0
            rows.append(m.reshape(1, -1))
        mask = torch.BoolTensor(np.concatenate(rows, 0))
    a, b = cnn_code.unsqueeze(0), rnn_code.unsqueeze(0)
    na, nb = a.norm(2, dim=2, keepdim=True), b.norm(2, dim=2, keepdim=True)
    s = torch.bmm(a, b.transpose(1, 2)) / torch.bmm(na, nb.transpose(1, 2)).clamp(min=eps) * g3
    s = s.squeeze()
    if mask is not None:
        s.data.masked_fill_(mask.to(s.device), -float("inf"))      # losses.py:29-30 (`masks.cuda()` under args.CUDA)
    return F.cross_entropy(s, labels), F.cross_entropy(s.t(), labels)


def arc_focal_port(x, weight, label, s=30.0, m=0.5, gamma=2.0, easy_margin=False):
    """models/metrics.py:42-60 followed by models/losses.py:321-325."""
    cosine = F.linear(F.normalize(x), F.normalize(weight))
    sine = torch.sqrt((1.0 - cosine.pow(2)).clamp(0, 1))
    phi = cosine * math.cos(m) - sine * math.sin(m)
    if easy_margin:
        phi = torch.where(cosine > 0, phi, cosine)
    else:
        phi = torch.where(cosine > math.cos(math.pi - m), phi, cosine - math.sin(math.pi - m) * m)
    one_hot = torch.zeros_like(cosine)
    one_hot.scatter_(1, label.view(-1, 1).long(), 1)
    out = (one_hot * phi + (1.0 - one_hot) * cosine) * s
    logp = F.cross_entropy(out, label)
    p = torch.exp(-logp)
    return out, ((1 - p) ** gamma * logp).mean()


def text_heading_port(tokens, weights, biases, bert_words_num):
    """models/models.py:170-232 op for op: three Conv2d + ReLU, the B x T Python loop of stack / amax calls of
    get_each_word_feature (the reference's `torch.cuda.FloatTensor(...)` copy of the last word = detach + clone),
    max_pool1d / mean / normalize.  weights: [F, 1, K, 768] x 3.  Returns (words [B, F, T], sent [B, F])."""
    x = tokens.unsqueeze(1)
    xs = [F.relu(F.conv2d(x, w, b)).squeeze(3) for w, b in zip(weights, biases)]
    bs = xs[0].size(0)
    a, b, c = (t.transpose(2, 1) for t in xs)
    code = []
    seq = bert_words_num - 1 - 3
    for i in range(bs):
        t = [torch.amax(torch.stack((a[i, j], b[i, j], c[i, j])), dim=0) for j in range(seq)]
        t += [torch.amax(torch.stack((a[i, seq], b[i, seq])), dim=0)]
        t += [a[i, seq + 1].detach().clone()]
        code.append(torch.stack(t))
    code = F.normalize(torch.stack(code), p=2, dim=2)
    pooled = [F.max_pool1d(t, t.size(2)).squeeze(2) for t in xs]
    sent = F.normalize(torch.stack(pooled).mean(dim=0), p=2, dim=1)
    return code.transpose(1, 2), sent
