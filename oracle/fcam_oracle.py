"""CPU oracle for the FCAM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement (float64 by default) of the reference's algorithm for the
word-region loss, the sentence/global/CLIP cosine losses and the ArcFace/MagFace
margin heads, with hand-derived backward passes.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package; the product path
(``text_guided_face_recognition_b200``) never does and fails loudly when the CUDA
library is missing.

Parity pin: the reference has no tests or golden vectors of its own (SURVEY.md
section 4), so this oracle is pinned against outputs of the reference's own
PyTorch code run in the build container: ``tests/golden/make_golden.py`` imports
``/root/reference/models/{attention,losses,metrics,magface}.py``, runs them
(forward + autograd) on seeded inputs and commits the results as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
below against those fixtures.

Canonical layouts used here (the reference's logical layouts are permutations of
these):
    ctx    [Bc, R, D]   region features (reference: img_features [B, D, ih, iw])
    words  [Bq, T, D]   word features   (reference: words_emb    [B, D, T])
Each function cites the reference file:line it restates.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "func_attention", "func_attention_bwd", "wordregion_sim", "wordregion_sim_bwd", "pair_ce", "pair_ce_bwd",
    "words_loss", "words_loss_grads", "cosine_scores", "cosine_scores_bwd", "class_mask",
    "sent_loss", "sent_loss_grads", "global_loss", "clip_loss", "arc_margin", "arc_margin_bwd",
    "cross_entropy_mean", "focal_loss", "focal_loss_bwd", "mag_linear", "mag_loss",
    "mag_head_grads", "linear_margin",
]


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _softmax(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def _logsumexp(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return np.squeeze(m, axis=axis) + np.log(np.sum(np.exp(x - m), axis=axis))


# --------------------------------------------------------------------------------------
# func_attention   (reference: models/attention.py:10-43)
# --------------------------------------------------------------------------------------
def func_attention(query, context, gamma1):
    """query [B, D, T], context [B, D, ih, iw] -> (weightedContext [B, D, T], attn [B, T, ih, iw]).

    models/attention.py:27  S[b, r, t] = sum_d ctx[b, d, r] * q[b, d, t]
    models/attention.py:28-29  softmax over t;  :35-36  times gamma1, softmax over r
    models/attention.py:41  W[b, d, t] = sum_r ctx[b, d, r] * A2[b, t, r]
    """
    query = np.asarray(query)
    context = np.asarray(context)
    B, D, T = query.shape
    ih, iw = context.shape[2], context.shape[3]
    c = context.reshape(B, D, ih * iw)                      # [B, D, R]
    s = np.einsum("bdr,bdt->brt", c, query)                 # [B, R, T]
    a1 = _softmax(s, axis=2)                                # over words
    a2 = _softmax(gamma1 * np.swapaxes(a1, 1, 2), axis=2)   # [B, T, R] over regions
    w = np.einsum("bdr,btr->bdt", c, a2)
    return w, a2.reshape(B, T, ih, iw)


def func_attention_bwd(query, context, gamma1, g_wc, g_attn):
    """(dquery, dcontext) for upstream gradients of both outputs of func_attention."""
    query = np.asarray(query, dtype=np.float64)
    context = np.asarray(context, dtype=np.float64)
    B, D, T = query.shape
    ih, iw = context.shape[2], context.shape[3]
    R = ih * iw
    c = context.reshape(B, D, R)
    gw = np.asarray(g_wc, dtype=np.float64)
    ga = np.asarray(g_attn, dtype=np.float64).reshape(B, T, R)
    s = np.einsum("bdr,bdt->brt", c, query)
    a1 = _softmax(s, axis=2)
    a2 = _softmax(gamma1 * np.swapaxes(a1, 1, 2), axis=2)
    da2 = ga + np.einsum("bdt,bdr->btr", gw, c)
    dc = np.einsum("bdt,btr->bdr", gw, a2)
    dz2 = a2 * (da2 - np.sum(a2 * da2, axis=2, keepdims=True))
    da1 = gamma1 * np.swapaxes(dz2, 1, 2)                         # [B, R, T]
    ds = a1 * (da1 - np.sum(a1 * da1, axis=2, keepdims=True))
    dq = np.einsum("bdr,brt->bdt", c, ds)
    dc += np.einsum("bdt,brt->bdr", query, ds)
    return dq, dc.reshape(context.shape)


# --------------------------------------------------------------------------------------
# word-region similarity matrix  (reference: models/losses.py:73-114, 122)
# --------------------------------------------------------------------------------------
def cosine_similarity(x1, x2, dim=1, eps=1e-8):
    """models/losses.py:12-16: sum(x1*x2, dim) / clamp(|x1|*|x2|, min=eps), squeezed."""
    x1, x2 = np.asarray(x1, dtype=np.float64), np.asarray(x2, dtype=np.float64)
    w12 = np.sum(x1 * x2, axis=dim)
    w1 = np.sqrt(np.sum(x1 * x1, axis=dim))
    w2 = np.sqrt(np.sum(x2 * x2, axis=dim))
    return np.squeeze(w12 / np.maximum(w1 * w2, eps))


def cosine_similarity_bwd(x1, x2, g, dim=1, eps=1e-8):
    """(dx1, dx2) for an upstream gradient g of the un-squeezed result (the clamp passes no gradient below eps)."""
    x1, x2 = np.asarray(x1, dtype=np.float64), np.asarray(x2, dtype=np.float64)
    g = np.expand_dims(np.asarray(g, dtype=np.float64).reshape(np.sum(x1 * x2, axis=dim).shape), dim)
    w12 = np.sum(x1 * x2, axis=dim, keepdims=True)
    w1 = np.sqrt(np.sum(x1 * x1, axis=dim, keepdims=True))
    w2 = np.sqrt(np.sum(x2 * x2, axis=dim, keepdims=True))
    prod = w1 * w2
    live = prod > eps
    inv = 1.0 / np.maximum(prod, eps)
    with np.errstate(divide="ignore", invalid="ignore"):
        k1 = np.where(live & (w1 > 0), w12 * inv / (w1 * w1), 0.0)
        k2 = np.where(live & (w2 > 0), w12 * inv / (w2 * w2), 0.0)
    return g * (x2 * inv - k1 * x1), g * (x1 * inv - k2 * x2)


def _lens(cap_lens, Bq, T):
    if cap_lens is None:
        return [T] * Bq
    return [int(v) for v in np.asarray(cap_lens).reshape(-1)]


def _pair_forward(ctx, q, g1, g2, eps):
    """All images against one caption.  ctx [Bc, R, D], q [T, D]."""
    s = np.einsum("brd,td->btr", ctx, q)                    # [Bc, T, R]
    a1 = _softmax(s, axis=1)                                # over words  (attention.py:28-29)
    a2 = _softmax(g1 * a1, axis=2)                          # over regions (attention.py:35-36)
    w = np.einsum("btr,brd->btd", a2, ctx)                  # [Bc, T, D]  (attention.py:41)
    nq = np.sqrt(np.sum(q * q, axis=1))[None, :]            # [1, T]
    nw = np.sqrt(np.sum(w * w, axis=2))                     # [Bc, T]
    prod = nq * nw
    den = np.maximum(prod, eps)                             # losses.py:12-16
    cos = np.einsum("td,btd->bt", q, w) / den
    return s, a1, a2, w, nq, nw, prod, den, cos


def wordregion_sim(ctx, words, cap_lens, gamma1, gamma2, gamma3, eps=1e-8, return_attn=True):
    """sim[b, i] = gamma3 * log sum_t exp(gamma2 * cos(q_it, W_bit))  (losses.py:104-109, 122).

    Returns (sim [Bc, Bq], attn_diag) where attn_diag[i] is A2 of pair (b=i, i), shape
    [T_i, R] (losses.py:97), only for i < min(Bc, Bq).
    """
    ctx = np.asarray(ctx, dtype=np.float64)
    words = np.asarray(words, dtype=np.float64)
    Bc, R, D = ctx.shape
    Bq, T, _ = words.shape
    lens = _lens(cap_lens, Bq, T)
    sim = np.zeros((Bc, Bq))
    attn = []
    for i in range(Bq):
        q = words[i, : lens[i]]
        _, _, a2, _, _, _, _, _, cos = _pair_forward(ctx, q, gamma1, gamma2, eps)
        sim[:, i] = gamma3 * np.log(np.sum(np.exp(gamma2 * cos), axis=1))
        if return_attn and i < Bc:
            attn.append(a2[i].copy())
    return sim, attn


def wordregion_sim_bwd(ctx, words, cap_lens, gamma1, gamma2, gamma3, gsim, eps=1e-8):
    """Gradients of sum(gsim * sim) w.r.t. ctx and words (closed form)."""
    ctx = np.asarray(ctx, dtype=np.float64)
    words = np.asarray(words, dtype=np.float64)
    gsim = np.asarray(gsim, dtype=np.float64)
    Bc, R, D = ctx.shape
    Bq, T, _ = words.shape
    lens = _lens(cap_lens, Bq, T)
    dctx = np.zeros_like(ctx)
    dwords = np.zeros_like(words)
    for i in range(Bq):
        q = words[i, : lens[i]]
        s, a1, a2, w, nq, nw, prod, den, cos = _pair_forward(ctx, q, gamma1, gamma2, eps)
        p = _softmax(gamma2 * cos, axis=1)                           # [Bc, T]
        dcos = gsim[:, i][:, None] * gamma3 * gamma2 * p             # [Bc, T]
        live = prod > eps                                            # clamp(min=eps) passes grad iff unclamped
        inv_den = 1.0 / den
        # d cos / d w  and  d cos / d q (direct)
        with np.errstate(divide="ignore", invalid="ignore"):
            cw = np.where(live, cos / np.maximum(nw * nw, 1e-300), 0.0)
            cq = np.where(live, cos / np.maximum(nq * nq, 1e-300), 0.0)
        dw = dcos[:, :, None] * (q[None] * inv_den[:, :, None] - cw[:, :, None] * w)
        dq = np.einsum("bt,btd->td", dcos * inv_den, w) - np.einsum("bt,td->td", dcos * cq, q)
        da2 = np.einsum("btd,brd->btr", dw, ctx)
        dctx += np.einsum("btr,btd->brd", a2, dw)
        dz2 = a2 * (da2 - np.sum(a2 * da2, axis=2, keepdims=True))
        da1 = gamma1 * dz2
        ds = a1 * (da1 - np.sum(a1 * da1, axis=1, keepdims=True))
        dq += np.einsum("btr,brd->td", ds, ctx)
        dctx += np.einsum("btr,td->brd", ds, q)
        dwords[i, : lens[i]] = dq
    return dctx, dwords


# --------------------------------------------------------------------------------------
# two-direction cross entropy over a [B, B] score matrix
#   (reference: models/losses.py:49-53, 128-132, 348-350)
# --------------------------------------------------------------------------------------
def pair_ce(sim, labels=None):
    """loss0 = CE(sim, labels) (rows), loss1 = CE(sim^T, labels) (columns); mean reduction."""
    sim = np.asarray(sim, dtype=np.float64)
    B = sim.shape[0]
    assert sim.shape == (B, B)
    lab = np.arange(B) if labels is None else np.asarray(labels).astype(np.int64)
    lse_r = _logsumexp(sim, axis=1)
    lse_c = _logsumexp(sim, axis=0)
    loss0 = float(np.mean(lse_r - sim[np.arange(B), lab]))
    loss1 = float(np.mean(lse_c - sim[lab, np.arange(B)]))
    return loss0, loss1


def pair_ce_bwd(sim, labels=None, g0=1.0, g1=1.0):
    """d(g0*loss0 + g1*loss1)/d sim.  -inf entries get exactly zero gradient."""
    sim = np.asarray(sim, dtype=np.float64)
    B = sim.shape[0]
    lab = np.arange(B) if labels is None else np.asarray(labels).astype(np.int64)
    pr = _softmax(sim, axis=1)
    pc = _softmax(sim, axis=0)
    oh_r = np.zeros_like(sim)
    oh_r[np.arange(B), lab] = 1.0
    oh_c = np.zeros_like(sim)
    oh_c[lab, np.arange(B)] = 1.0
    return (g0 * (pr - oh_r) + g1 * (pc - oh_c)) / B


def words_loss(ctx, words, labels, cap_lens, gamma1, gamma2, gamma3, eps=1e-8):
    """(loss0, loss1, attn_diag, sim) -- reference models/losses.py:61-135 (class_ids ignored there)."""
    sim, attn = wordregion_sim(ctx, words, cap_lens, gamma1, gamma2, gamma3, eps)
    l0, l1 = pair_ce(sim, labels)
    return l0, l1, attn, sim


def words_loss_grads(ctx, words, labels, cap_lens, gamma1, gamma2, gamma3, g0=1.0, g1=1.0, eps=1e-8):
    sim, _ = wordregion_sim(ctx, words, cap_lens, gamma1, gamma2, gamma3, eps, return_attn=False)
    gsim = pair_ce_bwd(sim, labels, g0, g1)
    return wordregion_sim_bwd(ctx, words, cap_lens, gamma1, gamma2, gamma3, gsim, eps)


# --------------------------------------------------------------------------------------
# sentence / global / CLIP losses  (reference: models/losses.py:19-57, 268-309, 329-351)
# --------------------------------------------------------------------------------------
def class_mask(class_ids):
    """mask[i, j] = (class_ids[j] == class_ids[i]) and i != j   (losses.py:20-30)."""
    c = np.asarray(class_ids).reshape(-1)
    m = c[None, :] == c[:, None]
    np.fill_diagonal(m, False)
    return m


def cosine_scores(x, y, scale, normalise=True, eps=1e-8):
    """scores[i, j] = scale * <x_i, y_j> / max(|x_i| |y_j|, eps)   (losses.py:38-43, 338-343);
    normalise=False gives the ClipLoss logits scale * x @ y^T (losses.py:293)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    dots = x @ y.T
    if not normalise:
        return scale * dots
    nx = np.sqrt(np.sum(x * x, axis=1))
    ny = np.sqrt(np.sum(y * y, axis=1))
    return dots / np.maximum(nx[:, None] * ny[None, :], eps) * scale


def cosine_scores_bwd(x, y, scale, gscores, normalise=True, eps=1e-8):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    g = np.asarray(gscores, dtype=np.float64) * scale
    if not normalise:
        return g @ y, g.T @ x
    nx = np.sqrt(np.sum(x * x, axis=1))
    ny = np.sqrt(np.sum(y * y, axis=1))
    prod = nx[:, None] * ny[None, :]
    den = np.maximum(prod, eps)
    live = prod > eps
    dots = x @ y.T
    gd = g / den                                    # through the numerator
    dx = gd @ y
    dy = gd.T @ x
    # through the denominator (only where the clamp is inactive)
    k = np.where(live, -g * dots / (den * den), 0.0)          # d/d(prod)
    with np.errstate(divide="ignore", invalid="ignore"):
        dx += (np.sum(k * ny[None, :], axis=1) / np.maximum(nx, 1e-300))[:, None] * x
        dy += (np.sum(k * nx[:, None], axis=0) / np.maximum(ny, 1e-300))[:, None] * y
    return dx, dy


def sent_loss(x, y, labels, class_ids, gamma3, eps=1e-8):
    """(loss0, loss1, scores) -- reference models/losses.py:19-57 with x = cnn_code, y = rnn_code."""
    sc = cosine_scores(x, y, gamma3, True, eps)
    if class_ids is not None:
        sc = np.where(class_mask(class_ids), -np.inf, sc)     # losses.py:47-48
    l0, l1 = pair_ce(sc, labels)
    return l0, l1, sc


def sent_loss_grads(x, y, labels, class_ids, gamma3, g0=1.0, g1=1.0, eps=1e-8):
    sc = cosine_scores(x, y, gamma3, True, eps)
    if class_ids is not None:
        sc = np.where(class_mask(class_ids), -np.inf, sc)
    g = pair_ce_bwd(sc, labels, g0, g1)                       # zero at the -inf entries
    return cosine_scores_bwd(x, y, gamma3, g, True, eps)


def global_loss(x, y, eps=1e-8, temp3=10.0):
    """loss0 + loss1, labels = arange   (losses.py:329-351)."""
    l0, l1, _ = sent_loss(x, y, None, None, temp3, eps)
    return l0 + l1


def global_loss_grads(x, y, eps=1e-8, temp3=10.0):
    """(dx, dy) of global_loss."""
    return sent_loss_grads(x, y, None, None, temp3, 1.0, 1.0, eps)


def clip_loss(text, image, logit_scale=1.0):
    """(CE(scale * img @ txt^T) + CE(scale * txt @ img^T)) / 2   (losses.py:292-309)."""
    sc = cosine_scores(image, text, logit_scale, normalise=False)
    l0, l1 = pair_ce(sc, None)
    return 0.5 * (l0 + l1)


# --------------------------------------------------------------------------------------
# ArcFace margin head + focal loss  (reference: models/metrics.py:17-60, models/losses.py:313-325)
# --------------------------------------------------------------------------------------
def _normalize_rows(x, eps=1e-12):
    n = np.sqrt(np.sum(x * x, axis=1, keepdims=True))
    return x / np.maximum(n, eps), n


def arc_margin(x, weight, label, s=30.0, m=0.50, easy_margin=False):
    """Dense logits [B, C]  (metrics.py:42-60).  weight is [C, Din]."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64)
    label = np.asarray(label).astype(np.int64).reshape(-1)
    xh, _ = _normalize_rows(x)
    wh, _ = _normalize_rows(w)
    cos = xh @ wh.T                                                  # metrics.py:44
    sin = np.sqrt(np.clip(1.0 - cos * cos, 0.0, 1.0))                # :45
    phi = cos * math.cos(m) - sin * math.sin(m)                      # :46
    if easy_margin:
        phi = np.where(cos > 0, phi, cos)                            # :48
    else:
        phi = np.where(cos > math.cos(math.pi - m), phi, cos - math.sin(math.pi - m) * m)  # :50
    out = cos.copy()
    rows = np.arange(x.shape[0])
    out[rows, label] = phi[rows, label]                              # :53-56
    return out * s                                                   # :57


def arc_margin_bwd(x, weight, label, glogits, s=30.0, m=0.50, easy_margin=False):
    """(dx, dweight) for upstream gradient glogits [B, C]."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64)
    g = np.asarray(glogits, dtype=np.float64)
    label = np.asarray(label).astype(np.int64).reshape(-1)
    xh, nx = _normalize_rows(x)
    wh, nw = _normalize_rows(w)
    cos = xh @ wh.T
    rows = np.arange(x.shape[0])
    ct = cos[rows, label]
    one_m = 1.0 - ct * ct
    inside = (one_m >= 0.0) & (one_m <= 1.0)
    sin_t = np.sqrt(np.clip(one_m, 0.0, 1.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        dphi = math.cos(m) + np.where(inside, ct / sin_t, 0.0) * math.sin(m)
    if easy_margin:
        dphi = np.where(ct > 0, dphi, 1.0)
    else:
        dphi = np.where(ct > math.cos(math.pi - m), dphi, 1.0)
    dcos = g * s
    dcos[rows, label] *= dphi
    dxh = dcos @ wh
    dwh = dcos.T @ xh
    dx = (dxh - np.sum(dxh * xh, axis=1, keepdims=True) * xh) / np.maximum(nx, 1e-12)
    dw = (dwh - np.sum(dwh * wh, axis=1, keepdims=True) * wh) / np.maximum(nw, 1e-12)
    return dx, dw


def cross_entropy_mean(logits, target):
    logits = np.asarray(logits, dtype=np.float64)
    target = np.asarray(target).astype(np.int64).reshape(-1)
    lse = _logsumexp(logits, axis=1)
    return float(np.mean(lse - logits[np.arange(logits.shape[0]), target]))


def focal_loss(logits, target, gamma=0.0):
    """(1 - exp(-CE))**gamma * CE with CE the *batch-mean* cross entropy  (losses.py:321-325)."""
    logp = cross_entropy_mean(logits, target)
    p = math.exp(-logp)
    return (1.0 - p) ** gamma * logp


def focal_loss_bwd(logits, target, gamma=0.0, gout=1.0):
    logits = np.asarray(logits, dtype=np.float64)
    target = np.asarray(target).astype(np.int64).reshape(-1)
    B = logits.shape[0]
    logp = cross_entropy_mean(logits, target)
    p = math.exp(-logp)
    dl = (1.0 - p) ** gamma
    if gamma != 0:
        dl += gamma * (1.0 - p) ** (gamma - 1.0) * p * logp
    sm = _softmax(logits, axis=1)
    sm[np.arange(B), target] -= 1.0
    return sm * (gout * dl / B)


# --------------------------------------------------------------------------------------
# MagFace head  (reference: models/magface.py:56-61, 69-108, 111-136)
# --------------------------------------------------------------------------------------
def linear_margin(x_norm, l_a, u_a, l_margin, u_margin):
    """SoftmaxBuilder._margin  (magface.py:56-61)."""
    return (u_margin - l_margin) / (u_a - l_a) * (x_norm - l_a) + l_margin


def mag_linear(x, weight, l_a, u_a, l_margin, u_margin, scale=64.0, easy_margin=True):
    """([cos_theta*s, cos_theta_m*s], x_norm [B,1]) with weight [Din, C]  (magface.py:83-108)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64)
    x_norm = np.clip(np.sqrt(np.sum(x * x, axis=1, keepdims=True)), l_a, u_a)      # :87
    mar = linear_margin(x_norm, l_a, u_a, l_margin, u_margin)                      # :88
    cm, sm = np.cos(mar), np.sin(mar)
    wn = w / np.maximum(np.sqrt(np.sum(w * w, axis=0, keepdims=True)), 1e-12)      # :92
    xh, _ = _normalize_rows(x)
    cos = np.clip(xh @ wn, -1.0, 1.0)                                              # :93-94
    sin = np.sqrt(1.0 - cos * cos)                                                 # :95
    cos_m = cos * cm - sin * sm                                                    # :96
    if easy_margin:
        cos_m = np.where(cos > 0, cos_m, cos)                                      # :98
    else:
        mm = np.sin(math.pi - mar) * mar
        th = np.cos(math.pi - mar)
        cos_m = np.where(cos > th, cos_m, cos - mm)                                # :100-103
    return [scale * cos, scale * cos_m], x_norm


def mag_loss(logits_pair, target, x_norm, u_a):
    """(loss, loss_g, one_hot)  (magface.py:124-136)."""
    cos, cos_m = (np.asarray(a, dtype=np.float64) for a in logits_pair)
    target = np.asarray(target).astype(np.int64).reshape(-1)
    x_norm = np.asarray(x_norm, dtype=np.float64)
    loss_g = float(np.mean(x_norm / (u_a ** 2) + 1.0 / x_norm))
    one_hot = np.zeros_like(cos)
    one_hot[np.arange(cos.shape[0]), target] = 1.0
    out = one_hot * cos_m + (1.0 - one_hot) * cos
    return cross_entropy_mean(out, target), loss_g, one_hot


def mag_head_grads(x, weight, target, l_a, u_a, l_margin, u_margin, scale=64.0, easy_margin=True,
                   g_loss=1.0, g_lossg=0.0):
    """(dx, dweight) of g_loss*loss + g_lossg*loss_g through MagLinear + MagLoss."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64)
    target = np.asarray(target).astype(np.int64).reshape(-1)
    B = x.shape[0]
    rows = np.arange(B)
    raw = np.sqrt(np.sum(x * x, axis=1, keepdims=True))
    x_norm = np.clip(raw, l_a, u_a)
    slope = (u_margin - l_margin) / (u_a - l_a)
    mar = slope * (x_norm - l_a) + l_margin
    cm, sm = np.cos(mar), np.sin(mar)
    nwc = np.sqrt(np.sum(w * w, axis=0, keepdims=True))
    wn = w / np.maximum(nwc, 1e-12)
    xh, nx = _normalize_rows(x)
    craw = xh @ wn
    cos = np.clip(craw, -1.0, 1.0)
    sin = np.sqrt(1.0 - cos * cos)
    cos_m = cos * cm - sin * sm
    if easy_margin:
        use = cos > 0
        alt_dm = np.zeros_like(cos)
    else:
        th = np.cos(math.pi - mar)
        use = cos > th
        # d/dmar of (cos - sin(pi-mar)*mar) = cos(pi-mar)*mar - sin(pi-mar)
        alt_dm = (np.cos(math.pi - mar) * mar - np.sin(math.pi - mar)) * np.ones_like(cos)
    out = scale * cos
    sel = np.where(use, cos_m, cos if easy_margin else cos - np.sin(math.pi - mar) * mar)
    out[rows, target] = scale * sel[rows, target]
    p = _softmax(out, axis=1)
    p[rows, target] -= 1.0
    gout = p * (g_loss / B)                                  # d/d out
    # back through the blend: non-target columns -> cos*scale ; target column -> sel*scale
    dcos = gout * scale
    ct = cos[rows, target]
    st = sin[rows, target]
    with np.errstate(divide="ignore", invalid="ignore"):
        dsel_dcos = np.where(use[rows, target], cm[:, 0] + (ct / st) * sm[:, 0], 1.0)
    dsel_dmar = np.where(use[rows, target], -ct * sm[:, 0] - st * cm[:, 0], alt_dm[rows, target])
    gt = gout[rows, target] * scale
    dcos[rows, target] = gt * dsel_dcos
    dmar = (gt * dsel_dmar)[:, None]                         # [B, 1]
    # clamp(-1, 1) passes gradient inside the closed interval
    dcos = np.where((craw >= -1.0) & (craw <= 1.0), dcos, 0.0)
    dxh = dcos @ wn.T
    dwn = xh.T @ dcos
    dx = (dxh - np.sum(dxh * xh, axis=1, keepdims=True) * xh) / np.maximum(nx, 1e-12)
    dw = (dwn - np.sum(dwn * wn, axis=0, keepdims=True) * wn) / np.maximum(nwc, 1e-12)
    # x_norm path: margin and the g regulariser
    dxn = dmar * slope + g_lossg * (1.0 / (u_a ** 2) - 1.0 / (x_norm * x_norm)) / B
    dxn = np.where((raw >= l_a) & (raw <= u_a), dxn, 0.0)
    dx += dxn * x / np.maximum(raw, 1e-300)
    return dx, dw
