"""fp64 numpy restatement of the reference's TextHeading (models/models.py:170-232) -- TEST INFRASTRUCTURE.

Pinned to fixtures generated from the reference itself (tests/golden/make_golden_texthead.py ->
tests/golden/texthead_*.npz; checked by tests/test_oracle_golden.py).  Never imported by the product package.

Layouts: tokens [B, L, E] with L = bert_words_num - 1; conv weights [F, K, E] (the reference's [F, 1, K, E]
squeezed), K = 2, 3, 4; words [B, T, F] (the reference returns the transpose), T = bert_words_num - 2.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-12          # F.normalize


def conv_relu(tokens, w, b):
    """relu(Conv2d(1, F, (K, E))(tokens))  (models.py:183-185) -> [B, L-K+1, F]."""
    tokens = np.asarray(tokens, np.float64)
    w = np.asarray(w, np.float64)
    B, L, E = tokens.shape
    F, K, _ = w.shape
    win = np.stack([tokens[:, k:L - K + 1 + k, :] for k in range(K)], axis=2)      # [B, L-K+1, K, E]
    out = np.einsum("bjke,fke->bjf", win, w)
    if b is not None:
        out = out + np.asarray(b, np.float64)
    return np.maximum(out, 0.0), win


def _code(acts, seq):
    """get_each_word_feature before the normalisation (models.py:197-210): [B, seq+2, F] and the candidates."""
    a, b, c = acts
    rows = [np.maximum(np.maximum(a[:, j], b[:, j]), c[:, j]) for j in range(seq)]
    rows.append(np.maximum(a[:, seq], b[:, seq]))
    rows.append(a[:, seq + 1])
    return np.stack(rows, axis=1)


def forward(tokens, ws, bs, bert_words_num):
    """(words [B, T, F], sent [B, F]) -- models.py:218-231 (words before the final transpose)."""
    seq = bert_words_num - 4
    acts = [conv_relu(tokens, w, b)[0] for w, b in zip(ws, bs)]
    code = _code(acts, seq)
    words = code / np.maximum(np.linalg.norm(code, axis=2, keepdims=True), EPS)
    m = np.stack([a.max(axis=1) for a in acts]).mean(axis=0)                         # max_pool1d, mean over the convs
    sent = m / np.maximum(np.linalg.norm(m, axis=1, keepdims=True), EPS)
    return words, sent


def _normalize_bwd(v, g):
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    vh = v / np.maximum(n, EPS)
    return np.where(n > EPS, (g - np.sum(g * vh, axis=-1, keepdims=True) * vh) / np.maximum(n, EPS), g / EPS)


def backward(tokens, ws, bs, bert_words_num, gwords, gsent):
    """([dW_K [F, K, E]], [db_K [F]]) for upstream gwords [B, T, F] / gsent [B, F] (either may be None).
    torch.amax shares the gradient evenly between ties; max_pool1d takes the first maximum; the last word is
    detached by the reference's torch.cuda.FloatTensor copy (models.py:206)."""
    seq = bert_words_num - 4
    conv = [conv_relu(tokens, w, b) for w, b in zip(ws, bs)]
    acts = [c[0] for c in conv]
    G = [np.zeros_like(a) for a in acts]
    if gwords is not None:
        code = _code(acts, seq)
        dcode = _normalize_bwd(code, np.asarray(gwords, np.float64))
        for j in range(seq + 1):                                                     # word seq + 1: no gradient
            cands = acts[:3] if j < seq else acts[:2]
            vals = np.stack([c[:, j] for c in cands])                                # [n, B, F]
            hit = vals == vals.max(axis=0, keepdims=True)
            share = dcode[:, j] / hit.sum(axis=0)
            for k in range(len(cands)):
                G[k][:, j] += np.where(hit[k], share, 0.0)
    if gsent is not None:
        m = np.stack([a.max(axis=1) for a in acts]).mean(axis=0)
        dm = _normalize_bwd(m, np.asarray(gsent, np.float64)) / 3.0
        for k, a in enumerate(acts):
            jm = a.argmax(axis=1)                                                    # first maximum, [B, F]
            bi, fi = np.meshgrid(np.arange(a.shape[0]), np.arange(a.shape[2]), indexing="ij")
            np.add.at(G[k], (bi, jm, fi), dm)
    dws, dbs = [], []
    for k, (a, win) in enumerate(conv):
        g = G[k] * (a > 0)                                                           # ReLU
        dws.append(np.einsum("bjf,bjke->fke", g, win))
        dbs.append(g.sum(axis=(0, 1)))
    return dws, dbs
