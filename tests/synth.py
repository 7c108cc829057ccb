"""Seeded synthetic inputs for the FCAM hot path.

numpy ``RandomState`` (MT19937) is used instead of torch's generator so that the
golden-fixture script, the CPU tests, the GPU tests and ``bench.py`` regenerate
bit-identical inputs everywhere, independent of the torch version.

Canonical layouts: ctx [B, R, D] (unit-L2 rows), words [B, T, D].
"""
from __future__ import annotations

import numpy as np


def _unit(a, axis=-1):
    n = np.sqrt(np.sum(a.astype(np.float64) ** 2, axis=axis, keepdims=True))
    return (a / n).astype(np.float32)


def wordregion_inputs(B, T, R, D, flavour="BERT", seed=100, ragged=False):
    """ctx [B,R,D] unit rows; words [B,T,D] (BERT: unit rows; LSTM: 0.5*randn); cap_lens or None."""
    rs = np.random.RandomState(seed)
    ctx = _unit(rs.randn(B, R, D).astype(np.float32))
    if flavour == "BERT":
        words = _unit(rs.randn(B, T, D).astype(np.float32))
        cap_lens = None
    else:
        words = (0.5 * rs.randn(B, T, D)).astype(np.float32)
        cap_lens = rs.randint(max(2, T // 3), T + 1, size=B).astype(np.int64) if ragged \
            else np.full(B, T, dtype=np.int64)
        if ragged:
            cap_lens[0] = T            # keep the padded width exercised
    return ctx, words, cap_lens


def sentence_inputs(B, D, seed=100, collisions=False):
    rs = np.random.RandomState(seed + 1)
    img = _unit(rs.randn(B, D).astype(np.float32))
    txt = _unit(rs.randn(B, D).astype(np.float32))
    if collisions:
        class_ids = rs.randint(0, max(2, B // 2), size=B).astype(np.int64)
    else:
        class_ids = np.arange(B, dtype=np.int64)
    return img, txt, class_ids


def margin_inputs(B, Din, C, seed=100, mag=False):
    rs = np.random.RandomState(seed + 2)
    x = rs.randn(B, Din).astype(np.float32)
    if mag:
        x = (x * rs.uniform(0.5, 6.0, size=(B, 1))).astype(np.float32)
        w = rs.uniform(-1, 1, size=(Din, C)).astype(np.float32)
        w = _unit(w, axis=0)                       # MagLinear init renormalises columns
    else:
        bound = np.sqrt(6.0 / (Din + C))           # xavier_uniform
        w = rs.uniform(-bound, bound, size=(C, Din)).astype(np.float32)
    label = rs.randint(0, C, size=B).astype(np.int64)
    return x, w, label


def texthead_inputs(B, bert_words_num, F, E=768, seed=100):
    """BERT-like token features [B, L, E] (L = bert_words_num - 1) and the three n-gram convolutions of
    Bert_Word_Mapping: weights [F, K, E] and biases [F] for K = 2, 3, 4 (Conv2d default init range)."""
    rs = np.random.RandomState(seed + 3)
    L = bert_words_num - 1
    tokens = rs.randn(B, L, E).astype(np.float32)
    ws, bs = [], []
    for K in (2, 3, 4):
        bound = 1.0 / np.sqrt(K * E)
        ws.append(rs.uniform(-bound, bound, size=(F, K, E)).astype(np.float32))
        bs.append(rs.uniform(-bound, bound, size=(F,)).astype(np.float32))
    return tokens, ws, bs
