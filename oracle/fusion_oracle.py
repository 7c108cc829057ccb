"""CPU oracle for the FCFM fusion net `Working` (eval-mode forward) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

SURVEY.md section 8(f) row f4 (the producer of the 640-d fused embeddings that verification scoring consumes, BASELINE
configs[4]).  No CUDA kernel exists for this row yet: this restatement and its fixtures
(tests/golden/fusion_working_*.npz, generated from the reference by tests/golden/make_golden_fusion.py) are the
checker the kernel will be built against.  Plain numpy, float64; each step cites /root/reference/models/fusion_nets.py.

forward(params, img [B,256,14,14], word [B,256,T], gl_img [B,256], sent [B,256]) -> [B, 640]
`params` holds the module's state_dict as numpy arrays (BatchNorm in eval mode: running statistics).
"""
from __future__ import annotations

import numpy as np


def _conv3x3_valid(x, w, b):
    """nn.Conv2d(256, 36, 3, padding=0) (fusion_nets.py:226): [B,Cin,H,W] -> [B,Cout,H-2,W-2]."""
    B, Cin, H, W = x.shape
    out = np.zeros((B, w.shape[0], H - 2, W - 2))
    for dy in range(3):
        for dx in range(3):
            out += np.einsum("bchw,oc->bohw", x[:, :, dy:dy + H - 2, dx:dx + W - 2], w[:, :, dy, dx])
    return out + b[None, :, None, None]


def _maxpool2(x):
    """nn.MaxPool2d(2) (fusion_nets.py:225): floor mode."""
    B, C, H, W = x.shape
    x = x[:, :, : H // 2 * 2, : W // 2 * 2].reshape(B, C, H // 2, 2, W // 2, 2)
    return x.max(axis=(3, 5))


def _bn_eval(x, p, name, eps=1e-5):
    g, b, m, v = (p[f"{name}.{k}"] for k in ("weight", "bias", "running_mean", "running_var"))
    return (x - m[None, :, None, None]) / np.sqrt(v[None, :, None, None] + eps) * g[None, :, None, None] + b[None, :, None, None]


def _layernorm(x, w, b, eps=1e-5):
    """nn.LayerNorm over all non-batch dimensions of x (biased variance)."""
    flat = x.reshape(x.shape[0], -1)
    mu = flat.mean(1, keepdims=True)
    var = flat.var(1, keepdims=True)
    return ((flat - mu) / np.sqrt(var + eps)).reshape(x.shape) * w[None] + b[None]


def _softmax(x):
    e = np.exp(x - x.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def self_attention(p, x, y, prefix="sa", scale=1):
    """SelfAttention.forward(x = image, y = text) (fusion_nets.py:82-118): 1x1 projections, key^T.query / sqrt(C/scale),
    softmax over the text positions, attention.value, back to [N,C,W,H]."""
    conv1 = lambda t, n: np.einsum("bchw,oc->bohw", t, p[f"{prefix}.{n}.weight"][:, :, 0, 0]) + p[f"{prefix}.{n}.bias"][None, :, None, None]
    q = conv1(y, "query_proj")
    N, C, W, H = q.shape
    q = q.reshape(N, C, H * W)
    k = conv1(x, "key_proj").reshape(N, C, -1).transpose(0, 2, 1)            # [N, HW, C]
    att = _softmax(k @ q / np.sqrt(x.shape[1] / scale))                      # [N, HW, HW]
    v = conv1(x, "value_proj")
    Cv = v.shape[1]
    v = v.reshape(N, Cv, -1).transpose(0, 2, 1)                              # [N, HW, C]
    resp = (att @ v).transpose(0, 2, 1)                                      # [N, C, HW]
    return resp.reshape(N, Cv, y.shape[2], y.shape[3])


def working_forward(p, img, word, gl_img, sent):
    """Working.forward (fusion_nets.py:234-258), eval mode."""
    p = {k: np.asarray(v, np.float64) for k, v in p.items()}
    img, word, gl_img, sent = (np.asarray(a, np.float64) for a in (img, word, gl_img, sent))
    x = _maxpool2(np.maximum(_conv3x3_valid(img, p["conv.weight"], p["conv.bias"]), 0.0))       # :235
    x = _bn_eval(x, p, "bn_img")                                                                   # :236
    w = word.transpose(0, 2, 1) @ p["projection.weight"].T + p["projection.bias"]                 # :239  [B,T,36]
    w = w.transpose(0, 2, 1) @ w / np.sqrt(36)                                                     # :240  [B,36,36]
    w = w.reshape(w.shape[0], w.shape[1], 6, 6)                                                    # :241
    w = _bn_eval(w, p, "bn_word")                                                                  # :242
    iw = self_attention(p, x, w)                                                                   # :247
    iw = _layernorm(iw, p["ln.weight"], p["ln.bias"])                                              # :248
    iw = _maxpool2(iw).reshape(iw.shape[0], -1)                                                    # :249-250  [B,324]
    iw = iw @ p["linear.weight"].T + p["linear.bias"]                                              # :254  [B,128]
    g = _layernorm(gl_img, p["ln_gl_image.weight"], p["ln_gl_image.bias"])                         # :255
    s = _layernorm(sent, p["ln_sent.weight"], p["ln_sent.bias"])                                   # :256
    return np.concatenate((iw, g, s), axis=1)                                                      # :257


def _bn_train(x, p, name, eps=1e-5):
    """BatchNorm2d in training mode: batch statistics over (B, H, W), biased variance (what normalises)."""
    g, b = p[f"{name}.weight"], p[f"{name}.bias"]
    m = x.mean(axis=(0, 2, 3), keepdims=True)
    v = x.var(axis=(0, 2, 3), keepdims=True)
    return (x - m) / np.sqrt(v + eps) * g[None, :, None, None] + b[None, :, None, None]


def imim_forward(p, img, training=False):
    """IMIM.forward (reference models/models.py:380-405), eval mode (running statistics) or
    training mode (batch statistics; pinned by tests/golden/imim_train.npz): the producer of the
    word-region loss's region features.  img [B,256,14,14] -> [B,256,14,14] logical, unit L2 norm over the channel axis
    at every position (memory order of the reference's result: channels-last, models.py:401-404)."""
    p = {k: np.asarray(v, np.float64) for k, v in p.items()}
    x = np.asarray(img, np.float64)
    x = _bn_train(x, p, "bn_img") if training else _bn_eval(x, p, "bn_img")                     # :394
    x = self_attention(p, x, x)                                                                  # :395 (scale = 1)
    x = _layernorm(x, p["ln.weight"], p["ln.bias"])                                              # :396
    conv1 = lambda t, n: np.einsum("bchw,oc->bohw", t, p[f"{n}.weight"][:, :, 0, 0]) + p[f"{n}.bias"][None, :, None, None]
    x = np.maximum(conv1(x, "conv1x1_1"), 0.0)                                                   # :398
    x = np.maximum(conv1(x, "conv1x1_2"), 0.0)                                                   # :399
    x = x.transpose(0, 2, 3, 1)                                                                  # :401
    x = x @ p["project_local.projection.weight"].T + p["project_local.projection.bias"]         # ProjectionHead :112
    x = x / np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), 1e-12)                           # :119, then again :403
    x = x / np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), 1e-12)
    return x.transpose(0, 3, 1, 2)                                                               # :404
