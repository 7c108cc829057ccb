"""Round-2 fixtures of the FCFM fusion net from the reference's own `Working` module (models/fusion_nets.py:217-258,
UNMODIFIED, CPU, fp32) in TRAINING mode with autograd -- build container only:

    python tests/golden/make_golden_fusion_r2.py

  fusion_working_train.npz   B = 5, T = 22: out, gradients of the 26 parameters and the four inputs, updated running stats.
                             Parameters = those of fusion_working_bert22.npz; inputs from fusion_inputs (seeded numpy).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def fusion_inputs(B, T, seed):
    rs = np.random.RandomState(seed)
    unit = lambda a, ax: (a / np.linalg.norm(a, axis=ax, keepdims=True)).astype(np.float32)
    img = unit(rs.randn(B, 14, 14, 256), -1).transpose(0, 3, 1, 2).copy()          # IMIM-like: unit over channels
    word = unit(rs.randn(B, T, 256), -1).transpose(0, 2, 1).copy()                 # [B,256,T]
    gl = unit(rs.randn(B, 256), -1)
    sent = unit(rs.randn(B, 256), -1)
    gout = rs.randn(B, 640).astype(np.float32)
    return img, word, gl, sent, gout


def main():
    sys.modules.setdefault("torchsummary", types.SimpleNamespace(summary=lambda *a, **k: None))
    sys.path.insert(0, "/root/reference")
    from models.fusion_nets import Working  # noqa: E402
    torch.manual_seed(100)
    net = Working(channel_dim=256)
    base = np.load(os.path.join(HERE, "fusion_working_bert22.npz"))
    sd = net.state_dict()
    with torch.no_grad():
        for k in sd:
            if ("p:" + k) in base.files:
                sd[k].copy_(torch.from_numpy(base["p:" + k]))
    net.train()
    B, T = 5, 22
    img, word, gl, sent, gout = fusion_inputs(B, T, 9)
    leaves = [torch.from_numpy(a).requires_grad_(True) for a in (img, word, gl, sent)]
    out = net(*leaves)
    out.backward(torch.from_numpy(gout))
    data = {"out": out.detach().numpy(), "dimg": leaves[0].grad.numpy(), "dword": leaves[1].grad.numpy(),
            "dgl": leaves[2].grad.numpy(), "dsent": leaves[3].grad.numpy()}
    for k, v in net.named_parameters():
        data["g:" + k] = v.grad.numpy()
    for k, v in net.state_dict().items():
        if "running" in k:
            data["s:" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "fusion_working_train.npz"), **data)
    print("fusion_working_train", out.shape, float(out.abs().mean()), len([k for k in data if k.startswith("g:")]))


if __name__ == "__main__":
    main()
