"""Generate tests/golden/fusion_working_*.npz from the reference's own `Working` module (models/fusion_nets.py:217-258)
run in eval mode on CPU (torchsummary, imported by the module and absent here, is shimmed).

    python tests/golden/make_golden_fusion.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.modules.setdefault("torchsummary", types.SimpleNamespace(summary=lambda *a, **k: None))
sys.path.insert(0, "/root/reference")
from models.fusion_nets import Working  # noqa: E402


def main():
    for name, B, T, seed in (("small", 2, 6, 100), ("bert22", 2, 22, 101)):
        torch.manual_seed(seed)
        net = Working(channel_dim=256)
        with torch.no_grad():                     # non-trivial BatchNorm statistics and affine terms
            for bn in (net.bn_img, net.bn_word):
                bn.running_mean.normal_(0, 0.3)
                bn.running_var.uniform_(0.5, 2.0)
                bn.weight.uniform_(0.5, 1.5)
                bn.bias.normal_(0, 0.2)
            for ln in (net.ln, net.ln_gl_image, net.ln_sent):
                ln.weight.uniform_(0.5, 1.5)
                ln.bias.normal_(0, 0.2)
        net.eval()
        g = torch.Generator().manual_seed(seed)
        img = torch.nn.functional.normalize(torch.randn(B, 14, 14, 256, generator=g), dim=-1).permute(0, 3, 1, 2)
        word = torch.nn.functional.normalize(torch.randn(B, T, 256, generator=g), dim=2).transpose(1, 2)
        gl = torch.nn.functional.normalize(torch.randn(B, 256, generator=g), dim=1)
        sent = torch.nn.functional.normalize(torch.randn(B, 256, generator=g), dim=1)
        with torch.no_grad():
            out = net(img, word, gl, sent)
        data = {"img": img.contiguous().numpy(), "word": word.contiguous().numpy(), "gl_img": gl.numpy(), "sent": sent.numpy(),
                "out": out.numpy()}
        for k, v in net.state_dict().items():
            if not k.endswith("num_batches_tracked"):
                data["p:" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, f"fusion_working_{name}.npz"), **data)
        print(name, out.shape, float(out.abs().mean()))


if __name__ == "__main__":
    main()
