"""Generate tests/golden/imim_small.npz from the reference's own IMIM module (models/models.py:380-405) in eval mode on
CPU (torchsummary, imported by models.models and absent here, is shimmed; transformers is present).

    python tests/golden/make_golden_imim.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.modules.setdefault("torchsummary", types.SimpleNamespace(summary=lambda *a, **k: None))
sys.path.insert(0, "/root/reference")
from models.models import IMIM  # noqa: E402


def main():
    torch.manual_seed(100)
    net = IMIM(types.SimpleNamespace(aux_feat_dim_per_granularity=256), channel_dim=256)
    with torch.no_grad():
        net.bn_img.running_mean.normal_(0, 0.3)
        net.bn_img.running_var.uniform_(0.5, 2.0)
        net.bn_img.weight.uniform_(0.5, 1.5)
        net.bn_img.bias.normal_(0, 0.2)
        net.ln.weight.uniform_(0.5, 1.5)
        net.ln.bias.normal_(0, 0.2)
    net.eval()
    g = torch.Generator().manual_seed(100)
    img = torch.randn(1, 256, 14, 14, generator=g)
    with torch.no_grad():
        out = net(img)
    data = {"img": img.numpy(), "out": out.contiguous().numpy(), "out_strides": np.array(out.stride())}
    for k, v in net.state_dict().items():
        if not k.endswith("num_batches_tracked") and not k.startswith("project_local.fc"):
            data["p:" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "imim_small.npz"), **data)
    print(out.shape, out.stride(), float(out.abs().mean()))


if __name__ == "__main__":
    main()
