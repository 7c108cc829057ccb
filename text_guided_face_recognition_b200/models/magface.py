"""Mirror of the reference's models/magface.py hot-path symbols: MagLinear and MagLoss.
(SoftmaxBuilder / load_features wire a backbone and are out of scope)"""
import numpy as np
import torch
import torch.nn.functional as F  # noqa: F401
from torch.nn import Parameter

from ._backend import ops


class MagLinear(torch.nn.Module):
    """Parallel fc for Mag loss, reference models/magface.py:69-108.

    forward(x, m, l_a, u_a) with `m` a callable margin(x_norm) returns
    ([scale*cos_theta, scale*cos_theta_m], x_norm).  The clamped-norm / margin scalars are [B,1]
    PyTorch ops (so any callable `m` keeps working and gradients reach x through it); the
    normalisation, cos-theta GEMM and the dense margin transform run in libtgfr_b200.so.
    `weight` keeps the reference's [in_features, out_features] shape.
    """

    def __init__(self, in_features, out_features, scale=64.0, easy_margin=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = Parameter(torch.empty(in_features, out_features))
        self.weight.data.uniform_(-1, 1).renorm_(2, 1, 1e-5).mul_(1e5)
        self.scale = scale
        self.easy_margin = easy_margin

    def forward(self, x, m, l_a, u_a):
        x_norm = torch.norm(x, dim=1, keepdim=True).clamp(l_a, u_a)       # magface.py:87
        ada_margin = m(x_norm)                                            # :88
        cos_theta, cos_theta_m = ops.mag_logits(x, self.weight, ada_margin, self.scale, self.easy_margin)
        return [cos_theta, cos_theta_m], x_norm


class MagLoss(torch.nn.Module):
    """MagFace loss, reference models/magface.py:111-136.  Returns (loss, loss_g, one_hot)."""

    def __init__(self, l_a, u_a, l_margin, u_margin, scale=64.0):
        super().__init__()
        self.l_a = l_a
        self.u_a = u_a
        self.scale = scale
        self.cut_off = np.cos(np.pi / 2 - l_margin)
        self.large_value = 1 << 10

    def calc_loss_G(self, x_norm):
        g = 1 / (self.u_a ** 2) * x_norm + 1 / (x_norm)
        return torch.mean(g)

    def forward(self, input, target, x_norm):
        loss_g = self.calc_loss_G(x_norm)
        cos_theta, cos_theta_m = input
        # label column from cos_theta_m, every other column from cos_theta (magface.py:131-134), mean CE (:135):
        # one pass over the two logit tensors in libtgfr_b200.so; one_hot is written by the same kernel
        loss, one_hot = ops.mag_ce(cos_theta, cos_theta_m, target)
        return loss.mean(), loss_g, one_hot
