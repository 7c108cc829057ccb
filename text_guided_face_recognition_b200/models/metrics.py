"""Mirror of the reference's models/metrics.py.  Hot-path symbol: ArcMarginProduct (the only head
the reference instantiates, src/train_encoders_bert.py:140,162, src/fusion_bert.py:106).  The
other heads ride the same fused cosine-logits kernel with their margins applied in PyTorch."""
import math

import torch
import torch.nn as nn
from torch.nn import Parameter

from ._backend import ops


def l2_norm(input, axis=1):
    """Reference models/metrics.py:10-13."""
    return torch.div(input, torch.norm(input, 2, axis, True))


class ArcMarginProduct(nn.Module):
    r"""Large margin arc distance, reference models/metrics.py:17-60.

    forward(input [B,in_features], label [B]) -> dense logits [B,out_features]:
    s*cos(theta+m) on the label column (with the reference's threshold / easy-margin rule),
    s*cos(theta) elsewhere.  Normalisation, the cos-theta GEMM, the margin and the scale run in
    libtgfr_b200.so; `weight` keeps the reference's name and [out_features, in_features] shape.
    """

    def __init__(self, in_features, out_features, s=30.0, m=0.50, easy_margin=False):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.s = s
        self.m = m
        self.weight = Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)
        self.easy_margin = easy_margin
        self.cos_m = math.cos(m)
        self.sin_m = math.sin(m)
        self.th = math.cos(math.pi - m)
        self.mm = math.sin(math.pi - m) * m

    def forward(self, input, label):
        if input.dim() != 2 or input.size(1) != self.in_features:
            raise RuntimeError(f"ArcMarginProduct: expected input [B,{self.in_features}], got {tuple(input.shape)}")
        return ops.arc_logits(input, self.weight, label, self.s, self.m, self.easy_margin)

    def fused_loss(self, input, label, gamma=0.0):
        """FocalLoss(gamma)(self(input, label), label) -- the reference's `criterion(metric_fc(x, y), y)` pair
        (src/train_encoders_bert.py:290-305) -- without materialising the [B, out_features] logits: margin, online
        softmax and the softmax gradient run in the epilogues of the tensor-core cos-theta GEMM.  An extension of
        the reference API (its forward must return dense logits); gamma=0 is plain nn.CrossEntropyLoss."""
        if input.dim() != 2 or input.size(1) != self.in_features:
            raise RuntimeError(f"ArcMarginProduct: expected input [B,{self.in_features}], got {tuple(input.shape)}")
        return ops.arc_fused_focal(input, self.weight, label, self.s, self.m, self.easy_margin, gamma)


class AddMarginProduct(nn.Module):
    r"""CosFace head, reference models/metrics.py:63-102: s*(cos(theta) - m) on the label column."""

    def __init__(self, in_features, out_features, s=30.0, m=0.40):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.s = s
        self.m = m
        self.weight = Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, input, label):
        logits = ops.cos_logits(input, self.weight, self.s)
        shift = torch.zeros_like(logits).scatter_(1, label.view(-1, 1).long(), self.s * self.m)
        return logits - shift

    def __repr__(self):
        return (f"{self.__class__.__name__}(in_features={self.in_features}, out_features={self.out_features}, "
                f"s={self.s}, m={self.m})")


class SphereProduct(nn.Module):
    r"""SphereFace head, reference models/metrics.py:105-165 (annealed cos(m*theta) margin)."""

    _CHEBYSHEV = (
        lambda x: x ** 0,
        lambda x: x ** 1,
        lambda x: 2 * x ** 2 - 1,
        lambda x: 4 * x ** 3 - 3 * x,
        lambda x: 8 * x ** 4 - 8 * x ** 2 + 1,
        lambda x: 16 * x ** 5 - 20 * x ** 3 + 5 * x,
    )

    def __init__(self, in_features, out_features, m=4):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.m = m
        self.base, self.gamma, self.power, self.LambdaMin = 1000.0, 0.12, 1, 5.0
        self.iter = 0
        self.weight = Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, input, label):
        self.iter += 1
        self.lamb = max(self.LambdaMin, self.base * (1 + self.gamma * self.iter) ** (-self.power))
        cos_theta = ops.cos_logits(input, self.weight, 1.0, clamp=True)
        cos_m_theta = self._CHEBYSHEV[self.m](cos_theta)
        k = (self.m * cos_theta.detach().acos() / 3.14159265).floor()
        phi_theta = ((-1.0) ** k) * cos_m_theta - 2 * k
        one_hot = torch.zeros_like(cos_theta).scatter_(1, label.view(-1, 1).long(), 1)
        output = one_hot * (phi_theta - cos_theta) / (1 + self.lamb) + cos_theta
        return output * torch.norm(input, 2, 1).view(-1, 1)

    def __repr__(self):
        return (f"{self.__class__.__name__}(in_features={self.in_features}, out_features={self.out_features}, "
                f"m={self.m})")


class AdaFace(nn.Module):
    """AdaFace head, reference models/metrics.py:170-247 (norm-adaptive angular + additive margin)."""

    def __init__(self, embedding_size, classnum, m=0.4, h=0.333, s=64., t_alpha=1.0):
        super().__init__()
        self.classnum = classnum
        self.kernel = Parameter(torch.empty(embedding_size, classnum))
        self.kernel.data.uniform_(-1, 1).renorm_(2, 1, 1e-5).mul_(1e5)
        self.m, self.eps, self.h, self.s, self.t_alpha = m, 1e-3, h, s, t_alpha
        self.register_buffer('t', torch.zeros(1))
        self.register_buffer('batch_mean', torch.ones(1) * 20)
        self.register_buffer('batch_std', torch.ones(1) * 100)

    def forward(self, embbedings, norms, label):
        # the reference multiplies the (already unit-norm) embeddings by the column-normalised kernel
        kernel_norm = l2_norm(self.kernel, axis=0)
        cosine = torch.mm(embbedings, kernel_norm).clamp(-1 + self.eps, 1 - self.eps)
        safe_norms = torch.clip(norms, min=0.001, max=100).clone().detach()
        with torch.no_grad():
            self.batch_mean = safe_norms.mean() * self.t_alpha + (1 - self.t_alpha) * self.batch_mean
            self.batch_std = safe_norms.std() * self.t_alpha + (1 - self.t_alpha) * self.batch_std
        scaler = torch.clip((safe_norms - self.batch_mean) / (self.batch_std + self.eps) * self.h, -1, 1)
        one_hot = torch.zeros(label.size(0), cosine.size(1), device=cosine.device)
        one_hot.scatter_(1, label.reshape(-1, 1), 1.0)
        theta_m = torch.clip(cosine.acos() + one_hot * (-self.m * scaler), min=self.eps, max=math.pi - self.eps)
        cosine = theta_m.cos() - one_hot * (self.m + self.m * scaler)
        return cosine * self.s
