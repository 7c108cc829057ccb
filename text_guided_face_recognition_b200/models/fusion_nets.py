"""Mirror of the reference's models/fusion_nets.py for the FCFM fusion net (hot symbol: Working).

`Working` keeps the reference's constructor, sub-module and parameter names (a reference `state_dict` loads unchanged)
and its forward signature.  In eval mode without gradients -- the verification path, utils/modules.py:141-147 -- the
whole forward is one CUDA kernel (csrc/fcfm.cu).  In training mode (BatchNorm batch statistics, running-stat update), or
whenever autograd is recording, the forward and the backward run the batch-wide kernels of csrc/fcfm_train.cu: gradients
reach all 26 parameter tensors and the four inputs, as the fusion training step src/fusion_bert.py:205-233 needs.
"""
import torch
import torch.nn as nn

from ._backend import ops


class SelfAttention(nn.Module):
    """Parameter container of reference fusion_nets.py:82-118 (its arithmetic runs inside the fused kernel)."""

    def __init__(self, channel_dim, scale=2):
        super().__init__()
        self.inplanes = channel_dim
        self.query_proj = nn.Conv2d(channel_dim, channel_dim // scale, 1)
        self.key_proj = nn.Conv2d(channel_dim, channel_dim // scale, 1)
        self.value_proj = nn.Conv2d(channel_dim, channel_dim, 1)


class Working(nn.Module):
    """FCFM fusion, reference models/fusion_nets.py:217-258: (local image features, word features, global image
    feature, sentence feature) -> 640-d fused embedding."""

    def __init__(self, channel_dim):
        super().__init__()
        channel_dim = 36                       # the reference overrides its argument (fusion_nets.py:220)
        self.bn_img = nn.BatchNorm2d(channel_dim)
        self.bn_word = nn.BatchNorm2d(channel_dim)
        self.projection = nn.Linear(256, channel_dim)
        self.sa = SelfAttention(channel_dim, scale=1)
        self.conv = nn.Conv2d(256, channel_dim, kernel_size=(3, 3), padding=0)
        self.ln = nn.LayerNorm([channel_dim, 6, 6])
        self.ln_gl_image = nn.LayerNorm([256])
        self.ln_sent = nn.LayerNorm([256])
        self.linear = nn.Linear(324, 128)

    def forward(self, img, word, gl_img, sent):
        wants_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or any(
            t.requires_grad for t in (img, word, gl_img, sent)))
        if not self.training and not wants_grad:
            state = {k: v for k, v in self.state_dict().items() if not k.endswith("num_batches_tracked")}
            return ops.fcfm_working(img, word, gl_img, sent, state)
        named = dict(self.named_parameters())
        params = []
        for name in ops.FCFM_TRAIN_PARAM_ORDER:
            t = named[name]
            params.append(t.reshape(t.shape[0], -1) if t.dim() == 4 else (t.reshape(-1) if t.dim() == 3 else t))
        momentum = self.bn_img.momentum
        if self.training:
            for bn in (self.bn_img, self.bn_word):
                if bn.track_running_stats and bn.num_batches_tracked is not None:
                    bn.num_batches_tracked.add_(1)
        stats = (self.bn_img.running_mean, self.bn_img.running_var, self.bn_word.running_mean, self.bn_word.running_var)
        return ops.fcfm_working_train(img, word, gl_img, sent, params, stats, self.training,
                                      momentum if momentum is not None else 0.1, self.bn_img.eps)
