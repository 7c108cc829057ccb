"""Mirror of the reference's models/fusion_nets.py for the FCFM fusion net (hot symbol: Working).

`Working` keeps the reference's constructor, sub-module and parameter names (a reference `state_dict` loads unchanged)
and its forward signature.  In eval mode -- the verification path, utils/modules.py:141-147 -- the whole forward is one
CUDA kernel (csrc/fcfm.cu).  Training mode is not provided this round: it raises instead of computing something else.
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
        if self.training:
            raise NotImplementedError("Working: only the eval-mode forward (verification path) runs on the B200 kernels; "
                                      "call .eval() -- the training-mode forward/backward is not provided yet")
        state = {k: v for k, v in self.state_dict().items() if not k.endswith("num_batches_tracked")}
        return ops.fcfm_working(img, word, gl_img, sent, state)
