"""Mirror of the reference's image head (models/models.py:96-119 ProjectionHead, :328-338 ImageHeading, :380-405 IMIM;
SelfAttention models/fusion_nets.py:82-118): same class names, constructor arguments, sub-module / parameter names
and shapes (a reference state_dict loads unchanged), forward AND backward in libtgfr_b200.so
(csrc/imim.cu: tgfr_imim_fwd/bwd, tgfr_proj_head_fwd/bwd).  SURVEY.md 8(f) row f3.

The sub-modules below only HOLD the parameters (their own forward is never called): IMIM.forward hands the 16
tensors to one fused operator that returns the reference's result -- logical [B,256,14,14] with memory order
[B,14,14,256], unit L2 norm over the channels, i.e. exactly what words_loss reads as `img_features`.
"""
import torch
import torch.nn as nn

from ._backend import ops


class ProjectionHead(nn.Module):
    """normalize(Linear(x)) (models/models.py:96-119; gelu / fc / dropout exist in the reference but are unused)."""

    def __init__(self, input_dim, projection_dim, dropout=0.4):
        super().__init__()
        self.projection = nn.Linear(input_dim, projection_dim)
        self.gelu = nn.GELU()
        self.fc = nn.Linear(projection_dim, projection_dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        return ops.proj_head(x, self.projection.weight, self.projection.bias)


class SelfAttention(nn.Module):
    """Parameter holder of models/fusion_nets.py:82-118 (1x1 query / key / value projections)."""

    def __init__(self, channel_dim, scale=2):
        super().__init__()
        self.inplanes = channel_dim
        self.query_proj = nn.Conv2d(self.inplanes, self.inplanes // scale, 1)
        self.key_proj = nn.Conv2d(self.inplanes, self.inplanes // scale, 1)
        self.value_proj = nn.Conv2d(self.inplanes, self.inplanes, 1)
        self.sqrt_dim = (channel_dim / scale) ** 0.5


class IMIM(nn.Module):
    def __init__(self, args, channel_dim):
        super().__init__()
        if channel_dim != 256 or int(args.aux_feat_dim_per_granularity) != 256:
            raise NotImplementedError("IMIM: the kernels are built for channel_dim = aux_feat_dim_per_granularity = 256 "
                                      "(the reference's only configuration)")
        self.channel_dim = channel_dim
        self.project_local = ProjectionHead(input_dim=256, projection_dim=args.aux_feat_dim_per_granularity)
        self.bn_img = nn.BatchNorm2d(self.channel_dim)
        self.sa = SelfAttention(channel_dim=self.channel_dim, scale=1)
        self.conv1x1_1 = nn.Conv2d(self.channel_dim, self.channel_dim // 2, kernel_size=(1, 1))
        self.relu = nn.ReLU()
        self.conv1x1_2 = nn.Conv2d(self.channel_dim // 2, self.channel_dim, kernel_size=(1, 1))
        self.ln = nn.LayerNorm([self.channel_dim, 14, 14])

    def _params(self):
        sd = dict(self.named_parameters())
        out = []
        for name in ops.IMIM_PARAM_ORDER:
            t = sd[name]
            out.append(t.reshape(t.shape[0], -1) if t.dim() == 4 else (t.reshape(-1) if t.dim() == 3 else t))
        return out

    def forward(self, img):
        B, C, H, W = img.shape
        if (H, W) != (14, 14):
            raise RuntimeError(f"IMIM: LayerNorm([256,14,14]) needs 14 x 14 feature maps, got {H} x {W}")
        bn = self.bn_img
        use_batch = self.training or bn.running_mean is None
        momentum = bn.momentum
        if self.training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
            if momentum is None:
                momentum = 1.0 / float(bn.num_batches_tracked)
        out = ops.imim(img, self._params(), bn.running_mean if bn.track_running_stats else None,
                       bn.running_var if bn.track_running_stats else None, use_batch, momentum if momentum is not None else 0.1,
                       bn.eps)
        return out.view(B, H, W, C).permute(0, 3, 1, 2)          # models.py:401-404: channels-last memory, NCHW logical


class ImageHeading(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.project_global = ProjectionHead(input_dim=512, projection_dim=args.aux_feat_dim_per_granularity)
        self.imim = IMIM(args, channel_dim=256)

    def forward(self, global_image, local_image):
        local_image = self.imim(local_image)
        global_image = self.project_global(global_image)       # batch_size x 256
        return global_image, local_image
