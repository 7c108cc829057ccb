"""Mirror of the reference's models/attention.py (hot-path symbol: func_attention)."""
import torch

from ._backend import ops


def func_attention(query, context, gamma1):
    """AttnGAN word->region attention, reference models/attention.py:10-43.

    query: [B, D, T]; context: [B, D, ih, iw].  Returns (weightedContext [B, D, T],
    attn [B, T, ih, iw]).  Both outputs are differentiable w.r.t. both inputs; the scores,
    both softmaxes and the context contraction run in one CUDA kernel per direction.
    """
    if query.dim() != 3 or context.dim() != 4:
        raise RuntimeError(f"func_attention expects query [B,D,T] and context [B,D,ih,iw], got "
                           f"{tuple(query.shape)} and {tuple(context.shape)}")
    B, D, T = query.shape
    if context.size(0) != B or context.size(1) != D:
        raise RuntimeError(f"func_attention: shape mismatch {tuple(query.shape)} vs {tuple(context.shape)}")
    ih, iw = context.size(2), context.size(3)
    feats = context.permute(0, 2, 3, 1).reshape(B, ih * iw, D)     # a view for both a0 layouts
    wc, attn = ops.func_attention_canonical(feats, query.transpose(1, 2), gamma1)
    return wc.transpose(1, 2), attn.view(B, T, ih, iw)


# SpatialAttention / ChannelAttention (models/attention.py:46-131) are AttnGAN generator
# leftovers with no call sites in the reference; they are out of scope.
class SpatialAttention(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("SpatialAttention is dead code in the reference and out of scope")


class ChannelAttention(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("ChannelAttention is dead code in the reference and out of scope")
