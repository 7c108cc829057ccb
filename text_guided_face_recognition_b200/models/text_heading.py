"""TextHeading / Bert_Word_Mapping of the reference (models/models.py:170-232) on libtgfr_b200.so.

Same class names, constructor arguments, parameter names and shapes (`bwm.convs1.{0,1,2}.{weight,bias}`, a
reference `state_dict` loads unchanged) and the same return value `(words_emb [B, F, T], sent_emb [B, F])`.
The reference's `models/models.py` holds many unrelated classes (BERT / LSTM encoders, ImageHeading ...), so this
module does not shadow it: a maintainer swaps the one import,

    from models.text_heading import TextHeading        # instead of: from models.models import TextHeading

The three n-gram convolutions, the shifted max over them (a Python loop of B x T stack / amax calls in the
reference, models.py:197-213, with a CUDA-only `torch.cuda.FloatTensor` copy), the max-pool / mean sentence
feature and both L2 normalisations run in three strided fp32 products and one fused kernel.
`words_emb` comes back as a transposed view of a [B, T, F] tensor -- exactly the memory layout the reference
produces and the one the word-region kernels read without a copy.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._backend import ops

__all__ = ["Bert_Word_Mapping", "TextHeading"]


class Bert_Word_Mapping(nn.Module):
    """Parameter container of the three n-gram convolutions (models.py:170-186)."""

    def __init__(self, feat_dim):
        super().__init__()
        Ks = [2, 3, 4]
        self.convs1 = nn.ModuleList([nn.Conv2d(1, feat_dim, (K, 768)) for K in Ks])   # 768: hard-coded in the reference
        self.dropout = nn.Dropout(0.1)                                                   # defined, never applied

    def forward(self, words_emb):
        """[relu(conv_K(words_emb)) [B, F, L-K+1] for K in (2, 3, 4)] -- kept for callers that want the raw maps."""
        x = words_emb.unsqueeze(1)
        return [torch.relu(conv(x)).squeeze(3) for conv in self.convs1]


class TextHeading(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.feat_dim = args.aux_feat_dim_per_granularity
        self.bwm = Bert_Word_Mapping(self.feat_dim)
        self.args = args

    def forward(self, words_emb, sent_emb):
        """words_emb: BERT token features [B, bert_words_num - 1, 768] ([CLS] removed); sent_emb is ignored, as in
        the reference (its projection is commented out, models.py:222-224)."""
        convs = self.bwm.convs1
        words, sent = ops.text_heading(words_emb, [c.weight for c in convs], [c.bias for c in convs],
                                       self.args.bert_words_num)
        return words.transpose(1, 2), sent
