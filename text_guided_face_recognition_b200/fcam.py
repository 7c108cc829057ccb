"""One FCAM contrastive step with the sentence-level chain in the shadow of the word-region kernels.

The reference's training loop calls `sent_loss` and `words_loss` one after the other on one stream
(src/train_encoders_bert.py:274-283).  The sentence loss is a chain of ~10 tiny kernels (row norms, a B x B
product, two cross entropies) that needs no shared memory, while the word-region kernels run one fat CTA per SM;
issued on a second stream the small kernels fit beside the fat CTAs and cost nothing.  Autograd replays every
backward node on the stream of its forward, so the same overlap holds in the backward pass, and the fork / join
is plain event work, so the whole step still captures into one CUDA graph (graphs.GraphedStep).

This is an optional convenience on top of the drop-in `models.losses` functions; the results are identical.
"""
from __future__ import annotations

import torch

__all__ = ["run_overlapped", "fcam_losses", "FcamTrainStep"]

_side_streams: dict = {}


def _side_stream(device):
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=device)
    return st


def _record(obj, stream):
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, (tuple, list)):
        for o in obj:
            _record(o, stream)


def run_overlapped(side_fn, main_fn, device=None):
    """Run `side_fn()` on a side stream and `main_fn()` on the current stream; both see everything enqueued on the
    current stream so far, and the current stream waits for the side work before this returns.
    Returns (side_result, main_result)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    main = torch.cuda.current_stream(device)
    side = _side_stream(device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        side_out = side_fn()
    main_out = main_fn()
    main.wait_stream(side)
    _record(side_out, main)
    return side_out, main_out


def fcam_losses(words_features, words_emb, img_code, txt_code, labels, cap_lens, class_ids, batch_size, args):
    """(w_loss0, w_loss1, att_maps, s_loss0, s_loss1): `words_loss` + `sent_loss` of models/losses.py:19-135 with
    the sentence chain overlapped (same arguments as the two reference calls)."""
    from .models import losses
    (s0, s1), (w0, w1, att) = run_overlapped(
        lambda: losses.sent_loss(img_code, txt_code, labels, class_ids, batch_size, args),
        lambda: losses.words_loss(words_features, words_emb, labels, cap_lens, class_ids, batch_size, args),
        device=words_features.device)
    return w0, w1, att, s0, s1


class FcamTrainStep:
    """The loss block of the reference's `Train.train` (src/train_encoders_bert.py:254-326) assembled from this
    package's operators, one process per GPU (BASELINE configs[3]; SURVEY.md 8(d) "Config 4"):

        tokens --TextHeading (no_grad: the text side is detached, utils/dataset_utils.py:36-46)--> words_emb, sent_emb
        (global, local) backbone features --ImageHeading (IMIM + ProjectionHead, trainable)--> img_features, words_features
        words_loss + sent_loss (class-id mask) + lambda_clip * global_loss      row-sharded over the ranks
        lambda_id * (FocalLoss(text_cls(sent_emb)) + FocalLoss(image_cls(img_features)))   class-sharded ArcFace heads
        total.backward();  gradients of the replicated image_head summed over the ranks (allreduce_gradients)

    The frozen encoders (IResNet-50, BERT) are NOT part of it: their outputs are the step's inputs (SURVEY.md section 2:
    out of scope).  With world size 1 the same operators run un-sharded.  `__call__` runs one step and returns the
    total loss; every kernel and collective is enqueued on the current stream, so a step captures into one CUDA graph."""

    def __init__(self, args, num_classes, device, group=None, lambda_id=100.0, lambda_clip=2.0, gamma=2.0, seed=100):
        import torch.distributed as dist
        from . import distributed as tdist
        from .models import metrics
        from .models.image_heading import ImageHeading
        from .models.text_heading import TextHeading
        self.args, self.group = args, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.lambda_id, self.lambda_clip, self.gamma = lambda_id, lambda_clip, gamma
        gen = torch.Generator().manual_seed(seed)              # identical replicas on every rank
        cpu_state = torch.random.get_rng_state()
        torch.random.set_rng_state(gen.get_state())
        self.image_head = ImageHeading(args).to(device).train()
        self.text_head = TextHeading(args).to(device).eval()
        torch.random.set_rng_state(cpu_state)
        D = int(args.aux_feat_dim_per_granularity)
        if self.world > 1:
            self.image_cls = tdist.ShardedArcMarginProduct(D, num_classes, s=30.0, m=0.5, group=group).to(device)
            self.text_cls = tdist.ShardedArcMarginProduct(D, num_classes, s=35.0, m=0.5, group=group).to(device)
        else:
            self.image_cls = metrics.ArcMarginProduct(D, num_classes, s=30.0, m=0.5).to(device)
            self.text_cls = metrics.ArcMarginProduct(D, num_classes, s=35.0, m=0.5).to(device)
        self.gammas = (float(args.TRAIN.SMOOTH.GAMMA1), float(args.TRAIN.SMOOTH.GAMMA2), float(args.TRAIN.SMOOTH.GAMMA3))

    def parameters(self):
        import itertools
        return itertools.chain(self.image_head.parameters(), self.image_cls.parameters(), self.text_cls.parameters())

    def zero_grad(self):
        for p in self.parameters():
            p.grad = None

    def __call__(self, global_feat, local_feat, tokens, class_ids):
        """global_feat [B,512], local_feat [B,256,14,14], tokens [B, bert_words_num - 1, 768], class_ids int64 [B]
        (this rank's samples).  Returns the total loss (identical on every rank)."""
        from . import distributed as tdist
        from . import ops
        g1, g2, g3 = self.gammas
        B = global_feat.shape[0]
        with torch.no_grad():
            words_emb, sent_emb = self.text_head(tokens, None)              # [B,256,T] (memory [B,T,256]), [B,256]
        img_features, words_features = self.image_head(global_feat, local_feat)
        feats = words_features.permute(0, 2, 3, 1).reshape(B, -1, words_features.shape[1])       # [B,196,256] view
        words = words_emb.transpose(1, 2)                                                         # [B,T,256] view
        if self.world > 1:
            w0, w1, _ = tdist.words_loss_sharded(feats, words, None, g1, g2, g3, group=self.group)
            s0, s1 = tdist.sent_loss_sharded(img_features, sent_emb, class_ids, g3, group=self.group)
            c0, c1 = tdist.sent_loss_sharded(img_features, sent_emb, None, 10.0, group=self.group)   # global_loss
            tid = self.text_cls.loss(sent_emb, class_ids, gamma=self.gamma)
            iid = self.image_cls.loss(img_features, class_ids, gamma=self.gamma)
        else:
            sim, _ = ops.wordregion_sim(feats, words, None, g1, g2, g3, 1e-8, None, False, 0)
            w0, w1 = ops.pair_ce(sim)
            ids = class_ids.to(torch.int64)
            s0, s1 = ops.pair_ce(ops.cosine_scores(img_features, sent_emb, g3, True, 1e-8, ids, ids))
            c0, c1 = ops.pair_ce(ops.cosine_scores(img_features, sent_emb, 10.0, True, 1e-8))
            tid = self.text_cls.fused_loss(sent_emb, class_ids, gamma=self.gamma)
            iid = self.image_cls.fused_loss(img_features, class_ids, gamma=self.gamma)
        total = w0 + w1 + s0 + s1 + self.lambda_id * (tid + iid) + self.lambda_clip * (c0 + c1)
        total.backward()
        if self.world > 1:
            tdist.allreduce_gradients(self.image_head.parameters(), group=self.group)
        return total
