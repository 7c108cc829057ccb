"""Sharding of the FCAM hot path over one NVSwitch box: one process per GPU, NCCL via
torch.distributed.  The reference has no distributed code at all (only
nn.DataParallel with the losses un-sharded on device 0); this replaces that.

  contrastive losses : every rank owns a row block of faces (its local region / image features never
                       move), all-gathers the caption-side embeddings (words [B,T,D], sentences
                       [B,D]) and computes its [B_local, B_global] block of the score matrix.
                       One all-gather of the per-column (max, sum-exp) pairs closes the column-wise
                       cross entropy; the backward needs one reduce-scatter of d(words) and only when
                       the text side requires grad.
  replica gradients  : the trainable heads (image_head: 692 864 fp32 parameters in the reference) are plain
                       replicas; their gradients are summed with bucketed all-reduces (`allreduce_gradients`),
                       which replaces nn.DataParallel's gather-to-device-0 reduce-add.
  verification       : the pair list is split across ranks (no data-path collective: every pair is independent);
                       the [N] scores and labels are gathered once (`gather_ragged`) for the ROC.
  margin head        : class-sharded partial FC.  Features/labels are all-gathered over the data
                       parallel batch, each rank holds W[c0:c1, :], the softmax statistics are
                       all-reduced (max, then sum-exp and target logit), dX is reduce-scattered.

The collective plumbing below (gather/scatter autograd functions, statistic merging, class ranges)
is device agnostic so that the N>1 logic is unit-tested with gloo on CPU; the arithmetic on each
rank's block always goes through libtgfr_b200.so.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import ptr, stream_ptr

__all__ = [
    "all_gather_rows", "gather_stats", "merge_column_stats", "class_range", "words_loss_sharded", "sent_loss_sharded",
    "ShardedArcMarginProduct", "sharded_focal_ce", "allreduce_gradients", "gather_ragged", "score_pairs_sharded",
]


def _world(group):
    return dist.get_world_size(group), dist.get_rank(group)


# ---------------------------------------------------------------------------------------------
# differentiable all-gather along dim 0 (backward: reduce-scatter of the gradient)
# ---------------------------------------------------------------------------------------------
class _AllGatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        n, _ = _world(group)
        x = x.contiguous()
        out = torch.empty((n * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=group)
        ctx.group = group
        ctx.rows = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        n, _ = _world(ctx.group)
        g = g.contiguous()
        out = torch.empty((ctx.rows,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        if g.is_cuda:
            dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=ctx.group)
        else:  # gloo (CPU tests) has no reduce_scatter: all-reduce then slice
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
            r = dist.get_rank(ctx.group)
            out.copy_(g[r * ctx.rows:(r + 1) * ctx.rows])
        return out, None


def all_gather_rows(x, group=None):
    """[B_local, ...] -> [world * B_local, ...]; equal B_local on every rank."""
    if x.requires_grad:
        return _AllGatherRows.apply(x, group)
    with torch.no_grad():
        return _AllGatherRows.apply(x, group)


def gather_stats(rows, group=None):
    """The exchange of the statistic merge: every rank contributes K rows [M]; returns [n, K, M] (rank-major).
    Device agnostic (unit-tested with gloo); the arithmetic of the merge is `_merge_stats`' CUDA kernel."""
    n, _ = _world(group)
    K, M = len(rows), rows[0].shape[0]
    mine = torch.stack(rows).contiguous()                            # [K, M]
    gathered = torch.empty((n * K, M), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    return gathered.view(n, K, M)


def _merge_stats(rows, group):
    """rows: list of K (2 or 3) [M] fp32 CUDA tensors (max, sum-exp[, summed value]) of this shard -> the K merged
    [M] tensors.  ONE all-gather of [K, M] per rank, then one merge kernel -- not an all-reduce(max) followed by an
    all-reduce(sum): over NVSwitch the exchange is launch-latency bound, so fewer collectives win."""
    _lib.ensure_device(rows[0].device)
    gathered = gather_stats([r.float() for r in rows], group)
    n, K, M = gathered.shape
    out = torch.empty((K, M), dtype=torch.float32, device=gathered.device)
    ops._call("tgfr_merge_softmax_stats", gathered.data_ptr(), n, K, M, out.data_ptr(), stream_ptr())
    return tuple(out[k] for k in range(K))


def merge_column_stats(colmax, colsum, group=None):
    """Combine per-rank column (max, sum exp(s - max)) pairs of a row-sharded score matrix.
    Returns the global (max, sum) per column."""
    return _merge_stats([colmax, colsum], group)


def class_range(num_classes, world, rank):
    """[c0, c1) owned by `rank`: the first (num_classes % world) ranks get one extra class."""
    base, extra = divmod(num_classes, world)
    c0 = rank * base + min(rank, extra)
    return c0, c0 + base + (1 if rank < extra else 0)


# ---------------------------------------------------------------------------------------------
# two-direction cross entropy over a row block [B_local, B_global] of the global score matrix
# ---------------------------------------------------------------------------------------------
class _PairCESharded(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, group):
        _lib.ensure_device(scores.device)
        n, rank = _world(group)
        scores = scores.contiguous()
        Bx, By = scores.shape
        off = rank * Bx
        dev = scores.device
        stats = torch.empty(2 * Bx + 2 * By, dtype=torch.float32, device=dev)
        rowlse, colmax, colsum, diag = stats[:Bx], stats[Bx:Bx + By], stats[Bx + By:Bx + 2 * By], stats[Bx + 2 * By:]
        st = stream_ptr()
        ops._call("tgfr_pair_ce_stats", scores.data_ptr(), Bx, By, off, rowlse.data_ptr(), colmax.data_ptr(),
                  colsum.data_ptr(), diag.data_ptr(), st)
        gmax, gsum = merge_column_stats(colmax, colsum, group)
        gmax, gsum = gmax.contiguous(), gsum.contiguous()
        losses = torch.empty(2, dtype=torch.float32, device=dev)
        collse = torch.empty(By, dtype=torch.float32, device=dev)
        ops._call("tgfr_pair_ce_finish", rowlse.data_ptr(), gmax.data_ptr(), gsum.data_ptr(), diag.data_ptr(),
                  Bx, By, off, 1.0 / By, losses.data_ptr(), collse.data_ptr(), stream_ptr())
        dist.all_reduce(losses, op=dist.ReduceOp.SUM, group=group)
        ctx.save_for_backward(scores, stats, collse)
        ctx.off = off
        return losses[0].clone(), losses[1].clone()

    @staticmethod
    def backward(ctx, g0, g1):
        scores, stats, collse = ctx.saved_tensors
        Bx, By = scores.shape
        g = torch.stack([g0.float().reshape(()), g1.float().reshape(())]).contiguous()
        gs = torch.empty_like(scores)
        ops._call("tgfr_pair_ce_bwd", scores.data_ptr(), stats.data_ptr(), collse.data_ptr(), g[0:].data_ptr(),
                  g[1:].data_ptr(), Bx, By, ctx.off, 1.0 / By, gs.data_ptr(), stream_ptr())
        return gs, None


def words_loss_sharded(feats, words, cap_lens, gamma1, gamma2, gamma3, group=None, precision=None,
                       want_attn=False):
    """Row-sharded words_loss.  feats [B_local,R,D] (this rank's faces), words [B_local,T,D]
    (this rank's captions), cap_lens int [B_local] or None.  Returns the GLOBAL (loss0, loss1) over
    the world*B_local batch -- identical on every rank -- and this rank's diagonal attention maps.
    Backward gives the exact gradient of the global loss w.r.t. the local tensors."""
    n, rank = _world(group)
    words_all = all_gather_rows(words, group)
    lens_all = None
    if cap_lens is not None:
        lens_all = all_gather_rows(cap_lens.to(device=feats.device, dtype=torch.int32), group)
    sim, attn = ops.wordregion_sim(feats, words_all, lens_all, gamma1, gamma2, gamma3, 1e-8, precision,
                                   want_attn, rank * feats.shape[0])
    loss0, loss1 = _PairCESharded.apply(sim, group)
    return loss0, loss1, attn


def sent_loss_sharded(img, txt, class_ids, gamma3, group=None, eps=1e-8):
    """Row-sharded sent_loss / global_loss (class_ids=None).  img, txt: [B_local, D]."""
    n, rank = _world(group)
    txt_all = all_gather_rows(txt, group)
    ids_x = ids_y = None
    if class_ids is not None:
        ids_x = class_ids.to(device=img.device, dtype=torch.int64).contiguous().view(-1)
        ids_y = all_gather_rows(ids_x, group)
    scores = _CosineScoresOff.apply(img.float(), txt_all.float(), float(gamma3), float(eps), ids_x, ids_y,
                                    rank * img.shape[0])
    return _PairCESharded.apply(scores, group)


class _CosineScoresOff(torch.autograd.Function):
    """ops._CosineScores with a diagonal offset (row b of the block pairs with column b + off)."""

    @staticmethod
    def forward(ctx, x, y, scale, eps, ids_x, ids_y, off):
        _lib.ensure_device(x.device)
        x, y = x.contiguous(), y.contiguous()
        Bx, D = x.shape
        By = y.shape[0]
        scores = torch.empty((Bx, By), dtype=torch.float32, device=x.device)
        xn = torch.empty(Bx, dtype=torch.float32, device=x.device)
        yn = torch.empty(By, dtype=torch.float32, device=x.device)
        ops._call("tgfr_cosine_scores_fwd", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), Bx, By, D,
                  scale, 1, eps, ptr(ids_x), ptr(ids_y), off, scores.data_ptr(), xn.data_ptr(), yn.data_ptr(),
                  stream_ptr())
        ctx.save_for_backward(x, y, xn, yn)
        ctx.cfg = (scale, eps)
        return scores

    @staticmethod
    def backward(ctx, gs):
        x, y, xn, yn = ctx.saved_tensors
        scale, eps = ctx.cfg
        Bx, D = x.shape
        By = y.shape[0]
        gs = gs.float().contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        lib = _lib.load()
        wsb = lib.tgfr_cosine_workspace_bytes(Bx, By, D)
        ws = ops._workspace(wsb, x.device)
        ops._call("tgfr_cosine_scores_bwd", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), Bx, By, D,
                  scale, 1, eps, xn.data_ptr(), yn.data_ptr(), gs.data_ptr(), ptr(dx), ptr(dy), ptr(ws), wsb,
                  stream_ptr())
        return dx, dy, None, None, None, None, None


# ---------------------------------------------------------------------------------------------
# class-sharded margin head with a fused, never-gathered softmax cross entropy
# ---------------------------------------------------------------------------------------------
class _FocalCESharded(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, gamma, class_off, group):
        _lib.ensure_device(logits.device)
        logits = logits.contiguous()
        B, C = logits.shape
        dev = logits.device
        stats = torch.empty(4 * B, dtype=torch.float32, device=dev)
        rowmax, rowsum, tgt, lse = stats[:B], stats[B:2 * B], stats[2 * B:3 * B], stats[3 * B:]
        ops._call("tgfr_ce_rows_stats", logits.data_ptr(), logits.stride(0), target.data_ptr(), B, C, class_off,
                  rowmax.data_ptr(), rowsum.data_ptr(), tgt.data_ptr(), stream_ptr())
        gmax, gsum, gtgt = _merge_row_stats(rowmax, rowsum, tgt, group)
        out = torch.empty(3, dtype=torch.float32, device=dev)
        ops._call("tgfr_focal_finish", gmax.data_ptr(), gsum.data_ptr(), gtgt.data_ptr(), B, gamma, out.data_ptr(),
                  lse.data_ptr(), stream_ptr())
        ctx.save_for_backward(logits, target, stats, out)
        ctx.class_off = class_off
        return out[1].clone()

    @staticmethod
    def backward(ctx, gout):
        logits, target, stats, out = ctx.saved_tensors
        B, C = logits.shape
        lse = stats[3 * B:]
        gout = gout.float().reshape(1).contiguous()
        gl = torch.empty((B, C), dtype=torch.float32, device=logits.device)
        ops._call("tgfr_ce_rows_bwd", logits.data_ptr(), logits.stride(0), target.data_ptr(), lse.data_ptr(),
                  out[2:].data_ptr(), gout.data_ptr(), B, C, ctx.class_off, gl.data_ptr(), gl.stride(0),
                  stream_ptr())
        return gl, None, None, None, None


def _merge_row_stats(rowmax, rowsum, tgt, group):
    """Per-shard online-softmax statistics -> global ones (max, rescaled sum, target logit; non-owners hold 0):
    one all-gather of [3, B] per rank + one merge kernel."""
    gmax, gsum, gtgt = _merge_stats([rowmax, rowsum, tgt], group)
    return gmax.contiguous(), gsum.contiguous(), gtgt.contiguous()


def sharded_focal_ce(logits_shard, target_all, gamma, class_off, group=None):
    """Focal loss of the batch-mean CE over class-sharded logits [B_global, C_local]."""
    target_all = target_all.view(-1).to(device=logits_shard.device, dtype=torch.int64).contiguous()
    return _FocalCESharded.apply(logits_shard.float(), target_all, float(gamma), int(class_off), group)


class ShardedArcMarginProduct(torch.nn.Module):
    """ArcMarginProduct with the class dimension split over the process group (partial FC).

    `weight` holds rows [c0, c1) of the reference's [out_features, in_features] parameter;
    load_full_weight / full_weight convert from / to the reference's state_dict tensor.
    forward(input_local, label_local) -> (logits_shard [B_global, C_local], labels_all);
    loss(input_local, label_local, gamma) runs the fused focal cross entropy without ever
    materialising the [B, C] logits on one device.
    """

    def __init__(self, in_features, out_features, s=30.0, m=0.50, easy_margin=False, group=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.s, self.m, self.easy_margin, self.group = s, m, easy_margin, group
        n, rank = _world(group)
        self.c0, self.c1 = class_range(out_features, n, rank)
        full = torch.empty(out_features, in_features)
        gen = torch.Generator().manual_seed(100)          # identical init on every rank, then slice
        bound = (6.0 / (in_features + out_features)) ** 0.5
        full.uniform_(-bound, bound, generator=gen)
        self.weight = torch.nn.Parameter(full[self.c0:self.c1].clone())

    @torch.no_grad()
    def load_full_weight(self, full):
        self.weight.copy_(full[self.c0:self.c1])

    @torch.no_grad()
    def full_weight(self):
        n, _ = _world(self.group)
        parts = [None] * n
        dist.all_gather_object(parts, self.weight.detach().cpu(), group=self.group)
        return torch.cat(parts, 0)

    def forward(self, input, label):
        x_all = all_gather_rows(input, self.group)
        lab_all = all_gather_rows(label.view(-1).to(device=input.device, dtype=torch.int64), self.group)
        logits = ops.arc_logits(x_all, self.weight, lab_all, self.s, self.m, self.easy_margin, self.c0)
        return logits, lab_all

    def loss(self, input, label, gamma=2.0):
        """Focal cross entropy of the sharded head.  On the tensor-core path the shard's logits are never written
        (fused GEMM epilogues, ops.arc_fused_focal); TGFR_HEAD_PRECISION=fp32 materialises the [B, C_local] shard."""
        if ops.head_precision(self.in_features) == _lib.PREC_TC:
            x_all = all_gather_rows(input, self.group)
            lab_all = all_gather_rows(label.view(-1).to(device=input.device, dtype=torch.int64), self.group)
            return ops.arc_fused_focal(x_all, self.weight, lab_all, self.s, self.m, self.easy_margin, gamma, self.c0,
                                       merge=lambda mx, sm, tg: _merge_row_stats(mx, sm, tg, self.group))
        logits, lab_all = self.forward(input, label)
        return sharded_focal_ce(logits, lab_all, gamma, self.c0, self.group)


# ---------------------------------------------------------------------------------------------
# gradients of the replicated trainable modules (the reference wraps them in nn.DataParallel)
# ---------------------------------------------------------------------------------------------
def _check_membership(plist, n, group):
    fp = torch.tensor([float(len(plist)), float(sum(p.numel() for p in plist))], dtype=torch.float64,
                      device=plist[0].device)
    fp_sum = fp.clone()
    dist.all_reduce(fp_sum, op=dist.ReduceOp.SUM, group=group)
    fp_max = fp.clone()
    dist.all_reduce(fp_max, op=dist.ReduceOp.MAX, group=group)
    if not (torch.equal(fp_sum, fp * n) and torch.equal(fp_max, fp)):
        raise RuntimeError("allreduce_gradients: the ranks disagree on the set of trainable parameters "
                           f"(local: {int(fp[0])} tensors, {int(fp[1])} elements)")


def allreduce_gradients(params, group=None, bucket_bytes=16 << 20, average=False, check=True):
    """Sum (or average) the .grad of replicated parameters over the group with a few large all-reduces.

    Every parameter with requires_grad takes part, in the order given (identical on every rank for replicas); a
    parameter whose .grad is None on this rank (an unused branch, a conditional head) contributes zeros and
    receives the sum, so the flat buffers have the same size and layout everywhere whatever each rank's autograd
    graph touched.  The first bucket carries a fingerprint (parameter count and total element count): ranks that
    disagree raise instead of adding up mismatched gradients.
    Gradients are packed into flat buckets of at most `bucket_bytes` (per dtype), every bucket is reduced with
    one asynchronous all-reduce -- over NVSwitch the cost is launch latency, not link count, so buckets are
    sized to keep the launches few -- and unpacked in place once all of them have completed.
    The sharded losses above already return gradients of the GLOBAL (batch-mean) loss, so the default is a
    plain sum; `average=True` divides by the world size (per-rank mean losses).  Returns the number of buckets."""
    plist = [p for p in params if p.requires_grad]
    if not plist:
        return 0
    n, _ = _world(group)
    # same membership on every rank?  (count, elements) summed over the group must equal n x the local values.
    # The comparison reads the result on the host, so it is skipped while a CUDA graph is being captured
    # (run one eager step first) or with check=False.
    capturing = plist[0].is_cuda and torch.cuda.is_current_stream_capturing()
    if check and not capturing:
        _check_membership(plist, n, group)
    order = sorted(range(len(plist)), key=lambda i: (str(plist[i].dtype), i))
    buckets, cur, cur_bytes = [], [], 0
    for i in order:
        p = plist[i]
        nbytes = p.numel() * p.element_size()
        if cur and (cur[0].dtype != p.dtype or cur_bytes + nbytes > bucket_bytes):
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(p)
        cur_bytes += nbytes
    buckets.append(cur)
    flats, works = [], []
    for b in buckets:
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in b])
        flats.append(flat)
        works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for b, flat, w in zip(buckets, flats, works):
        w.wait()
        if average:
            flat.div_(n)
        off = 0
        for p in b:
            piece = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = piece.clone()
            else:
                p.grad.copy_(piece)
            off += p.numel()
    return len(buckets)


# ---------------------------------------------------------------------------------------------
# verification scoring: pair-sharded cosine, one gather of the scores for the ROC
# ---------------------------------------------------------------------------------------------
def gather_ragged(x, group=None):
    """Concatenate every rank's 1-D tensor (lengths may differ) in rank order, on every rank."""
    world, _ = _world(group)
    x = x.contiguous().view(-1)
    if world == 1:
        return x
    n = torch.tensor([x.numel()], dtype=torch.int64, device=x.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(v.item()) for v in sizes]
    cap = max(max(sizes), 1)
    padded = torch.zeros(cap, dtype=x.dtype, device=x.device)
    padded[: x.numel()] = x
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:k] for p, k in zip(parts, sizes)])


def score_pairs_sharded(batches, args=None, group=None):
    """utils/modules.py:150-166 with the pair list split across ranks: `batches` yields THIS rank's
    (out1 [n, D], out2 [n, D], pair_label [n]); every rank returns the summary of the whole list."""
    from .utils import modules
    preds, labels = [], []
    for out1, out2, pair_label in batches:
        preds.append(modules.pair_scores(out1, out2))
        labels.append(pair_label.to(device=out1.device, dtype=torch.int64).view(-1))
    dev = preds[0].device if preds else torch.device("cuda", torch.cuda.current_device())
    preds = torch.cat(preds) if preds else torch.zeros(0, device=dev)
    labels = torch.cat(labels) if labels else torch.zeros(0, dtype=torch.int64, device=dev)
    return modules.calculate_scores(gather_ragged(preds, group), gather_ragged(labels, group), args)
