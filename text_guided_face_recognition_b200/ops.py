"""torch.autograd bindings of the C-ABI kernels (one Function per differentiable operator).

PyTorch is used for device memory, streams and autograd bookkeeping only; every FLOP of the
hot path runs in libtgfr_b200.so.  Nothing here falls back to eager PyTorch math.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import PREC_FP32, PREC_TC, check, ptr, stream_ptr

__all__ = [
    "default_precision", "tc_supported", "wordregion_sim", "pair_ce", "cosine_scores", "arc_logits", "focal_ce",
    "mag_logits", "mag_ce", "cosine_rows", "func_attention_canonical", "launch_counter", "arc_fused_focal", "text_heading",
    "pair_cosine", "roc_counts", "row_argmax", "fcfm_working", "fcfm_working_train", "imim", "proj_head",
]


class _LaunchCounter:
    """Counts C-ABI calls (each issues >= 1 kernel); bench.py reports it as gpu_launches."""
    n = 0


launch_counter = _LaunchCounter


_noticed = set()


def _notice_once(var: str, what: str) -> None:
    """The drop-in mirrors default to fp16-operand tensor-core arithmetic, which is NOT what the reference computes bit
    for bit (losses within 1e-4, gradients within 1e-3, logits within 1e-2; a `cos > th` margin switch can flip for a sample
    sitting on the threshold).  A process that did not choose says so once, on stderr via `warnings` (TGFR_QUIET=1 silences)."""
    if var in os.environ or var in _noticed or os.environ.get("TGFR_QUIET") == "1":
        return
    _noticed.add(var)
    import warnings
    warnings.warn(f"tgfr_b200: {what} run on the tensor cores with fp16 operands / fp32 accumulation (TF32-class: losses within "
                  f"1e-4, gradients within 1e-3 of the reference's fp32). Set {var}=fp32 for the exact-fp32 kernels, or {var}=tc "
                  f"to acknowledge this default.", stacklevel=3)


def default_precision() -> int:
    """TGFR_WORDREGION_PRECISION = tc (default: tcgen05, fp16 operands / fp32 accumulate) | fp32 (SIMT)."""
    _notice_once("TGFR_WORDREGION_PRECISION", "the word-region loss kernels")
    v = os.environ.get("TGFR_WORDREGION_PRECISION", "tc").lower()
    return PREC_FP32 if v in ("fp32", "simt", "0") else PREC_TC


def head_precision(Din: int) -> int:
    """TGFR_HEAD_PRECISION = tc (default: tcgen05 cos-theta / gradient GEMMs, fp16 operands, fp32 accumulate) | fp32."""
    if Din >= 8:
        _notice_once("TGFR_HEAD_PRECISION", "the margin-head products (ArcMarginProduct / MagLinear)")
    v = os.environ.get("TGFR_HEAD_PRECISION", "tc").lower()
    if v in ("fp32", "simt", "0") or Din < 8:
        return PREC_FP32
    return PREC_TC


def _head_ws(B, C, Din, prec, device):
    lib = _lib.load()
    wsb = lib.tgfr_margin_workspace_bytes(B, C, Din, prec)
    return _workspace(wsb, device), wsb


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.float32 else t.float()


# kernels launched by one call of each entry point (memsets not counted)
_KERNELS_PER_CALL = {
    "tgfr_wordregion_fwd": 3, "tgfr_wordregion_bwd": 1, "tgfr_attention_fwd": 1, "tgfr_attention_bwd": 1,
    "tgfr_cosine_scores_fwd": 3, "tgfr_cosine_scores_bwd": 4, "tgfr_pair_ce_stats": 1, "tgfr_pair_ce_finish": 1,
    "tgfr_pair_ce_bwd": 1, "tgfr_cos_logits_fwd": 3, "tgfr_arc_margin_apply": 1, "tgfr_arc_margin_bwd": 5,
    "tgfr_mag_margin_fwd": 1, "tgfr_mag_margin_bwd": 1, "tgfr_cos_logits_bwd": 4, "tgfr_ce_rows_stats": 1,
    "tgfr_focal_finish": 1, "tgfr_ce_rows_bwd": 1, "tgfr_mag_ce_stats": 1, "tgfr_mag_ce_bwd": 1, "tgfr_arc_fused_fwd": 6, "tgfr_arc_fused_bwd": 5, "tgfr_texthead_fwd": 12, "tgfr_texthead_bwd": 5,
    "tgfr_pair_cosine": 1, "tgfr_roc_curve": 20, "tgfr_row_argmax": 1, "tgfr_imim_fwd": 34, "tgfr_imim_bwd": 39, "tgfr_matmul_split": 5, "tgfr_fcfm_working_fwd_tc": 8,
    "tgfr_proj_head_fwd": 2, "tgfr_proj_head_bwd": 5, "tgfr_fcfm_train_fwd": 21, "tgfr_fcfm_train_bwd": 45,
}


def _call(name, *args):
    lib = _lib.load()
    _LaunchCounter.n += _KERNELS_PER_CALL.get(name, 1)
    check(getattr(lib, name)(*args), name)


def _workspace(nbytes: int, device) -> torch.Tensor | None:
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device) if nbytes else None


# ---------------------------------------------------------------------------------------------
# word-region similarity matrix
# ---------------------------------------------------------------------------------------------
class _WordRegionSim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, words, cap_lens, g1, g2, g3, eps, precision, want_attn, diag_off):
        # feats [Bc,R,D], words [Bq,T,D] -- arbitrary strides; cap_lens int32 [Bq] or None
        _lib.ensure_device(feats.device)
        Bc, R, D = feats.shape
        Bq, T, _ = words.shape
        sim = torch.empty((Bc, Bq), dtype=torch.float32, device=feats.device)
        attn = torch.empty((Bc, T, R), dtype=torch.float32, device=feats.device) if want_attn else None
        lib = _lib.load()
        wsb = lib.tgfr_wordregion_workspace_bytes(Bc, Bq, T, R, D, precision)
        ws = _workspace(wsb, feats.device)
        # forward -> backward image of the word softmax / attention (fp16 records): the backward then skips the score
        # GEMM and every exponential.  TGFR_WORDREGION_SAVE=wu keeps only the Wu tiles (2.5x fewer bytes; the backward
        # recomputes the scores), TGFR_WORDREGION_SAVE=0 (or a buffer above the cap) selects full recomputation.
        # The library sizes the buffer for the selected layout and recognises the layout by that size.
        saved, svb = None, 0
        if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) and _save_enabled():
            svb = lib.tgfr_wordregion_saved_bytes(Bc, Bq, T, R, D, precision)
            if 0 < svb <= _SAVE_CAP_BYTES:
                saved = torch.empty(svb, dtype=torch.uint8, device=feats.device)
            else:
                svb = 0
        _call("tgfr_wordregion_fwd", feats.data_ptr(), *feats.stride(), words.data_ptr(), *words.stride(),
              ptr(cap_lens), Bc, Bq, T, R, D, g1, g2, g3, eps, sim.data_ptr(), ptr(attn), diag_off,
              precision, ptr(ws), wsb, ptr(saved), svb, stream_ptr())
        ctx.save_for_backward(feats, words, cap_lens, saved)
        ctx.cfg = (g1, g2, g3, eps, precision)
        if attn is not None:
            ctx.mark_non_differentiable(attn)
        return sim, attn

    @staticmethod
    def backward(ctx, gsim, _gattn):
        feats, words, cap_lens, saved = ctx.saved_tensors
        g1, g2, g3, eps, precision = ctx.cfg
        Bc, R, D = feats.shape
        Bq, T, _ = words.shape
        need_c, need_q = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gsim = _f32(gsim).contiguous()
        dctx = torch.empty((Bc, R, D), dtype=torch.float32, device=feats.device) if need_c else None
        dwords = torch.empty((Bq, T, D), dtype=torch.float32, device=feats.device) if need_q else None
        lib = _lib.load()
        wsb = lib.tgfr_wordregion_workspace_bytes(Bc, Bq, T, R, D, precision)
        ws = _workspace(wsb, feats.device)
        _call("tgfr_wordregion_bwd", feats.data_ptr(), *feats.stride(), words.data_ptr(), *words.stride(),
              ptr(cap_lens), Bc, Bq, T, R, D, g1, g2, g3, eps, gsim.data_ptr(), ptr(dctx), ptr(dwords),
              precision, ptr(ws), wsb, ptr(saved), 0 if saved is None else saved.numel(), stream_ptr())
        return dctx, dwords, None, None, None, None, None, None, None, None


_SAVE_CAP_BYTES = 24 << 30


def _save_enabled() -> bool:
    return os.environ.get("TGFR_WORDREGION_SAVE", "1").lower() not in ("0", "no", "off", "false")


def tc_supported(T, R, D) -> bool:
    """Shapes the tcgen05 kernels take (csrc/wordregion_tc.cu make_plan); others run the fp32 SIMT kernels."""
    return D % 64 == 0 and 64 <= D <= 256 and 1 <= R <= 256 and 1 <= T <= 32


def wordregion_sim(feats, words, cap_lens, g1, g2, g3, eps=1e-8, precision=None, want_attn=True, diag_off=0):
    """sim [Bc,Bq] (differentiable) and the diagonal attention maps [Bc,T,R] (or None).

    precision=None: TGFR_WORDREGION_PRECISION (default tc) when the shape fits the tensor-core kernel,
    else the fp32 SIMT CUDA kernel.  An explicit PREC_TC on an unsupported shape raises."""
    if precision is None:
        precision = default_precision()
        if precision == PREC_TC and not tc_supported(words.shape[1], feats.shape[1], feats.shape[2]):
            precision = PREC_FP32
    if cap_lens is not None:
        cap_lens = cap_lens.to(device=feats.device, dtype=torch.int32).contiguous()
    return _WordRegionSim.apply(_f32(feats), _f32(words), cap_lens, float(g1), float(g2), float(g3), float(eps),
                                int(precision), bool(want_attn), int(diag_off))


# ---------------------------------------------------------------------------------------------
# two-direction cross entropy over a square score matrix (single device)
# ---------------------------------------------------------------------------------------------
class _PairCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores):
        _lib.ensure_device(scores.device)
        scores = scores.contiguous()
        Bx, By = scores.shape
        dev = scores.device
        stats = torch.empty(2 * Bx + 2 * By, dtype=torch.float32, device=dev)
        rowlse, colmax, colsum, diag = stats[:Bx], stats[Bx:Bx + By], stats[Bx + By:Bx + 2 * By], stats[Bx + 2 * By:]
        losses = torch.empty(2, dtype=torch.float32, device=dev)
        collse = torch.empty(By, dtype=torch.float32, device=dev)
        st = stream_ptr()
        _call("tgfr_pair_ce_stats", scores.data_ptr(), Bx, By, 0, rowlse.data_ptr(), colmax.data_ptr(),
              colsum.data_ptr(), diag.data_ptr(), st)
        _call("tgfr_pair_ce_finish", rowlse.data_ptr(), colmax.data_ptr(), colsum.data_ptr(), diag.data_ptr(),
              Bx, By, 0, 1.0 / Bx, losses.data_ptr(), collse.data_ptr(), st)
        ctx.save_for_backward(scores, stats, collse)
        return losses[0].clone(), losses[1].clone()

    @staticmethod
    def backward(ctx, g0, g1):
        scores, stats, collse = ctx.saved_tensors
        Bx, By = scores.shape
        g = torch.stack([_f32(g0).reshape(()), _f32(g1).reshape(())]).contiguous()
        gs = torch.empty_like(scores)
        _call("tgfr_pair_ce_bwd", scores.data_ptr(), stats.data_ptr(), collse.data_ptr(), g[0:].data_ptr(),
              g[1:].data_ptr(), Bx, By, 0, 1.0 / Bx, gs.data_ptr(), stream_ptr())
        return gs


def pair_ce(scores):
    """(loss0, loss1): CE(scores, arange) and CE(scores^T, arange), mean reduction."""
    if scores.dim() != 2 or scores.shape[0] != scores.shape[1]:
        raise ValueError(f"pair_ce expects a square [B,B] matrix, got {tuple(scores.shape)}")
    return _PairCE.apply(_f32(scores))


# ---------------------------------------------------------------------------------------------
# cosine score matrix (sent_loss / global_loss / ClipLoss)
# ---------------------------------------------------------------------------------------------
class _CosineScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, scale, normalise, eps, ids_x, ids_y):
        _lib.ensure_device(x.device)
        x, y = x.contiguous(), y.contiguous()
        Bx, D = x.shape
        By = y.shape[0]
        dev = x.device
        scores = torch.empty((Bx, By), dtype=torch.float32, device=dev)
        xn = torch.empty(Bx, dtype=torch.float32, device=dev)
        yn = torch.empty(By, dtype=torch.float32, device=dev)
        _call("tgfr_cosine_scores_fwd", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), Bx, By, D, scale,
              int(normalise), eps, ptr(ids_x), ptr(ids_y), 0, scores.data_ptr(), xn.data_ptr(), yn.data_ptr(),
              stream_ptr())
        ctx.save_for_backward(x, y, xn, yn)
        ctx.cfg = (scale, normalise, eps)
        return scores

    @staticmethod
    def backward(ctx, gs):
        x, y, xn, yn = ctx.saved_tensors
        scale, normalise, eps = ctx.cfg
        Bx, D = x.shape
        By = y.shape[0]
        gs = _f32(gs).contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        lib = _lib.load()
        wsb = lib.tgfr_cosine_workspace_bytes(Bx, By, D)
        ws = _workspace(wsb, x.device)
        _call("tgfr_cosine_scores_bwd", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), Bx, By, D, scale,
              int(normalise), eps, xn.data_ptr(), yn.data_ptr(), gs.data_ptr(), ptr(dx), ptr(dy), ptr(ws), wsb,
              stream_ptr())
        return dx, dy, None, None, None, None, None


def cosine_scores(x, y, scale, normalise=True, eps=1e-8, ids_x=None, ids_y=None):
    return _CosineScores.apply(_f32(x), _f32(y), float(scale), bool(normalise), float(eps), ids_x, ids_y)


# ---------------------------------------------------------------------------------------------
# ArcFace logits
# ---------------------------------------------------------------------------------------------
class _ArcLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, label, s, m, easy, class_off):
        _lib.ensure_device(x.device)
        x = x.contiguous()
        B, Din = x.shape
        C = weight.shape[0]
        dev = x.device
        out = torch.empty((B, C), dtype=torch.float32, device=dev)
        xn = torch.empty(B, dtype=torch.float32, device=dev)
        wn = torch.empty(C, dtype=torch.float32, device=dev)
        cos_t = torch.empty(B, dtype=torch.float32, device=dev)
        st = stream_ptr()
        prec = head_precision(Din)
        ws, wsb = _head_ws(B, C, Din, prec, dev)
        _call("tgfr_cos_logits_fwd", x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(0),
              weight.stride(1), B, C, Din, s, 0, out.data_ptr(), out.stride(0), xn.data_ptr(), wn.data_ptr(),
              prec, ptr(ws), wsb, st)
        _call("tgfr_arc_margin_apply", out.data_ptr(), out.stride(0), label.data_ptr(), B, C, class_off, s, m,
              int(easy), cos_t.data_ptr(), st)
        ctx.save_for_backward(x, weight, label, xn, wn, cos_t)
        ctx.cfg = (s, m, easy, class_off, prec)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, label, xn, wn, cos_t = ctx.saved_tensors
        s, m, easy, class_off, prec = ctx.cfg
        B, Din = x.shape
        C = weight.shape[0]
        g = _f32(g)
        if g.stride(1) != 1:
            g = g.contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(weight, memory_format=torch.contiguous_format)
        ws, wsb = _head_ws(B, C, Din, prec, x.device)
        _call("tgfr_arc_margin_bwd", x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(0),
              weight.stride(1), xn.data_ptr(), wn.data_ptr(), label.data_ptr(), cos_t.data_ptr(), g.data_ptr(),
              g.stride(0), B, C, Din, class_off, s, m, int(easy), ptr(dx), dw.data_ptr(), prec, ptr(ws), wsb,
              stream_ptr())
        # dw was written with weight's strides; both are [C,Din] contiguous here
        return dx, dw, None, None, None, None, None


def arc_logits(x, weight, label, s, m, easy_margin=False, class_off=0):
    label = label.view(-1).to(device=x.device, dtype=torch.int64).contiguous()
    return _ArcLogits.apply(_f32(x), _f32(weight).contiguous(), label, float(s), float(m), bool(easy_margin),
                            int(class_off))


# ---------------------------------------------------------------------------------------------
# fused ArcFace logits + focal cross entropy: the [B, C] logits are never materialised
# ---------------------------------------------------------------------------------------------
class _ArcFusedFocal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, label, s, m, easy, gamma, class_off, merge):
        # merge: None, or a callable (rowmax, rowsum, tgt) -> global (rowmax, rowsum, tgt) over class shards
        _lib.ensure_device(x.device)
        x = x.contiguous()
        B, Din = x.shape
        C = weight.shape[0]
        dev = x.device
        lib = _lib.load()
        stats = torch.empty(7 * B + C, dtype=torch.float32, device=dev)
        xn, rowmax, rowsum, tgt, cos_t, lse = (stats[k * B:(k + 1) * B] for k in range(6))
        wn = stats[7 * B:]
        wsb = lib.tgfr_arc_fused_workspace_bytes(B, C, Din)
        svb = lib.tgfr_arc_fused_saved_bytes(B, C, Din)
        ws = _workspace(wsb, dev)
        saved = torch.empty(svb, dtype=torch.uint8, device=dev)
        st = stream_ptr()
        _call("tgfr_arc_fused_fwd", x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(0), weight.stride(1),
              label.data_ptr(), B, C, Din, class_off, s, m, int(easy), xn.data_ptr(), wn.data_ptr(), rowmax.data_ptr(),
              rowsum.data_ptr(), tgt.data_ptr(), cos_t.data_ptr(), ptr(ws), wsb, saved.data_ptr(), svb, st)
        if merge is not None:
            rowmax, rowsum, tgt = merge(rowmax, rowsum, tgt)
        out = torch.empty(3, dtype=torch.float32, device=dev)
        _call("tgfr_focal_finish", rowmax.data_ptr(), rowsum.data_ptr(), tgt.data_ptr(), B, gamma, out.data_ptr(),
              lse.data_ptr(), stream_ptr())
        ctx.save_for_backward(x, weight, label, stats, out, saved)
        ctx.cfg = (s, m, easy, class_off)
        return out[1].clone()

    @staticmethod
    def backward(ctx, gout):
        x, weight, label, stats, out, saved = ctx.saved_tensors
        s, m, easy, class_off = ctx.cfg
        B, Din = x.shape
        C = weight.shape[0]
        xn, lse, wn = stats[:B], stats[5 * B:6 * B], stats[7 * B:]
        gout = _f32(gout).reshape(1).contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(weight, memory_format=torch.contiguous_format)
        lib = _lib.load()
        wsb = lib.tgfr_arc_fused_workspace_bytes(B, C, Din)
        ws = _workspace(wsb, x.device)
        _call("tgfr_arc_fused_bwd", x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(0), weight.stride(1),
              label.data_ptr(), xn.data_ptr(), wn.data_ptr(), lse.data_ptr(), out[2:].data_ptr(), gout.data_ptr(),
              B, C, Din, class_off, s, m, int(easy), ptr(dx), dw.data_ptr(), ptr(ws), wsb, saved.data_ptr(),
              saved.numel(), stream_ptr())
        return dx, dw, None, None, None, None, None, None, None


def arc_fused_focal(x, weight, label, s, m, easy_margin=False, gamma=0.0, class_off=0, merge=None):
    """FocalLoss(gamma)(ArcMarginProduct(x, label), label) (metrics.py:42-60 + losses.py:313-325) without the [B,C]
    logits: margin, online-softmax statistics and the softmax gradient live in the epilogues of the tcgen05
    cos-theta GEMM.  weight [C,Din]; class_off / merge serve the class-sharded head (distributed.py)."""
    label = label.view(-1).to(device=x.device, dtype=torch.int64).contiguous()
    return _ArcFusedFocal.apply(_f32(x), _f32(weight).contiguous(), label, float(s), float(m), bool(easy_margin),
                                float(gamma), int(class_off), merge)


# ---------------------------------------------------------------------------------------------
# row-wise cross entropy + focal transform of its batch mean
# ---------------------------------------------------------------------------------------------
class _FocalCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, gamma):
        _lib.ensure_device(logits.device)
        if logits.stride(1) != 1:
            logits = logits.contiguous()
        B, C = logits.shape
        dev = logits.device
        stats = torch.empty(4 * B, dtype=torch.float32, device=dev)
        rowmax, rowsum, tgt, lse = stats[:B], stats[B:2 * B], stats[2 * B:3 * B], stats[3 * B:]
        out = torch.empty(3, dtype=torch.float32, device=dev)
        st = stream_ptr()
        _call("tgfr_ce_rows_stats", logits.data_ptr(), logits.stride(0), target.data_ptr(), B, C, 0,
              rowmax.data_ptr(), rowsum.data_ptr(), tgt.data_ptr(), st)
        _call("tgfr_focal_finish", rowmax.data_ptr(), rowsum.data_ptr(), tgt.data_ptr(), B, gamma, out.data_ptr(),
              lse.data_ptr(), st)
        ctx.save_for_backward(logits, target, stats, out)
        return out[1].clone()

    @staticmethod
    def backward(ctx, gout):
        logits, target, stats, out = ctx.saved_tensors
        B, C = logits.shape
        lse = stats[3 * B:]
        gout = _f32(gout).reshape(1).contiguous()
        gl = torch.empty((B, C), dtype=torch.float32, device=logits.device)
        _call("tgfr_ce_rows_bwd", logits.data_ptr(), logits.stride(0), target.data_ptr(), lse.data_ptr(),
              out[2:].data_ptr(), gout.data_ptr(), B, C, 0, gl.data_ptr(), gl.stride(0), stream_ptr())
        return gl, None, None


def focal_ce(logits, target, gamma=0.0):
    """(1 - exp(-CE))**gamma * CE with CE = mean cross entropy over the batch; gamma=0 -> plain CE."""
    if logits.dim() != 2:
        raise ValueError(f"focal_ce expects [B,C] logits, got {tuple(logits.shape)}")
    target = target.view(-1).to(device=logits.device, dtype=torch.int64).contiguous()
    return _FocalCE.apply(_f32(logits), target, float(gamma))


# ---------------------------------------------------------------------------------------------
# MagFace logits: (scale*cos, scale*cos(theta+m)) with per-row margins
# ---------------------------------------------------------------------------------------------
class _MagLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, margin, scale, easy):
        _lib.ensure_device(x.device)
        x = x.contiguous()
        B, Din = x.shape
        C = weight.shape[1]                      # MagLinear.weight is [Din, C]
        dev = x.device
        cos_s = torch.empty((B, C), dtype=torch.float32, device=dev)
        cos_m = torch.empty((B, C), dtype=torch.float32, device=dev)
        xn = torch.empty(B, dtype=torch.float32, device=dev)
        wn = torch.empty(C, dtype=torch.float32, device=dev)
        mar = margin.reshape(-1).contiguous()
        st = stream_ptr()
        prec = head_precision(Din)
        ws, wsb = _head_ws(B, C, Din, prec, dev)
        _call("tgfr_cos_logits_fwd", x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(1),
              weight.stride(0), B, C, Din, scale, 1, cos_s.data_ptr(), cos_s.stride(0), xn.data_ptr(),
              wn.data_ptr(), prec, ptr(ws), wsb, st)
        _call("tgfr_mag_margin_fwd", cos_s.data_ptr(), mar.data_ptr(), B, C, scale, int(easy), cos_m.data_ptr(), st)
        ctx.save_for_backward(x, weight, mar, cos_s, xn, wn)
        ctx.cfg = (scale, easy, margin.shape, prec)
        return cos_s, cos_m

    @staticmethod
    def backward(ctx, g_cos, g_cosm):
        x, weight, mar, cos_s, xn, wn = ctx.saved_tensors
        scale, easy, mshape, prec = ctx.cfg
        B, Din = x.shape
        C = weight.shape[1]
        dev = x.device
        g_cos = None if g_cos is None else _f32(g_cos).contiguous()
        g_cosm = None if g_cosm is None else _f32(g_cosm).contiguous()
        gtotal = torch.empty((B, C), dtype=torch.float32, device=dev)
        gmar = torch.empty(B, dtype=torch.float32, device=dev)
        st = stream_ptr()
        _call("tgfr_mag_margin_bwd", cos_s.data_ptr(), mar.data_ptr(), ptr(g_cos), ptr(g_cosm), B, C, scale,
              int(easy), gtotal.data_ptr(), gmar.data_ptr(), st)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(weight, memory_format=torch.contiguous_format)
        ws, wsb = _head_ws(B, C, Din, prec, dev)
        _call("tgfr_cos_logits_bwd", x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(1),
              weight.stride(0), xn.data_ptr(), wn.data_ptr(), cos_s.data_ptr(), cos_s.stride(0),
              gtotal.data_ptr(), gtotal.stride(0), B, C, Din, scale, 1, ptr(dx), dw.data_ptr(), prec, ptr(ws), wsb, st)
        return dx, dw, gmar.reshape(mshape), None, None


def mag_logits(x, weight, margin, scale, easy_margin=True):
    return _MagLogits.apply(_f32(x), _f32(weight).contiguous(), _f32(margin), float(scale), bool(easy_margin))


class _CosineRows(torch.autograd.Function):
    """cosine_similarity(x1, x2, dim=1, eps) of models/losses.py:12-16 on [N, D] rows."""

    @staticmethod
    def forward(ctx, x1, x2, eps):
        _lib.ensure_device(x1.device)
        N, D = x1.shape
        out = torch.empty(N, dtype=torch.float32, device=x1.device)
        stats = torch.empty((N, 3), dtype=torch.float32, device=x1.device)
        _call("tgfr_cosine_rows_fwd", ptr(x1), x1.stride(0), x1.stride(1), ptr(x2), x2.stride(0), x2.stride(1), N, D, eps,
              ptr(out), ptr(stats), stream_ptr())
        ctx.save_for_backward(x1, x2, stats)
        ctx.eps = eps
        return out

    @staticmethod
    def backward(ctx, g):
        x1, x2, stats = ctx.saved_tensors
        N, D = x1.shape
        g = _f32(g).contiguous()
        d1 = torch.empty((N, D), dtype=torch.float32, device=x1.device) if ctx.needs_input_grad[0] else None
        d2 = torch.empty((N, D), dtype=torch.float32, device=x1.device) if ctx.needs_input_grad[1] else None
        _call("tgfr_cosine_rows_bwd", ptr(x1), x1.stride(0), x1.stride(1), ptr(x2), x2.stride(0), x2.stride(1), N, D,
              ctx.eps, ptr(stats), ptr(g), ptr(d1), ptr(d2), stream_ptr())
        return d1, d2, None


def cosine_rows(x1, x2, eps=1e-8):
    """sum(x1 x2, 1) / max(|x1| |x2|, eps) for two [N, D] CUDA tensors (any strides) -> [N], differentiable."""
    if x1.dim() != 2 or x1.shape != x2.shape:
        raise RuntimeError(f"cosine_rows: expected two [N, D] tensors of one shape, got {tuple(x1.shape)} / {tuple(x2.shape)}")
    return _CosineRows.apply(_f32(x1), _f32(x2), float(eps))


class _MagCE(torch.autograd.Function):
    """MagLoss's blend + mean cross entropy (magface.py:131-135) without the blended [B,C] tensor."""

    @staticmethod
    def forward(ctx, cos_s, cos_m, target):
        _lib.ensure_device(cos_s.device)
        if cos_s.stride(1) != 1 or cos_m.stride() != cos_s.stride():
            cos_s, cos_m = cos_s.contiguous(), cos_m.contiguous()
        B, C = cos_s.shape
        dev = cos_s.device
        stats = torch.empty(4 * B, dtype=torch.float32, device=dev)
        rowmax, rowsum, tgt, lse = stats[:B], stats[B:2 * B], stats[2 * B:3 * B], stats[3 * B:]
        out = torch.empty(3, dtype=torch.float32, device=dev)
        one_hot = torch.empty((B, C), dtype=torch.float32, device=dev)
        st = stream_ptr()
        _call("tgfr_mag_ce_stats", cos_s.data_ptr(), cos_m.data_ptr(), cos_s.stride(0), target.data_ptr(), B, C,
              rowmax.data_ptr(), rowsum.data_ptr(), tgt.data_ptr(), one_hot.data_ptr(), st)
        _call("tgfr_focal_finish", rowmax.data_ptr(), rowsum.data_ptr(), tgt.data_ptr(), B, 0.0, out.data_ptr(),
              lse.data_ptr(), st)
        ctx.save_for_backward(cos_s, cos_m, target, stats)
        ctx.mark_non_differentiable(one_hot)
        return out[1].clone(), one_hot

    @staticmethod
    def backward(ctx, gout, _g_onehot):
        cos_s, cos_m, target, stats = ctx.saved_tensors
        B, C = cos_s.shape
        lse = stats[3 * B:]
        gout = _f32(gout).reshape(1).contiguous()
        g_cos = torch.empty((B, C), dtype=torch.float32, device=cos_s.device)
        g_cosm = torch.empty((B, C), dtype=torch.float32, device=cos_s.device)
        _call("tgfr_mag_ce_bwd", cos_s.data_ptr(), cos_m.data_ptr(), cos_s.stride(0), target.data_ptr(), lse.data_ptr(),
              gout.data_ptr(), B, C, g_cos.data_ptr(), g_cosm.data_ptr(), stream_ptr())
        return g_cos, g_cosm, None


def mag_ce(cos_s, cos_m, target):
    """(mean CE of the MagLoss-blended logits, one_hot [B,C]) -- label column from cos_m, the rest from cos_s."""
    if cos_s.dim() != 2 or cos_s.shape != cos_m.shape:
        raise ValueError(f"mag_ce expects two [B,C] tensors, got {tuple(cos_s.shape)} / {tuple(cos_m.shape)}")
    target = target.view(-1).to(device=cos_s.device, dtype=torch.int64).contiguous()
    return _MagCE.apply(_f32(cos_s), _f32(cos_m), target)


# ---------------------------------------------------------------------------------------------
# TextHeading: BERT tokens -> word / sentence features (n-gram convolutions, shifted max, L2 norm)
# ---------------------------------------------------------------------------------------------
class _TextHeading(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, w2, w3, w4, b2, b3, b4, words_num):
        _lib.ensure_device(tokens.device)
        tokens = tokens.contiguous()
        B, L, E = tokens.shape
        F = w2.shape[0]
        dev = tokens.device
        lib = _lib.load()
        words = torch.empty((B, words_num - 2, F), dtype=torch.float32, device=dev)
        sent = torch.empty((B, F), dtype=torch.float32, device=dev)
        svb = lib.tgfr_texthead_saved_bytes(B, L, E, F)
        saved = torch.empty(svb, dtype=torch.uint8, device=dev)
        ws_ = [w.contiguous() for w in (w2, w3, w4)]
        _call("tgfr_texthead_fwd", tokens.data_ptr(), ws_[0].data_ptr(), ws_[1].data_ptr(), ws_[2].data_ptr(),
              ptr(b2), ptr(b3), ptr(b4), B, L, E, F, words_num, words.data_ptr(), sent.data_ptr(), saved.data_ptr(), svb,
              stream_ptr())
        ctx.save_for_backward(tokens, saved)
        ctx.cfg = (words_num, F, tuple(w.shape for w in (w2, w3, w4)), b2 is not None)
        return words, sent

    @staticmethod
    def backward(ctx, gwords, gsent):
        tokens, saved = ctx.saved_tensors
        words_num, F, wshapes, has_bias = ctx.cfg
        B, L, E = tokens.shape
        dev = tokens.device
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("TextHeading: gradient w.r.t. the BERT tokens is not provided (the reference "
                                      "trains the head on a frozen encoder)")
        gwords = None if gwords is None else _f32(gwords).contiguous()
        gsent = None if gsent is None else _f32(gsent).contiguous()
        dws = [torch.empty(s_, dtype=torch.float32, device=dev) for s_ in wshapes]
        dbs = [torch.empty(F, dtype=torch.float32, device=dev) for _ in range(3)]
        lib = _lib.load()
        wsb = lib.tgfr_texthead_workspace_bytes(B, L, E, F)
        ws = _workspace(wsb, dev)
        _call("tgfr_texthead_bwd", tokens.data_ptr(), ptr(gwords), ptr(gsent), B, L, E, F, words_num,
              dws[0].data_ptr(), dws[1].data_ptr(), dws[2].data_ptr(), dbs[0].data_ptr(), dbs[1].data_ptr(),
              dbs[2].data_ptr(), ptr(ws), wsb, saved.data_ptr(), saved.numel(), stream_ptr())
        if not has_bias:
            dbs = [None, None, None]
        return (None, *dws, *dbs, None)


def text_heading(tokens, weights, biases, bert_words_num):
    """(words [B, T, F] unit rows, sent [B, F] unit rows) of models/models.py:170-232.
    weights: the three Conv2d(1, F, (K, 768)) weights [F, 1, K, E]; biases: three [F] tensors (or Nones)."""
    w2, w3, w4 = (_f32(w) for w in weights)
    b2, b3, b4 = (None if b is None else _f32(b).contiguous() for b in biases)
    return _TextHeading.apply(_f32(tokens), w2, w3, w4, b2, b3, b4, int(bert_words_num))


# ---------------------------------------------------------------------------------------------
# stand-alone func_attention on canonical layouts
# ---------------------------------------------------------------------------------------------
class _FuncAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, query, g1):
        # feats [B,R,D], query [B,T,D] (arbitrary strides) -> wc [B,T,D], attn [B,T,R]
        _lib.ensure_device(feats.device)
        B, R, D = feats.shape
        T = query.shape[1]
        wc = torch.empty((B, T, D), dtype=torch.float32, device=feats.device)
        attn = torch.empty((B, T, R), dtype=torch.float32, device=feats.device)
        _call("tgfr_attention_fwd", feats.data_ptr(), *feats.stride(), query.data_ptr(), *query.stride(), B, T, R,
              D, g1, wc.data_ptr(), attn.data_ptr(), stream_ptr())
        ctx.save_for_backward(feats, query)
        ctx.g1 = g1
        return wc, attn

    @staticmethod
    def backward(ctx, g_wc, g_attn):
        feats, query = ctx.saved_tensors
        B, R, D = feats.shape
        T = query.shape[1]
        g_wc = None if g_wc is None else _f32(g_wc).contiguous()
        g_attn = None if g_attn is None else _f32(g_attn).contiguous()
        dctx = torch.empty((B, R, D), dtype=torch.float32, device=feats.device) if ctx.needs_input_grad[0] else None
        dq = torch.empty((B, T, D), dtype=torch.float32, device=feats.device) if ctx.needs_input_grad[1] else None
        _call("tgfr_attention_bwd", feats.data_ptr(), *feats.stride(), query.data_ptr(), *query.stride(), B, T, R,
              D, ctx.g1, ptr(g_wc), ptr(g_attn), ptr(dctx), ptr(dq), stream_ptr())
        return dctx, dq, None


def func_attention_canonical(feats, query, gamma1):
    return _FuncAttention.apply(_f32(feats), _f32(query), float(gamma1))


# ---------------------------------------------------------------------------------------------
# plain cosine logits s * cos(x, w_c) with weight [C, Din] (AddMargin / Sphere / AdaFace heads)
# ---------------------------------------------------------------------------------------------
class _CosLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, s, clamp, w_class_dim):
        _lib.ensure_device(x.device)
        x = x.contiguous()
        B, Din = x.shape
        C = weight.shape[w_class_dim]
        dev = x.device
        out = torch.empty((B, C), dtype=torch.float32, device=dev)
        xn = torch.empty(B, dtype=torch.float32, device=dev)
        wn = torch.empty(C, dtype=torch.float32, device=dev)
        prec = head_precision(Din)
        ws, wsb = _head_ws(B, C, Din, prec, dev)
        _call("tgfr_cos_logits_fwd", x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(w_class_dim),
              weight.stride(1 - w_class_dim), B, C, Din, s, int(clamp), out.data_ptr(), out.stride(0),
              xn.data_ptr(), wn.data_ptr(), prec, ptr(ws), wsb, stream_ptr())
        ctx.save_for_backward(x, weight, out, xn, wn)
        ctx.cfg = (s, clamp, w_class_dim, prec)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, out, xn, wn = ctx.saved_tensors
        s, clamp, wd, prec = ctx.cfg
        B, Din = x.shape
        C = weight.shape[wd]
        g = _f32(g).contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(weight, memory_format=torch.contiguous_format)
        ws, wsb = _head_ws(B, C, Din, prec, x.device)
        _call("tgfr_cos_logits_bwd", x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(wd),
              weight.stride(1 - wd), xn.data_ptr(), wn.data_ptr(), out.data_ptr(), out.stride(0), g.data_ptr(),
              g.stride(0), B, C, Din, s, int(clamp), ptr(dx), dw.data_ptr(), prec, ptr(ws), wsb, stream_ptr())
        return dx, dw, None, None, None


def cos_logits(x, weight, s=1.0, clamp=False, w_class_dim=0):
    """s * normalize(x) @ normalize(weight)^T; weight is [C,Din] (w_class_dim=0) or [Din,C] (=1)."""
    return _CosLogits.apply(_f32(x), _f32(weight).contiguous(), float(s), bool(clamp), int(w_class_dim))


# ---------------------------------------------------------------------------------------------
# verification / identification scoring (utils/modules.py:40-88,150-166; no autograd: evaluation only)
# ---------------------------------------------------------------------------------------------
def pair_cosine(x1, x2, eps=1e-6):
    """nn.CosineSimilarity(dim=1, eps)(x1, x2) of utils/modules.py:150-151 for [N, D] CUDA tensors -> [N] fp32."""
    if x1.dim() != 2 or x1.shape != x2.shape:
        raise RuntimeError(f"pair_cosine: expected two [N, D] tensors of one shape, got {tuple(x1.shape)} and {tuple(x2.shape)}")
    _lib.ensure_device(x1.device)
    if x2.device != x1.device:
        raise RuntimeError("pair_cosine: tensors on different devices")
    x1, x2 = _f32(x1.detach()), _f32(x2.detach())
    N, D = x1.shape
    out = torch.empty(N, dtype=torch.float32, device=x1.device)
    with torch.cuda.device(x1.device):
        _call("tgfr_pair_cosine", ptr(x1), x1.stride(0), x1.stride(1), ptr(x2), x2.stride(0), x2.stride(1), N, D,
              float(eps), ptr(out), stream_ptr())
    return out


def roc_counts(scores, labels, drop_intermediate=True):
    """Integer part of sklearn.metrics.roc_curve(labels, scores) on the device: (thresholds fp32 [M], fps int64 [M],
    tps int64 [M], distinct) for fp32 CUDA scores [N] and integer labels [N] (positive = 1), without the leading
    (inf, 0, 0) point.  Raises ValueError on NaN scores (as scikit-learn's input validation does)."""
    _lib.ensure_device(scores.device)
    scores = _f32(scores.detach()).contiguous().view(-1)
    labels = labels.detach().to(device=scores.device, dtype=torch.int64).contiguous().view(-1)
    N = scores.numel()
    if labels.numel() != N:
        raise ValueError(f"roc_counts: {N} scores but {labels.numel()} labels")
    dev = scores.device
    thr = torch.empty(max(N, 1), dtype=torch.float32, device=dev)
    fps = torch.empty(max(N, 1), dtype=torch.int64, device=dev)
    tps = torch.empty(max(N, 1), dtype=torch.int64, device=dev)
    counts = torch.empty(3, dtype=torch.int64, device=dev)
    wsb = _lib.load().tgfr_roc_workspace_bytes(N)
    ws = _workspace(wsb, dev)
    with torch.cuda.device(dev):
        _call("tgfr_roc_curve", ptr(scores), ptr(labels), N, 1 if drop_intermediate else 0, ptr(thr), ptr(fps), ptr(tps),
              ptr(counts), ptr(ws), wsb, stream_ptr())
    m, nan, distinct = (int(v) for v in counts.tolist())       # the one device -> host sync of the curve
    if nan:
        raise ValueError("Input y_score contains NaN.")
    return thr[:m], fps[:m], tps[:m], distinct


def row_argmax(scores):
    """np.argmax(scores, axis=1) (first maximum; NaN counts as the maximum) for a [rows, cols] CUDA tensor -> int64 [rows]."""
    if scores.dim() != 2:
        raise RuntimeError(f"row_argmax: expected [rows, cols], got {tuple(scores.shape)}")
    _lib.ensure_device(scores.device)
    scores = _f32(scores.detach())
    if scores.stride(1) != 1:
        scores = scores.contiguous()
    rows, cols = scores.shape
    out = torch.empty(rows, dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        _call("tgfr_row_argmax", ptr(scores), scores.stride(0), rows, cols, ptr(out), stream_ptr())
    return out


# ---------------------------------------------------------------------------------------------
# FCFM fusion net `Working`, eval-mode forward (models/fusion_nets.py:217-258)
# ---------------------------------------------------------------------------------------------
FCFM_PARAM_ORDER = (
    "conv.weight", "conv.bias", "bn_img.weight", "bn_img.bias", "bn_img.running_mean", "bn_img.running_var",
    "projection.weight", "projection.bias", "bn_word.weight", "bn_word.bias", "bn_word.running_mean", "bn_word.running_var",
    "sa.query_proj.weight", "sa.query_proj.bias", "sa.key_proj.weight", "sa.key_proj.bias", "sa.value_proj.weight",
    "sa.value_proj.bias", "ln.weight", "ln.bias", "linear.weight", "linear.bias", "ln_gl_image.weight", "ln_gl_image.bias",
    "ln_sent.weight", "ln_sent.bias",
)


def _fcfm_conv_on_tensor_cores(B) -> bool:
    """TGFR_FCFM_CONV=tc|simt forces the convolution path; by default batches of 64 samples and more take the tensor cores
    (below that the one-launch per-sample kernel wins on launch count)."""
    mode = os.environ.get("TGFR_FCFM_CONV", "").lower()
    if mode in ("tc", "simt"):
        return mode == "tc"
    return B >= 64


def fcfm_working(img, word, gl_img, sent, state):
    """Working.forward in eval mode: img [B,256,14,14] (any strides), word [B,256,T], gl_img / sent [B,256] -> [B,640].
    `state` maps the reference module's state_dict names (FCFM_PARAM_ORDER) to CUDA tensors.  No autograd."""
    import ctypes
    _lib.ensure_device(img.device)
    if img.dim() != 4 or tuple(img.shape[1:]) != (256, 14, 14):
        raise RuntimeError(f"fcfm_working: expected img [B,256,14,14], got {tuple(img.shape)}")
    B = img.shape[0]
    if word.dim() != 3 or word.shape[0] != B or word.shape[1] != 256:
        raise RuntimeError(f"fcfm_working: expected word [B,256,T], got {tuple(word.shape)}")
    if tuple(gl_img.shape) != (B, 256) or tuple(sent.shape) != (B, 256):
        raise RuntimeError(f"fcfm_working: expected gl_img / sent [B,256], got {tuple(gl_img.shape)} / {tuple(sent.shape)}")
    img, word = _f32(img.detach()), _f32(word.detach())
    gl_img, sent = _f32(gl_img.detach()), _f32(sent.detach())
    if gl_img.stride(1) != 1:
        gl_img = gl_img.contiguous()
    if sent.stride(1) != 1:
        sent = sent.contiguous()
    params = [_f32(state[name].detach()).contiguous() for name in FCFM_PARAM_ORDER]
    arr = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    out = torch.empty((B, 640), dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        if B > 0 and _fcfm_conv_on_tensor_cores(B):
            wsb = _lib.load().tgfr_fcfm_working_workspace_bytes(B)
            ws = _workspace(wsb, img.device)
            _call("tgfr_fcfm_working_fwd_tc", ptr(img), img.stride(0), img.stride(1), img.stride(2), img.stride(3), ptr(word),
                  word.stride(0), word.stride(1), word.stride(2), ptr(gl_img), gl_img.stride(0), ptr(sent), sent.stride(0),
                  ctypes.cast(arr, ctypes.c_void_p), len(params), B, word.shape[2], ptr(out), out.stride(0), ptr(ws), wsb,
                  stream_ptr())
        else:
            _call("tgfr_fcfm_working_fwd", ptr(img), img.stride(0), img.stride(1), img.stride(2), img.stride(3), ptr(word),
                  word.stride(0), word.stride(1), word.stride(2), ptr(gl_img), gl_img.stride(0), ptr(sent), sent.stride(0),
                  ctypes.cast(arr, ctypes.c_void_p), len(params), B, word.shape[2], ptr(out), out.stride(0), stream_ptr())
    return out


# ---------------------------------------------------------------------------------------------
# ImageHeading (models/models.py:328-338, 380-405): IMIM local branch and the global ProjectionHead
# ---------------------------------------------------------------------------------------------
IMIM_PARAM_ORDER = (
    "bn_img.weight", "bn_img.bias", "sa.query_proj.weight", "sa.query_proj.bias", "sa.key_proj.weight", "sa.key_proj.bias",
    "sa.value_proj.weight", "sa.value_proj.bias", "ln.weight", "ln.bias", "conv1x1_1.weight", "conv1x1_1.bias",
    "conv1x1_2.weight", "conv1x1_2.bias", "project_local.projection.weight", "project_local.projection.bias",
)


def _ptr_array(tensors):
    import ctypes
    arr = (ctypes.c_void_p * len(tensors))(*[ptr(t) for t in tensors])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


class _Imim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, running_mean, running_var, training, momentum, eps, *params):
        _lib.ensure_device(x.device)
        B, C, H, Wd = x.shape
        P = H * Wd
        dev = x.device
        lib = _lib.load()
        prm = [_f32(p_.detach()).contiguous() for p_ in params]
        out = torch.empty((B, P, C), dtype=torch.float32, device=dev)
        svb = lib.tgfr_imim_saved_bytes(B, P)
        saved = torch.empty(svb, dtype=torch.uint8, device=dev)
        xs = x if x.stride(2) == Wd * x.stride(3) else x.contiguous()     # the (h, w) axes must merge into one position axis
        sb, sc, sp = xs.stride(0), xs.stride(1), xs.stride(3)
        arr, parr = _ptr_array(prm)
        _call("tgfr_imim_fwd", ptr(xs), sb, sc, sp, parr, len(prm), B, P, int(training), float(momentum), float(eps),
              ptr(running_mean), ptr(running_var), ptr(out), ptr(saved), svb, stream_ptr())
        ctx.save_for_backward(xs, out, saved, *prm)
        ctx.cfg = (B, C, H, Wd, sb, sc, sp, bool(training))
        return out

    @staticmethod
    def backward(ctx, gout):
        xs, out, saved, *prm = ctx.saved_tensors
        B, C, H, Wd, sb, sc, sp, training = ctx.cfg
        P = H * Wd
        dev = out.device
        lib = _lib.load()
        gout = _f32(gout).contiguous()
        dprm = [torch.empty_like(p_) for p_ in prm]
        dx = torch.empty((B, C, H, Wd), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        wsb = lib.tgfr_imim_workspace_bytes(B, P)
        ws = _workspace(wsb, dev)
        arr, parr = _ptr_array(prm)
        darr, dparr = _ptr_array(dprm)
        _call("tgfr_imim_bwd", ptr(gout), ptr(out), ptr(xs), sb, sc, sp, parr, len(prm), B, P, int(training), ptr(saved),
              saved.numel(), dparr, ptr(dx), ptr(ws), wsb, stream_ptr())
        return (dx, None, None, None, None, None, *dprm)


def imim(x, params, running_mean, running_var, training, momentum=0.1, eps=1e-5):
    """IMIM.forward: x [B,256,14,14] -> [B, 196, 256] unit rows (the reference's result in ITS memory order);
    params: the 16 tensors of IMIM_PARAM_ORDER.  Differentiable w.r.t. x and the parameters."""
    if x.dim() != 4 or x.shape[1] != 256:
        raise RuntimeError(f"imim: expected x [B,256,H,W], got {tuple(x.shape)}")
    return _Imim.apply(_f32(x), running_mean, running_var, bool(training), float(momentum), float(eps), *params)


class _ProjHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        _lib.ensure_device(x.device)
        if x.stride(1) != 1:
            x = x.contiguous()
        weight = weight.contiguous()
        M, K = x.shape
        N = weight.shape[0]
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
        znorm = torch.empty(M, dtype=torch.float32, device=x.device)
        _call("tgfr_proj_head_fwd", ptr(x), x.stride(0), ptr(weight), ptr(bias), M, N, K, ptr(out), ptr(znorm), stream_ptr())
        ctx.save_for_backward(x, weight, out, znorm)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, out, znorm = ctx.saved_tensors
        M, K = x.shape
        N = weight.shape[0]
        g = _f32(g).contiguous()
        dz = torch.empty((M, N), dtype=torch.float32, device=x.device)
        dx = torch.empty((M, K), dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(weight)
        db = torch.empty(N, dtype=torch.float32, device=x.device)
        _call("tgfr_proj_head_bwd", ptr(g), ptr(out), ptr(znorm), ptr(x), x.stride(0), ptr(weight), M, N, K, ptr(dz), ptr(dx),
              ptr(dw), ptr(db), stream_ptr())
        return dx, dw, (db if ctx.has_bias else None)


def matmul_split(a, b, mode=0, alpha=1.0, bias=None, relu=False, splits=1, nterms=3):
    """The tensor-core contraction IMIM's layers run on (C ABI tgfr_matmul_split; no autograd): a, b fp32, 2-D or 3-D
    (batch of densely packed samples).  mode 0: a [M,K] @ b [N,K].T; mode 1: a [M,K] @ b [K,N]; mode 2: a [K,M].T @ b [K,N].
    nterms = 3 accumulates the fp16 hi / lo split products (fp32-class accuracy), nterms = 1 the hi parts only."""
    _lib.ensure_device(a.device)
    a, b = _f32(a).contiguous(), _f32(b).contiguous()
    batch = a.shape[0] if a.dim() == 3 else 1
    am, bm = a.shape[-2:], b.shape[-2:]
    M, K = (am[1], am[0]) if mode == 2 else (am[0], am[1])
    N = bm[0] if mode == 0 else bm[1]
    if (bm[1] if mode == 0 else bm[0]) != K or (b.shape[0] if b.dim() == 3 else 1) != batch:
        raise ValueError(f"matmul_split: shapes {tuple(a.shape)} / {tuple(b.shape)} do not contract in mode {mode}")
    out = torch.empty((batch, M, N) if a.dim() == 3 else (M, N), dtype=torch.float32, device=a.device)
    lib = _lib.load()
    wsb = lib.tgfr_matmul_split_workspace_bytes(mode, M, N, K, batch)
    ws = _workspace(wsb, a.device)
    _call("tgfr_matmul_split", mode, a.data_ptr(), am[1], b.data_ptr(), bm[1], out.data_ptr(), N, M, N, K, batch,
          float(alpha), ptr(None if bias is None else _f32(bias).contiguous()), int(relu), int(splits), int(nterms),
          ptr(ws), wsb, stream_ptr())
    return out


def proj_head(x, weight, bias):
    """normalize(x @ weight.T + bias, dim=-1) (ProjectionHead.forward, models/models.py:111-119) for x [..., K]."""
    lead = x.shape[:-1]
    out = _ProjHead.apply(_f32(x).reshape(-1, x.shape[-1]), _f32(weight), None if bias is None else _f32(bias).contiguous())
    return out.reshape(*lead, weight.shape[0])


# ---------------------------------------------------------------------------------------------
# FCFM fusion net `Working`, training-mode forward + backward (models/fusion_nets.py:217-258 under autograd)
# ---------------------------------------------------------------------------------------------
FCFM_TRAIN_PARAM_ORDER = (
    "conv.weight", "conv.bias", "bn_img.weight", "bn_img.bias", "projection.weight", "projection.bias", "bn_word.weight",
    "bn_word.bias", "sa.query_proj.weight", "sa.query_proj.bias", "sa.key_proj.weight", "sa.key_proj.bias",
    "sa.value_proj.weight", "sa.value_proj.bias", "ln.weight", "ln.bias", "linear.weight", "linear.bias",
    "ln_gl_image.weight", "ln_gl_image.bias", "ln_sent.weight", "ln_sent.bias",
)


class _FcfmTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, word, gl_img, sent, run_stats, training, momentum, eps, *params):
        _lib.ensure_device(img.device)
        B, T = img.shape[0], word.shape[2]
        dev = img.device
        lib = _lib.load()
        if word.stride(2) != 1:
            word = word.contiguous()
        gl_img = gl_img if gl_img.stride(1) == 1 else gl_img.contiguous()
        sent = sent if sent.stride(1) == 1 else sent.contiguous()
        prm = [_f32(p_.detach()).contiguous() for p_ in params]
        out = torch.empty((B, 640), dtype=torch.float32, device=dev)
        svb = lib.tgfr_fcfm_train_saved_bytes(B, T)
        saved = torch.empty(svb, dtype=torch.uint8, device=dev)
        arr, parr = _ptr_array(prm)
        rarr, rparr = _ptr_array(list(run_stats))
        _call("tgfr_fcfm_train_fwd", ptr(img), *img.stride(), ptr(word), word.stride(0), word.stride(1), ptr(gl_img),
              gl_img.stride(0), ptr(sent), sent.stride(0), parr, len(prm), B, T, int(training), float(momentum), float(eps),
              rparr, ptr(out), out.stride(0), ptr(saved), svb, stream_ptr())
        ctx.save_for_backward(word, gl_img, sent, saved, *prm)
        ctx.cfg = (B, T, bool(training))
        return out

    @staticmethod
    def backward(ctx, gout):
        word, gl_img, sent, saved, *prm = ctx.saved_tensors
        B, T, training = ctx.cfg
        dev = word.device
        lib = _lib.load()
        gout = _f32(gout)
        if gout.stride(1) != 1:
            gout = gout.contiguous()
        need = ctx.needs_input_grad
        dimg = torch.empty((B, 256, 14, 14), dtype=torch.float32, device=dev) if need[0] else None
        dword = torch.empty((B, 256, T), dtype=torch.float32, device=dev) if need[1] else None
        dgl = torch.empty((B, 256), dtype=torch.float32, device=dev) if need[2] else None
        dsent = torch.empty((B, 256), dtype=torch.float32, device=dev) if need[3] else None
        dprm = [torch.empty_like(p_) for p_ in prm]
        wsb = lib.tgfr_fcfm_train_workspace_bytes(B, T)
        ws = _workspace(wsb, dev)
        arr, parr = _ptr_array(prm)
        darr, dparr = _ptr_array(dprm)
        _call("tgfr_fcfm_train_bwd", ptr(gout), gout.stride(0), ptr(word), word.stride(0), word.stride(1), ptr(gl_img),
              gl_img.stride(0), ptr(sent), sent.stride(0), parr, len(prm), B, T, int(training), ptr(saved), saved.numel(), dparr,
              ptr(dimg), ptr(dword), ptr(dgl), ptr(dsent), ptr(ws), wsb, stream_ptr())
        return (dimg, dword, dgl, dsent, None, None, None, None, *dprm)


def fcfm_working_train(img, word, gl_img, sent, params, run_stats, training=True, momentum=0.1, eps=1e-5):
    """Working.forward under autograd: img [B,256,14,14], word [B,256,T], gl_img / sent [B,256] -> [B,640];
    params: the 22 tensors of FCFM_TRAIN_PARAM_ORDER (4-D / 3-D weights flattened to 2-D / 1-D);
    run_stats: (bn_img.running_mean, bn_img.running_var, bn_word.running_mean, bn_word.running_var), updated in place."""
    if img.dim() != 4 or tuple(img.shape[1:]) != (256, 14, 14):
        raise RuntimeError(f"fcfm_working_train: expected img [B,256,14,14], got {tuple(img.shape)}")
    B = img.shape[0]
    if word.dim() != 3 or word.shape[0] != B or word.shape[1] != 256:
        raise RuntimeError(f"fcfm_working_train: expected word [B,256,T], got {tuple(word.shape)}")
    if tuple(gl_img.shape) != (B, 256) or tuple(sent.shape) != (B, 256):
        raise RuntimeError("fcfm_working_train: expected gl_img / sent [B,256]")
    return _FcfmTrain.apply(_f32(img), _f32(word), _f32(gl_img), _f32(sent), tuple(run_stats), bool(training),
                            float(momentum), float(eps), *params)
