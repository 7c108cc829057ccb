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

__all__ = ["run_overlapped", "fcam_losses"]

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
