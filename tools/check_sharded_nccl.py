"""Multi-GPU parity of the sharded FCAM path over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/check_sharded_nccl.py [--B 32]

Every rank owns B faces / captions / sentence pairs / head samples.  The sharded losses and the gradients of the
local tensors are compared with the fp64 oracle evaluated on the gathered global batch (rank 0 prints one
"sharded-nccl ok" line; any mismatch raises on the rank that sees it).  The oracle is the checker only.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import synth  # noqa: E402
from oracle import fcam_oracle as O  # noqa: E402
from text_guided_face_recognition_b200 import distributed as tdist  # noqa: E402

G = (4.0, 5.0, 10.0)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=12)
    ap.add_argument("--D", type=int, default=128)
    ap.add_argument("--C", type=int, default=1003)      # not divisible by 2/4/8: the remainder classes are exercised
    a = ap.parse_args()
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, T, R, D = a.B, a.T, 49, a.D
    Bg = B * world

    # the global batch is generated identically everywhere; a rank keeps its row block
    ctx, words, cap = synth.wordregion_inputs(Bg, T, R, D, "LSTM", seed=7, ragged=True)
    img, txt, cid = synth.sentence_inputs(Bg, D, seed=7, collisions=True)
    sl = slice(rank * B, (rank + 1) * B)
    c = torch.from_numpy(ctx[sl]).to(dev).requires_grad_(True)
    w = torch.from_numpy(words[sl]).to(dev).requires_grad_(True)
    x = torch.from_numpy(img[sl]).to(dev).requires_grad_(True)
    y = torch.from_numpy(txt[sl]).to(dev).requires_grad_(True)
    lens = torch.from_numpy(cap[sl]).to(dev)
    ids = torch.from_numpy(cid[sl]).to(dev)

    report = {}
    for prec_name, prec, ltol, gtol in (("fp32", 0, 2e-5, 1e-4), ("tc", 1, 1e-4, 1e-3)):
        for t in (c, w, x, y):
            t.grad = None
        l0, l1, att = tdist.words_loss_sharded(c, w, lens, *G, precision=prec, want_attn=True)
        s0, s1 = tdist.sent_loss_sharded(x, y, ids, G[2])
        (l0 + l1 + s0 + s1).backward()
        r0, r1, r_att, _ = O.words_loss(ctx, words, None, cap, *G)
        # this rank's diagonal attention maps (rows pair with captions rank * B_local + b): exact in fp32 mode, within
        # 2e-3 of a map's largest entry when the tensor-core forward emits them
        att_np = att.cpu().numpy()
        for bl in range(att_np.shape[0]):
            ref_map = r_att[sl.start + bl]
            n_w = ref_map.shape[0]
            tol_map = 1e-5 if prec_name == "fp32" else 2e-3 * float(np.max(ref_map))
            assert np.max(np.abs(att_np[bl, :n_w] - ref_map)) < tol_map, (prec_name, "att", bl)
            assert not att_np[bl, n_w:].any()
        q0, q1, _ = O.sent_loss(img, txt, None, cid, G[2])
        dctx, dwords = O.words_loss_grads(ctx, words, None, cap, *G)
        dimg, dtxt = O.sent_loss_grads(img, txt, None, cid, G[2])
        for got, ref, name in ((l0, r0, "w0"), (l1, r1, "w1"), (s0, q0, "s0"), (s1, q1, "s1")):
            err = abs(got.item() - ref) / abs(ref)
            assert err < ltol, (prec_name, name, got.item(), ref)
            report[f"{prec_name}.{name}"] = err
        for got, ref, name in ((c.grad, dctx[sl], "dctx"), (w.grad, dwords[sl], "dwords"),
                               (x.grad, dimg[sl], "dimg"), (y.grad, dtxt[sl], "dtxt")):
            err = rel(got.cpu().numpy(), ref)
            assert err < gtol, (prec_name, name, err)
            report[f"{prec_name}.{name}"] = err

    # class-sharded ArcFace head with the fused (never gathered) focal cross entropy
    Din, C = 2 * D, a.C
    xn, wn, lab = synth.margin_inputs(Bg, Din, C, seed=7)
    head = tdist.ShardedArcMarginProduct(Din, C, s=30.0, m=0.5).to(dev)
    head.load_full_weight(torch.from_numpy(wn).to(dev))
    hx = torch.from_numpy(xn[sl]).to(dev).requires_grad_(True)
    hl = torch.from_numpy(lab[sl]).to(dev)
    ref_logits = O.arc_margin(xn, wn, lab, 30.0, 0.5, False)
    ref_loss = O.focal_loss(ref_logits, lab, 2.0)
    dx_ref, dw_ref = O.arc_margin_bwd(xn, wn, lab, O.focal_loss_bwd(ref_logits, lab, 2.0), 30.0, 0.5, False)
    # fp32: the [B, C_local] shard is materialised; tc: fused GEMM epilogues, no logits anywhere
    for hp in ("fp32", "tc"):
        os.environ["TGFR_HEAD_PRECISION"] = hp
        hx.grad = None
        head.weight.grad = None
        loss = head.loss(hx, hl, gamma=2.0)
        loss.backward()
        err = abs(loss.item() - ref_loss) / abs(ref_loss)
        assert err < 1e-4, ("head loss", hp, loss.item(), ref_loss)
        e1 = rel(hx.grad.cpu().numpy(), dx_ref[sl])
        e2 = rel(head.weight.grad.cpu().numpy(), dw_ref[head.c0:head.c1])
        assert e1 < 1e-3 and e2 < 1e-3, ("head grads", hp, e1, e2)
        report[f"head.{hp}.loss"], report[f"head.{hp}.dx"], report[f"head.{hp}.dw"] = err, e1, e2

    # pair-sharded verification scoring: ragged shards of the pair list, one gather, the summary on every rank
    import contextlib
    import io
    from oracle import scoring_oracle as SO
    rs = np.random.RandomState(21)
    NPAIR, DF = 5003, 640
    plab = (np.arange(NPAIR) % 10 == 0).astype(np.int64)
    ide = rs.randn(NPAIR, DF).astype(np.float32)
    p1 = ide + 4.0 * rs.randn(NPAIR, DF).astype(np.float32)
    p2 = np.where(plab[:, None] == 1, ide, rs.randn(NPAIR, DF).astype(np.float32)) + 4.0 * rs.randn(NPAIR, DF).astype(np.float32)
    cuts = np.linspace(0, NPAIR, world + 1).astype(int)
    cuts[1:-1] += 3                                         # uneven shards
    mine = slice(cuts[rank], cuts[rank + 1])
    with contextlib.redirect_stdout(io.StringIO()):
        out = tdist.score_pairs_sharded([(torch.from_numpy(p1[mine]).to(dev), torch.from_numpy(p2[mine]).to(dev),
                                          torch.from_numpy(plab[mine]).to(dev))])
    from text_guided_face_recognition_b200 import ops
    all_scores = ops.pair_cosine(torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)).cpu().numpy()
    want = SO.calculate_scores(all_scores, plab)
    assert out["auc"] == want["auc"] and out["eer"] == want["eer"] and out["tpr_at_fpr"] == want["tpr_at_fpr"], (out, want)
    report["verif.auc"] = out["auc"]

    dist.barrier()
    if rank == 0:
        print("sharded-nccl ok world=%d " % world + " ".join(f"{k}={v:.1e}" for k, v in report.items()), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
