"""World-size-2 tests of the N>1 plumbing on CPU (gloo): the differentiable all-gather, the merge of
per-rank column statistics that closes the column-wise cross entropy of a row-sharded score matrix, the
class partition of the sharded head and the (max, sum-exp, target) all-reduce of its softmax.

The arithmetic on each rank's block is done here with numpy (the oracle) because the product kernels
are CUDA only; what is under test is the collective protocol in
text_guided_face_recognition_b200/distributed.py, which is device agnostic."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, port, fn_name, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        globals()[fn_name](rank)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _run(fn_name, tmp_path):
    mp.spawn(_worker, args=(_free_port(), fn_name, str(tmp_path)), nprocs=WORLD, join=True)
    for r in range(WORLD):
        assert os.path.exists(os.path.join(str(tmp_path), f"ok{r}"))


# ------------------------------------------------------------------------------------------------
def _case_all_gather_rows(rank):
    from text_guided_face_recognition_b200 import distributed as D
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(3, 4, generator=g) for _ in range(WORLD)]
    ws = [torch.randn(WORLD * 3, 4, generator=g) for _ in range(WORLD)]
    x = xs[rank].clone().requires_grad_(True)
    y = D.all_gather_rows(x)
    assert torch.equal(y.detach(), torch.cat(xs))
    (y * ws[rank]).sum().backward()
    # backward = reduce-scatter: every rank's weight block for MY rows, summed
    want = sum(w[rank * 3:(rank + 1) * 3] for w in ws)
    assert torch.allclose(x.grad, want, atol=1e-6)
    # no-grad input (labels / caption lengths): plain gather, integer dtype preserved
    ids = torch.arange(3, dtype=torch.int64) + 10 * rank
    got = D.all_gather_rows(ids)
    assert got.dtype == torch.int64 and got.tolist() == [0, 1, 2, 10, 11, 12]


def _case_column_stats_and_pair_ce(rank):
    from oracle import fcam_oracle as O
    from text_guided_face_recognition_b200 import distributed as D
    rs = np.random.RandomState(3)
    Bl = 5
    B = Bl * WORLD
    scores = (rs.randn(B, B) * 6).astype(np.float64)
    scores[1, 7] = -np.inf                       # a masked pair (sent_loss class collision)
    scores[7, 1] = -np.inf
    blk = scores[rank * Bl:(rank + 1) * Bl]      # this rank's row block [Bl, B]
    cmax = np.max(blk, axis=0)
    csum = np.sum(np.exp(blk - cmax), axis=0)
    # the exchange is ONE all-gather of every rank's [2, B] statistics (distributed.gather_stats); the merge
    # arithmetic is a CUDA kernel in the product (tgfr_merge_softmax_stats) and restated here in numpy
    g = D.gather_stats([torch.from_numpy(cmax), torch.from_numpy(csum)]).numpy()          # [WORLD, 2, B]
    assert g.shape == (WORLD, 2, B) and np.array_equal(g[rank, 0], cmax)
    gm = g[:, 0].max(axis=0)
    gmax = torch.from_numpy(gm)
    gsum = torch.from_numpy((g[:, 1] * np.exp(g[:, 0] - gm)).sum(axis=0))
    full_max = np.max(scores, axis=0)
    full_sum = np.sum(np.exp(scores - full_max), axis=0)
    assert np.allclose(gmax.numpy(), full_max) and np.allclose(gsum.numpy(), full_sum, rtol=1e-12)
    # the two losses: rows complete locally, columns through the merged statistics, partial sums all-reduced
    rowlse = np.log(np.sum(np.exp(blk - blk.max(1, keepdims=True)), axis=1)) + blk.max(1)
    diag = np.array([blk[b, rank * Bl + b] for b in range(Bl)])
    collse = gmax.numpy() + np.log(gsum.numpy())
    part = torch.tensor([np.sum(rowlse - diag) / B, np.sum(collse[rank * Bl:(rank + 1) * Bl] - diag) / B])
    dist.all_reduce(part)
    r0, r1 = O.pair_ce(scores)
    assert abs(part[0].item() - r0) < 1e-12 * abs(r0) and abs(part[1].item() - r1) < 1e-12 * abs(r1)


def _case_class_sharded_softmax(rank):
    from text_guided_face_recognition_b200 import distributed as D
    C, B = 10177, 6
    ranges = [D.class_range(C, 8, r) for r in range(8)]
    assert ranges[0][0] == 0 and ranges[-1][1] == C
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1
    rs = np.random.RandomState(9)
    Cs = 23
    logits = (rs.randn(B, Cs) * 8).astype(np.float64)
    labels = rs.randint(0, Cs, size=B)
    c0, c1 = D.class_range(Cs, WORLD, rank)
    shard = logits[:, c0:c1]
    rowmax = torch.from_numpy(shard.max(1))
    rowsum = torch.from_numpy(np.exp(shard - shard.max(1, keepdims=True)).sum(1))
    tgt = torch.tensor([shard[b, labels[b] - c0] if c0 <= labels[b] < c1 else 0.0 for b in range(B)])
    # the exchange the class-sharded head performs: ONE all-gather of (max, sum-exp, target) per rank
    g = D.gather_stats([rowmax, rowsum, tgt.double()])                                    # [WORLD, 3, B]
    gmax = g[:, 0].max(dim=0).values
    gsum = (g[:, 1] * torch.exp(g[:, 0] - gmax)).sum(dim=0)
    ce = (gmax + torch.log(gsum) - g[:, 2].sum(dim=0)).mean().item()
    m = logits.max(1, keepdims=True)
    want = np.mean(m[:, 0] + np.log(np.exp(logits - m).sum(1)) - logits[np.arange(B), labels])
    assert abs(ce - want) < 1e-12 * abs(want)


def _case_allreduce_gradients(rank):
    """Bucketed all-reduce of replica gradients: mixed dtypes, a parameter without grad, several buckets."""
    from text_guided_face_recognition_b200 import distributed as D
    gen = torch.Generator().manual_seed(11)
    shapes = [(256, 256), (256,), (17, 3), (640, 256), (5,)]
    base = [torch.randn(*s_, generator=gen, dtype=torch.float64) for s_ in shapes]
    params = []
    for k, t in enumerate(base):
        p = torch.nn.Parameter(torch.zeros_like(t, dtype=torch.float32 if k != 2 else torch.float64))
        p.grad = (t * (rank + 1)).to(p.dtype)             # rank r contributes (r + 1) * base
        params.append(p)
    # a parameter whose gradient exists on rank 1 only (an unused branch on rank 0): rank 0 contributes zeros and
    # receives the sum -- the flat buffers have one layout on every rank (the old code packed
    # `[p.grad for p in params if p.grad is not None]` per rank and hung / mis-added here)
    lone = torch.nn.Parameter(torch.zeros(3))
    if rank == 1:
        lone.grad = torch.tensor([1.0, 2.0, 3.0])
    params.insert(1, lone)
    frozen = torch.nn.Parameter(torch.zeros(4), requires_grad=False)      # frozen: never touched
    params.append(frozen)
    nb = D.allreduce_gradients(params, bucket_bytes=300_000)
    assert nb >= 3                                        # two dtypes, and the fp32 gradients exceed one bucket
    for p, t in zip([q for q in params if q is not lone and q is not frozen], base):
        assert torch.allclose(p.grad.double(), 3.0 * t, rtol=1e-6, atol=1e-6)
    assert torch.equal(lone.grad, torch.tensor([1.0, 2.0, 3.0])) and frozen.grad is None
    D.allreduce_gradients(params, average=True)
    for p, t in zip([q for q in params if q is not lone and q is not frozen], base):
        assert torch.allclose(p.grad.double(), 3.0 * t, rtol=1e-6, atol=1e-6)   # mean of two equal tensors
    # ranks that disagree on the parameter set raise instead of reducing mismatched buffers
    extra = params + ([torch.nn.Parameter(torch.zeros(2))] if rank == 0 else [])
    try:
        D.allreduce_gradients(extra)
        raise AssertionError("membership mismatch not detected")
    except RuntimeError as e:
        assert "disagree" in str(e)


@pytest.mark.parametrize("case", ["_case_all_gather_rows", "_case_column_stats_and_pair_ce",
                                  "_case_class_sharded_softmax", "_case_allreduce_gradients"])
def test_gloo_world2(case, tmp_path):
    _run(case, tmp_path)


# ------------------------------------------------------------------------------------------------
def _case_gather_ragged(rank):
    """Pair-sharded verification scoring: ranks hold different numbers of scores; the gather keeps rank order,
    and the ROC of the gathered list equals the ROC of the unsplit list (oracle arithmetic on CPU)."""
    from oracle import scoring_oracle as SO
    from text_guided_face_recognition_b200 import distributed as D
    rs = np.random.RandomState(3)
    labels = (rs.rand(101) < 0.3).astype(np.int64)
    scores = (np.round(rs.randn(101) * 4) / 8 + 0.3 * labels).astype(np.float32)
    cut = [0, 37, 101]                                     # ragged shards
    mine = slice(cut[rank], cut[rank + 1])
    got_s = D.gather_ragged(torch.from_numpy(scores[mine]))
    got_l = D.gather_ragged(torch.from_numpy(labels[mine]))
    assert torch.equal(got_s, torch.from_numpy(scores)) and torch.equal(got_l, torch.from_numpy(labels))
    a, b = SO.roc_counts(got_l.numpy(), got_s.numpy()), SO.roc_counts(labels, scores)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    empty = D.gather_ragged(torch.zeros(0 if rank == 0 else 2))
    assert empty.numel() == 2


def test_gather_ragged_world2(tmp_path):
    _run("_case_gather_ragged", tmp_path)
