#!/usr/bin/env python
"""Benchmark of the FCAM hot path (BASELINE.json metric: caption-face pairs/sec of the word-region +
sentence loss fwd+bwd; margin-head samples/sec reported beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N = 1 : BASELINE configs[1] -- words_loss + sent_loss forward + backward (both gradients),
        B=128 captions x 128 faces, T=22 words (BERT flavour), R=14x14 regions, D=256, fp32 inputs.
N > 1 : launched by torchrun, one rank per GPU; every rank keeps 128 faces + 128 captions, the
        captions are all-gathered (NCCL) and each rank computes its [128, 128*N] block of the
        global score matrix ("weak": fixed local batch, global batch 128*N).
One JSON line is printed by rank 0.  `--impl reference` times the CPU port of the reference's
PyTorch implementation (oracle/ref_port.py) on the host cores instead.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import synth  # noqa: E402

B_LOCAL, T, IH, IW, D = 128, 22, 14, 14, 256
R = IH * IW
GAMMAS = (4.0, 5.0, 10.0)
HEAD = dict(B=512, Din=512, C=10177, s=30.0, m=0.5, gamma=2.0)
# fwd 4TRD (2 GEMMs) + bwd 8TRD (4 GEMMs) per pair with both gradients; 10TRD when only the
# face-side gradient is live, which is what the reference's training scripts run (text side detached,
# utils/dataset_utils.py:42-46).  Recomputation inside the backward kernel is NOT counted.
def flops_per_pair(grads):
    return (12 if grads == "both" else 10) * T * R * D
L2_BYTES = 126 * 2 ** 20


GRADS_DESC = {"both": "face- and text-side gradients",
              "ctx": "face-side gradient only, text detached as in the reference's training scripts"}


def make_args():
    ns = types.SimpleNamespace
    return ns(en_type="BERT", bert_words_num=T + 2, CUDA=True, device="cuda",
              TRAIN=ns(SMOOTH=ns(GAMMA1=GAMMAS[0], GAMMA2=GAMMAS[1], GAMMA3=GAMMAS[2])))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the dominant kernel, from the committed
    `ncu --set full` summary of this round (profiles/r2_final_ncu_raw.txt, same shapes as this bench)."""
    path = os.path.join(ROOT, "profiles", "r2_final_ncu_raw.txt")
    try:
        rd = wr = None
        hit = False
        for line in open(path):
            if line.startswith("== "):
                hit = kernel_substr in line
            elif hit and line.startswith("dram__bytes_read.sum ="):
                rd = float(line.split("=")[1])
            elif hit and line.startswith("dram__bytes_write.sum ="):
                wr = float(line.split("=")[1])
        if rd is None or wr is None:
            return None
        return (rd + wr) * 1e6                                   # ncu prints MB
    except OSError:
        return None


HEAD_STEP_KERNELS = ("norm_f16_pair", "gemm_tc_kernel", "ce_grad_from_cos", "normalize_bwd", "ce_merge_partials", "focal_finish")


def ncu_traffic_sum(kernel_substrs, fname="r2_head_step_ncu_raw.txt"):
    """Sum of dram read + write bytes over every launch of one head step whose kernel name contains one of the
    substrings, from the committed `ncu --set full` capture of tools/head_step.py (None when absent)."""
    path = os.path.join(ROOT, "profiles", fname)
    try:
        total, hit, seen = 0.0, False, False
        for line in open(path):
            if line.startswith("== "):
                hit = any(k in line for k in kernel_substrs)
            elif hit and (line.startswith("dram__bytes_read.sum =") or line.startswith("dram__bytes_write.sum =")):
                total += float(line.split("=")[1]) * 1e6
                seen = True
        return total if seen else None
    except OSError:
        return None


class ClockSampler:
    """SM clock / throttle-reason sampler running during the timed region.

    NVML is polled from a thread (a few ms period, cheap driver queries); an nvidia-smi subprocess is the
    fallback.  nvidia-smi must not be *started* inside a short timed region: its start-up holds driver locks
    for tens of milliseconds and stalls kernel launches."""
    REASONS = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
               ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
               ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))

    def __init__(self, index, period_s=0.004):
        import threading
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _poll(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append((sm, pw))
                for name, const in self.REASONS:
                    if mask & getattr(nv, const):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        import threading
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        if self.samples:
            power = [p for _, p in self.samples]
            thr = (max(power) + min(power)) / 2
            load = [s for s, p in self.samples if p >= thr] or [s for s, _ in self.samples]
            self.result = {"sm_mhz": statistics.median(load), "sm_max_mhz": self._max, "reasons": sorted(self.reasons),
                           "samples": len(self.samples), "source": "nvml"}
        else:
            self.result = self._smi_once()

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0]
            f = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]),
                    "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")], "samples": 1,
                    "source": "nvidia-smi after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


# ------------------------------------------------------------------------------------------------
# reference arm: CPU port of the reference implementation
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(ctx, words, img, txt, cid, n_caps=None, grads="both"):
    """One fwd+bwd of words_loss + sent_loss with the reference's op sequence on CPU.
    n_caps < B restricts the per-caption loop to the first n_caps captions (bounded sample)."""
    from oracle import ref_port as P
    B = ctx.shape[0]
    c = torch.from_numpy(ctx).clone().requires_grad_(True)
    w = torch.from_numpy(words).clone().requires_grad_(grads == "both")
    a = torch.from_numpy(img).clone().requires_grad_(True)
    b = torch.from_numpy(txt).clone().requires_grad_(True)
    labels = torch.arange(B)
    t0 = time.perf_counter()
    c_ref = c.view(B, IH, IW, D).permute(0, 3, 1, 2)
    w_ref = w.transpose(1, 2)
    if n_caps is None or n_caps >= B:
        l0, l1, _ = P.words_loss_port(c_ref, w_ref, labels, None, *GAMMAS)
        pairs = B * B
    else:
        # the reference's cost is linear in the caption loop: run n_caps iterations of it
        cols = []
        for i in range(n_caps):
            word = w_ref[i, :, :T].unsqueeze(0).contiguous().repeat(B, 1, 1)
            wctx, _ = P.attention_port(word, c_ref, GAMMAS[0])
            row = P._cos(word.transpose(1, 2).contiguous().view(B * T, -1),
                         wctx.transpose(1, 2).contiguous().view(B * T, -1)).view(B, T)
            cols.append(torch.log(row.mul(GAMMAS[1]).exp().sum(dim=1, keepdim=True)))
        sim = torch.cat(cols, 1) * GAMMAS[2]
        l0 = torch.nn.functional.cross_entropy(sim, torch.arange(B) % n_caps)
        l1 = torch.nn.functional.cross_entropy(sim.t(), torch.arange(n_caps))
        pairs = B * n_caps
    s0, s1 = P.sent_loss_port(a, b, labels, cid, GAMMAS[2])
    (l0 + l1 + s0 + s1).backward()
    return time.perf_counter() - t0, pairs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = B_LOCAL
    ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=100)
    img, txt, cid = synth.sentence_inputs(B, D, seed=100)
    # size the per-step sample so that the whole run stays within ~3 minutes
    t_probe, pairs = cpu_reference_step(ctx, words, img, txt, cid, n_caps=8, grads=args.grads)
    per_cap = t_probe / 8
    budget = 150.0 / max(1, args.steps + args.warmup)
    n_caps = int(max(4, min(B, budget / max(per_cap, 1e-6))))
    for _ in range(args.warmup):
        cpu_reference_step(ctx, words, img, txt, cid, n_caps, args.grads)
    times, pairs = [], 0
    for _ in range(args.steps):
        dt, pairs = cpu_reference_step(ctx, words, img, txt, cid, n_caps, args.grads)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = pairs / (ms * 1e-3)
    sample = (f"{n_caps} of {B} captions x {B} faces per step ({pairs} pairs), T={T}, R={R}, D={D}, fwd+bwd "
              f"({GRADS_DESC[args.grads]}), + sent_loss B={B}")
    line = {
        "impl": "reference", "metric": "fcam_words+sent_loss_fwd_bwd_pairs_per_sec", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: words_loss+sent_loss fwd+bwd ({GRADS_DESC[args.grads]}), {B} faces x {B} "
                               f"captions, T={T}, R={R}, D={D} (bounded sample)",
                   "global_batch": B, "sample": sample, "grads": args.grads},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# one-shot parity of the exact code path that is timed (oracle = checker only), printed as "parity" in the line
# ------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def parity_check(world, rank, dev, precision):
    """Sharded (N > 1: NCCL all-gather / statistic merge / reduce-scatter) or single-GPU losses against the fp64
    oracle on the gathered batch.  Word-region + sentence loss at the config-4 caption length (T = 30 -> Tp = 32),
    R = 196, D = 256, 6 faces/captions per rank; the class-sharded fused ArcFace head at config 3 (512 x 10 177,
    global B = 512, the class remainder 10 177 mod N included).  Rank 0 evaluates the oracle; gradients of every
    rank are gathered to it.  Returns the dict on rank 0 (None elsewhere); never raises on a mismatch -- the
    verdict is in "ok"."""
    import torch.distributed as dist
    from oracle import fcam_oracle as O
    from text_guided_face_recognition_b200 import _lib, ops
    from text_guided_face_recognition_b200 import distributed as tdist
    from text_guided_face_recognition_b200.models import metrics
    Bp, Tp_, Rp_, Dp_ = 6, 30, R, D
    Bg = Bp * world
    ctx, words, _ = synth.wordregion_inputs(Bg, Tp_, Rp_, Dp_, "BERT", seed=31)
    img, txt, cid = synth.sentence_inputs(Bg, Dp_, seed=31, collisions=True)
    sl = slice(rank * Bp, (rank + 1) * Bp)
    c = torch.from_numpy(ctx[sl]).to(dev).requires_grad_(True)
    w = torch.from_numpy(words[sl]).to(dev).requires_grad_(True)
    a = torch.from_numpy(img[sl]).to(dev).requires_grad_(True)
    b = torch.from_numpy(txt[sl]).to(dev).requires_grad_(True)
    ids = torch.from_numpy(cid[sl]).to(dev)
    if world > 1:
        l0, l1, _ = tdist.words_loss_sharded(c, w, None, *GAMMAS, precision=precision)
        s0, s1 = tdist.sent_loss_sharded(a, b, ids, GAMMAS[2])
    else:
        sim, _ = ops.wordregion_sim(c, w, None, *GAMMAS, 1e-8, precision, False, 0)
        l0, l1 = ops.pair_ce(sim)
        s0, s1 = ops.pair_ce(ops.cosine_scores(a, b, GAMMAS[2], True, 1e-8, ids, ids))
    (l0 + l1 + s0 + s1).backward()

    def gather(t):
        if world == 1:
            return t.detach().cpu().numpy()
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t.detach().contiguous())
        return out.cpu().numpy()
    g_c, g_w, g_a, g_b = gather(c.grad), gather(w.grad), gather(a.grad), gather(b.grad)

    # class-sharded fused head, config 3
    h = HEAD
    xn, wn, lab = synth.margin_inputs(h["B"], h["Din"], h["C"], seed=100)
    Bl = h["B"] // world
    hs = slice(rank * Bl, (rank + 1) * Bl)
    hx = torch.from_numpy(xn[hs]).to(dev).requires_grad_(True)
    hl = torch.from_numpy(lab[hs]).to(dev)
    if world > 1:
        head = tdist.ShardedArcMarginProduct(h["Din"], h["C"], s=h["s"], m=h["m"]).to(dev)
        head.load_full_weight(torch.from_numpy(wn).to(dev))
        hloss = head.loss(hx, hl, gamma=h["gamma"])
        c0, c1 = head.c0, head.c1
    else:
        head = metrics.ArcMarginProduct(h["Din"], h["C"], s=h["s"], m=h["m"]).to(dev)
        with torch.no_grad():
            head.weight.copy_(torch.from_numpy(wn))
        hloss = head.fused_loss(hx, hl, gamma=h["gamma"])
        c0, c1 = 0, h["C"]
    hloss.backward()
    g_hx = gather(hx.grad)
    rows = -(-h["C"] // world)
    pad = torch.zeros(rows, h["Din"], device=dev)
    pad[:c1 - c0] = head.weight.grad
    g_hw = gather(pad)
    bounds = gather(torch.tensor([[c0, c1]], device=dev, dtype=torch.int64))
    if rank != 0:
        return None
    ltol, gtol = (2e-5, 1e-4) if precision == _lib.PREC_FP32 else (1e-4, 1e-3)
    r0, r1, _, _ = O.words_loss(ctx, words, None, None, *GAMMAS)
    q0, q1, _ = O.sent_loss(img, txt, None, cid, GAMMAS[2])
    dctx, dwords = O.words_loss_grads(ctx, words, None, None, *GAMMAS)
    dimg, dtxt = O.sent_loss_grads(img, txt, None, cid, GAMMAS[2])
    ref_logits = O.arc_margin(xn, wn, lab, h["s"], h["m"], False)
    ref_hloss = O.focal_loss(ref_logits, lab, h["gamma"])
    dx_ref, dw_ref = O.arc_margin_bwd(xn, wn, lab, O.focal_loss_bwd(ref_logits, lab, h["gamma"]), h["s"], h["m"], False)
    dw_got = np.concatenate([g_hw[r * rows:r * rows + int(bounds[r, 1] - bounds[r, 0])] for r in range(world)])
    loss_err = {"words_loss0": abs(l0.item() - r0) / abs(r0), "words_loss1": abs(l1.item() - r1) / abs(r1),
                "sent_loss0": abs(s0.item() - q0) / abs(q0), "sent_loss1": abs(s1.item() - q1) / abs(q1),
                "head_loss": abs(hloss.item() - ref_hloss) / abs(ref_hloss)}
    grad_err = {"d_ctx": _rel(g_c, dctx), "d_words": _rel(g_w, dwords), "d_img": _rel(g_a, dimg), "d_txt": _rel(g_b, dtxt),
                "head_dx": _rel(g_hx, dx_ref), "head_dw": _rel(dw_got, dw_ref)}
    ok = all(v < ltol for v in loss_err.values()) and all(v < gtol for v in grad_err.values())
    return {"ok": bool(ok), "world": world, "checker": "oracle/fcam_oracle.py (fp64) on the gathered batch, rank 0",
            "workload": f"words_loss+sent_loss: {Bp} faces/captions per rank (global {Bg}), T={Tp_}, R={Rp_}, D={Dp_}, "
                        f"class-id collisions; class-sharded fused ArcFace+focal head {h['B']} x {h['C']} x {h['Din']} "
                        f"(classes per rank {[int(bounds[r, 1] - bounds[r, 0]) for r in range(world)]})",
            "loss_rel_err": loss_err, "grad_rel_err": grad_err, "tol": {"loss": ltol, "grad": gtol}}


def time_step(fn, flush, reps, use_graph, world=1, dev=None):
    """ms per call of `fn` (CUDA-graph replay when use_graph), L2 flushed before every call, CUDA events on the launch
    stream, max over ranks."""
    import torch.distributed as dist
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    run = fn
    if use_graph:
        from text_guided_face_recognition_b200.graphs import GraphedStep
        run = GraphedStep(fn)
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        flush.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, run


def head_roofline(ms, pk, traffic=None):
    """Tensor and HBM fractions of one fused margin-head step (6 Din C flops per sample; fp32 bytes:
    W read twice + dW written once + X / dX)."""
    h = HEAD
    flops = 6 * h["Din"] * h["C"] * h["B"]
    nbytes = 3 * 4 * h["C"] * h["Din"] + 2 * 4 * h["B"] * h["Din"]
    tf, gbs = flops / (ms * 1e-3) / 1e12, nbytes / (ms * 1e-3) / 1e9
    ft, fh = tf / pk["tf_burst"], gbs / pk["hbm"]
    return {"bound": "tensor" if ft >= fh else "hbm", "achieved": tf if ft >= fh else gbs,
            "peak": pk["tf_burst"] if ft >= fh else pk["hbm"], "unit": "TFLOP/s" if ft >= fh else "GB/s",
            "frac": max(ft, fh), "traffic": traffic,
            "tensor": {"achieved": tf, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": ft,
                       "algorithmic_flops_per_step": flops},
            "hbm": {"achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": fh, "algorithmic_bytes_per_step": nbytes},
            "note": "the whole step (all launches of fwd+bwd), not one kernel: the head is latency/launch bound at "
                    "this size, so both fractions are reported"}


def sharded_head_leg(world, rank, dev, use_graph, flush, reps):
    """BASELINE configs[2] on N ranks: global B = 512 samples (512 / N per rank, all-gathered), 10 177 classes split by
    class (partial FC), fused margin + softmax statistics in the GEMM epilogues, ONE all-gather of the (max, sum-exp,
    target) rows, dX reduce-scattered.  Strong scaling: the total work is fixed."""
    from text_guided_face_recognition_b200 import distributed as tdist
    h = HEAD
    xn, wn, lab = synth.margin_inputs(h["B"], h["Din"], h["C"], seed=100)
    Bl = h["B"] // world
    hs = slice(rank * Bl, (rank + 1) * Bl)
    head = tdist.ShardedArcMarginProduct(h["Din"], h["C"], s=h["s"], m=h["m"]).to(dev)
    head.load_full_weight(torch.from_numpy(wn).to(dev))
    x = torch.from_numpy(xn[hs]).to(dev).requires_grad_(True)
    labt = torch.from_numpy(lab[hs]).to(dev)

    def step():
        x.grad = None
        head.weight.grad = None
        head.loss(x, labt, gamma=h["gamma"]).backward()
    ms, _ = time_step(step, flush, reps, use_graph, world, dev)
    return {"metric": "arcface_focal_fwd_bwd_samples_per_sec", "value": h["B"] / (ms * 1e-3), "unit": "samples/s",
            "ms_per_step": ms, "scaling": "strong", "n_gpus": world,
            "config": dict(h, classes_per_rank=head.c1 - head.c0, samples_per_rank=Bl),
            "collectives": "all-gather X [512,512] + labels, all-gather of [3,512] softmax statistics, "
                           "reduce-scatter dX [512,512]",
            "launch": "CUDA graph replay (NCCL captured)" if use_graph else "eager"}


def config4_leg(world, rank, dev, use_graph, flush, reps):
    """BASELINE configs[3] ("Full FCAM train step ... global B = 1024 over 8 x B200"): fcam.FcamTrainStep -- TextHeading
    (no_grad), ImageHeading (IMIM + ProjectionHead, trainable), words / sent / global losses row-sharded with the captions
    all-gathered, two class-sharded ArcFace + focal heads over 10 177 identities, backward, all-reduce of the image
    head's gradients.  128 faces + captions per rank, 32 BERT tokens -> T = 30 words.  The frozen encoders (IResNet-50,
    BERT) are out of scope: their outputs are the step's (synthetic) inputs."""
    import types as _t
    from text_guided_face_recognition_b200.fcam import FcamTrainStep
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_imim_r2 import imim_inputs
    Bl, bwn, C = 128, 32, HEAD["C"]
    ns = _t.SimpleNamespace
    a4 = ns(aux_feat_dim_per_granularity=256, bert_words_num=bwn, en_type="BERT",
            TRAIN=ns(SMOOTH=ns(GAMMA1=GAMMAS[0], GAMMA2=GAMMAS[1], GAMMA3=GAMMAS[2])))
    step = FcamTrainStep(a4, C, dev)
    x, _, xg, _, _, _ = imim_inputs(Bl, 100 + rank)
    tok, _, _ = synth.texthead_inputs(Bl, bwn, 256, seed=100 + rank)
    gfeat, lfeat = torch.from_numpy(xg).to(dev), torch.from_numpy(x).to(dev)
    tokens = torch.from_numpy(tok).to(dev)
    cid = (torch.arange(Bl, device=dev) + rank * Bl) % C

    def run():
        step.zero_grad()
        return step(gfeat, lfeat, tokens, cid)
    ms, _ = time_step(run, flush, reps, use_graph, world, dev)
    Bg = Bl * world
    T4 = bwn - 2
    wr_flops = 10 * T4 * R * D * Bg * Bg                       # word-region loss, face-side gradient
    imim_flops = 3 * 2 * 83.9e6 * Bg                           # IMIM fwd + bwd: 83.9 M MACs per sample forward
    head_flops = 2 * 6 * 256 * C * Bg
    return {"metric": "fcam_full_step_pairs_per_sec", "value": Bg * Bg / (ms * 1e-3), "unit": "pairs/s",
            "samples_per_sec": Bg / (ms * 1e-3), "ms_per_step": ms, "n_gpus": world, "scaling": "weak",
            "config": {"workload": "configs[3]: TextHeading (no_grad) + ImageHeading + words/sent/global losses + two "
                                   "ArcFace+focal heads + backward + image-head gradient all-reduce; frozen IResNet-50 / "
                                   "BERT excluded (their outputs are the inputs)",
                       "global_batch": Bg, "local_batch": Bl, "T": T4, "R": R, "D": D, "classes": C,
                       "launch": "one CUDA graph replay per step" if use_graph else "eager"},
            "algorithmic_tflops": (wr_flops + imim_flops + head_flops) / (ms * 1e-3) / 1e12,
            "flops_breakdown": {"wordregion": wr_flops, "imim": imim_flops, "heads": head_flops}}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from text_guided_face_recognition_b200 import _lib, fcam, ops
    from text_guided_face_recognition_b200 import distributed as tdist
    from text_guided_face_recognition_b200.models import losses, metrics

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    margs = make_args()
    B = B_LOCAL
    Bg = B * world
    precision = ops.default_precision()

    # rotating input sets so that every step reads inputs that are not L2 resident
    bytes_per_set = 4 * (B * R * D + B * T * D + 2 * B * D)
    n_sets = max(2, int(np.ceil(2.0 * L2_BYTES / bytes_per_set)))
    sets, host_sets = [], []
    for k in range(n_sets):
        ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=100 + 17 * k + 1000 * rank)
        img, txt, cid = synth.sentence_inputs(B, D, seed=100 + 17 * k + 1000 * rank)
        if k < 2:
            host_sets.append(tuple(torch.from_numpy(a).pin_memory() for a in (ctx, words, img, txt)))
        sets.append(tuple(torch.from_numpy(a).to(dev).requires_grad_(j != 1 or args.grads == "both")
                          for j, a in enumerate((ctx, words, img, txt))))
    labels = torch.arange(B, device=dev)
    cid_dev = torch.arange(B, device=dev) + rank * B       # distinct classes: mask path runs, nothing masked

    # one-shot parity of the path about to be timed (all ranks take part; the verdict goes into the JSON line)
    try:
        parity = parity_check(world, rank, dev, precision)
    except Exception as e:                                 # a crash of the checker must not lose the measurement
        parity = {"ok": False, "error": f"{type(e).__name__}: {e}"[:300]} if rank == 0 else None
    barrier_early = (lambda: (dist.barrier(), torch.cuda.synchronize())) if world > 1 else torch.cuda.synchronize
    barrier_early()

    def step(c, w, a, b):
        for t in (c, w, a, b):
            t.grad = None
        c_ref = c.view(B, IH, IW, D).permute(0, 3, 1, 2)
        # the sentence-loss chain (tiny kernels) runs on a side stream beside the word-region kernels (fcam.py)
        if world == 1:
            l0, l1, _, s0, s1 = fcam.fcam_losses(c_ref, w.transpose(1, 2), a, b, labels, None, cid_dev, B, margs)
        else:
            (s0, s1), (l0, l1, _) = fcam.run_overlapped(
                lambda: tdist.sent_loss_sharded(a, b, cid_dev, GAMMAS[2]),
                lambda: tdist.words_loss_sharded(c, w, None, *GAMMAS, precision=precision), device=dev)
        total = l0 + l1 + s0 + s1
        total.backward()
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # One CUDA graph per input set (single GPU): the eager Python / autograd launch path (~0.9 ms per step) is
    # slower than the device work (~0.75 ms), so the step is captured once per rotating input set and replayed.
    # Multi-GPU steps contain NCCL collectives; NCCL launches are capturable, so they are graphed the same way.
    # All graphs share one memory pool (they are replayed one at a time).
    use_graph = not args.no_graph
    runners = None

    def build_runners():
        from text_guided_face_recognition_b200.graphs import GraphedStep
        pool = torch.cuda.graph_pool_handle()                 # a fresh pool per build: it dies with its graphs
        out = []
        for k in range(n_sets):
            st_ = sets[k]
            out.append(GraphedStep(lambda st_=st_: step(*st_), pool=pool))
        return out

    def run_step(k):
        return runners[k % n_sets]() if runners is not None else step(*sets[k % n_sets])

    ops.launch_counter.n = 0
    step(*sets[0])
    launches_per_step = ops.launch_counter.n                # kernels one step issues (captured or eager alike)
    if use_graph:
        runners = build_runners()
    for k in range(args.warmup):
        run_step(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for k in range(args.steps):
            run_step(args.warmup + k)
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms = ms_total / args.steps
    pairs = B * Bg * world                         # all ranks' [B, Bg] blocks = Bg^2
    value = pairs / (ms * 1e-3)

    # ---- end to end through the public API with HOST inputs (pinned) + D2H of the loss, every step.
    # The step's inputs are copied host -> device on a copy stream into the alternate device buffer set while the
    # previous step computes (ordinary input prefetching; every copy and every loss read is inside the timed region).
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_copy(k):
        slot = k % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])            # the step that last used this buffer set is done
            with torch.no_grad():
                for h, d_ in zip(host_sets[slot], sets[slot]):
                    d_.copy_(h, non_blocking=True)
            copied[slot].record(copy_stream)

    def e2e_run(n):
        for ev in consumed:
            ev.record(main_stream)
        issue_copy(0)
        for k in range(n):
            if k + 1 < n:
                issue_copy(k + 1)
            slot = k % 2
            main_stream.wait_event(copied[slot])
            total = run_step(slot)
            consumed[slot].record(main_stream)
            loss_host.copy_(total.detach().reshape(1), non_blocking=True)
            main_stream.synchronize()                          # the caller reads the loss every step
            _ = float(loss_host[0])

    e2e_run(3)
    barrier()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_ms /= args.steps
    e2e = {"value": pairs / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": bytes_per_set,
           "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
           "note": "H2D of step k+1 on a copy stream overlaps step k; loss read back and synchronised every step"}

    # ---- the other gradient configuration beside the headline one (short loop, same timing rules)
    other = "both" if args.grads == "ctx" else "ctx"
    for st_ in sets:
        st_[1].requires_grad_(other == "both")
    n_other = max(3, min(args.steps, 50))
    if use_graph:
        runners = None
        torch.cuda.synchronize()
        runners = build_runners()
    for k in range(n_sets + 3):                               # every input set once: gradient buffers get allocated
        run_step(k)
    barrier()
    e0.record()
    for k in range(n_other):
        run_step(3 + k)
    e1.record()
    barrier()
    other_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([other_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        other_ms = float(t.item())
    other_ms /= n_other
    for st_ in sets:
        st_[1].requires_grad_(args.grads == "both")

    # ---- class-sharded margin head on all ranks (N > 1): the head's own curve beside the contrastive step;
    # ---- the assembled configs[3] step on every rank
    sharded_head = None
    flush_all = torch.empty(L2_BYTES * 2, dtype=torch.uint8, device=dev)
    if world > 1:
        sharded_head = sharded_head_leg(world, rank, dev, use_graph, flush_all, max(3, min(args.steps, 20)))
    try:
        config4 = config4_leg(world, rank, dev, use_graph, flush_all, max(3, min(args.steps, 10)))
    except Exception as e:                                  # never lose the headline line to a side leg
        config4 = {"error": f"{type(e).__name__}: {e}"[:300]}
    del flush_all

    line = {
        "metric": "fcam_words+sent_loss_fwd_bwd_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if precision == _lib.PREC_FP32 else "f16-operand/f32-accumulate",
        "data": "synthetic",
        "config": {"workload": f"configs[1]: words_loss+sent_loss fwd+bwd ({GRADS_DESC[args.grads]}), {B} faces x {Bg} "
                               f"captions per rank, T={T}, R={R}, D={D}",
                   "global_batch": Bg, "local_batch": B, "grads": args.grads, "parallelism": f"row-sharded x{world}",
                   "precision": "fp32-simt" if precision == _lib.PREC_FP32 else "tcgen05",
                   "l2": f"rotating {n_sets} input sets ({n_sets * bytes_per_set >> 20} MiB > 126 MiB L2)",
                   "launch": "one CUDA graph replay per step (graphs.GraphedStep)" if use_graph else "eager"},
        "clocks": clk.result, "e2e": e2e, "gpu_launches": launches,
        "other_grads": {"grads": other, "value": pairs / (other_ms * 1e-3), "unit": "pairs/s", "ms_per_step": other_ms,
                        "steps": n_other},
        "parity": parity,
        "full_step": config4,
    }
    if sharded_head is not None:
        line["margin_head_sharded"] = sharded_head

    if rank == 0:
        # ---- dominant kernel: word-region backward, timed alone with CUDA events on the launch stream
        pk = peaks()
        c, w = sets[0][0].detach(), sets[1][1].detach()
        wall = torch.cat([w] * world) if world > 1 else w
        gsim = torch.randn(B, Bg, device=dev) / Bg
        dctx = torch.empty(B, R, D, device=dev)
        dwords = torch.empty(Bg, T, D, device=dev) if args.grads == "both" else None
        lib = _lib.load()
        wsb = lib.tgfr_wordregion_workspace_bytes(B, Bg, T, R, D, precision)
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        # the step's backward reads the forward's saved fp16 attention image: produce it with one forward call
        svb = lib.tgfr_wordregion_saved_bytes(B, Bg, T, R, D, precision) if ops._save_enabled() else 0
        saved = torch.empty(max(svb, 1), dtype=torch.uint8, device=dev)
        sim_tmp = torch.empty(B, Bg, device=dev)
        _lib.check(lib.tgfr_wordregion_fwd(c.data_ptr(), *c.stride(), wall.data_ptr(), *wall.stride(), 0, B, Bg, T, R, D,
                                           *GAMMAS, 1e-8, sim_tmp.data_ptr(), 0, 0, precision, ws.data_ptr(), wsb,
                                           saved.data_ptr() if svb else 0, svb, st), "fwd")

        def bwd_call():
            _lib.check(lib.tgfr_wordregion_bwd(c.data_ptr(), *c.stride(), wall.data_ptr(), *wall.stride(), 0, B, Bg,
                                               T, R, D, *GAMMAS, 1e-8, gsim.data_ptr(), dctx.data_ptr(),
                                               _lib.ptr(dwords), precision, ws.data_ptr(), wsb,
                                               saved.data_ptr() if svb else 0, svb, st), "bwd")
        def fwd_call():
            _lib.check(lib.tgfr_wordregion_fwd(c.data_ptr(), *c.stride(), wall.data_ptr(), *wall.stride(), 0, B, Bg, T, R, D,
                                               *GAMMAS, 1e-8, sim_tmp.data_ptr(), 0, 0, precision, ws.data_ptr(), wsb,
                                               saved.data_ptr() if svb else 0, svb, st), "fwd")
        flush = torch.empty(L2_BYTES * 2, dtype=torch.uint8, device=dev)

        def time_alone(call):
            ks = []
            for k in range(3 + max(3, min(args.steps, 10))):
                flush.zero_()
                e0.record()
                call()
                e1.record()
                torch.cuda.synchronize()
                if k >= 3:
                    ks.append(e0.elapsed_time(e1))
            return sum(ks) / len(ks)

        # the two word-region launches, each timed alone (L2 flushed); `roofline` describes the slower of the two.
        # Algorithmic flops: forward 4 T R D per pair, backward 6 (d ctx) or 8 (both gradients).
        f_ms, b_ms = time_alone(fwd_call), time_alone(bwd_call)
        f_flops = 4 * T * R * D * B * Bg
        b_flops = (8 if args.grads == "both" else 6) * T * R * D * B * Bg
        tc_mode = world == 1 and args.grads == "ctx" and precision == _lib.PREC_TC
        kernels = {
            "wordregion_fwd": {"kernel_ms": f_ms, "algorithmic_flops_per_launch": f_flops,
                               "achieved": f_flops / (f_ms * 1e-3) / 1e12,
                               "traffic": ncu_traffic("wr_tc_fwd3_kernel") if tc_mode else None},
            "wordregion_bwd": {"kernel_ms": b_ms, "algorithmic_flops_per_launch": b_flops,
                               "achieved": b_flops / (b_ms * 1e-3) / 1e12,
                               "traffic": ncu_traffic("wr_tc_bwd2_kernel") if tc_mode else None},
        }
        dom = max(kernels, key=lambda k: kernels[k]["kernel_ms"])
        for v in kernels.values():
            v["frac"] = v["achieved"] / pk["tf_burst"]
        line["roofline"] = {"bound": "tensor", "achieved": kernels[dom]["achieved"], "peak": pk["tf_burst"],
                            "unit": "TFLOP/s", "frac": kernels[dom]["frac"], "traffic": kernels[dom]["traffic"],
                            "traffic_source": "profiles/r2_final_ncu_raw.txt (ncu --set full, same shapes)",
                            "kernel": dom, "kernel_ms": kernels[dom]["kernel_ms"],
                            "algorithmic_flops_per_launch": kernels[dom]["algorithmic_flops_per_launch"],
                            "peak_source": pk["src"] + " bf16 burst (kernel timed alone)",
                            "kernels": kernels,
                            "step_tflops": flops_per_pair(args.grads) * B * Bg / (ms * 1e-3) / 1e12,
                            "step_frac_of_peak": flops_per_pair(args.grads) * B * Bg / (ms * 1e-3) / 1e12 / pk["tf_sus"]}

    if rank == 0 and world > 1:
        # N > 1: the margin head of this line IS the class-sharded one (all ranks timed it above); whole-job peaks
        pk = peaks()
        sharded_head["roofline"] = head_roofline(sharded_head["ms_per_step"],
                                                 dict(pk, tf_burst=pk["tf_burst"] * world, hbm=pk["hbm"] * world))
        line["margin_head"] = sharded_head
        print(json.dumps(line), flush=True)
    if rank == 0 and world == 1:
        # ---- margin head (BASELINE configs[2]) reported beside the headline metric
        h = HEAD
        xn, wn, lab = synth.margin_inputs(h["B"], h["Din"], h["C"], seed=100)
        head = metrics.ArcMarginProduct(h["Din"], h["C"], s=h["s"], m=h["m"]).to(dev)
        with torch.no_grad():
            head.weight.copy_(torch.from_numpy(wn))
        x = torch.from_numpy(xn).to(dev).requires_grad_(True)
        labt = torch.from_numpy(lab).to(dev)
        crit = losses.FocalLoss(gamma=h["gamma"])

        def head_step():
            x.grad = None
            head.weight.grad = None
            crit(head(x, labt), labt).backward()
        for _ in range(3):
            head_step()
        torch.cuda.synchronize()
        head_run = head_step
        if use_graph:
            from text_guided_face_recognition_b200.graphs import GraphedStep
            head_run = GraphedStep(head_step)
        hs = []
        for _ in range(max(3, min(args.steps, 10))):
            flush.zero_()
            e0.record()
            head_run()
            e1.record()
            torch.cuda.synchronize()
            hs.append(e0.elapsed_time(e1))
        h_ms = sum(hs) / len(hs)
        hflops = 6 * h["Din"] * h["C"] * h["B"]
        line["margin_head"] = {"metric": "arcface_focal_fwd_bwd_samples_per_sec", "value": h["B"] / (h_ms * 1e-3),
                               "unit": "samples/s", "ms_per_step": h_ms, "config": h,
                               "precision": os.environ.get("TGFR_HEAD_PRECISION", "tc"),
                               "launch": "CUDA graph replay" if use_graph else "eager",
                               "tflops": hflops / (h_ms * 1e-3) / 1e12,
                               "frac_of_tensor_peak": hflops / (h_ms * 1e-3) / 1e12 / pk["tf_burst"]}

        # the same head step through ArcMarginProduct.fused_loss (no [B, C] logits: an extension of the reference API)
        def fused_step():
            x.grad = None
            head.weight.grad = None
            head.fused_loss(x, labt, gamma=h["gamma"]).backward()
        for _ in range(3):
            fused_step()
        torch.cuda.synchronize()
        fused_run = fused_step
        if use_graph:
            from text_guided_face_recognition_b200.graphs import GraphedStep
            fused_run = GraphedStep(fused_step)
        hs = []
        for _ in range(max(3, min(args.steps, 10))):
            flush.zero_()
            e0.record()
            fused_run()
            e1.record()
            torch.cuda.synchronize()
            hs.append(e0.elapsed_time(e1))
        hf_ms = sum(hs) / len(hs)
        line["margin_head"]["fused_loss"] = {"value": h["B"] / (hf_ms * 1e-3), "unit": "samples/s", "ms_per_step": hf_ms,
                                             "tflops": hflops / (hf_ms * 1e-3) / 1e12,
                                             "frac_of_tensor_peak": hflops / (hf_ms * 1e-3) / 1e12 / pk["tf_burst"],
                                             "what": "ArcMarginProduct.fused_loss: margin + online softmax + softmax "
                                                     "gradient in the tcgen05 GEMM epilogues, logits never written"}
        line["margin_head"]["roofline"] = head_roofline(hf_ms, pk, ncu_traffic_sum(HEAD_STEP_KERNELS))
        line["margin_head"]["roofline"]["step"] = "fused_loss"

        # ---- MagFace head at the same size: MagLinear(512, 10177, scale=64) + MagLoss (two dense [B,C] logit tensors
        # are part of the reference API, models/magface.py:69-136), fwd + bwd of loss + 35 loss_g
        from text_guided_face_recognition_b200.models import magface
        mx_np, mw_np, mlab_np = synth.margin_inputs(h["B"], h["Din"], h["C"], seed=100, mag=True)
        mhead = magface.MagLinear(h["Din"], h["C"], scale=64.0, easy_margin=True).to(dev)
        with torch.no_grad():
            mhead.weight.copy_(torch.from_numpy(mw_np))
        mcrit = magface.MagLoss(10.0, 110.0, 0.45, 0.8, 64.0)
        mx = torch.from_numpy(mx_np * 4.0).to(dev).requires_grad_(True)
        mlab = torch.from_numpy(mlab_np).to(dev)

        def mag_step():
            mx.grad = None
            mhead.weight.grad = None
            lg, xnorm = mhead(mx, lambda v: (0.8 - 0.45) / (110.0 - 10.0) * (v - 10.0) + 0.45, 10.0, 110.0)
            ml, mg, _ = mcrit(lg, mlab, xnorm)
            (ml + 35.0 * mg).backward()
        mag_ms, _ = time_step(mag_step, flush, max(3, min(args.steps, 10)), use_graph)
        mag_bytes = 3 * 4 * h["C"] * h["Din"] + 2 * 4 * h["B"] * h["Din"] + 7 * 4 * h["B"] * h["C"]
        line["mag_head"] = {"metric": "magface_fwd_bwd_samples_per_sec", "value": h["B"] / (mag_ms * 1e-3),
                            "unit": "samples/s", "ms_per_step": mag_ms,
                            "config": {"B": h["B"], "Din": h["Din"], "C": h["C"], "scale": 64.0, "l_a": 10, "u_a": 110,
                                       "l_margin": 0.45, "u_margin": 0.8, "easy_margin": True},
                            "tflops": hflops / (mag_ms * 1e-3) / 1e12,
                            "roofline": {"bound": "hbm", "achieved": mag_bytes / (mag_ms * 1e-3) / 1e9, "peak": pk["hbm"],
                                         "unit": "GB/s", "frac": mag_bytes / (mag_ms * 1e-3) / 1e9 / pk["hbm"],
                                         "algorithmic_bytes_per_step": mag_bytes,
                                         "note": "dense API: cos / cos_m / one_hot written (3), cos + cos_m read by the CE "
                                                 "(2), two dense gradients written (2) = 7 x 4BC bytes on top of W / dW / X"}}
        del mhead, mx

        # ---- image head local branch (IMIM) at the configs[1] batch: fwd + bwd, training mode
        import types as _types2
        from text_guided_face_recognition_b200.models.image_heading import ImageHeading
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        from make_golden_imim_r2 import imim_inputs
        ih = ImageHeading(_types2.SimpleNamespace(aux_feat_dim_per_granularity=D)).to(dev).train()
        ix_np, ig_np, _, _, _, _ = imim_inputs(B, 100)
        # the upstream gradient in the layout the step produces it in: d ctx [B, R, D] of the word-region backward, i.e.
        # channels-last memory behind the logical [B, 256, 14, 14] (an NCHW-contiguous one would add a 30 us transpose copy)
        ix = torch.from_numpy(ix_np).to(dev)
        ig = torch.from_numpy(ig_np).to(dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)

        def imim_step():
            for p_ in ih.imim.parameters():
                p_.grad = None
            ih.imim(ix).backward(ig)
        im_ms, _ = time_step(imim_step, flush, max(3, min(args.steps, 10)), use_graph)
        im_flops = 3 * 2 * 83.9e6 * B
        im_mode = os.environ.get("TGFR_IMIM_PRECISION", "split")
        if im_mode == "fp32":
            im_peak, im_bound = 148 * 128 * 2 * 1.965e9 / 1e12, "fp32-simt"
            im_note = ("TGFR_IMIM_PRECISION=fp32: register-blocked fp32 SIMT GEMMs; peak = 148 SMs x 128 FMA lanes x 2 x "
                       "1.965 GHz (no measured fp32 peak in MEASURED_PEAKS.json)")
        else:
            im_peak, im_bound = pk["tf_sus"], "tensor"
            im_note = ("contractions on tcgen05 as fp16 hi/lo split products (3 MMA terms per algorithmic product, fp32-class "
                       "accuracy): executed tensor flops = 3 x algorithmic; the rest of the step is the operand splits and "
                       "the BatchNorm / LayerNorm / softmax passes (HBM bound)")
        line["image_head"] = {"metric": "imim_fwd_bwd_samples_per_sec", "value": B / (im_ms * 1e-3), "unit": "samples/s",
                              "ms_per_step": im_ms, "dtype": "f32" if im_mode == "fp32" else "f16x3 split (f32 accumulate)",
                              "config": {"B": B, "positions": R, "channels": D, "mode": "training (batch statistics)"},
                              "roofline": {"bound": im_bound, "achieved": im_flops / (im_ms * 1e-3) / 1e12,
                                           "peak": im_peak, "unit": "TFLOP/s", "frac": im_flops / (im_ms * 1e-3) / 1e12 / im_peak,
                                           "algorithmic_flops_per_step": im_flops,
                                           "note": im_note + "; 83.9 M MACs per sample forward, x3 with the backward"}}
        del ih, ix, ig

        # ---- configs[4]: FCFM fusion (eval kernel) + verification scoring in ONE timed call: 6 000 faces x 10 captions
        # = 60 000 pairs; a pair fuses (face_a, caption_a) and (face_b, caption_b) -> two 640-d embeddings -> cosine ->
        # exact ROC / AUC / EER / TPR@FPR (utils/modules.py:129-166).  Faces and caption features are resident in HBM.
        import contextlib as _ctx
        import io as _io
        from text_guided_face_recognition_b200.models.fusion_nets import Working
        from text_guided_face_recognition_b200.utils import modules as scoring
        fus = Working(channel_dim=256).to(dev).eval()
        NF, NCAP, TF = 6000, 10, 22
        gen5 = torch.Generator(device="cpu").manual_seed(5)
        faces_l = torch.nn.functional.normalize(torch.randn(NF, 14, 14, D, generator=gen5), dim=-1).to(dev).permute(0, 3, 1, 2)
        faces_g = torch.nn.functional.normalize(torch.randn(NF, D, generator=gen5), dim=1).to(dev)
        caps_w = [torch.nn.functional.normalize(torch.randn(NF, TF, D, generator=gen5), dim=2).to(dev).transpose(1, 2)
                  for _ in range(2)]
        caps_s = [torch.nn.functional.normalize(torch.randn(NF, D, generator=gen5), dim=1).to(dev) for _ in range(2)]
        plab5 = (torch.arange(NF * NCAP) % 10 == 0).long().to(dev)
        perm = torch.roll(torch.arange(NF), 1).to(dev)

        def fusion_verif():
            with torch.no_grad():
                o1, o2 = [], []
                for cset in range(NCAP):
                    o1.append(fus(faces_l, caps_w[cset & 1], faces_g, caps_s[cset & 1]))
                    o2.append(fus(faces_l, caps_w[(cset + 1) & 1], faces_g, caps_s[(cset + 1) & 1])[perm])
                with _ctx.redirect_stdout(_io.StringIO()):
                    return scoring.score_pairs([(torch.cat(o1), torch.cat(o2), plab5)])
        fusion_verif()
        torch.cuda.synchronize()
        e0.record()
        fusion_verif()
        e1.record()
        torch.cuda.synchronize()
        fv_ms = e0.elapsed_time(e1)
        conv_flops = 2 * 2 * NF * NCAP * 144 * 36 * 2304              # the 3x3 convolution of both fusions of every pair
        fv_bytes = 2 * NF * NCAP * 4 * (D * 196 + D * TF + 2 * D + 640)
        line["fusion_verification"] = {
            "metric": "fusion_verification_pairs_per_sec", "value": NF * NCAP / (fv_ms * 1e-3), "unit": "pairs/s",
            "ms_per_call": fv_ms, "dtype": "f32 (convolution: f16x3 split, f32 accumulate)",
            "config": {"workload": "configs[4]: 6000 faces x 10 captions = 60000 pairs; 2 x Working (eval forward) per pair "
                                   "-> pair cosine -> exact ROC / AUC / EER / TPR@FPR, one call", "pairs": NF * NCAP,
                       "fusion_samples": 2 * NF * NCAP, "T": TF},
            "fusion_samples_per_sec": 2 * NF * NCAP / (fv_ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": fv_bytes / (fv_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": fv_bytes / (fv_ms * 1e-3) / 1e9 / pk["hbm"], "algorithmic_bytes_per_call": fv_bytes,
                         "algorithmic_flops_per_call": conv_flops,
                         "conv_tflops": conv_flops / (fv_ms * 1e-3) / 1e12,
                         "note": "per fused sample the inputs are 256x196 + 256xT + 512 floats and the output 640; the 3x3 "
                                 "convolution (11.9 M MACs per sample) runs as an implicit tcgen05 GEMM on fp16 hi/lo split "
                                 "copies (3 MMA terms, fp32-class), the rest in one per-sample kernel; measured shares in "
                                 "DESIGN.md 4.6"}}
        del fus, faces_l, faces_g, caps_w, caps_s

        # ---- TextHeading (the BERT 768 -> 256 word / sentence projection of configs[1])
        import types as _types
        from text_guided_face_recognition_b200.models.text_heading import TextHeading
        bwn = T + 2
        tok_np, tw, tb = synth.texthead_inputs(B, bwn, D, seed=100)
        th = TextHeading(_types.SimpleNamespace(aux_feat_dim_per_granularity=D, bert_words_num=bwn)).to(dev)
        with torch.no_grad():
            for conv, w_, b_ in zip(th.bwm.convs1, tw, tb):
                conv.weight.copy_(torch.from_numpy(w_).unsqueeze(1))
                conv.bias.copy_(torch.from_numpy(b_))
        tok = torch.from_numpy(tok_np).to(dev)
        gw_up, gs_up = torch.randn(B, D, T, device=dev), torch.randn(B, D, device=dev)

        def th_step():
            for conv in th.bwm.convs1:
                conv.weight.grad = conv.bias.grad = None
            wo, so = th(tok, None)
            torch.autograd.backward([wo, so], [gw_up, gs_up])
        th_ms, _ = time_step(th_step, flush, 10, use_graph)           # one CUDA-graph replay per step, L2 flushed between
        th_flops = 2 * 2 * B * (bwn - 1) * (2 + 3 + 4) * 768 * D       # three products forward, three backward
        line["text_heading"] = {"metric": "text_heading_fwd_bwd_captions_per_sec", "value": B / (th_ms * 1e-3),
                                "unit": "captions/s", "ms_per_step": th_ms, "dtype": "f16 hi+lo split operands (3 accumulated tcgen05 terms, ~22 bits) / f32 accumulate",
                                "tflops": th_flops / (th_ms * 1e-3) / 1e12, "launch": "CUDA graph replay" if use_graph else "eager",
                                "config": {"B": B, "bert_words_num": bwn, "E": 768, "F": D}}

        # ---- verification scoring (configs[4] in its pair-list form, utils/modules.py:150-166):
        # 6000 faces x 10 captions = 60 000 pairs of 640-d fused embeddings -> cosine -> ROC -> AUC / EER / TPR@FPR
        import contextlib as _ctx
        import io as _io
        from text_guided_face_recognition_b200.utils import modules as scoring
        NP, DF = 60000, 640
        gen = torch.Generator(device="cpu").manual_seed(100)
        pl_np = (torch.arange(NP) % 10 == 0).long()
        ident = torch.randn(NP, DF, generator=gen)
        o1 = (ident + 4.0 * torch.randn(NP, DF, generator=gen)).to(dev)
        o2 = (torch.where(pl_np[:, None] == 1, ident, torch.randn(NP, DF, generator=gen))
              + 4.0 * torch.randn(NP, DF, generator=gen)).to(dev)
        pl = pl_np.to(dev)
        del ident

        def verif_step():
            with _ctx.redirect_stdout(_io.StringIO()):
                return scoring.score_pairs([(o1, o2, pl)])
        for _ in range(3):
            verif = verif_step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            verif_step()
        e1.record()
        torch.cuda.synchronize()
        v_ms = e0.elapsed_time(e1) / 10
        cos_ms, roc_ms = [], []
        sc = ops.pair_cosine(o1, o2)
        for _ in range(10):
            flush.zero_()
            e0.record()
            ops.pair_cosine(o1, o2)
            e1.record()
            torch.cuda.synchronize()
            cos_ms.append(e0.elapsed_time(e1))
            e0.record()
            ops.roc_counts(sc, pl)
            e1.record()
            torch.cuda.synchronize()
            roc_ms.append(e0.elapsed_time(e1))
        cos_ms, roc_ms = float(np.median(cos_ms)), float(np.median(roc_ms))
        NS = 1 << 24                                                             # sort throughput at a size that fills the GPU
        big_s, big_l = torch.rand(NS, device=dev) * 2 - 1, (torch.rand(NS, device=dev) < 0.1).long()
        ops.roc_counts(big_s, big_l)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            ops.roc_counts(big_s, big_l)
        e1.record()
        torch.cuda.synchronize()
        big_ms = e0.elapsed_time(e1) / 3
        del big_s, big_l
        cos_bytes = 2 * NP * DF * 4 + NP * 4
        roc_bytes_per_key = 12 + 4 * (4 + 2 * 5) + 5 + 20                         # keys pass, 4 radix passes, curve, points
        line["verification"] = {
            "metric": "verification_pairs_per_sec", "value": NP / (v_ms * 1e-3), "unit": "pairs/s", "ms_per_step": v_ms,
            "dtype": "f32 scores / u32 keys", "config": {"pairs": NP, "D": DF, "what": "pair cosine + exact ROC "
                       "(sklearn roc_curve semantics) + AUC/EER/TPR@FPR, embeddings resident in HBM"},
            "result": verif,
            "roofline": {"bound": "hbm", "kernel": "cosine_rows_vec_kernel", "achieved": cos_bytes / (cos_ms * 1e-3) / 1e9,
                         "peak": pk["hbm"], "unit": "GB/s", "frac": cos_bytes / (cos_ms * 1e-3) / 1e9 / pk["hbm"],
                         "kernel_ms": cos_ms, "algorithmic_bytes_per_launch": cos_bytes,
                         "traffic": ncu_traffic("cosine_rows_vec_kernel")},
            "roc": {"ms_60000": roc_ms, "ms_2p24": big_ms, "keys_per_sec_2p24": NS / (big_ms * 1e-3),
                    "algorithmic_bytes_per_key": roc_bytes_per_key,
                    "achieved_gbs_2p24": NS * roc_bytes_per_key / (big_ms * 1e-3) / 1e9,
                    "frac_of_hbm_2p24": NS * roc_bytes_per_key / (big_ms * 1e-3) / 1e9 / pk["hbm"],
                    "note": "whole tgfr_roc_curve call incl. workspace allocation and the count read-back"}}
        if world == 1 and not args.no_cpu_baseline:
            # the reference's own third-party calls on the host cores: torch CosineSimilarity + sklearn roc_curve / auc
            from sklearn import metrics as _skm
            torch.set_num_threads(os.cpu_count() or 1)
            h1, h2, hl = o1.cpu(), o2.cpu(), pl_np.tolist()
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 3.0:
                pred = torch.nn.CosineSimilarity(dim=1, eps=1e-6)(h1, h2).tolist()
                f_, t_, _ = _skm.roc_curve(hl, pred)
                _skm.auc(f_, t_)
                reps += 1
            dt = (time.perf_counter() - t0) / reps
            line["verification"]["cpu_baseline"] = {"value": NP / dt, "unit": "pairs/s", "cores": torch.get_num_threads(),
                                                    "kind": "reference", "sample": f"{reps} full passes of torch "
                                                    "CosineSimilarity + sklearn roc_curve + auc over the 60 000 pairs"}
        del o1, o2

        # ---- CPU baseline beside it: bounded sample of the same workload on the host cores
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=100)
            img, txt, cid = synth.sentence_inputs(B, D, seed=100)
            cpu_reference_step(ctx, words, img, txt, cid, n_caps=4, grads=args.grads)          # warm-up
            t_used, n_pairs, n_caps = 0.0, 0, 32
            while t_used < 10.0:
                dt, p = cpu_reference_step(ctx, words, img, txt, cid, n_caps=n_caps, grads=args.grads)
                t_used += dt
                n_pairs += p
            line["cpu_baseline"] = {"value": n_pairs / t_used, "unit": "pairs/s", "cores": torch.get_num_threads(),
                                    "kind": "port",
                                    "sample": f"{n_pairs} pairs ({n_caps} of {B} captions x {B} faces per pass), "
                                              f"{t_used:.1f} s of oracle/ref_port.py on the host"}
            # the same op-for-op port of the reference on THIS GPU (CUDA fp32, TF32 off): the like-for-like number
            # SURVEY 8(d) asks for beside the CPU one (the reference itself does not travel to the GPU box)
            from oracle import ref_port as P
            tf32 = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            gc_, gw_, ga_, gb_ = (torch.from_numpy(a_).to(dev) for a_ in (ctx, words, img, txt))
            gc_.requires_grad_(True), ga_.requires_grad_(True), gb_.requires_grad_(True)
            gw_.requires_grad_(args.grads == "both")

            def ref_gpu_step():
                for t_ in (gc_, gw_, ga_, gb_):
                    t_.grad = None
                l0_, l1_, _ = P.words_loss_port(gc_.view(B, IH, IW, D).permute(0, 3, 1, 2), gw_.transpose(1, 2), labels,
                                                None, *GAMMAS)
                s0_, s1_ = P.sent_loss_port(ga_, gb_, labels, cid, GAMMAS[2])
                (l0_ + l1_ + s0_ + s1_).backward()
            ref_gpu_step()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                ref_gpu_step()
            e1.record()
            torch.cuda.synchronize()
            rg_ms = e0.elapsed_time(e1) / 3
            torch.backends.cuda.matmul.allow_tf32 = tf32
            line["reference_on_gpu"] = {"value": B * B / (rg_ms * 1e-3), "unit": "pairs/s", "ms_per_step": rg_ms,
                                        "kind": "port", "dtype": "f32 (TF32 off)",
                                        "what": "oracle/ref_port.py (the reference's PyTorch op sequence) on this B200, "
                                                "full configs[1] step, eager"}
            # margin head (configs[2]) on the host cores and on this GPU through the same port
            hx = torch.from_numpy(xn).requires_grad_(True)
            hw = torch.from_numpy(wn).requires_grad_(True)
            hl = torch.from_numpy(lab)
            P.arc_focal_port(hx, hw, hl, h["s"], h["m"], h["gamma"])[1].backward()
            t0 = time.perf_counter()
            n_h = 0
            while time.perf_counter() - t0 < 3.0:
                hx.grad = hw.grad = None
                P.arc_focal_port(hx, hw, hl, h["s"], h["m"], h["gamma"])[1].backward()
                n_h += 1
            h_cpu = (time.perf_counter() - t0) / n_h
            gx, gwt, gl = hx.detach().to(dev).requires_grad_(True), hw.detach().to(dev).requires_grad_(True), hl.to(dev)
            for _ in range(3):
                gx.grad = gwt.grad = None
                P.arc_focal_port(gx, gwt, gl, h["s"], h["m"], h["gamma"])[1].backward()
            torch.cuda.synchronize()
            eh0, eh1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eh0.record()
            for _ in range(10):
                gx.grad = gwt.grad = None
                P.arc_focal_port(gx, gwt, gl, h["s"], h["m"], h["gamma"])[1].backward()
            eh1.record()
            torch.cuda.synchronize()
            head_ref_gpu_ms = eh0.elapsed_time(eh1) / 10          # read out before any other event is re-recorded
            # TextHeading through the port on this GPU (the reference's B x T Python loop of stack / amax calls)
            tws = [torch.from_numpy(w_).unsqueeze(1).to(dev).requires_grad_(True) for w_ in tw]
            tbs = [torch.from_numpy(b_).to(dev).requires_grad_(True) for b_ in tb]

            def th_ref():
                for t_ in tws + tbs:
                    t_.grad = None
                wo, so = P.text_heading_port(tok, tws, tbs, bwn)
                torch.autograd.backward([wo, so], [gw_up, gs_up])
            th_ref()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(2):
                th_ref()
            e1.record()
            torch.cuda.synchronize()
            line["text_heading"]["reference_on_gpu"] = {"value": B / (e0.elapsed_time(e1) / 2 * 1e-3), "unit": "captions/s",
                                                        "ms_per_step": e0.elapsed_time(e1) / 2, "kind": "port"}
            line["margin_head"]["cpu_baseline"] = {"value": h["B"] / h_cpu, "unit": "samples/s", "kind": "port",
                                                   "cores": torch.get_num_threads(), "sample": f"{n_h} full steps"}
            line["margin_head"]["reference_on_gpu"] = {"value": h["B"] / (head_ref_gpu_ms * 1e-3), "unit": "samples/s",
                                                       "ms_per_step": head_ref_gpu_ms, "kind": "port",
                                                       "dtype": "f32 (TF32 off)",
                                                       "what": "oracle/ref_port.py arc_focal_port (the reference's op "
                                                               "sequence: normalize, linear, one_hot blend, CE) on this "
                                                               "B200, eager, 10 steps"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # release the captured graphs before the communicator goes away, then leave without the NCCL teardown
        # (destroy_process_group() was seen to block when graphs holding captured collectives had existed)
        runners = None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--grads", default="ctx", choices=["ctx", "both"],
                    help="gradients of the word-region loss: ctx = face side only (the reference's training step), both")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
