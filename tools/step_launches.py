"""configs[1] loss step (words_loss + sent_loss fwd + bwd, face-side gradient) as bench.py runs it -- one CUDA-graph replay per
step -- for an ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches.csv python tools/step_launches.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import synth  # noqa: E402
from text_guided_face_recognition_b200 import fcam  # noqa: E402
from text_guided_face_recognition_b200.graphs import GraphedStep  # noqa: E402

B, T, R, D = 128, 22, 196, 256
ns = types.SimpleNamespace
args = ns(en_type="BERT", bert_words_num=T + 2, CUDA=True, device="cuda", TRAIN=ns(SMOOTH=ns(GAMMA1=4.0, GAMMA2=5.0, GAMMA3=10.0)))
ctx, words, _ = synth.wordregion_inputs(B, T, R, D, "BERT", seed=100)
img, txt, _ = synth.sentence_inputs(B, D, seed=100)
c = torch.from_numpy(ctx).cuda().requires_grad_(True)
w = torch.from_numpy(words).cuda()
a = torch.from_numpy(img).cuda().requires_grad_(True)
b = torch.from_numpy(txt).cuda()
labels = torch.arange(B, device="cuda")
cid = torch.arange(B, device="cuda")


def step():
    c.grad = None
    a.grad = None
    w0, w1, _, s0, s1 = fcam.fcam_losses(c.view(B, 14, 14, D).permute(0, 3, 1, 2), w.transpose(1, 2), a, b, labels, None, cid, B, args)
    (w0 + w1 + s0 + s1).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
g = GraphedStep(step)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    g()
torch.cuda.synchronize()
print("ok")
