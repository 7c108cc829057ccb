"""IMIM (image head local branch) fwd + bwd at the configs[1] batch: ms per step."""
import os, sys, types, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_golden_imim_r2 import imim_inputs
from test_imim import make_head
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
x, gout, xg, gg, wg, bg = imim_inputs(B, 11)
head = make_head(os.path.join(ROOT, 'tests', 'golden'), wg, bg).train()
xt = torch.from_numpy(x).cuda()
go = torch.from_numpy(gout).cuda()
def step():
    for p in head.parameters(): p.grad = None
    head.imim(xt).backward(go)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f'IMIM fwd+bwd B={B}: {ms:.3f} ms  ({3 * 2 * 83.9e6 * B / ms / 1e9:.1f} TFLOP/s algorithmic)')
with torch.no_grad():
    for _ in range(3): head.imim(xt)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): head.imim(xt)
    e1.record(); torch.cuda.synchronize()
print(f'IMIM fwd only: {e0.elapsed_time(e1) / 10:.3f} ms')
