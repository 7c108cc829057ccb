"""One IMIM fwd + bwd at B = 128 for ncu launch lists."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_golden_imim_r2 import imim_inputs
from test_imim import make_head
B = 128
x, gout, xg, gg, wg, bg = imim_inputs(B, 11)
head = make_head(os.path.join(ROOT, 'tests', 'golden'), wg, bg).train()
xt = torch.from_numpy(x).cuda(); go = torch.from_numpy(gout).cuda()
for _ in range(2):
    for p in head.parameters(): p.grad = None
    head.imim(xt).backward(go)
torch.cuda.synchronize(); print('ok')
