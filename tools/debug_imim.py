import os, sys, types, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_golden_imim_r2 import imim_inputs
from test_imim import make_head, load, rel
gd = os.path.join(ROOT, 'tests', 'golden')
g = load(gd, 'imim_train')
x, gout, xg, gg, wg, bg = imim_inputs(3, 7)
head = make_head(gd, wg, bg).train()
xt = torch.from_numpy(x).cuda().requires_grad_(True)
loc = head.imim(xt)
loc.backward(torch.from_numpy(gout).cuda())
print('out', np.abs(loc.detach().cpu().numpy() - g['out']).max())
print('dx', rel(xt.grad.cpu().numpy(), g['dx']))
for name, p in head.imim.named_parameters():
    if p.grad is not None:
        print(name, rel(p.grad.cpu().numpy(), g['g:' + name]), float(np.linalg.norm(g['g:' + name])))
