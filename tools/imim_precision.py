"""IMIM fwd + bwd at B = 128 in the three product modes vs the float64 run of the reference module
(tests/golden/imim_config2_f64.npz): relative error of the output slice, d input and every parameter gradient."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_golden_imim_r2 import imim_inputs
from test_imim import make_head
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'imim_config2_f64.npz'))
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))
x, gout, xg, gg, wg, bg = imim_inputs(128, 11)
for mode in ('fp32', 'split'):
    os.environ['TGFR_IMIM_PRECISION'] = mode
    head = make_head(os.path.join(ROOT, 'tests', 'golden'), wg, bg).train()
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    loc = head.imim(xt)
    loc.backward(torch.from_numpy(gout).cuda())
    o = loc.detach().contiguous().cpu().numpy()
    res = {'out': rel(o[:2], g['out_head']), 'dx': rel(xt.grad.cpu().numpy()[:2], g['dx_head'])}
    for name, p in head.imim.named_parameters():
        if name.startswith('project_local.fc') or float(g['n:' + name]) < 1e-3:
            continue
        got = p.grad.cpu().numpy()
        sl = got if got.size <= 512 else got.reshape(got.shape[0], -1)[:4]
        res[name] = rel(sl, g['g:' + name])
    print(mode, ' '.join(f'{k.replace("project_local.projection", "proj")}={v:.1e}' for k, v in res.items()))
