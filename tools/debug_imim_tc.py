"""Where do the split-mode IMIM gradients leave the fp32-mode ones?  Captures the backward workspace of both modes."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_golden_imim_r2 import imim_inputs
from test_imim import make_head
from text_guided_face_recognition_b200 import ops
B, P = 128, 196
M = B * P
x, gout, xg, gg, wg, bg = imim_inputs(B, 11)
caught = {}
orig = ops._workspace
def spy(n, dev):
    t = orig(n, dev)
    caught.setdefault('ws', []).append(t)
    return t
ops._workspace = spy
res = {}
for mode in ('fp32', 'split'):
    os.environ['TGFR_IMIM_PRECISION'] = mode
    caught.clear()
    head = make_head(os.path.join(ROOT, 'tests', 'golden'), wg, bg).train()
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    loc = head.imim(xt)
    loc.backward(torch.from_numpy(gout).cuda())
    torch.cuda.synchronize()
    ws = max(caught['ws'], key=lambda t: t.numel()).view(torch.float32)
    off = 0
    parts = {}
    for name, n in (('dxn', M * 256), ('dH2', M * 256), ('dH1', M * 128), ('dO', M * 256), ('dS', B * P * P), ('dQKV', M * 768)):
        parts[name] = ws[off:off + n].double().cpu().numpy()
        off += n
    res[mode] = parts
for name in res['fp32']:
    a, b = res['fp32'][name], res['split'][name]
    cols = {'dxn': 256, 'dH2': 256, 'dH1': 128, 'dO': 256, 'dS': P, 'dQKV': 768}[name]
    a2, b2 = a.reshape(-1, cols), b.reshape(-1, cols)
    d = b2 - a2
    print(name, 'rel', np.linalg.norm(d) / np.linalg.norm(a2), 'colsum rel', np.linalg.norm(d.sum(0)) / np.linalg.norm(a2.sum(0)),
          'mean err / rms', d.mean() / np.sqrt((a2 ** 2).mean()), 'mask flips', int(((a2 == 0) != (b2 == 0)).sum()),
          'worst col', int(np.argmax(np.abs(d.sum(0)))), 'row-block errs', [float(np.linalg.norm(d[i:i + M // 4]) / np.linalg.norm(a2[i:i + M // 4])) for i in range(0, d.shape[0], max(d.shape[0] // 4, 1))][:4])
for name in ('dH2', 'dQKV'):
    cols = {'dH2': 256, 'dQKV': 768}[name]
    a2, b2 = res['fp32'][name].reshape(-1, cols), res['split'][name].reshape(-1, cols)
    e = np.array([np.linalg.norm(b2[i:i + 128] - a2[i:i + 128]) / max(np.linalg.norm(a2[i:i + 128]), 1e-30) for i in range(0, M, 128)])
    bad = np.nonzero(e > 1e-5)[0]
    print(name, 'bad 128-row tiles:', bad[:40], 'count', len(bad), 'their errs', e[bad][:8])
    if len(bad):
        t = bad[0]
        d = b2[t * 128:(t + 1) * 128] - a2[t * 128:(t + 1) * 128]
        print(' within tile', t, ': per-row err (first 16 rows)', np.linalg.norm(d, axis=1)[:16] / np.linalg.norm(a2[t * 128:(t + 1) * 128], axis=1)[:16])
        print(' per 32-col chunk', [float(np.linalg.norm(d[:, c:c + 32]) / np.linalg.norm(a2[t * 128:(t + 1) * 128, c:c + 32])) for c in range(0, cols, 32)][:8])
