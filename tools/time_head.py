"""Time the margin-head step (config 3) as CUDA-graph replays: dense-logits API and fused_loss."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import synth
from text_guided_face_recognition_b200.models import losses, metrics
from text_guided_face_recognition_b200.graphs import GraphedStep
B, Din, C = 512, 512, 10177
xn, wn, lab = synth.margin_inputs(B, Din, C, seed=100)
head = metrics.ArcMarginProduct(Din, C, s=30., m=0.5).cuda()
with torch.no_grad():
    head.weight.copy_(torch.from_numpy(wn))
x = torch.from_numpy(xn).cuda().requires_grad_(True)
labt = torch.from_numpy(lab).cuda()
crit = losses.FocalLoss(gamma=2.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
def dense():
    x.grad = None; head.weight.grad = None
    crit(head(x, labt), labt).backward()
def fused():
    x.grad = None; head.weight.grad = None
    head.fused_loss(x, labt, gamma=2.0).backward()
for name, fn in (('dense', dense), ('fused', fused)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = GraphedStep(fn)
    ts = []
    for _ in range(20):
        flush.zero_(); e0.record(); g(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f'{name}: median {ts[10]*1e3:.1f} us  min {ts[0]*1e3:.1f} us  -> {B/ts[10]*1e3/1e6:.2f} M samples/s')
