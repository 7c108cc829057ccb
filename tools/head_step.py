"""One ArcFace + focal-loss step at BASELINE config 3 (B=512, Din=512, C=10177), for ncu launch lists."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import synth
from text_guided_face_recognition_b200.models import losses, metrics
xn, wn, lab = synth.margin_inputs(512, 512, 10177, seed=100)
head = metrics.ArcMarginProduct(512, 10177, s=30., m=0.5).cuda()
with torch.no_grad():
    head.weight.copy_(torch.from_numpy(wn))
x = torch.from_numpy(xn).cuda().requires_grad_(True)
labt = torch.from_numpy(lab).cuda()
crit = losses.FocalLoss(gamma=2.0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
fused = len(sys.argv) > 2 and sys.argv[2] == 'fused'
for _ in range(n):
    x.grad = None; head.weight.grad = None
    if fused:
        head.fused_loss(x, labt, gamma=2.0).backward()
    else:
        crit(head(x, labt), labt).backward()
torch.cuda.synchronize()
print('ok')
