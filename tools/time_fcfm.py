"""Time the fused FCFM `Working` eval forward (csrc/fcfm.cu): B = 6000 and 60 000 samples, T = 30 words."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from text_guided_face_recognition_b200.models.fusion_nets import Working
torch.manual_seed(0)
torch.set_grad_enabled(False)          # the verification path (utils/modules.py:129 runs under no_grad)
net = Working(256).cuda().eval()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for B in (6000, 60000):
    img = torch.randn(B, 14, 14, 256, device='cuda').permute(0, 3, 1, 2)
    word = torch.randn(B, 30, 256, device='cuda').transpose(1, 2)
    gl = torch.randn(B, 256, device='cuda'); sent = torch.randn(B, 256, device='cuda')
    net(img, word, gl, sent); torch.cuda.synchronize()
    e0.record()
    for _ in range(3): net(img, word, gl, sent)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flops = B * 2 * (36 * 144 * 2304 + 30 * 36 * 256 + 36 * 36 * 30 + 5 * 36 * 36 * 36 + 128 * 324)
    print(f'B={B}: {ms:.3f} ms  {B/ms*1e3/1e6:.2f} M samples/s  {flops/ms/1e9:.1f} TFLOP/s fp32  input {B*(256*196+256*30+512)*4/ms/1e6:.0f} GB/s')
    del img, word
