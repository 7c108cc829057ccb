"""One tensor-core word-region forward + backward (face-side gradient) at BASELINE config 2, for ncu."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import synth
from text_guided_face_recognition_b200 import _lib, ops
B, T, R, D = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 22, 196, 256
ctx, words, _ = synth.wordregion_inputs(B, T, R, D, 'BERT', 100)
f = torch.from_numpy(ctx).cuda().requires_grad_(True)
w = torch.from_numpy(words).cuda()
for _ in range(2):
    f.grad = None
    sim = ops.wordregion_sim(f, w, None, 4., 5., 10., precision=_lib.PREC_TC, want_attn=False)[0]
    sim.backward(torch.ones_like(sim) / B)
torch.cuda.synchronize()
print('ok')
