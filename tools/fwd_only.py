"""One tensor-core word-region forward at BASELINE config 2 (with or without the forward->backward records), for ncu."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import synth
from text_guided_face_recognition_b200 import _lib, ops
B, T, R, D = 128, int(sys.argv[2]) if len(sys.argv) > 2 else 22, 196, 256
ctx, words, _ = synth.wordregion_inputs(B, T, R, D, 'BERT', 100)
save = len(sys.argv) > 1 and sys.argv[1] == 'save'
f = torch.from_numpy(ctx).cuda().requires_grad_(save)
w = torch.from_numpy(words).cuda()
for _ in range(2):
    sim = ops.wordregion_sim(f, w, None, 4., 5., 10., precision=_lib.PREC_TC, want_attn=False)[0]
torch.cuda.synchronize()
print('ok', float(sim.sum()))
