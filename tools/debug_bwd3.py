import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import synth
from oracle import fcam_oracle as O
from text_guided_face_recognition_b200 import _lib, ops
B, T, R, D = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 22, 196, 256
ctx, words, _ = synth.wordregion_inputs(B, T, R, D, 'BERT', 100)
f = torch.from_numpy(ctx).cuda().requires_grad_(True)
w = torch.from_numpy(words).cuda()
sim, _ = ops.wordregion_sim(f, w, None, 4., 5., 10., precision=_lib.PREC_TC, want_attn=False)
torch.cuda.synchronize(); print('fwd ok')
l0, l1 = ops.pair_ce(sim)
(l0 + l1).backward()
torch.cuda.synchronize(); print('bwd ok')
ref, _ = O.words_loss_grads(ctx, words, None, None, 4., 5., 10.)
got = f.grad.cpu().numpy()
print('rel err', np.linalg.norm(got - ref) / np.linalg.norm(ref))
