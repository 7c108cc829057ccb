"""One pass of the scoring kernels at sizes that fill the GPU, for ncu: pair cosine 600 000 x 640, exact ROC of 2^24 scores."""
import sys, torch
sys.path.insert(0, '/root/repo')
from text_guided_face_recognition_b200 import ops
a = torch.randn(600000, 640, device='cuda'); b = torch.randn(600000, 640, device='cuda')
ops.pair_cosine(a, b)
N = 1 << 24
s = torch.rand(N, device='cuda') * 2 - 1; l = (torch.rand(N, device='cuda') < 0.1).long()
ops.roc_counts(s, l)
torch.cuda.synchronize(); print('ok')
