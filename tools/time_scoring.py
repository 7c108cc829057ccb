"""Time the scoring kernels alone (CUDA events, L2-sized inputs): pair cosine GB/s and ROC keys/s."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from text_guided_face_recognition_b200 import ops
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for N, D in ((60000, 640), (600000, 640), (2000000, 512)):
    a = torch.randn(N, D, device='cuda'); b = torch.randn(N, D, device='cuda')
    ms = timeit(lambda: ops.pair_cosine(a, b))
    print(f'pair_cosine N={N} D={D}: {ms*1e3:.1f} us, {2*N*D*4/ms/1e6:.0f} GB/s')
    del a, b
for N in (60000, 1 << 20, 1 << 24, 1 << 27):
    s = (torch.rand(N, device='cuda') * 2 - 1); l = (torch.rand(N, device='cuda') < 0.1).long()
    ms = timeit(lambda: ops.roc_counts(s, l), n=5, warm=2)
    print(f'roc_counts N={N}: {ms:.3f} ms, {N/ms/1e6:.2f} G keys/s (incl. host sync + allocs)')
    del s, l
x = torch.rand(6000, 10, device='cuda')
print('row_argmax 6000x10 us', timeit(lambda: ops.row_argmax(x)) * 1e3)
