"""Phase trace of the tensor-core word-region kernels (clock64 stamps of CTA 0; see tgfr_debug_set_trace)."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import synth
from text_guided_face_recognition_b200 import _lib, ops
B, T, R, D = 128, 22, 196, 256
ctx, words, _ = synth.wordregion_inputs(B, T, R, D, 'BERT', 100)
w = torch.from_numpy(words).cuda()
f = torch.from_numpy(ctx).cuda().requires_grad_(True)
lib = _lib.load()
trace = torch.zeros(16 * 32, dtype=torch.int64, device='cuda')
def fwd(): return ops.wordregion_sim(f, w, None, 4., 5., 10., precision=_lib.PREC_TC, want_attn=False)[0]
sim = fwd(); g = torch.randn_like(sim) / B
sim.backward(g, retain_graph=True); torch.cuda.synchronize()
names = {
    # pipelined forward (wr_tc_fwd3_kernel): A = word-softmax warps, B = cosine / V warps
    'fwd': {2: 'A:tile0 start', 3: 'A:tile0 done', 4: 'A:tile1 done', 5: 'B:Wu ready', 6: 'B:pass1 done', 7: 'B:V pass done',
            17: 'mma:E0 ready', 18: 'mma:G2a+G1 issued', 19: 'mma:G2b issued'},
    # record-free backward (wr_tc_bwd3_kernel): one line per (face, tile, caption group) item
    'bwd': {2: 'S|X ready', 3: 'ops written', 8: 'acc done', 9: 'drained', 17: 'mma:Wu ready', 18: 'mma:ops ready',
            19: 'mma:G5 issued'},
}
for which in ('fwd', 'bwd'):
    trace.zero_()
    _lib.check(lib.tgfr_debug_set_trace(trace.data_ptr()), 'trace')
    if which == 'fwd': fwd()
    else: f.grad = None; sim.backward(g, retain_graph=True)
    torch.cuda.synchronize()
    _lib.check(lib.tgfr_debug_set_trace(0), 'trace')
    t = trace.cpu().numpy().reshape(16, 32)
    base = t[:, [2, 3, 17]][t[:, [2, 3, 17]] > 0].min()
    print('====', which)
    for n in range(2, 8):
        evs = sorted((t[n, e] - base, e) for e in range(32) if t[n, e] > 0 and not (8 <= e <= 12 or 24 <= e <= 26))
        print('unit', n, ' '.join(f"{names[which].get(e, e)}@{int(c)}" for c, e in evs))
    print('cycles per unit (epilogue thread):', np.diff(t[2:14, 3]))

