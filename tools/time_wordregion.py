import sys, torch, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import synth
from text_guided_face_recognition_b200 import _lib, ops
B,T,R,D=128,22,196,256
ctx,words,_=synth.wordregion_inputs(B,T,R,D,'BERT',100)
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
precs = ((_lib.PREC_TC,'tc'),) if len(sys.argv) > 1 and sys.argv[1] == 'tc' else ((_lib.PREC_TC,'tc'),(_lib.PREC_FP32,'fp32'))
for prec,name in precs:
    for mode in ('ctx', 'words', 'both'):
        f=torch.from_numpy(ctx).cuda().requires_grad_(mode != 'words')
        w=torch.from_numpy(words).cuda().requires_grad_(mode != 'ctx')
        def fwd(): return ops.wordregion_sim(f,w,None,4.,5.,10.,precision=prec,want_attn=False)[0]
        if mode == 'ctx':
            for _ in range(3): fwd()
            torch.cuda.synchronize(); e0.record()
            for _ in range(10): fwd()
            e1.record(); torch.cuda.synchronize()
            ms=e0.elapsed_time(e1)/10
            print(name,'fwd ms',ms,'TFLOP/s (alg 4TRD)',4*T*R*D*B*B/ms/1e9)
        sim=fwd(); g=torch.randn_like(sim)/B
        def bwd():
            f.grad=None; w.grad=None; sim.backward(g,retain_graph=True)
        for _ in range(3): bwd()
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): bwd()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        k = {'ctx': 6, 'words': 4, 'both': 8}[mode]
        print(name,f'bwd({mode}) ms',ms,f'TFLOP/s (alg {k}TRD)',k*T*R*D*B*B/ms/1e9)
# diagonal attention maps: forward with and without them (the difference is the attention-only kernel)
f = torch.from_numpy(ctx).cuda(); w = torch.from_numpy(words).cuda()
for want in (False, True):
    def fwd2(): return ops.wordregion_sim(f, w, None, 4., 5., 10., precision=_lib.PREC_TC, want_attn=want)[0]
    for _ in range(3): fwd2()
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): fwd2()
    e1.record(); torch.cuda.synchronize()
    print('tc fwd (no records), want_attn =', want, 'ms', e0.elapsed_time(e1) / 20)
