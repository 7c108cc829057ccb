import sys, torch, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import synth
from oracle import fcam_oracle as O
from text_guided_face_recognition_b200 import _lib, ops
B,T,R,D=(int(a) for a in sys.argv[1:5]) if len(sys.argv)>4 else (8,22,196,256)
fl = sys.argv[5] if len(sys.argv)>5 else 'BERT'
ctx,words,cap=synth.wordregion_inputs(B,T,R,D,fl,100,ragged=(fl=='LSTM'))
f=torch.from_numpy(ctx).cuda().requires_grad_(True); w=torch.from_numpy(words).cuda()
capt=None if cap is None else torch.from_numpy(cap).cuda()
sim,_=ops.wordregion_sim(f,w,capt,4.,5.,10.,precision=_lib.PREC_TC,want_attn=False)
l0,l1=ops.pair_ce(sim); (l0+l1).backward(); torch.cuda.synchronize()
got=f.grad.cpu().numpy()
ref,_=O.words_loss_grads(ctx,words,None,cap,4.,5.,10.)
bad=~np.isfinite(got)
print('nonfinite frac',bad.mean())
if bad.any():
    idx=np.argwhere(bad)
    print('b:',np.unique(idx[:,0]),'r range',idx[:,1].min(),idx[:,1].max(),'d range',idx[:,2].min(),idx[:,2].max())
    print('r unique',np.unique(idx[:,1])[:40]); print('d unique',np.unique(idx[:,2])[:70])
g=np.where(bad,0,got); rf=np.where(bad,0,ref)
print('rel err on finite',np.linalg.norm(g-rf)/np.linalg.norm(rf))
for b in range(min(B,4)):
    for (r0,r1) in ((0,128),(128,R)):
        if r1<=r0: continue
        for q in range(D//64):
            a=g[b,r0:r1,64*q:64*q+64]; c=rf[b,r0:r1,64*q:64*q+64]
            print(f'b{b} rows{r0}-{r1} q{q} rel',np.linalg.norm(a-c)/max(np.linalg.norm(c),1e-30), 'ratio', (a*c).sum()/max((c*c).sum(),1e-30))
