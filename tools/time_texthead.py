"""Time TextHeading fwd+bwd at config 2 (B=128, bert_words_num=24, 768 -> 256)."""
import sys, types, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import synth
from text_guided_face_recognition_b200.models.text_heading import TextHeading
B, T, D = 128, 22, 256
bwn = T + 2
tok_np, tw, tb = synth.texthead_inputs(B, bwn, D, seed=100)
th = TextHeading(types.SimpleNamespace(aux_feat_dim_per_granularity=D, bert_words_num=bwn)).cuda()
with torch.no_grad():
    for conv, w_, b_ in zip(th.bwm.convs1, tw, tb):
        conv.weight.copy_(torch.from_numpy(w_).unsqueeze(1)); conv.bias.copy_(torch.from_numpy(b_))
tok = torch.from_numpy(tok_np).cuda()
gw, gs = torch.randn(B, D, T, device='cuda'), torch.randn(B, D, device='cuda')
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
def fwd():
    return th(tok, None)
def step():
    for conv in th.bwm.convs1: conv.weight.grad = conv.bias.grad = None
    wo, so = th(tok, None)
    torch.autograd.backward([wo, so], [gw, gs])
for name, fn in (('fwd', fwd), ('fwd+bwd', step)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, 'ms', e0.elapsed_time(e1) / 10)
