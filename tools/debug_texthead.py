import os, sys, types, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import synth
from oracle import texthead_oracle as TO
from text_guided_face_recognition_b200.models.text_heading import TextHeading
B, wn, F = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 24, 256
tokens, ws, bs = synth.texthead_inputs(B, wn, F, seed=7)
rng = np.random.RandomState(3)
gw, gs = rng.randn(B, wn - 2, F).astype(np.float32), rng.randn(B, F).astype(np.float32)
rw, rs_ = TO.forward(tokens, ws, bs, wn)
dws, dbs = TO.backward(tokens, ws, bs, wn, gw, gs)
def rel(a, b): return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
for mode in ('tc', 'fp32'):
    os.environ['TGFR_TEXTHEAD_PRECISION'] = mode
    th = TextHeading(types.SimpleNamespace(aux_feat_dim_per_granularity=F, bert_words_num=wn)).cuda()
    with torch.no_grad():
        for conv, w_, b_ in zip(th.bwm.convs1, ws, bs):
            conv.weight.copy_(torch.from_numpy(w_).unsqueeze(1)); conv.bias.copy_(torch.from_numpy(b_))
    words, sent = th(torch.from_numpy(tokens).cuda(), None)
    print(mode, 'fwd max abs', np.max(np.abs(words.transpose(1, 2).detach().cpu().numpy() - rw)), np.max(np.abs(sent.detach().cpu().numpy() - rs_)))
    ((words.transpose(1, 2) * torch.from_numpy(gw).cuda()).sum() + (sent * torch.from_numpy(gs).cuda()).sum()).backward()
    for k, conv in enumerate(th.bwm.convs1):
        got = conv.weight.grad.squeeze(1).cpu().numpy()
        rows = np.linalg.norm(got - dws[k], axis=1) / np.linalg.norm(dws[k], axis=1)
        print(mode, 'conv', k, 'dw rel', rel(got, dws[k]), 'db rel', rel(conv.bias.grad.cpu().numpy(), dbs[k]),
              'feature rows with rel > 1e-4:', int((rows > 1e-4).sum()), 'worst', float(rows.max()), 'median', float(np.median(rows)))
