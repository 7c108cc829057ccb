"""Drop-in proof: the loss block of the reference's `Train.train`
(/root/reference/src/train_encoders_bert.py:267-323) running on the mirror.

The fixture `tests/golden/train_block_bert.npz` was produced by `tests/golden/make_golden_r2.py`, which reads
those very source lines from the reference tree and executes them against the reference's own modules on CPU.
Here the same block -- the same calls, in the same order, with the same argument expressions and the same
`from models.losses import ...` / `from models import metrics, losses` imports the reference driver uses
(train_encoders_bert.py:19,25), resolved by putting the mirror package first on sys.path -- runs on the GPU
through libtgfr_b200.so.  Compared: total_loss, the logged partial sums, and every gradient the block produces
(image_head's two outputs and the two classifier weights).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BLOCK = r'''
import os, sys, types
import numpy as np
import torch
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "text_guided_face_recognition_b200"))     # the mirror shadows `models`
sys.modules.setdefault("easydict", types.ModuleType("easydict"))
from models.losses import sent_loss, words_loss, CMPLoss, ClipLoss, global_loss   # train_encoders_bert.py:19
from models import metrics, losses                                                 # :25
from make_golden_r2 import train_block_inputs

g = np.load(os.path.join(ROOT, "tests", "golden", "train_block_bert.npz"))
B, T, D, C = int(g["B"]), int(g["T"]), int(g["D"]), int(g["C"])
ctx, words, img, txt, cid, w_img, w_txt = train_block_inputs(B, T, D, C)
ns = types.SimpleNamespace
args = ns(en_type="BERT", bert_words_num=T + 2, CUDA=True, device="cuda",
          TRAIN=ns(SMOOTH=ns(GAMMA1=4.0, GAMMA2=5.0, GAMMA3=10.0)),
          is_DAMSM=True, is_WRA=False, is_ident_loss=True, is_CLIP=True, is_CMP=False, lambda_clip=2.0, lambda_id=100,
          model_type="arcface", batch_size=B)


class Train:                                     # what build_models() leaves on `self` (train_encoders_bert.py:139-192)
    pass


self = Train()
self.args = args
self.image_cls = metrics.ArcMarginProduct(D, C, s=30, m=0.5, easy_margin=False).cuda()
self.text_cls = metrics.ArcMarginProduct(D, C, s=35, m=0.5, easy_margin=False).cuda()
with torch.no_grad():
    self.image_cls.weight.copy_(torch.from_numpy(w_img))
    self.text_cls.weight.copy_(torch.from_numpy(w_txt))
self.image_cls = torch.nn.DataParallel(self.image_cls, device_ids=[0])      # wrapped exactly like the driver does
self.text_cls = torch.nn.DataParallel(self.text_cls, device_ids=[0])
self.ident_loss = losses.FocalLoss(gamma=2)

words_features = torch.from_numpy(ctx).cuda().view(B, 14, 14, D).permute(0, 3, 1, 2).requires_grad_(True)
img_features = torch.from_numpy(img).cuda().requires_grad_(True)
words_emb = torch.from_numpy(words).cuda().transpose(1, 2)
sent_emb = torch.from_numpy(txt).cuda()
labels = torch.arange(B).cuda()
class_ids = torch.from_numpy(cid)
batch_size, cap_lens = B, None
total_damsm_loss = w_total_loss = s_total_loss = total_cl_loss = 0

# ---- the block (same statements as train_encoders_bert.py:270-323 with is_WRA / is_CMP off) ----
total_loss = 0
if self.args.is_DAMSM == True:
    w_loss0, w_loss1, attn_maps = words_loss(words_features, words_emb, labels,
                                             cap_lens, class_ids.numpy(), batch_size, self.args)
    s_loss0, s_loss1 = sent_loss(img_features, sent_emb, labels,
                                 class_ids.numpy(), batch_size, self.args)
    damsm_loss = w_loss0 + w_loss1 + s_loss0 + s_loss1
    total_damsm_loss += damsm_loss.item()
    total_loss += damsm_loss
    w_total_loss += ((w_loss0 + w_loss1).data).item()
    s_total_loss += ((s_loss0 + s_loss1).data).item()
if self.args.is_ident_loss == True:
    class_ids = class_ids.cuda()
    output = self.text_cls(sent_emb, class_ids)
    tid_loss = self.ident_loss(output, class_ids)
    if self.args.model_type == "arcface":
        output = self.image_cls(img_features, class_ids)
    iid_loss = self.ident_loss(output, class_ids)
    total_loss += self.args.lambda_id * tid_loss
    total_loss += self.args.lambda_id * iid_loss
if self.args.is_CLIP == True:
    cl_loss = global_loss(img_features, sent_emb)
    total_loss += self.args.lambda_clip * cl_loss
    total_cl_loss += self.args.lambda_clip * cl_loss
total_loss.backward()
# ---- end of the block ----
torch.cuda.synchronize()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


assert len(attn_maps) == B and tuple(attn_maps[0].shape) == (1, T, 14, 14)
tol_l, tol_g = (2e-5, 2e-4) if PREC == "fp32" else (1e-4, 1e-3)
checks = {
    "total_loss": abs(total_loss.item() - float(g["total_loss"])) / abs(float(g["total_loss"])),
    "damsm": abs(total_damsm_loss - float(g["damsm"])) / abs(float(g["damsm"])),
    "w_total": abs(w_total_loss - float(g["w_total"])) / abs(float(g["w_total"])),
    "s_total": abs(s_total_loss - float(g["s_total"])) / abs(float(g["s_total"])),
    "cl": abs(total_cl_loss.item() - float(g["cl"])) / abs(float(g["cl"])),
}
for k, v in checks.items():
    assert v < tol_l, (k, v)
grads = {
    "d_words_features": rel(words_features.grad.cpu().numpy(), g["d_words_features"]),
    "d_img_features": rel(img_features.grad.cpu().numpy(), g["d_img_features"]),
    "d_image_cls": rel(self.image_cls.module.weight.grad.cpu().numpy(), g["d_image_cls"]),
    "d_text_cls": rel(self.text_cls.module.weight.grad.cpu().numpy(), g["d_text_cls"]),
}
for k, v in grads.items():
    assert v < tol_g, (k, v)
print("dropin-loop ok", PREC, checks, grads)
'''


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "tc"])
def test_reference_train_block_runs_on_the_mirror(prec):
    env = dict(os.environ, TGFR_WORDREGION_PRECISION=prec, TGFR_HEAD_PRECISION=prec)
    code = f"ROOT = {ROOT!r}\nPREC = {prec!r}\n" + BLOCK
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", env=env, timeout=600)
    assert out.returncode == 0 and "dropin-loop ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
