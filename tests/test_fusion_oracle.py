"""FCFM fusion net `Working` (reference models/fusion_nets.py:217-258): the numpy oracle against
fixtures generated from the reference module in eval mode (tests/golden/make_golden_fusion.py); the one-launch eval
kernel (csrc/fcfm.cu) and the training-mode forward + backward (csrc/fcfm_train.cu) against the reference's outputs and
autograd gradients (tests/golden/make_golden_fusion_r2.py)."""
import os

import numpy as np
import pytest

from oracle import fusion_oracle as FO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["small", "bert22"])
def test_working_oracle_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, f"fusion_working_{name}.npz"))
    params = {k[2:]: g[k] for k in g.files if k.startswith("p:")}
    out = FO.working_forward(params, g["img"], g["word"], g["gl_img"], g["sent"])
    assert out.shape == g["out"].shape == (g["img"].shape[0], 640)
    np.testing.assert_allclose(out, g["out"], atol=2e-5, rtol=0)
    # the two LayerNorm'd pass-through blocks are exactly normalised rows (before the affine terms)
    tail = (out[:, 128:384] - params["ln_gl_image.bias"]) / params["ln_gl_image.weight"]
    np.testing.assert_allclose(tail.mean(1), 0.0, atol=1e-9)


def _module_from_fixture(g):
    import torch
    from text_guided_face_recognition_b200.models.fusion_nets import Working
    net = Working(channel_dim=256)
    state = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p:")}
    missing = net.load_state_dict(state, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys)
    return net.cuda().eval()


@pytest.mark.gpu
@pytest.mark.parametrize("conv", ["simt", "tc"])
@pytest.mark.parametrize("name", ["small", "bert22"])
def test_gpu_working_forward_matches_reference(name, conv, monkeypatch):
    """The eval forward (csrc/fcfm.cu through the models/fusion_nets.py mirror) against the reference's own output and the
    fp64 oracle; image features in both memory layouts (contiguous and IMIM's channels-last); the convolution inside the
    per-sample kernel (simt) and as the implicit tensor-core GEMM that large batches take (tc)."""
    import torch
    monkeypatch.setenv("TGFR_FCFM_CONV", conv)
    g = np.load(os.path.join(GOLDEN, f"fusion_working_{name}.npz"))
    net = _module_from_fixture(g)
    params = {k[2:]: g[k] for k in g.files if k.startswith("p:")}
    ref64 = FO.working_forward(params, g["img"], g["word"], g["gl_img"], g["sent"])
    img = torch.from_numpy(g["img"]).cuda()
    word = torch.from_numpy(g["word"]).cuda()
    gl, sent = torch.from_numpy(g["gl_img"]).cuda(), torch.from_numpy(g["sent"]).cuda()
    for im in (img, img.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)):
        for wd in (word, word.transpose(1, 2).contiguous().transpose(1, 2)):
            with torch.no_grad():                                              # verification path: the one-launch kernel
                out = net(im, wd, gl, sent).cpu().numpy()
            assert out.shape == (img.shape[0], 640)
            np.testing.assert_allclose(out, g["out"], atol=1e-4, rtol=0)       # the reference (CPU fp32)
            np.testing.assert_allclose(out, ref64, atol=1e-4, rtol=0)          # the fp64 oracle
            out_ag = net(im, wd, gl, sent)                                     # eval mode under autograd: batch-wide kernels
            assert out_ag.requires_grad
            np.testing.assert_allclose(out_ag.detach().cpu().numpy(), g["out"], atol=1e-4, rtol=0)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["contiguous", "channels_last"])
def test_gpu_working_training_matches_reference_autograd(layout):
    """Training mode (BatchNorm batch statistics): output, the gradients of all 26 parameters and of the four inputs, and
    the updated running statistics against the reference module under autograd (fusion_working_train.npz)."""
    import sys
    import torch
    sys.path.insert(0, GOLDEN)
    from make_golden_fusion_r2 import fusion_inputs
    base = np.load(os.path.join(GOLDEN, "fusion_working_bert22.npz"))
    g = np.load(os.path.join(GOLDEN, "fusion_working_train.npz"))
    net = _module_from_fixture(base).train()
    img, word, gl, sent, gout = fusion_inputs(5, 22, 9)
    it = torch.from_numpy(img).cuda()
    if layout == "channels_last":
        it = it.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    leaves = [it.requires_grad_(True)] + [torch.from_numpy(a).cuda().requires_grad_(True) for a in (word, gl, sent)]
    out = net(*leaves)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out"], atol=5e-5, rtol=0)
    out.backward(torch.from_numpy(gout).cuda())
    for t, name in zip(leaves, ("dimg", "dword", "dgl", "dsent")):
        assert _rel(t.grad.cpu().numpy(), g[name]) < 3e-4, (name, _rel(t.grad.cpu().numpy(), g[name]))
    for name, p in net.named_parameters():
        ref = g["g:" + name]
        got = p.grad.cpu().numpy()
        if np.linalg.norm(ref) < 1e-4:                    # sa.query_proj.bias: the softmax is invariant to it
            assert np.max(np.abs(got)) < 1e-4, name
        else:
            assert _rel(got, ref) < 3e-4, (name, _rel(got, ref))
    sd = net.state_dict()
    for k in g.files:
        if k.startswith("s:"):
            assert _rel(sd[k[2:]].cpu().numpy(), g[k]) < 1e-5, k
    assert int(sd["bn_img.num_batches_tracked"]) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("conv", ["simt", "tc"])
def test_gpu_working_batch_independence_and_scoring(conv, monkeypatch):
    """Size-independent property at a batch that fills the GPU: every sample's embedding equals the one computed alone
    (bit for bit in the per-sample kernel; within 2e-5 when the convolution runs as the tensor-core GEMM, whose fp16
    hi / lo split is scaled by the batch's largest entry); the fused embeddings feed the pair-cosine kernel directly
    (configs[4]: fusion -> cosine).  B = 4700 in tc mode crosses the 4096-sample chunk of the implicit GEMM."""
    import torch
    from text_guided_face_recognition_b200 import ops
    monkeypatch.setenv("TGFR_FCFM_CONV", conv)
    g = np.load(os.path.join(GOLDEN, "fusion_working_bert22.npz"))
    net = _module_from_fixture(g)
    gen = torch.Generator().manual_seed(9)
    B, T = (4700 if conv == "tc" else 600), 30
    img = torch.nn.functional.normalize(torch.randn(B, 14, 14, 256, generator=gen), dim=-1).permute(0, 3, 1, 2).cuda()
    word = torch.nn.functional.normalize(torch.randn(B, T, 256, generator=gen), dim=2).transpose(1, 2).cuda()
    gl = torch.nn.functional.normalize(torch.randn(B, 256, generator=gen), dim=1).cuda()
    sent = torch.nn.functional.normalize(torch.randn(B, 256, generator=gen), dim=1).cuda()
    with torch.no_grad():                          # the verification path runs under no_grad (utils/modules.py:129)
        out = net(img, word, gl, sent)
        assert torch.isfinite(out).all()
        for i in (0, 299, B - 1):
            one = net(img[i:i + 1], word[i:i + 1], gl[i:i + 1], sent[i:i + 1])
            if conv == "simt":
                assert torch.equal(one[0], out[i])
            else:
                assert float((one[0] - out[i]).abs().max()) < 2e-5
    params = {k[2:]: g[k] for k in g.files if k.startswith("p:")}
    ref = FO.working_forward(params, img[:4].cpu().numpy(), word[:4].cpu().numpy(), gl[:4].cpu().numpy(), sent[:4].cpu().numpy())
    np.testing.assert_allclose(out[:4].cpu().numpy(), ref, atol=1e-4, rtol=0)
    scores = ops.pair_cosine(out[:300], out[300:600])
    assert scores.shape == (300,) and bool((scores.abs() <= 1 + 1e-5).all())


def test_imim_oracle_matches_reference():
    """IMIM (reference models/models.py:380-405), eval mode: the oracle against the reference
    module's output.  Checker only -- no kernel of this row exists yet.  Also pins the data contract of SURVEY 8(a) row
    a0: unit-norm region vectors, channels-last memory."""
    g = np.load(os.path.join(GOLDEN, "imim_small.npz"))
    params = {k[2:]: g[k] for k in g.files if k.startswith("p:")}
    out = FO.imim_forward(params, g["img"])
    np.testing.assert_allclose(out, g["out"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(np.sqrt((out * out).sum(1)), 1.0, atol=1e-12)
    assert tuple(g["out_strides"]) == (50176, 1, 3584, 256)


@pytest.mark.gpu
def test_gpu_fusion_training_step_runs_on_the_mirror():
    """The fusion training step of the reference (src/fusion_bert.py:205-233): image_head -> Working -> metric_fc
    (ArcMarginProduct(640, C)) -> FocalLoss(2) -> backward, on the mirror modules; every trainable parameter of the three
    modules and both text inputs (`.requires_grad_()` leaves there) receive a finite gradient, and a second step after an
    SGD update lowers the loss (the gradients point downhill)."""
    import types
    import torch
    from text_guided_face_recognition_b200.models import losses, metrics
    from text_guided_face_recognition_b200.models.fusion_nets import Working
    from text_guided_face_recognition_b200.models.image_heading import ImageHeading
    torch.manual_seed(3)
    B, T, C = 16, 22, 40
    head = ImageHeading(types.SimpleNamespace(aux_feat_dim_per_granularity=256)).cuda().train()
    fusion = Working(channel_dim=256).cuda().train()
    metric_fc = metrics.ArcMarginProduct(640, C, s=30, m=0.5, easy_margin=False).cuda()
    criterion = losses.FocalLoss(gamma=2)
    gen = torch.Generator().manual_seed(3)
    imgs_g = torch.randn(B, 512, generator=gen).cuda()
    imgs_l = torch.randn(B, 256, 14, 14, generator=gen).cuda()
    words_emb = torch.nn.functional.normalize(torch.randn(B, T, 256, generator=gen), dim=2).transpose(1, 2).cuda().requires_grad_()
    sent_emb = torch.nn.functional.normalize(torch.randn(B, 256, generator=gen), dim=1).cuda().requires_grad_()
    label = torch.randint(0, C, (B,), generator=gen).cuda()
    params = list(head.parameters()) + list(fusion.parameters()) + list(metric_fc.parameters())
    opt = torch.optim.SGD([p for p in params], lr=0.05)
    vals = []
    for _ in range(3):
        img_feats, local_feats = head(imgs_g, imgs_l)
        output = fusion(local_feats, words_emb, img_feats, sent_emb)           # get_fusion_output, fusion_bert.py:190-203
        output = metric_fc(output, label)
        opt.zero_grad()
        loss = criterion(output, label)
        loss.backward()
        vals.append(loss.item())
        for n, p in list(head.named_parameters()) + list(fusion.named_parameters()) + list(metric_fc.named_parameters()):
            if n.startswith("imim.project_local.fc") or n.startswith("project_global.fc"):
                continue
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
        assert torch.isfinite(words_emb.grad).all() and torch.isfinite(sent_emb.grad).all()
        opt.step()
    assert vals[-1] < vals[0], vals
