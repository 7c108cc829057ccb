"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
exactly what include/tgfr_b200.h declares; the product path refuses CPU tensors loudly."""
import ctypes
import os
import re
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tgfr_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    from text_guided_face_recognition_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.LIB_PATH


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tgfr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("tgfr_wordregion_fwd", "tgfr_wordregion_bwd", "tgfr_cosine_scores_fwd", "tgfr_pair_ce_stats",
                 "tgfr_cos_logits_fwd", "tgfr_arc_margin_apply", "tgfr_arc_margin_bwd", "tgfr_ce_rows_stats",
                 "tgfr_mag_margin_fwd", "tgfr_attention_fwd"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in tgfr_b200.h but not exported"


def test_python_signatures_cover_the_header(lib_path):
    from text_guided_face_recognition_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.tgfr_version() >= 100
    assert lib.tgfr_wordregion_workspace_bytes(16, 16, 18, 196, 256, _lib.PREC_FP32) == 0


def test_header_arity_matches_ctypes(lib_path):
    from text_guided_face_recognition_b200 import _lib
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(args), f"{name}: header has {n} parameters, ctypes binding {len(args)}"


def test_no_cpu_fallback(lib_path):
    """CPU tensors must raise, never silently compute on the host."""
    from text_guided_face_recognition_b200 import _lib
    from text_guided_face_recognition_b200.models import losses, metrics
    args = types.SimpleNamespace(en_type="BERT", bert_words_num=7, CUDA=False,
                                 TRAIN=types.SimpleNamespace(SMOOTH=types.SimpleNamespace(GAMMA1=4.0, GAMMA2=5.0, GAMMA3=10.0)))
    img = torch.randn(4, 16, 3, 3)
    words = torch.randn(4, 16, 5)
    with pytest.raises(_lib.TgfrError):
        losses.words_loss(img, words, torch.arange(4), None, None, 4, args)
    with pytest.raises(_lib.TgfrError):
        losses.global_loss(torch.randn(4, 16), torch.randn(4, 16))
    with pytest.raises(_lib.TgfrError):
        metrics.ArcMarginProduct(16, 10)(torch.randn(4, 16), torch.arange(4))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "text_guided_face_recognition_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} references oracle/"


def test_dropin_import_as_top_level_models():
    """`sys.path.insert(0, <pkg dir>)` + `from models.losses import ...` (src/train_encoders_bert.py:15-25)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); "
            "from models.losses import sent_loss, words_loss, CMPLoss, ClipLoss, global_loss; "
            "from models import metrics, losses; from models.attention import func_attention; "
            "from models.magface import MagLinear, MagLoss; "
            "assert metrics.ArcMarginProduct and metrics.AddMarginProduct and metrics.SphereProduct and metrics.AdaFace; "
            "print('ok')") % os.path.join(ROOT, "text_guided_face_recognition_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
