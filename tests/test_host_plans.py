"""CPU-side checks of the host logic behind the C ABI's buffer sizing (no GPU, no compute calls): the two forward ->
backward layouts of the word-region kernels and how TGFR_WORDREGION_SAVE selects them, workspace / saved sizes of the
split-product front end, of IMIM and of the FCFM forward, and the Python-side switches of ops.py."""
import os

import pytest


@pytest.fixture(scope="module")
def lib():
    from text_guided_face_recognition_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def test_wordregion_saved_layouts_follow_the_switch(lib, monkeypatch):
    from text_guided_face_recognition_b200 import _lib, ops
    shape = (128, 128, 22, 196, 256)
    monkeypatch.delenv("TGFR_WORDREGION_SAVE", raising=False)
    rec = lib.tgfr_wordregion_saved_bytes(*shape, _lib.PREC_TC)
    assert ops._save_enabled()
    monkeypatch.setenv("TGFR_WORDREGION_SAVE", "1")
    assert lib.tgfr_wordregion_saved_bytes(*shape, _lib.PREC_TC) == rec
    monkeypatch.setenv("TGFR_WORDREGION_SAVE", "wu")
    wu = lib.tgfr_wordregion_saved_bytes(*shape, _lib.PREC_TC)
    assert ops._save_enabled()
    monkeypatch.setenv("TGFR_WORDREGION_SAVE", "0")
    assert not ops._save_enabled()
    # the library tells the layouts apart by size: the record layout must be strictly larger, both 256-byte multiples
    assert rec > wu > 0 and rec % 256 == 0 and wu % 256 == 0
    # records = V planes + A1|E per (face, caption, word, region); Wu layout = one plane set, nothing per region
    units, nw, Tp, Rp, D = 128 * 26, 120, 24, 208, 256
    assert rec >= units * nw * D * 2 + units * 5 * Rp * Tp * 4
    assert wu < units * nw * D * 2 + 32 * 2**20
    # the fp32 path keeps nothing
    assert lib.tgfr_wordregion_saved_bytes(*shape, _lib.PREC_FP32) == 0
    # sizes grow with the caption count of a row-sharded rank (Bq = 8 x 128)
    monkeypatch.setenv("TGFR_WORDREGION_SAVE", "1")
    assert lib.tgfr_wordregion_saved_bytes(128, 1024, 22, 196, 256, _lib.PREC_TC) > 7 * rec


def test_wordregion_plan_limits(lib):
    from text_guided_face_recognition_b200 import _lib
    # shapes outside the tensor-core plan report no buffers (the mirror then runs the fp32 kernels)
    assert lib.tgfr_wordregion_saved_bytes(8, 8, 40, 196, 256, _lib.PREC_TC) == 0       # T > 32
    assert lib.tgfr_wordregion_saved_bytes(8, 8, 22, 196, 200, _lib.PREC_TC) == 0       # D not a multiple of 64
    assert lib.tgfr_wordregion_workspace_bytes(8, 8, 30, 196, 256, _lib.PREC_TC) > 0    # T = 30 (Tp = 32) is planned


def test_split_product_and_module_buffer_sizes(lib):
    # tgfr_matmul_split: scales + hi / lo of both operands, K padded to 8 columns
    w = lib.tgfr_matmul_split_workspace_bytes(0, 196, 196, 256, 5)
    assert w >= 256 + 2 * 2 * 5 * 196 * 256 * 2 and w % 256 == 0
    assert lib.tgfr_matmul_split_workspace_bytes(1, 196, 256, 196, 5) >= 256 + 2 * 5 * (196 * 200 + 196 * 256) * 2
    # IMIM keeps fp32 activations plus their fp16 hi / lo copies; both grow linearly in the batch
    s1, s2 = lib.tgfr_imim_saved_bytes(1, 196), lib.tgfr_imim_saved_bytes(128, 196)
    assert s1 > 0 and s2 > 100 * s1 / 2 and s2 % 256 == 0
    assert lib.tgfr_imim_workspace_bytes(128, 196) > 128 * 196 * (3 * 256 + 128 + 196 + 768) * 4
    # FCFM tensor-core forward: one 4096-sample chunk of image copies at most, whatever the batch
    f1, f2, f3 = (lib.tgfr_fcfm_working_workspace_bytes(b) for b in (64, 4096, 60000))
    assert 0 < f1 < f2 == f3
    assert f2 >= 4096 * 196 * (2 * 256 * 2 + 36 * 4)
    assert lib.tgfr_fcfm_working_workspace_bytes(0) == 0


def test_python_side_switches(monkeypatch):
    from text_guided_face_recognition_b200 import ops
    monkeypatch.delenv("TGFR_FCFM_CONV", raising=False)
    assert not ops._fcfm_conv_on_tensor_cores(63) and ops._fcfm_conv_on_tensor_cores(64)
    monkeypatch.setenv("TGFR_FCFM_CONV", "simt")
    assert not ops._fcfm_conv_on_tensor_cores(60000)
    monkeypatch.setenv("TGFR_FCFM_CONV", "tc")
    assert ops._fcfm_conv_on_tensor_cores(1)
