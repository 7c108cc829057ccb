"""FCFM fusion net `Working` (reference models/fusion_nets.py:217-258): the numpy oracle against
fixtures generated from the reference module in eval mode (tests/golden/make_golden_fusion.py).  CPU only -- the CUDA
kernel of this row is not built yet; this pins the checker it will be built against."""
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
