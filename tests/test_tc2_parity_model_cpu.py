"""CPU model of the 2-D tensor-core family's ARITHMETIC (scripts/tc2_emulate_cpu.py): analysis with r and A_k rounded to
tf32 (round-to-nearest, ties away, as cvt.rna / the kernels' integer rounding), residual synthesis with z and B_k rounded,
fp32 accumulation, exact final D z - applied to the reference-generated golden vectors of the geometries those kernels
cover.  The bar is north_star's: max|xhat - reference| <= 1e-4.  The model reproduced the GPU-measured error of
tests/test_tc2_gpu.py::test_forward_parity_vs_oracle_cfg1b_like (3.19e-5 predicted, 3.11e-5 measured on B200), so this
test says on the CPU what tests/test_zz_golden_tc2_gpu.py will see on the GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

from tc2_emulate_cpu import forward, tf32          # noqa: E402
from util import load_case                           # noqa: E402


def test_tf32_rounding_is_round_to_nearest_ties_away():
    import torch
    x = torch.tensor([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -11 - 2.0 ** -20, -(1.0 + 2.0 ** -11), 3.0e-3, 0.0])
    r = tf32(x)
    assert r[0] == 1.0 and r[1] == 1.0 + 2.0 ** -10 and r[2] == 1.0 and r[3] == -(1.0 + 2.0 ** -10) and r[5] == 0.0
    assert ((r.view(torch.int32) & 0x1fff) == 0).all()                     # 13 low mantissa bits cleared
    assert (r - x).abs().max() <= 2.0 ** -11 * x.abs().max()


@pytest.mark.parametrize("name,bound", [("cdlnet2d_jdd_s1_w4", 7e-5), ("gdlnet_s1_c3", 4e-5), ("cdlnet2d_nonadaptive", 1e-4)])
def test_predicted_parity_on_reference_golden_vectors(name, bound):
    d = load_case(name)
    xhat, z = forward(d)
    ex = np.abs(xhat.numpy() - d["xhat"]).max()
    assert ex <= bound <= 1e-4, ex


@pytest.mark.parametrize("name", ["cdlnet2d_jdd_s1_w4", "gdlnet_s1_c3", "cdlnet2d_nonadaptive"])
def test_three_term_analysis_model(name):
    """The accurate mode (cdl_tc2_analysis_x3.cuh, precision "tf32x3"): u = r_hi W_hi + r_lo W_hi + r_hi W_lo.
    With it the remaining error is the residual synthesis's; it must not be worse than the single-pass analysis."""
    d = load_case(name)
    e1 = np.abs(forward(d)[0].numpy() - d["xhat"]).max()
    e3 = np.abs(forward(d, analysis="3term")[0].numpy() - d["xhat"]).max()
    assert e3 <= 7e-5 and e3 <= e1 * 1.05, (e1, e3)
