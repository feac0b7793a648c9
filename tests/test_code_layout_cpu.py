"""The plan-internal "quad-blocked channels-last" code layout of the tensor-core path (code_site_offset /
code_floats in cdlnet-video_b200/csrc/cdl_tc_analysis.cuh), restated in numpy: it must be a bijection of
(row, qw, subband) onto the buffer, keep every 8-site block of a row contiguous, and put 4 same-parity sites x
8 subbands into one 128-byte line (what makes the kernels' 256-bit accesses fill whole cache lines).
The header text is parsed for the constants so that a change there fails here."""
import os
import re

import numpy as np
import pytest

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cdlnet-video_b200", "csrc", "cdl_tc_analysis.cuh")
NA = 176


def code_site_offset(row, Qw, qw):
    Qw8 = (Qw + 7) >> 3
    return ((row * Qw8 + (qw >> 3)) * 2 + (qw & 1)) * (NA // 8) * 32 + ((qw & 7) >> 1) * 8


def test_header_constants_match():
    src = open(HDR).read()
    assert re.search(r"constexpr int kNA = 176;", src)
    assert re.search(r"constexpr int kCodeBlk = 32;", src)
    assert re.search(r"constexpr int kCodeGroup = \(kNA / 8\) \* kCodeBlk;", src)
    assert "((row * Qw8 + (size_t)(qw >> 3)) * 2 + (size_t)(qw & 1)) * kCodeGroup + (size_t)(((qw & 7) >> 1) * 8)" in src


@pytest.mark.parametrize("rows,Qw", [(3, 8), (2, 20), (5, 22), (1, 128), (2, 7)])
def test_layout_is_a_bijection(rows, Qw):
    Qw8 = (Qw + 7) >> 3
    total = rows * Qw8 * 8 * NA
    seen = np.zeros(total, dtype=np.int32)
    for row in range(rows):
        for qw in range(Qw):
            base = code_site_offset(row, Qw, qw)
            for m in range(NA):
                off = base + (m >> 3) * 32 + (m & 7)
                assert 0 <= off < total
                seen[off] += 1
    assert seen.max() == 1
    assert seen.sum() == rows * Qw * NA                       # the rest is the padding of ragged rows


def test_lines_and_blocks():
    Qw = 64
    for qw0 in range(0, Qw, 8):                               # an 8-site block of a row is one contiguous 8*176-float run
        offs = sorted(code_site_offset(0, Qw, qw0 + i) + (m >> 3) * 32 + (m & 7) for i in range(8) for m in range(NA))
        assert offs == list(range(offs[0], offs[0] + 8 * NA)) and offs[0] % (8 * NA) == 0
    for par in (0, 1):                                        # 4 same-parity sites x 8 subbands = one 128-byte line
        for b in range(NA // 8):
            line = {(code_site_offset(0, Qw, 8 + 2 * k + par) + b * 32 + j) * 4 // 128 for k in range(4) for j in range(8)}
            assert len(line) == 1
