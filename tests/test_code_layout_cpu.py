"""The plan-internal code layout of the video tensor-core path (code_site_offset / code_floats / code_enc in
cdlnet-video_b200/csrc/cdl_tc_analysis.cuh), restated in numpy: it must be a bijection of (row, qw, subband) onto the
buffer; the 8 same-parity sites of a 16-site block must form contiguous 128-byte chunks of [8 sites][4 subbands] (one
K-major UMMA core matrix - what lets the synthesis kernel's TMA box land as a legal tcgen05 operand, and what makes the
analysis epilogue's 128-bit accesses fill whole cache lines); and the pre-biased word encoding must make the tensor core's
truncation equal round-to-nearest(ties away) tf32 while staying exactly invertible.
The header text is parsed for the constants so that a change there fails here."""
import os
import re

import numpy as np
import pytest

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cdlnet-video_b200", "csrc", "cdl_tc_analysis.cuh")
NA, CHUNK = 176, 32
GROUP = (NA // 4) * CHUNK


def groups_per_row(Qw):
    return 2 * ((Qw + 15) >> 4)


def code_site_offset(row, Qw, qw):
    return (row * groups_per_row(Qw) + 2 * (qw >> 4) + (qw & 1)) * GROUP + ((qw & 15) >> 1) * 4


def test_header_constants_match():
    src = open(HDR).read()
    assert re.search(r"constexpr int kNA = 176;", src)
    assert re.search(r"constexpr int kCodeChunk = 32;", src)
    assert re.search(r"constexpr int kCodeK4 = kNA / 4;", src)
    assert re.search(r"constexpr int kCodeGroup = kCodeK4 \* kCodeChunk;", src)
    assert re.search(r"constexpr uint32_t kCodeBias = 0x1000u;", src)
    assert "return 2 * ((Qw + 15) >> 4);" in src
    assert "(row * (size_t)code_groups_per_row(Qw) + (size_t)(2 * (qw >> 4) + (qw & 1))) * kCodeGroup + (size_t)(((qw & 15) >> 1) * 4)" in src


@pytest.mark.parametrize("rows,Qw", [(3, 8), (2, 20), (5, 22), (1, 128), (2, 7), (2, 960 // 8)])
def test_layout_is_a_bijection(rows, Qw):
    total = rows * groups_per_row(Qw) * GROUP
    seen = np.zeros(total, dtype=np.int32)
    for row in range(rows):
        for qw in range(Qw):
            base = code_site_offset(row, Qw, qw)
            for m in range(NA):
                off = base + (m >> 2) * CHUNK + (m & 3)
                assert 0 <= off < total
                seen[off] += 1
    assert seen.max() == 1
    assert seen.sum() == rows * Qw * NA                       # the rest is the padding of ragged rows


def test_core_matrices():
    Qw = 64
    for b in range(Qw // 16):
        for par in (0, 1):
            sites = [16 * b + 2 * i + par for i in range(8)]
            offs0 = code_site_offset(0, Qw, sites[0])
            assert offs0 % GROUP == 0                          # a group starts at a 5632-byte boundary
            for k4 in range(NA // 4):                          # chunk k4 = [8 sites][4 subbands], 128 contiguous bytes
                offs = [code_site_offset(0, Qw, s) + k4 * CHUNK + e for s in sites for e in range(4)]
                assert offs == list(range(offs0 + k4 * CHUNK, offs0 + (k4 + 1) * CHUNK))
    # the two groups of a block are adjacent (one 11264-byte L2 prefetch per tile row of the analysis kernel)
    assert code_site_offset(0, Qw, 17) - code_site_offset(0, Qw, 16) == GROUP


def test_prebiased_words():
    """word = bits(z) + 0x1000:  truncate(word) to tf32 == rna_tf32(z)  and  word - 0x1000 == bits(z)"""
    rng = np.random.default_rng(0)
    z = np.concatenate([rng.standard_normal(100000).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 3, 100000).astype(np.float32),
                        np.array([0.0, -0.0, 1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -11 - 2.0 ** -23, -1.0 - 2.0 ** -11, 3.4e-38], dtype=np.float32)])
    bits = z.view(np.uint32)
    word = bits + np.uint32(0x1000)
    trunc = (word & np.uint32(0xFFFFE000)).view(np.float32)
    # round to nearest, ties away from zero, on the magnitude: add half an ulp of the 10-bit mantissa, drop 13 bits
    mag = np.abs(z.astype(np.float64))
    e = np.floor(np.log2(np.where(mag > 0, mag, 1.0)))
    ulp = 2.0 ** (e - 10)
    rna = np.sign(z) * np.floor(mag / ulp + 0.5) * ulp
    normal = mag >= 2.0 ** -120
    assert np.array_equal(trunc[normal].astype(np.float64), rna[normal])
    assert np.array_equal((word - np.uint32(0x1000)).view(np.float32).view(np.uint32), bits)       # exactly invertible
    assert trunc[z == 0].tolist() == [0.0, -0.0] or np.all(trunc[z == 0] == 0)                      # zero stays zero for the tensor core
