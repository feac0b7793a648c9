"""The analysis kernel's implicit-im2col operand (cdlnet-video_b200/csrc/cdl_tc_analysis.cuh), restated in numpy.

The kernel never builds the im2col matrix: the tensor core reads it from a TMA-loaded halo tile through an
OVERLAPPING K-major shared-memory descriptor (LBO = 16 B, SBO = 144 B).  This test rebuilds, for one CTA of a pair,
exactly what the hardware is told to read - tile boxes as the 5-D tensor map delivers them (h-parity major, zero
fill outside the clip, rank 1 from the copy shifted by two floats), core-matrix addressing as the descriptor encodes
it - and checks it against the textbook im2col of a stride-2, pad-3 7x7x7 correlation (model/net.py:137-139).
Constants are parsed from the header so that a change there fails here."""
import os
import re

import numpy as np
import pytest

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cdlnet-video_b200", "csrc", "cdl_tc_analysis.cuh")


def _const(name):
    m = re.search(r"constexpr int %s = ([^;]+);" % name, open(HDR).read())
    assert m, name
    return m.group(1).strip()


def test_geometry_constants():
    assert _const("kATile") == "16" and _const("kARW") == "36" and _const("kARD") == "7"
    assert _const("kARows") == "kATile + 3"
    assert _const("kABoxPitch") == "4800"
    src = open(HDR).read()
    assert "make_smem_desc_kmajor_noswz(smem_u32(sR), 16, kARW * 4)" in src          # LBO 16 B, SBO 144 B
    assert "const int w0 = rank ? 2 * qw0 : 2 * qw0 - 4" in src
    assert "const int fe = 2 * qh0 - 3, fo = 2 * qh0 - 2;" in src


def _tile_boxes(r, rank, qd, qh0, qw0):
    """What the two TMA loads of one tile put into shared memory: [h-parity][d 7][rows 19][w 36], zero filled."""
    Fd, Fh, Fw = r.shape
    src = r
    if rank:                                                   # the odd-parity CTA reads the copy shifted right by two floats
        src = np.zeros((Fd, Fh, Fw + 4), r.dtype)
        src[:, :, 2:Fw + 2] = r
    w0 = 2 * qw0 if rank else 2 * qw0 - 4
    out = np.zeros((2, 7, 19, 36), r.dtype)
    for hp, f0 in enumerate((2 * qh0 - 3, 2 * qh0 - 2)):       # box rows: fine h = f0, f0 + 2, ...
        for d in range(7):
            for hh in range(19):
                fd, fh = 2 * qd - 3 + d, f0 + 2 * hh
                for w in range(36):
                    fw = w0 + w
                    if 0 <= fd < Fd and 0 <= fh < Fh and 0 <= fw < src.shape[2]:
                        out[hp, d, hh, w] = src[fd, fh, fw]
    return out.reshape(2, -1)


@pytest.mark.parametrize("rank,qd,qh0,qw0", [(0, 0, 0, 0), (1, 0, 0, 0), (0, 2, 16, 16), (1, 3, 16, 0), (1, 1, 0, 16)])
def test_descriptor_reads_the_im2col_matrix(rank, qd, qh0, qw0):
    rng = np.random.default_rng(rank * 100 + qd * 10 + qh0 + qw0)
    Fd, Fh, Fw = 8, 40, 72                                     # ragged on purpose: tile rows/cols run past the clip
    r = rng.standard_normal((Fd, Fh, Fw)).astype(np.float32)
    box = _tile_boxes(r, rank, qd, qh0, qw0)
    for ks in range(49):                                       # one MMA K-step per (td, th) row
        td, th = divmod(ks, 7)
        start = (td * 19 + (th >> 1)) * 36                     # descriptor start inside the h-parity box (floats)
        for m in range(128):                                   # MMA row = TMEM lane = site
            g, i = divmod(m, 8)                                # 8-site core-matrix group (SBO = 36 floats), row in group (16 B)
            qh, qw = qh0 + g, qw0 + 2 * i + rank
            for j in range(8):                                 # K inside the step: two 16-byte halves, LBO = 4 floats
                got = box[th & 1, start + g * 36 + 4 * i + (j & 3) + 4 * (j >> 2)]
                fd, fh, fw = 2 * qd - 3 + td, 2 * qh - 3 + th, 2 * qw - 4 + j      # column 0 meets a zero filter; 1..7 = tw 0..6
                want = r[fd, fh, fw] if (0 <= fd < Fd and 0 <= fh < Fh and 0 <= fw < Fw) else 0.0
                assert got == want, (ks, m, j)
