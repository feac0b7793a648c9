"""The experimental 2-D (stride-1) tensor-core analysis kernel (cdlnet-video_b200/csrc/cdl_tc2_analysis.cuh), restated in numpy.

With stride 1 the 8-float windows of sites w and w+4 start 16 bytes apart, so 8 sites of EQUAL RESIDUE w mod 4 form
a legal K-major core matrix.  The kernel stages a 40 x 22 x C halo tile with TMA, two warps copy it into four operand
copies (copy rho shifted left by rho floats, rounded to tf32) and the tensor core reads each copy through the
overlapping descriptor (LBO = 16 B, SBO = 144 B) the video kernel uses.  This test rebuilds exactly what the hardware
is told to read - TMA box with zero fill, the shifter's indexing, the descriptor's core-matrix addressing, the filter
packing of k_pack_tc2_analysis, the epilogue's lane -> site map - and checks the resulting GEMM against
nn.Conv2d(C, M, 7, padding=3, bias=False) (reference model/net.py:32), ragged borders included.
Constants are parsed from the header so that a change there fails here."""
import os
import re

import numpy as np
import pytest
import torch

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cdlnet-video_b200", "csrc", "cdl_tc2_analysis.cuh")
SRC = open(HDR).read()


def _const(name):
    m = re.search(r"constexpr int %s = ([^;,]+)[;,]" % name, SRC)
    assert m, name
    return m.group(1).strip()


TH, TW, RW, ROWS, SW = 16, 32, 36, 22, 40


def test_geometry_constants():
    assert _const("kTH") == "16" and "kTW = 32" in SRC
    assert _const("kRW") == "36" and _const("kSW") == "40" and _const("kRows") == "kTH + kP - 1"
    assert "make_smem_desc_kmajor_noswz(smem_u32(sOp), 16, kRW * 4)" in SRC           # LBO 16 B, SBO 144 B
    assert "tma_load_4d(sStage + b * L.stage_pitch, &rmap, w0 - 4, h0 - (kP / 2), 0, n, &sfull[b]);" in SRC
    assert "if (k >= 0 && k < kRW) o[rho * cp - rho] = v;" in SRC
    assert "(uint32_t)((c * kRows + th) * kRW * 4)" in SRC
    # every descriptor start is 16-byte aligned: copy pitch and row pitch are multiples of 16 bytes
    for C in (1, 2, 3):
        assert (C * ROWS * RW * 4) % 16 == 0 and (RW * 4) % 16 == 0
    # TMA box: innermost extent a multiple of 16 bytes, start column w0 - 4 a multiple of 4 floats
    assert (SW * 4) % 16 == 0 and (TW - 4) % 4 == 0


def _stage(r, h0, w0):
    """The TMA box: [C][22 rows][40 floats], image column w0 - 4 + x, row h0 - 3 + y, zero fill outside the image."""
    C, H, W = r.shape
    out = np.zeros((C, ROWS, SW), r.dtype)
    for y in range(ROWS):
        for x in range(SW):
            h, w = h0 - 3 + y, w0 - 4 + x
            if 0 <= h < H and 0 <= w < W:
                out[:, y, x] = r[:, h, w]
    return out.reshape(-1)


def _shift(stage, C):
    """The shifter warps: op[rho][row][k] = stage[row][k + rho], k < 36 (flat indexing as in the kernel)."""
    cp = C * ROWS * RW
    op = np.full(4 * cp, np.nan, stage.dtype)
    for idx in range(C * ROWS * SW):
        row, x = divmod(idx, SW)
        if x < RW + 3:
            for rho in range(4):
                k = x - rho
                if 0 <= k < RW:
                    op[row * RW + x + rho * cp - rho] = stage[idx]
    assert not np.isnan(op).any()                              # every operand float is written
    return op


def _pack(w, Ng):
    """k_pack_tc2_analysis: (M,C,7,7) -> [7C k-steps][Ng/8 groups][2 k-chunks][8 rows][4]; column 0 is the zero pad."""
    M, C = w.shape[:2]
    out = np.zeros(7 * C * Ng * 8, w.dtype)
    for i in range(out.size):
        e, r8, kc, grp, ks = i % 4, (i // 4) % 8, (i // 32) % 2, (i // 64) % (Ng // 8), i // (Ng * 8)
        m, j, c, th = grp * 8 + r8, kc * 4 + e, ks // 7, ks % 7
        if m < M and j > 0:
            out[i] = w[m, c, th, j - 1]
    return out


def _b_matrix(pack, Ng, ks):
    """What the B descriptor (start + ks*Ng*32 B, LBO 128 B, SBO 256 B) reads: [Ng rows][8 k]."""
    base = ks * Ng * 8
    B = np.zeros((Ng, 8), pack.dtype)
    for n in range(Ng):
        for j in range(8):
            B[n, j] = pack[base + (n // 8) * 64 + (j // 4) * 32 + (n % 8) * 4 + (j % 4)]
    return B


def _a_matrix(op, C, rho, c, th):
    """What the A descriptor (copy rho, row c*22 + th; LBO 16 B, SBO 144 B) reads: [128 rows][8 k]."""
    start = rho * C * ROWS * RW + (c * ROWS + th) * RW
    A = np.zeros((128, 8), op.dtype)
    for m in range(128):
        g, i = divmod(m, 8)
        for j in range(8):
            A[m, j] = op[start + g * RW + 4 * i + (j & 3) + 4 * (j >> 2)]
    return A


@pytest.mark.parametrize("C,M,H,W,h0,w0", [(1, 32, 40, 72, 0, 0), (3, 64, 40, 72, 32, 64), (3, 20, 21, 44, 16, 32), (2, 64, 16, 32, 0, 0)])
def test_tile_gemm_equals_conv2d(C, M, H, W, h0, w0):
    rng = np.random.default_rng(C * 1000 + M + h0 + w0)
    r = rng.integers(-4, 5, size=(C, H, W)).astype(np.float32)          # exactly representable: any mismatch is indexing
    w = rng.integers(-4, 5, size=(M, C, 7, 7)).astype(np.float32)
    want = torch.nn.functional.conv2d(torch.from_numpy(r)[None], torch.from_numpy(w), padding=3)[0].numpy()
    Ng = (M + 15) // 16 * 16
    op = _shift(_stage(r, h0, w0), C)
    pack = _pack(w, Ng)
    for rho in range(4):
        D = np.zeros((128, Ng), np.float32)                     # accumulator rho: TMEM lane = MMA row, column = subband
        for c in range(C):
            for th in range(7):
                D += _a_matrix(op, C, rho, c, th) @ _b_matrix(pack, Ng, c * 7 + th).T
        for lane in range(128):                                 # epilogue: lane -> (row, i); site w = w0 + 4 i + rho
            hrow, i = divmod(lane, 8)
            h, ww = h0 + hrow, w0 + 4 * i + rho
            if h < H and ww < W:
                assert np.array_equal(D[lane, :M], want[:, h, ww]), (rho, lane)
            assert not D[lane, M:].any()


def test_shared_memory_budget():
    """SmemLayout for the largest supported geometry (C = 3, N = 64) fits one SM's 227 KB with room to spare."""
    C, Ng = 3, 64
    a128 = lambda v: (v + 127) // 128 * 128
    b = 7 * C * Ng * 32
    op = a128(b)
    stage = a128(op + 2 * 4 * C * ROWS * RW * 4)
    total = stage + 2 * a128(C * ROWS * SW * 4) + 2 * 64 * 4 + 128
    assert total == 43008 + 76032 + 2 * 10624 + 512 + 128
    assert total < 200 * 1024


# ------------------------------------------------------------------------------------------------------------
# residual synthesis (cdl_tc2_synthesis.cuh): GEMM + col2im
# ------------------------------------------------------------------------------------------------------------
SHDR = os.path.join(os.path.dirname(HDR), "cdl_tc2_synthesis.cuh")
SSRC = open(SHDR).read()


def test_synthesis_constants():
    assert "constexpr int kSTH = 4, kSTW = 32;" in SSRC and "constexpr int kSN = 176;" in SSRC
    assert "__shfl_sync(0xffffffffu, __uint_as_float(v[tw]), (lane - tw) & 31);" in SSRC
    assert "const int gw = w0 - kP / 2 + fx;" in SSRC and "const int gh0 = h0 - kP / 2 + Y0;" in SSRC
    # TMEM budget: two accumulators + two A slots
    assert 2 * 176 + 2 * 64 <= 512


def _pack_syn(w, Kg):
    """k_pack_tc2_synthesis: ConvTranspose2d weight (M,C,7,7) -> [Kg/8 k-steps][22 groups][2][8][4]."""
    M, C = w.shape[:2]
    out = np.zeros((Kg // 8) * 176 * 8, w.dtype)
    for i in range(out.size):
        e, r8, kc, grp, ks = i % 4, (i // 4) % 8, (i // 32) % 2, (i // 64) % 22, i // (176 * 8)
        n, m = grp * 8 + r8, ks * 8 + kc * 4 + e
        row, tw = n >> 3, n & 7
        c, th = row // 7, row % 7
        if row < 7 * C and tw < 7 and m < M:
            out[i] = w[m, c, th, tw]
    return out


# ------------------------------------------------------------------------------------------------------------
# col2im of k_tc2_synthesis: write-once private footprints, overlap-add in the flush
# ------------------------------------------------------------------------------------------------------------
V2SRC = SSRC


def test_synthesis_source_matches_model():
    # flush: a thread walks footprint column fx of channel fc; pc = pv + fc * kP * kFPitch + fx
    assert "const float* pc = pv + fc * kP * kFPitch + fx;" in V2SRC
    assert "if (th >= 0 && th < kP) v += pc[r * kPrivWarp + th * kFPitch];" in V2SRC
    assert "row[lane] = own;" in V2SRC and "if (lane < kP - 1) row[32 + lane] = spill;" in V2SRC
    assert "put_row(&cur[8 * q], priv + (4 * g + q) * kFPitch);" in V2SRC
    assert "tmem_ld32(dcol + 32 * (g + 1), nxt);" in V2SRC and "tmem_ld8(dcol + 160," in V2SRC
    # shared memory: filters + two private buffers + barriers
    assert 45056 + 2 * 4 * 21 * 40 * 4 + 128 < 100 * 1024


@pytest.mark.parametrize("C,M,H,W", [(1, 32, 9, 40), (3, 64, 12, 72), (2, 20, 7, 44)])
def test_synthesis_private_footprints_equal_conv_transpose2d(C, M, H, W):
    rng = np.random.default_rng(C * 100 + M + 7)
    z = rng.integers(-4, 5, size=(M, H, W)).astype(np.float32)
    w = rng.integers(-4, 5, size=(M, C, 7, 7)).astype(np.float32)
    want = torch.nn.functional.conv_transpose2d(torch.from_numpy(z)[None], torch.from_numpy(w), padding=3)[0].numpy()
    nrows = 7 * C
    out = np.zeros((C, H, W), np.float32)
    lanes = np.arange(32)
    for h0 in range(0, H, 4):
        for w0 in range(0, W, 32):
            A = np.zeros((128, M), np.float32)
            for lane in range(128):
                r, x = divmod(lane, 32)
                if h0 + r < H and w0 + x < W:
                    A[lane] = z[:, h0 + r, w0 + x]
            D = np.zeros((128, 176), np.float32)               # column (c*7 + th)*8 + tw (packing checked above)
            for c in range(C):
                for th in range(7):
                    D[:, (c * 7 + th) * 8:(c * 7 + th) * 8 + 7] = A @ w[:, c, th, :]
            priv = np.full((4, 21, 40), np.nan, np.float32)     # NaN = never written: the flush must not read such a cell
            for r in range(4):
                # the drain order of the kernel: 32-column loads g = 0..4 (rows 4g..4g+3), then row 20
                rows = [4 * g + q for g in range(5) if 4 * g < nrows for q in range(4) if 4 * g + q < nrows]
                if 20 < nrows:
                    rows.append(20)
                assert rows == list(range(nrows))
                for R in rows:
                    v = D[32 * r:32 * r + 32, 8 * R:8 * R + 8]
                    own, spill = v[:, 0].copy(), np.zeros(32, np.float32)
                    for tw in range(1, 7):
                        w_ = v[(lanes - tw) & 31, tw]
                        own += np.where(lanes >= tw, w_, 0)
                        spill += np.where(lanes < tw, w_, 0)
                    priv[r, R, :32] = own
                    priv[r, R, 32:38] = spill[:6]
            for c in range(C):                                  # the flush: at most 4 private rows meet in one output row
                for y in range(10):
                    for x in range(38):
                        v = np.float32(0)
                        for r in range(4):
                            th = y - r
                            if 0 <= th < 7:
                                v += priv[r, c * 7 + th, x]
                        gh, gw = h0 - 3 + y, w0 - 3 + x
                        if 0 <= gh < H and 0 <= gw < W:
                            out[c, gh, gw] += v
    assert not np.isnan(out).any() and np.array_equal(out, want)


# ------------------------------------------------------------------------------------------------------------
# 3-term analysis (cdl_tc2_analysis_x3.cuh, precision "tf32x3"): k_tc2_analysis with eight operand copies
# ------------------------------------------------------------------------------------------------------------
def test_analysis_x3_is_the_validated_kernel_plus_three_edits():
    x3 = open(os.path.join(os.path.dirname(HDR), "cdl_tc2_analysis_x3.cuh")).read()
    body = lambda t, name: t[t.index("__global__ void __launch_bounds__(kThreads, 1) " + name):]
    a, b = body(SRC, "k_tc2_analysis(").splitlines(), body(x3, "k_tc2_analysis_x3(").splitlines()
    import difflib
    changed = [l for l in difflib.unified_diff(a, b, lineterm="", n=0) if l[:1] in "+-" and l[:3] not in ("+++", "---")]
    assert 10 <= len(changed) <= 40, len(changed)              # a handful of edited lines, everything else identical
    assert "if (k >= 0 && k < kRW) { o[rho * cp - rho] = v; o[(4 + rho) * cp - rho] = vl; }" in x3
    assert "adesc0 + (uint64_t)((aoff + 4 * L.copy_pitch) >> 4), bdesc0 + (uint64_t)ks * bstep, idesc, 1);" in x3
    assert "adesc0 + (uint64_t)(aoff >> 4), bdesc_lo + (uint64_t)ks * bstep, idesc, 1);" in x3
    # shared memory: two filter banks + ONE buffer of eight copies + two staging buffers
    C, Ng = 3, 64
    a128 = lambda v: (v + 127) // 128 * 128
    total = a128(a128(2 * 7 * C * Ng * 32) + 8 * C * ROWS * RW * 4) + 2 * a128(C * ROWS * SW * 4) + 512 + 128
    assert total == 86016 + 76032 + 21248 + 640 and total <= 227 * 1024
    # every descriptor start stays inside the 14-bit (16-byte units) address field
    assert (86016 + 8 * C * ROWS * RW * 4) // 16 < 2 ** 14
