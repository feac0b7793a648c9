"""CPU model of the round-2 video synthesis kernel (cdlnet-video_b200/csrc/cdl_tc_synthesis.cuh), restated in numpy at the
level of what the hardware is told to do: the code layout and its TMA box (group order = TMEM lane order), the filter
packing (tap -> accumulator column), the dealing of tiles to CTAs as contiguous ranges, and - lane by lane - the col2im of
the 16 warps: lane -> site permutation, shuffle sources, edge-lane reductions, footprint ring slots, the 2-rows-per-tile
flush and the full flush at the end of a run.  The result must equal conv_transpose3d.  Arithmetic is exact here (fp64):
any difference is an indexing error."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

NA, K4, CHUNK = 176, 44, 32
GROUP = K4 * CHUNK
TILE_W, XW, XPL, XRING = 128, 264, 7, 7


def groups_per_row(Qw):
    return 2 * ((Qw + 15) >> 4)


def code_site_offset(row, Qw, qw):
    return (row * groups_per_row(Qw) + 2 * (qw >> 4) + (qw & 1)) * GROUP + ((qw & 15) >> 1) * 4


def to_code(z):
    """(N,M,Qd,Qh,Qw) -> internal layout (values, not pre-biased: the bias is arithmetic, checked in test_code_layout_cpu)"""
    N, M, Qd, Qh, Qw = z.shape
    code = np.zeros(N * Qd * Qh * groups_per_row(Qw) * GROUP)
    for n in range(N):
        for qd in range(Qd):
            for qh in range(Qh):
                row = (n * Qd + qd) * Qh + qh
                for qw in range(Qw):
                    base = code_site_offset(row, Qw, qw)
                    for m in range(M):
                        code[base + (m >> 2) * CHUNK + (m & 3)] = z[n, m, qd, qh, qw]
    return code


def tma_tile(code, row, g0, nrows, Qw):
    """A operand of one tile: rows = TMEM lanes 8*gl + i (16 groups), columns = 176 subbands; zero fill out of bounds"""
    G = groups_per_row(Qw)
    A = np.zeros((128, NA))
    if row >= nrows:
        return A
    for gl in range(16):
        gidx = g0 + gl
        if gidx >= G:
            continue
        base = (row * G + gidx) * GROUP
        for i in range(8):
            for m in range(NA):
                A[8 * gl + i, m] = code[base + (m >> 2) * CHUNK + i * 4 + (m & 3)]
    return A


def pack_filters(w):
    """(M,1,7,7,7) [m, td, th, tw] -> B[half][column j][k = m]; column j of half h = row h*25 + j//7 of the (th,td) list, tap j%7"""
    M = w.shape[0]
    B = np.zeros((2, 176, NA))
    for h in range(2):
        nrows = 25 if h == 0 else 24
        for j in range(7 * nrows):
            row, tw = h * 25 + j // 7, j % 7
            th, td = row // 7, row % 7
            B[h, j, :M] = w[:, 0, td, th, tw]
    return B


def lane_of(o):
    o &= 31
    return 8 * (2 * (o >> 4) + (o & 1)) + ((o & 15) >> 1)


def synthesize(z, w, od, Fd, nctas, sweep=False):
    N, M, Qd, Qh, Qw = z.shape
    Fh, Fw = 2 * Qh, 2 * Qw
    code = to_code(z)
    B = pack_filters(w)
    tiles_w = -(-Qw // TILE_W)
    nrows = N * Qd * Qh
    T = nrows * tiles_w
    out = np.zeros((N, Fd, Fh, Fw))
    lanes = np.arange(32)
    gq, i8 = lanes >> 3, lanes & 7
    o = 16 * (gq >> 1) + 2 * i8 + (gq & 1)
    assert sorted(o) == list(range(32))
    src1, src2, srcm = [np.array([lane_of(v + d) for v in o]) for d in (1, 2, 31)]
    m1, m2, mm = (o < 31).astype(float), (o < 30).astype(float), (o > 0).astype(float)
    parts = [(0, 0, 13, 0), (13, 0, 12, 7 * 13), (25, 1, 12, 0), (37, 1, 12, 7 * 12)]   # (first row, half, rows, first column)
    Fr = tiles_w * Qh                                               # tiles of one coarse frame
    for cta in range(nctas):
        if sweep:                                                   # syn_tile, sweep = 1: a share of ONE frame, swept over all (n, qd)
            a, ln = Fr * cta // nctas, Fr * (cta + 1) // nctas - Fr * cta // nctas
            count = ln * N * Qd
        else:
            a, count = T * cta // nctas, T * (cta + 1) // nctas - T * cta // nctas
        X = np.zeros((XRING, XPL, XW))
        S = np.zeros((XRING, XPL, 4, 8))                          # seam columns: [0..2] left spill, [4..5] right spill per quadrant
        for i in range(count):
            if sweep:
                fi, j = divmod(i, ln)
                wt, qh = divmod(a + j, Qh)
                qw0 = wt * TILE_W
                n, qd = divmod(fi, Qd)
                first = j == 0 or qh == 0
                last = j == ln - 1 or qh == Qh - 1
            else:
                tau = a + i
                col, qh = divmod(tau, Qh)
                qw0 = (col % tiles_w) * TILE_W
                col //= tiles_w
                qd, n = col % Qd, col // Qd
                first = i == 0 or qh == 0
                last = i == count - 1 or qh == Qh - 1
            row = (n * Qd + qd) * Qh + qh
            if first:
                assert not X.any() and not S.any()                # the ring is clean when a run starts
            A = tma_tile(code, row, (qw0 >> 4) * 2, nrows, Qw)
            D = [A @ B[0].T, A @ B[1].T]                          # [128 lanes, 176 columns] per half
            pbase = (2 * qh) % XRING
            for q in range(4):
                for row0, half, nr, c0 in parts:
                    for r in range(nr):
                        u = D[half][32 * q:32 * q + 32, c0 + 7 * r:c0 + 7 * r + 7]     # u[lane, tap]
                        rr = row0 + r
                        th, td = rr // 7, rr % 7
                        v = [u[:, k] for k in range(7)]
                        x0 = v[3] + v[1][src1] * m1 + v[5][srcm] * mm
                        x1 = v[4] + v[2][src1] * m1 + v[0][src2] * m2 + v[6][srcm] * mm
                        n0 = v[0][src1]
                        ps = (pbase + th) % XRING
                        cb = 4 + 64 * q + 2 * o
                        np.add.at(X[ps, td], cb, x0)
                        np.add.at(X[ps, td], cb + 1, x1)
                        l0, l31 = int(np.where(o == 0)[0][0]), int(np.where(o == 31)[0][0])
                        S[ps, td, q, 0] += v[0][l0]
                        S[ps, td, q, 1] += v[1][l0]
                        S[ps, td, q, 2] += v[2][l0] + n0[l0]
                        S[ps, td, q, 4] += v[5][l31]
                        S[ps, td, q, 5] += v[6][l31]
            for pi in range(XPL if last else 2):
                pr = 2 * qh + pi
                for td in range(XPL):
                    gd, gh = 2 * qd + td - od, pr - 3
                    for c4 in range(XW // 4):
                        gw = 2 * qw0 - 4 + 4 * c4
                        cell = X[pr % XRING, td, 4 * c4:4 * c4 + 4].copy()
                        X[pr % XRING, td, 4 * c4:4 * c4 + 4] = 0
                        if (c4 & 15) == 0 and c4 < 64:             # left spill of quadrant c4 / 16: columns 64 q + 1..3
                            cell[1:4] += S[pr % XRING, td, c4 >> 4, 0:3]
                            S[pr % XRING, td, c4 >> 4, 0:4] = 0
                        elif (c4 & 15) == 1 and c4 >= 17:          # right spill of quadrant (c4 - 17) / 16: columns 64 q + 68, 69
                            cell[0:2] += S[pr % XRING, td, (c4 - 17) >> 4, 4:6]
                            S[pr % XRING, td, (c4 - 17) >> 4, 4:8] = 0
                        if 0 <= gd < Fd and 0 <= gh < Fh and gw >= 0 and gw + 4 <= Fw:
                            out[n, gd, gh, gw:gw + 4] += cell
        assert not X.any() and not S.any()
    return out


@pytest.mark.parametrize("N,M,Qd,Qh,Qw,nctas", [(1, 5, 2, 3, 20, 3), (2, 3, 1, 4, 16, 5), (1, 4, 2, 2, 136, 4), (1, 2, 3, 5, 6, 2), (1, 2, 1, 3, 130, 7)])
def test_model_equals_conv_transpose3d(N, M, Qd, Qh, Qw, nctas):
    rng = np.random.default_rng(0)
    z = rng.integers(-3, 4, size=(N, M, Qd, Qh, Qw)).astype(np.float64)
    w = rng.integers(-2, 3, size=(M, 1, 7, 7, 7)).astype(np.float64)
    ref = F.conv_transpose3d(torch.from_numpy(z), torch.from_numpy(w), stride=2, padding=3, output_padding=1)[:, 0].numpy()
    got = synthesize(z, w, od=3, Fd=2 * Qd, nctas=nctas)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
    assert np.array_equal(synthesize(z, w, od=3, Fd=2 * Qd, nctas=nctas, sweep=True), ref)       # frame-synchronous tile order


def test_model_temporal_slab_offsets():
    """slab plan (cdl_desc_t.halo_front): od = Pd/2 - halo_front shifts the fine frames; the uncropped-in-time synthesis of a
    slab equals the matching frames of the transposed convolution without temporal cropping"""
    rng = np.random.default_rng(1)
    Qd, Qh, Qw, M = 3, 2, 16, 3
    z = rng.integers(-3, 4, size=(1, M, Qd, Qh, Qw)).astype(np.float64)
    w = rng.integers(-2, 3, size=(M, 1, 7, 7, 7)).astype(np.float64)
    full = F.conv_transpose3d(torch.from_numpy(z), torch.from_numpy(w), stride=2, padding=0, output_padding=1)[:, 0].numpy()   # frames -3 .. 2*Qd+3
    hf, hb = 3, 2                                                 # interior slab: resident fine frames [-3, 2*Qd + 2)
    got = synthesize(z, w, od=3 - hf, Fd=2 * Qd + hf + hb, nctas=2)
    ref = full[:, 0:2 * Qd + hf + hb, 3:3 + 2 * Qh, 3:3 + 2 * Qw]
    assert np.array_equal(got, ref)
