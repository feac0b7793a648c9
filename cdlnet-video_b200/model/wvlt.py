"""Wavelet filter banks with the reference's names (reference model/wvlt.py:5-42).  The reference asks PyWavelets for the
1-D bank at run time; the one wavelet its hot-path neighbour uses (`nle_mad`: 'bior4.4') is baked in here so that the
noise-level estimate needs no extra dependency.  Other names go to pywt when it is installed."""
import torch

# pywt.Wavelet('bior4.4').filter_bank = (dec_lo, dec_hi, rec_lo, rec_hi): the CDF 9/7 pair scaled by sqrt(2), zero-padded to 10 taps
_BIOR44 = (
    (0.0, 0.03782845550726404, -0.023849465019556843, -0.11062440441843718, 0.37740285561283066, 0.8526986790088938,
     0.37740285561283066, -0.11062440441843718, -0.023849465019556843, 0.03782845550726404),
    (0.0, -0.06453888262869706, 0.04068941760916406, 0.41809227322161724, -0.7884856164055829, 0.41809227322161724,
     0.04068941760916406, -0.06453888262869706, 0.0, 0.0),
    (0.0, -0.06453888262869706, -0.04068941760916406, 0.41809227322161724, 0.7884856164055829, 0.41809227322161724,
     -0.04068941760916406, -0.06453888262869706, 0.0, 0.0),
    (0.0, -0.03782845550726404, -0.023849465019556843, 0.11062440441843718, 0.37740285561283066, -0.8526986790088938,
     0.37740285561283066, 0.11062440441843718, -0.023849465019556843, -0.03782845550726404),
)


def filter_bank_1D(wname):
    """(analysis, synthesis) 1-D banks, each (2, L): low-pass row, high-pass row"""
    if wname == "bior4.4":
        bank = _BIOR44
    else:
        import pywt                          # optional dependency, as in the reference
        bank = pywt.Wavelet(wname).filter_bank
    fb = torch.tensor(bank).float()
    return fb[:2, :], fb[2:, :]


def outerprod(u, v):
    return torch.einsum('...i,...j->...ij', u, v)


def nonsep(w):
    """(2, L) 1-D bank -> (1, 4, L, L) 2-D bank in the order LL, LH, HL, HH; flipped so that correlation convolves"""
    rows = torch.cat([w[:1], w[:1], w[1:], w[1:]])
    cols = torch.cat([w, w])
    return outerprod(rows, cols)[None, :].flip(2, 3)


def filter_bank_2D(wname):
    """Wa (4,1,L,L) analysis bank, Ws (4,1,L,L) synthesis bank (flipped back)"""
    wa, ws = filter_bank_1D(wname)
    return nonsep(wa).transpose(0, 1), nonsep(ws).transpose(0, 1).flip(2, 3)
