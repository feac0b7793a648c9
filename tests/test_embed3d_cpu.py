"""The algebra behind the two-frame embedding route (default for config 1; CDL_EMBED3D=0 disables) (cdlnet-video_b200/model/net.py::_forward_embedded3d): a 2-D stride-2
7x7 network is the video network on a two-frame clip (image, zero) with each filter in the td = 3 slice of a 7x7x7
filter.  Checked here with torch's own convolutions (the reference's arithmetic): analysis, synthesis, and a whole
ISTA iteration agree with the 2-D operators, frame 1 stays zero."""
import torch
import torch.nn.functional as F


def lift(w):
    w3 = torch.zeros(w.shape[0], 1, 7, 7, 7, dtype=w.dtype)
    w3[:, :, 3] = w
    return w3


def test_embedded_operators_equal_the_2d_ones():
    torch.manual_seed(0)
    M, H, W = 5, 12, 16
    A, B = torch.randn(M, 1, 7, 7, dtype=torch.float64), torch.randn(M, 1, 7, 7, dtype=torch.float64)
    r = torch.randn(2, 1, H, W, dtype=torch.float64)
    r3 = torch.zeros(2, 1, 2, H, W, dtype=torch.float64)
    r3[:, :, 0] = r
    z2 = F.conv2d(r, A, stride=2, padding=3)
    z3 = F.conv3d(r3, lift(A), stride=2, padding=3)
    assert z3.shape[2] == 1 and torch.allclose(z3[:, :, 0], z2, atol=1e-12)
    x2 = F.conv_transpose2d(z2, B, stride=2, padding=3, output_padding=1)
    x3 = F.conv_transpose3d(z3, lift(B), stride=2, padding=3, output_padding=1)
    assert x3.shape[2:] == (2, H, W)
    assert torch.allclose(x3[:, :, 0], x2, atol=1e-12) and x3[:, :, 1].abs().max() == 0     # nothing leaks into frame 1
    # one full iteration z <- ST(z - A(B z - y), tau): the residual of frame 1 is 0 - 0, so it never feeds back
    y3 = torch.zeros_like(r3)
    y3[:, :, 0] = r
    st = lambda v, t: v.sign() * (v.abs() - t).clamp_min(0)
    n2 = st(z2 - F.conv2d(x2 - r, A, stride=2, padding=3), 0.1)
    n3 = st(z3 - F.conv3d(x3 - y3, lift(A), stride=2, padding=3), 0.1)
    assert torch.allclose(n3[:, :, 0], n2, atol=1e-12)
