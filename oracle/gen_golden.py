"""Generate golden input/output vectors from the UNMODIFIED reference.

Run in the build container only (`python oracle/gen_golden.py`); it imports
/root/reference through `oracle/_refshim.py`, runs the reference modules on CPU in
fp32 with seeded synthetic weights and inputs, and writes small `.npz` fixtures to
`tests/golden/`.  The fixtures are committed; nothing on the GPU box regenerates
them.  Every case stores inputs, weights and the reference's own outputs
(`xhat`, `z`, and where the reference's generator works, the per-iteration codes).
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

warnings.filterwarnings("ignore")
OUT = os.path.join(HERE, "..", "tests", "golden")


def _set_weights(net, scale, gen, gabor=False, neg_t=False):
    K = net.K
    with torch.no_grad():
        if gabor:
            for k in range(K):
                for mod in (net.A[k], net.B[k]):
                    mod.alpha.data = torch.randn(mod.alpha.shape, generator=gen) * scale
                    mod.a.data = torch.randn(mod.a.shape, generator=gen) * 0.5
                    mod.w0.data = torch.randn(mod.w0.shape, generator=gen)
                    mod.psi.data = torch.randn(mod.psi.shape, generator=gen)
        else:
            base = torch.randn(net.A[0].weight.shape, generator=gen) * scale
            for k in range(K):
                net.A[k].weight.data = base * (1 + 0.1 * torch.randn(base.shape, generator=gen))
                net.B[k].weight.data = base * (1 + 0.1 * torch.randn(base.shape, generator=gen))
        t = torch.rand(net.t.shape, generator=gen) * 0.02
        t[:, 1] *= 2.0
        if neg_t:
            t[:, 0] -= 0.008          # some thresholds negative
        net.t.data = t


def _dump(name, net, y, sigma, mask, gabor=False, trace=True, extra=None):
    net.eval()
    with torch.no_grad():
        xhat, z = net(y, sigma, mask=mask) if mask is not None else net(y, sigma)
        d = dict(y=y.numpy(), xhat=xhat.numpy(), z=z.numpy(), t=net.t.detach().numpy(),
                 s=np.int64(net.s), adaptive=np.int64(bool(net.adaptive)))
        if gabor:
            d["A"] = np.stack([m.get_filter(transpose=True).numpy() for m in net.A])
            d["B"] = np.stack([m.get_filter().numpy() for m in net.B])
            for nm in ("alpha", "a", "w0", "psi"):
                d["A_" + nm] = np.stack([getattr(m, nm).detach().numpy() for m in net.A])
                d["B_" + nm] = np.stack([getattr(m, nm).detach().numpy() for m in net.B])
        else:
            d["A"] = np.stack([m.weight.detach().numpy() for m in net.A])
            d["B"] = np.stack([m.weight.detach().numpy() for m in net.B])
        if sigma is None:
            d["sigma_none"] = np.int64(1)
        elif torch.is_tensor(sigma):
            d["sigma"] = sigma.numpy()
        else:
            d["sigma"] = np.float64(sigma)
        if mask is not None:
            d["mask"] = mask.numpy()
        if trace:
            gen = net.forward_generator(y, sigma, mask=mask) if mask is not None else net.forward_generator(y, sigma)
            items = [g.numpy() for g in gen]
            d["trace"] = np.stack(items[:-1])
        # preprocess outputs (bit-exact targets for pad / mean)
        ru = ref._ref_utils
        pp = ru.pre_process if y.dim() == 4 else ru.pre_process_3d
        yp, params, mp = pp(y, net.s, mask=mask if mask is not None else 1)
        d["yp"] = yp.numpy()
        d["mean"] = params[0].numpy()
        d["pad"] = np.array(params[1], dtype=np.int64)
        if mask is not None:
            d["mask_p"] = mp.numpy()
        if extra:
            d.update(extra)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: xhat{tuple(xhat.shape)} z{tuple(z.shape)} nnz={float((z != 0).float().mean()):.3f} "
          f"|xhat|max={float(xhat.abs().max()):.3f} -> {os.path.getsize(path) / 1024:.0f} KB")


def tc2_cases(ref):
    """Fixtures at geometries the 2-D tensor-core kernels cover (7x7, s = 1, C <= 3, M <= 64, W % 4 == 0), each with its
    own generator so that adding them leaves the cases below untouched:  python oracle/gen_golden.py tc2"""
    # 9. JDD like trained_nets/JDD_CDLNet-s0120 (s = 1, C = 3, Bayer mask, per-sample sigma), M = 20 -> GEMM N = 32,
    #    36 rows = 2.25 analysis tiles, 44 columns = 1.4 tiles
    g = torch.Generator().manual_seed(91)
    net = ref.CDLNet(K=4, M=20, P=7, s=1, C=3, t0=0, adaptive=True, init=False)
    _set_weights(net, 0.03, g)
    y = torch.rand(2, 3, 36, 44, generator=g)
    mask = ref._ref_root_utils.gen_bayer_mask(y)
    _dump("cdlnet2d_jdd_s1_w4", net, mask * y, torch.tensor([8.0, 16.0]).reshape(2, 1, 1, 1), mask, trace=False)
    # 10. Gabor dictionary, colour, stride 1 (BASELINE config 4's family), order 1, M = 16
    g = torch.Generator().manual_seed(92)
    net = ref.GDLNet(K=3, M=16, P=7, s=1, C=3, t0=0, order=1, adaptive=True, init=False)
    _refshim.fix_gdlnet(net)
    _set_weights(net, 0.025, g, gabor=True)      # stride 1 keeps 4x the energy of the stride-2 case above: half the amplitude
    _dump("gdlnet_s1_c3", net, torch.rand(1, 3, 24, 40, generator=g), 15.0, None, gabor=True, trace=False)


def nle_cases():
    """model/nle.py nle_mad of the unmodified reference (pywt stubbed with the baked bior4.4 table):  python oracle/gen_golden.py nle"""
    import cdl_oracle as O
    nle = _refshim.load_nle((O.BIOR44_DEC_LO, O.BIOR44_DEC_HI, O.BIOR44_REC_LO, O.BIOR44_REC_HI))
    g = torch.Generator().manual_seed(77)
    out = {}
    for i, (shape, sig) in enumerate([((2, 1, 64, 80), (25.0, 10.0)), ((1, 3, 33, 47), (15.0,)), ((3, 2, 10, 11), (5.0, 50.0, 1.0)),
                                      ((1, 1, 128, 128), (25.0,))]):
        yy, xx = torch.meshgrid(torch.arange(shape[-2]) / 9.0, torch.arange(shape[-1]) / 7.0, indexing="ij")
        clean = 0.5 + 0.25 * torch.sin(xx) * torch.cos(yy) + torch.zeros(*shape)                    # smooth: the HH band is all noise
        if i == 1:
            clean = clean + 0.2 * torch.rand(*shape, generator=g)                                  # textured: the estimate is biased up
        s = torch.tensor(sig).reshape(-1, 1, 1, 1) / 255.0
        y = clean + s * torch.randn(*shape, generator=g)
        with torch.no_grad():
            sh = nle.noise_level(y, method="MAD")
        assert tuple(sh.shape) == (shape[0], 1, 1, 1)
        out[f"y{i}"] = y.numpy()
        out[f"sigma_hat{i}"] = sh.reshape(-1).numpy()
        out[f"sigma_true{i}"] = s.reshape(-1).numpy()
    np.savez_compressed(os.path.join(OUT, "nle_mad.npz"), n=np.int64(4), **out)
    print("wrote nle_mad.npz", {k: v.shape for k, v in out.items() if k.startswith("sigma_hat")})


def csr_cases(ref):
    """The frame-recurrent CSR networks of the unmodified reference (model/net.py:363-567), three consecutive frames each:
    frame 0 without a neighbour, then with z_prev (CDLNet_CSR) / with every neighbour combination (CDLNet_CSRf2).
        python oracle/gen_golden.py csr"""
    def weights(net, gen, scale, names):
        base = torch.randn(net.A[0].weight.shape, generator=gen) * scale
        with torch.no_grad():
            for k in range(net.K):
                net.A[k].weight.data = base * (1 + 0.1 * torch.randn(base.shape, generator=gen))
                net.B[k].weight.data = base * (1 + 0.1 * torch.randn(base.shape, generator=gen))
                if hasattr(net, "A2"):
                    net.A2[k].weight.data = base * (1 + 0.1 * torch.randn(base.shape, generator=gen))
                    net.B2[k].weight.data = base * (1 + 0.1 * torch.randn(base.shape, generator=gen))
            for nm in names:
                p = getattr(net, nm)
                v = torch.rand(p.shape, generator=gen) * (0.02 if nm.startswith("t") else 0.6)
                v[:, 1] *= 2.0
                p.data = v
    out = {}
    # CDLNet_CSR: argscsr.json's family reduced (P = 9, s = 2), ragged extents
    g = torch.Generator().manual_seed(501)
    net = ref.CDLNet_CSR(K=4, M=10, P=9, s=2, C=1, t0=0, adaptive=True, init=False).eval()
    weights(net, g, 0.04, ("t", "t2", "g"))
    frames = torch.rand(3, 2, 1, 27, 30, generator=g)           # (frame, N, C, H, W)
    sig = torch.tensor([20.0, 35.0]).reshape(2, 1, 1, 1)
    with torch.no_grad():
        x0, z0 = net(frames[0], None, sig)
        x1, z1 = net(frames[1], z0, sig)
        x2, z2 = net(frames[2], z1, sig)
    out.update(csr_frames=frames.numpy(), csr_sigma=sig.numpy(), csr_x=np.stack([x0.numpy(), x1.numpy(), x2.numpy()]),
               csr_z=np.stack([z0.numpy(), z1.numpy(), z2.numpy()]), csr_s=np.int64(2))
    for nm in ("t", "t2", "g"):
        out["csr_" + nm] = getattr(net, nm).detach().numpy()
    for nm in ("A", "B", "A2", "B2"):
        out["csr_" + nm] = np.stack([m.weight.detach().numpy() for m in getattr(net, nm)])
    # CDLNet_CSRf2: stride 1, colour, Bayer-like mask on; every neighbour combination
    g = torch.Generator().manual_seed(502)
    net = ref.CDLNet_CSRf2(K=3, M=8, P=7, s=1, C=3, t0=0, adaptive=True, init=False).eval()
    weights(net, g, 0.03, ("t", "g1", "g2"))
    y = torch.rand(1, 3, 20, 24, generator=g)
    mask = ref._ref_root_utils.gen_bayer_mask(y)
    y = mask * y
    with torch.no_grad():
        xa, za = net(y, None, None, 15.0, mask=mask)
        zn1 = za * (1 + 0.3 * torch.randn(za.shape, generator=g)) + 0.01 * torch.randn(za.shape, generator=g) * (torch.rand(za.shape, generator=g) > 0.8)
        zn2 = za * (1 + 0.3 * torch.randn(za.shape, generator=g))
        xb, zb = net(y, zn1, None, 15.0, mask=mask)
        xc, zc = net(y, None, zn2, 15.0, mask=mask)
        xd, zd = net(y, zn1, zn2, 15.0, mask=mask)
    out.update(f2_y=y.numpy(), f2_mask=mask.numpy(), f2_sigma=np.float64(15.0), f2_zprev=zn1.numpy(), f2_zafter=zn2.numpy(),
               f2_x=np.stack([t_.numpy() for t_ in (xa, xb, xc, xd)]), f2_z=np.stack([t_.numpy() for t_ in (za, zb, zc, zd)]), f2_s=np.int64(1))
    for nm in ("t", "g1", "g2"):
        out["f2_" + nm] = getattr(net, nm).detach().numpy()
    for nm in ("A", "B"):
        out["f2_" + nm] = np.stack([m.weight.detach().numpy() for m in getattr(net, nm)])
    np.savez_compressed(os.path.join(OUT, "csr.npz"), **out)
    print("wrote csr.npz; nnz", float((z2 != 0).float().mean()), float((zd != 0).float().mean()),
          "effect of the neighbours on xhat:", float((xb - xa).abs().max()), float((xd - xa).abs().max()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["csr"]:
        csr_cases(_refshim.load())
        sys.exit(0)
    if sys.argv[1:] == ["nle"]:
        nle_cases()
        sys.exit(0)
    ref = _refshim.load()
    if sys.argv[1:] == ["tc2"]:
        tc2_cases(ref)
        sys.exit(0)
    g = torch.Generator().manual_seed(1234)
    R = lambda *sh: torch.rand(*sh, generator=g)

    # 1. 2D grayscale, stride 2, odd H and ragged W (both need stride padding), scalar sigma
    net = ref.CDLNet(K=4, M=12, P=7, s=2, C=1, t0=0, adaptive=True, init=False)
    _set_weights(net, 0.05, g)
    _dump("cdlnet2d_s2", net, R(2, 1, 33, 30), 25.0, None)

    # 2. JDD: 2D colour, stride 1, Bayer mask, per-sample sigma tensor (awgn's shape, utils.py:40-41)
    net = ref.CDLNet(K=3, M=8, P=7, s=1, C=3, t0=0, adaptive=True, init=False)
    _set_weights(net, 0.03, g)
    y = R(2, 3, 24, 26)
    mask = ref._ref_root_utils.gen_bayer_mask(y)
    _dump("cdlnet2d_jdd_mask", net, mask * y, torch.tensor([10.0, 18.0]).reshape(2, 1, 1, 1), mask)

    # 3. video, cubic 7^3, stride 2 (args3d.json hyper-parameter family)
    net = ref.CDLNetVideo(K=3, M=10, P=[7, 7, 7], s=2, C=1, t0=0, adaptive=True, init=False)
    _set_weights(net, 0.02, g)
    _dump("video_s2_p777", net, R(1, 1, 8, 20, 22), 25.0, None, trace=False)

    # 4. video, anisotropic [9,9,5] (args3dmri.json), all-odd extents (a valid unpad_3d class),
    #    per-sample sigma, some negative thresholds
    net = ref.CDLNetVideo(K=3, M=6, P=[9, 9, 5], s=2, C=1, t0=0, adaptive=True, init=False)
    _set_weights(net, 0.02, g, neg_t=True)
    _dump("video_s2_p995_odd", net, R(2, 1, 9, 15, 13), torch.tensor([20.0, 30.0]).reshape(2, 1, 1, 1, 1), None, trace=False)

    # 5. video, default (7,7,5), stride 1, sigma=None, two input channels
    net = ref.CDLNetVideo(K=2, M=6, P=(7, 7, 5), s=1, C=2, t0=0, adaptive=True, init=False)
    _set_weights(net, 0.02, g)
    _dump("video_s1_p775_c2", net, R(1, 2, 5, 12, 14), None, None, trace=False)

    # 6. Gabor dictionary, colour, stride 2, order 2
    net = ref.GDLNet(K=3, M=8, P=7, s=2, C=3, t0=0, order=2, adaptive=True, init=False)
    _refshim.fix_gdlnet(net)
    _set_weights(net, 0.05, g, gabor=True)
    _dump("gdlnet_s2_c3", net, R(1, 3, 20, 18), 15.0, None, gabor=True)

    # 7. non-adaptive net ignores sigma (model/net.py:82)
    net = ref.CDLNet(K=3, M=8, P=7, s=1, C=1, t0=0, adaptive=False, init=False)
    _set_weights(net, 0.05, g)
    _dump("cdlnet2d_nonadaptive", net, R(1, 1, 16, 16), 50.0, None)

    # 8. unpad_3d behaviour table for the 8 parity classes at s=2
    ru = ref._ref_utils
    rows = []
    for D in (16, 15):
        for H in (64, 63):
            for W in (64, 63):
                pad = ru.calc_pad_3D(D, H, W, 2)
                x = torch.zeros(1, 1, D + pad[4] + pad[5], H + pad[2] + pad[3], W + pad[0] + pad[1])
                rows.append([D, H, W, *pad, *ru.unpad_3d(x, pad).shape[2:]])
    np.savez_compressed(os.path.join(OUT, "unpad3d_table.npz"), table=np.array(rows, dtype=np.int64))
    print("unpad3d_table:", rows)
    tc2_cases(ref)
