"""The tensor-core kernels of the 2-D stride-1 networks (csrc/cdl_tc2_analysis.cuh, cdl_tc2_synthesis.cuh; BASELINE
configs 1b, 3, 4) against the exact fp32 CUDA-core kernels and the oracle.

The step tests use small-integer data, exactly representable in tf32 with exact fp32 sums: the tensor-core step must
then equal the exact fp32 CUDA-core step BIT FOR BIT, so any difference is an indexing error, not rounding.
CDL_TC2D selects the family at plan creation: 0 = fp32 kernels, 1 = tensor-core analysis only, 2 (default) = analysis +
residual synthesis."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _plans(N, C, M, K, H, W, mode=None, has_mask=False, maskpass=None, ana=None):
    from cdlnet_video_b200 import Plan
    os.environ.pop("CDL_TC2D", None)
    ref = Plan(2, N, C, M, K, (H, W), (7, 7), 1, has_mask=has_mask, precision="fp32")
    if mode is not None:
        os.environ["CDL_TC2D"] = mode
    if maskpass is not None:
        os.environ["CDL_TC2D_MASKPASS"] = "1" if maskpass else "0"
    prec = "tf32x3" if ana == "3" else "tf32"             # tf32x3 = the 3-term split analysis (CDL_PREC_TF32X3)
    try:
        tc = Plan(2, N, C, M, K, (H, W), (7, 7), 1, has_mask=has_mask, precision=prec)
    finally:
        os.environ.pop("CDL_TC2D", None)
        os.environ.pop("CDL_TC2D_MASKPASS", None)
    assert ref.precision == "fp32" and tc.precision == prec
    return ref, tc


@pytest.mark.parametrize("N,C,M,H,W", [(2, 3, 64, 40, 72), (1, 3, 20, 21, 44), (3, 2, 64, 128, 256), (1, 1, 32, 16, 32), (1, 1, 8, 16, 16), (2, 3, 48, 33, 100)])
def test_integer_data_bit_exact_vs_fp32_kernel(N, C, M, H, W):
    _analysis_bit_exact(N, C, M, H, W, None)


@pytest.mark.parametrize("N,C,M,H,W", [(2, 3, 64, 40, 72), (1, 3, 20, 21, 44), (3, 2, 64, 128, 256), (1, 1, 8, 16, 16)])
def test_analysis_x3_integer_data_bit_exact_vs_fp32_kernel(N, C, M, H, W):
    """The 3-term analysis (cdl_tc2_analysis_x3.cuh, precision "tf32x3"): on exactly representable data the lo
    parts vanish and the result must still equal the fp32 kernel's bit for bit (eight-copy shifter, single operand
    buffer, three MMAs per K-step)."""
    _analysis_bit_exact(N, C, M, H, W, "3")


def test_analysis_x3_forward_close_to_oracle():
    """With the 3-term analysis the forward error is the residual synthesis's alone: well under the single-pass 3.1e-5."""
    ex = _forward_cfg1b_like("1", ana="3")       # tensor-core analysis only (exact fp32 synthesis): ~1e-6 expected
    assert ex <= 1e-5, ex


def _analysis_bit_exact(N, C, M, H, W, ana):
    torch.manual_seed(N * 100 + C * 10 + M)
    dev = torch.device("cuda", 0)
    K = 2
    ref, tc = _plans(N, C, M, K, H, W, mode="1", ana=ana)
    A = [torch.randint(-4, 5, (M, C, 7, 7), device=dev).float() / 8 for _ in range(K)]
    t = torch.randint(0, 4, (K, 2, M), device=dev).float() / 4
    for pl in (ref, tc):
        pl.set_weights(A, A, t)
    r = torch.randint(-4, 5, (N, C, H, W), device=dev).float()
    c = torch.tensor([0.25 * (n + 1) for n in range(N)], device=dev)
    z0 = torch.randint(-8, 9, (N, M, H, W), device=dev).float()
    for first in (True, False):
        for k in range(K):
            za, zb = z0.clone(), z0.clone()
            ref.analysis_step(k, r, za, c=c, first=first)
            tc.analysis_step(k, r, zb, c=c, first=first)
            torch.cuda.synchronize()
            bad = (za != zb)
            if bad.any():                                   # where the mismatches sit tells which index map is wrong
                b = bad.float()
                by_res = [round(b[..., i::4].mean().item(), 3) for i in range(4)]           # w mod 4 = residue / accumulator
                by_row = [round(b[:, :, i::16].mean().item(), 3) for i in range(16)]         # h mod 16 = tile row
                by_blk = [round(b[:, i:i + 8].mean().item(), 3) for i in range(0, M, 8)]     # 8-subband block
                by_col = [round(b[..., i::32].mean().item(), 3) for i in range(0, 32, 4)]    # lane i (w mod 32, residue 0)
                idx = bad.nonzero()[:4].tolist()
                vals = [(za[tuple(i)].item(), zb[tuple(i)].item()) for i in idx]
                pytest.fail(f"first={first} k={k} bad={int(bad.sum())}/{bad.numel()} res={by_res} row={by_row} blk={by_blk} col={by_col} at={idx} ref/tc={vals}")


@pytest.mark.parametrize("N,C,M,H,W,use_mask,maskpass", [
    (2, 3, 64, 40, 72, True, None), (1, 3, 20, 21, 44, False, None), (3, 2, 64, 128, 256, False, None), (1, 1, 32, 16, 32, False, None),
    (1, 1, 8, 16, 16, False, None),
    (2, 3, 64, 40, 72, True, True), (1, 3, 48, 64, 128, True, True),       # JDD mask as an image pass (the default) ...
    (2, 3, 64, 40, 72, True, False), (1, 3, 48, 64, 128, True, False)])    # ... or inside the footprint flush (CDL_TC2D_MASKPASS=0)
def test_synthesis_integer_data_bit_exact_vs_fp32_kernel(N, C, M, H, W, use_mask, maskpass):
    """Residual synthesis mask * B z - yp on the tensor cores (default family) vs the exact fp32 kernel, exact data."""
    _synthesis_bit_exact(N, C, M, H, W, use_mask, maskpass)


def _synthesis_bit_exact(N, C, M, H, W, use_mask, maskpass):
    torch.manual_seed(N * 100 + C * 10 + M + 1)
    dev = torch.device("cuda", 0)
    K = 2
    ref, tc = _plans(N, C, M, K, H, W, has_mask=use_mask, maskpass=maskpass)
    Bw = [torch.randint(-4, 5, (M, C, 7, 7), device=dev).float() / 8 for _ in range(K)]
    t = torch.zeros(K, 2, M, device=dev)
    for pl in (ref, tc):
        pl.set_weights(Bw, Bw, t)
    z = torch.randint(-8, 9, (N, M, H, W), device=dev).float()
    z = z * (torch.rand_like(z) < 0.5)                          # sparse, like a real code
    yp = torch.randint(-4, 5, (N, C, H, W), device=dev).float()
    mp = (torch.rand(N, C, H, W, device=dev) < 0.5).float() if use_mask else None
    for k in range(K):
        oa, ob = torch.empty_like(yp), torch.empty_like(yp)
        ref.synthesis_step(k, z, oa, yp=yp, mask_p=mp, residual=True)
        tc.synthesis_step(k, z, ob, yp=yp, mask_p=mp, residual=True)
        torch.cuda.synchronize()
        bad = (oa != ob)
        if bad.any():
            b = bad.float()
            by_row = [round(b[:, :, i::4].mean().item(), 3) for i in range(4)]            # h mod 4 = tile row
            by_col = [round(b[..., i::32].mean().item(), 3) for i in range(0, 32, 4)]     # w mod 32
            by_c = [round(b[:, i].mean().item(), 3) for i in range(C)]
            idx = bad.nonzero()[:4].tolist()
            vals = [(oa[tuple(i)].item(), ob[tuple(i)].item()) for i in idx]
            pytest.fail(f"k={k} bad={int(bad.sum())}/{bad.numel()} row={by_row} col={by_col} c={by_c} at={idx} ref/tc={vals}")


@pytest.mark.parametrize("mode", ["1", None])
def test_forward_parity_vs_oracle_cfg1b_like(mode):
    """CDLNet(K=20, M=32, P=7, s=1) (root args.json, SURVEY cfg 1b) on a small image: max|xhat - oracle| <= 1e-4."""
    ex = _forward_cfg1b_like(mode)
    assert ex <= 1e-4, ex


def _forward_cfg1b_like(mode, ana=None):
    import cdl_oracle as O
    import cdlnet_video_b200 as cb
    torch.manual_seed(3)
    K, M = 20, 32
    net = cb.CDLNet(K=K, M=M, P=7, s=1, C=1, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(K):
            net.A[k].weight.mul_(0.015)     # ~ 0.7/sqrt(M*49): keeps the K-step iteration stable for a randn bank
            net.B[k].weight.copy_(net.A[k].weight * (1 + 0.05 * torch.randn_like(net.A[k].weight)))
        net.t.copy_(torch.rand_like(net.t) * 0.01)
    y = torch.rand(2, 1, 64, 96)
    xr, zr, *_ = O.forward_t(y, [m.weight.detach() for m in net.A], [m.weight.detach() for m in net.B], net.t.detach(), 1, 25.0, True, 1)
    net = net.cuda().eval()
    net.precision = "tf32x3" if ana == "3" else "tf32"
    if mode is not None:
        os.environ["CDL_TC2D"] = mode
    try:
        with torch.no_grad():
            xhat, z = net(y.cuda(), 25.0)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("CDL_TC2D", None)
    plan = net._last_plan
    assert plan.precision == net.precision
    ex = (xhat.cpu() - xr).abs().max().item()
    print(f"tc2 forward (CDL_TC2D={mode}, CDL_TC2D_ANA={ana}): max|xhat-oracle|={ex:.3e} max|z-oracle|={(z.cpu() - zr).abs().max().item():.3e}")
    return ex



def test_cuda_graph_replay_matches_eager_and_is_faster_on_a_small_image():
    """net.use_cuda_graph: the forward of a launch-bound input (config 1b: K=20, M=32, one 256x256 image, ~45 kernels) replayed
    from a CUDA graph gives the same xhat / z as the eager enqueue, also after the weights change (re-capture)."""
    import time
    import numpy as np
    import cdlnet_video_b200 as cb
    torch.manual_seed(3)
    K, M = 20, 32
    net = cb.CDLNet(K=K, M=M, P=7, s=1, C=1, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(K):
            net.A[k].weight.mul_(0.7 / np.sqrt(2.0 * M * 49))
            net.B[k].weight.copy_(net.A[k].weight * (1 + 0.05 * torch.randn_like(net.A[k].weight)))
        net.t.copy_(torch.rand_like(net.t) * 0.01)
    net = net.cuda().eval()
    net.precision = "fp32"                                   # deterministic family: the comparison is exact
    y = torch.rand(1, 1, 256, 256, device="cuda")

    def run(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.no_grad():
            for _ in range(n):
                out = net(y, 25.0)
        torch.cuda.synchronize()
        return out, (time.perf_counter() - t0) / n
    (x0, z0), _ = run(2)
    _, t_eager = run(20)
    net.use_cuda_graph = True
    (x1, z1), _ = run(2)
    _, t_graph = run(20)
    assert torch.equal(x0, x1) and torch.equal(z0, z1)
    print(f"cfg1b fp32 forward: eager {t_eager * 1e3:.3f} ms, graph replay {t_graph * 1e3:.3f} ms")
    with torch.no_grad():
        net.t.mul_(1.5)                                      # bumps the version counter: repack + re-capture
        x2, _ = net(y, 25.0)
        net.use_cuda_graph = False
        x3, _ = net(y, 25.0)
    assert torch.equal(x2, x3) and not torch.equal(x2, x1)


@pytest.mark.parametrize("N,C,M,H,W", [(2, 3, 64, 40, 72), (1, 1, 32, 33, 64), (1, 3, 20, 21, 44)])
def test_final_dictionary_synthesis_three_term_split_vs_fp32_kernel(N, C, M, H, W):
    """D z (residual = 0, k = 0) on the tensor cores as hi(z) hi(W) + lo(z) hi(W) + hi(z) lo(W): bit-exact on integer data
    (the low parts vanish), fp32-class on real data (the dropped lo*lo term is 2^-22 relative per product)."""
    torch.manual_seed(7 * N + M)
    dev = torch.device("cuda", 0)
    ref, tc = _plans(N, C, M, 2, H, W)
    t = torch.zeros(2, 2, M, device=dev)
    for integer in (True, False):
        if integer:
            Bw = [torch.randint(-4, 5, (M, C, 7, 7), device=dev).float() / 8 for _ in range(2)]
            z = torch.randint(-8, 9, (N, M, H, W), device=dev).float()
        else:
            Bw = [torch.randn(M, C, 7, 7, device=dev) * 0.05 for _ in range(2)]
            z = torch.randn(N, M, H, W, device=dev)
        z = z * (torch.rand_like(z) < 0.3)
        for pl in (ref, tc):
            pl.set_weights(Bw, Bw, t)
        oa = torch.empty(N, C, H, W, device=dev)
        ob = torch.empty_like(oa)
        n0 = tc.launch_count()
        ref.synthesis_step(0, z, oa, residual=False)
        tc.synthesis_step(0, z, ob, residual=False)
        torch.cuda.synchronize()
        assert tc.launch_count() - n0 == 3                               # three tensor-core launches (+ a memset), no fp32 kernel
        if integer:
            assert torch.equal(oa, ob)
        else:
            assert (oa - ob).abs().max().item() <= 4e-6 * max(1.0, oa.abs().max().item())
