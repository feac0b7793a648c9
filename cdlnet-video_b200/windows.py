"""16-frame-window evaluation of a long clip, the way the reference actually runs the video network
(reference analyze3d.py:100-128: the loader yields clips of exactly 16 frames, each is noised, masked and denoised on
its own, the PSNR is averaged over the windows).  SURVEY.md 8(f) N2.

Every window is an independent sample of the hot path - its own mean (model/utils.py:75-78), its own zero padding at
the window ends - so a clip of D frames becomes a batch of D/window clips for `net(y, sigma, mask)`: windows are
stacked along the batch axis `batch` at a time, which keeps one plan (one geometry) alive for the whole clip and lets
the kernels' persistent grids fill the GPU.  A ragged last window is denoised on its own (the reference would raise).
"""
import torch


def split_windows(D, window=16):
    """[(start, stop)] frame ranges: full windows first, then the ragged tail (if any)."""
    if window < 1:
        raise ValueError("window must be >= 1")
    spans = [(f, f + window) for f in range(0, D - window + 1, window)]
    tail = D - (D // window) * window
    if tail:
        spans.append((D - tail, D))
    return spans


def denoise_windows(net, clip, sigma=None, mask=1, window=16, batch=4):
    """clip (N,C,D,H,W) -> xhat (N,C,D,H,W), every `window` frames denoised as an independent sample.

    sigma: None | number | tensor with one entry per clip sample (N) - it is repeated per window.
    mask: 1 or a tensor shaped like `clip`.  The sparse codes are not returned (42 GB for a 240-frame 1080p clip)."""
    if clip.dim() != 5:
        raise ValueError("clip must be (N,C,D,H,W)")
    N, C, D, H, W = clip.shape
    out = torch.empty_like(clip)
    spans = split_windows(D, window)
    full = [s for s in spans if s[1] - s[0] == window]
    groups = [full[i:i + batch] for i in range(0, len(full), batch)] + [[s] for s in spans if s[1] - s[0] != window]
    per_sample = torch.is_tensor(sigma) and sigma.numel() == N and N > 1
    has_mask = torch.is_tensor(mask)
    for grp in groups:
        y = torch.cat([clip[:, :, a:b] for a, b in grp], dim=0)                     # (len(grp)*N, C, window, H, W)
        m = torch.cat([mask.expand_as(clip)[:, :, a:b] for a, b in grp], dim=0) if has_mask else mask
        s = sigma.reshape(N).repeat(len(grp)).reshape(-1, 1, 1, 1, 1) if per_sample else sigma
        with torch.no_grad():
            xhat, _ = net(y, s, mask=m)
        for i, (a, b) in enumerate(grp):
            out[:, :, a:b] = xhat[i * N:(i + 1) * N]
    return out


def noisy_forward(net, x, sigma, noise=None, bayer=False, mask=None, want_noisy=False):
    """The evaluation step of the reference's analyze loops (analyze3d.py:108-128, analyze.py) on a CLEAN device tensor:

        mask  = utils.gen_bayer_mask(x) if bayer else (mask or 1)
        noisy = mask * (x + noise * (sigma / 255))         # noise = torch.randn_like(x), the caller's draw
        xhat, z = net(noisy, sigma, mask=mask)

    with awgn + mask + pre_process fused into two passes over x (cdl_preprocess_noisy) instead of ~10; the forward itself
    is the module's native route.  sigma: number or per-sample tensor (N).  CUDA fp32 only - there is no fallback here."""
    if not (x.is_cuda and x.dtype == torch.float32):
        raise RuntimeError("noisy_forward needs a CUDA fp32 tensor (the fused input pipeline has no CPU route)")
    x = x.contiguous()
    N = x.shape[0]
    has_mask = bool(bayer) or torch.is_tensor(mask)
    m = mask.to(device=x.device, dtype=torch.float32).expand_as(x).contiguous() if torch.is_tensor(mask) else None
    c = net._c_vector(sigma, N, x.device)            # sigma / 255 rounded like the reference (None when not adaptive)
    cn = c
    if cn is None and noise is not None:             # non-adaptive nets still get noise at the requested level
        cn = (sigma.to(device=x.device, dtype=torch.float32).reshape(-1) / 255.0).expand(N).contiguous() if torch.is_tensor(sigma) \
            else torch.full((N,), float(sigma) / 255.0, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), torch.no_grad():
        key = net._weights_key()
        prec = net._precision_for(x, m if has_mask and m is not None else (torch.ones_like(x) if has_mask else None), c, key)
        plan = net._plan_for(x.shape, has_mask, x.device.index, prec)
        net._set_plan_weights(plan, key)
        net.__dict__["_last_plan"] = plan
        res = plan.preprocess_noisy(x, None if noise is None else noise.contiguous(), cn, m, bayer=bool(bayer), want_y=want_noisy)
        yp, mp, mean = res[:3]
        z, xp = plan.forward(yp, mp, c)
        xhat = plan.postprocess(xp, mean)
    return (xhat, z, res[3]) if want_noisy else (xhat, z)
