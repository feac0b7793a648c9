// cdl_input.cuh — the step in front of the hot path, fused with pre_process:
//
//     mask  = gen_bayer_mask(x) | gen_bayer_mask3d(x) | a tensor | 1                  (utils.py:13-27)
//     noisy = mask * (x + noise * (sigma / 255))                                       (utils.py:29-55 awgn / awgn3d;
//                                                                                       analyze.py, analyze3d.py:100-128)
//     yp, mean, mask_p = pre_process[_3d](noisy, s, mask)                             (model/utils.py:5-22, 70-87)
//
// The reference materialises randn, the scaled noise, the noisy clip, the masked clip and the centred clip (~10 passes
// over the clip).  Here the noisy sample is recomputed where it is needed: one pass for the per-sample sums (it has to
// finish before anything can be centred: the mean is global), one pass that writes yp (and mask_p).  `noise` is the
// caller's torch.randn_like draw (the reference's generator), so the result is comparable bit for bit; the Bayer mask is
// generated from the indices instead of being read.  Roundings follow the reference expression: t = noise * c, y = x + t,
// y = mask * y, each rounded to fp32.
#pragma once
#include "cdl_common.cuh"
#include "cdl_prepost.cuh"

namespace cdl {

enum { kMaskNone = 0, kMaskTensor = 1, kMaskBayer2D = 2 };

struct NoisySrc {
  const float* x;        // clean (N,C,D,H,W)
  const float* noise;    // same shape or nullptr (no noise added)
  const float* c;        // [N] sigma / 255 or nullptr
  const float* mask;     // kMaskTensor only
  int mode;
  int C, D, H, W;
};

// RGGB pattern of utils.gen_bayer_mask: R at (even, even), G at (even, odd) and (odd, even), B at (odd, odd)
__device__ __forceinline__ float bayer2d(int ch, int h, int w) {
  const int ph = h & 1, pw = w & 1;
  return (ch == 0 ? (!ph && !pw) : ch == 1 ? (ph != pw) : ch == 2 ? (ph && pw) : false) ? 1.0f : 0.0f;
}

// the masked noisy sample at flat index `idx` of sample n (idx over (C,D,H,W)); m returns the mask value there
__device__ __forceinline__ float noisy_at(const NoisySrc& s, int n, long long per_sample, long long idx, float& m) {
  const long long g = (long long)n * per_sample + idx;
  float v = __ldg(s.x + g);
  if (s.noise) v = __fadd_rn(v, __fmul_rn(__ldg(s.noise + g), s.c ? s.c[n] : 0.0f));
  m = 1.0f;
  if (s.mode == kMaskTensor) { m = __ldg(s.mask + g); v = __fmul_rn(m, v); }
  else if (s.mode == kMaskBayer2D) {
    const int w = (int)(idx % s.W); const long long r = idx / s.W;
    const int h = (int)(r % s.H); const int ch = (int)(r / ((long long)s.H * s.D));
    m = bayer2d(ch, h, w);
    v = __fmul_rn(m, v);
  }
  return v;
}

// stage 1 of the sums (same partial layout and fp64 accumulation as k_reduce_partial); optionally stores the noisy sample
__global__ void __launch_bounds__(kRedThreads) k_reduce_partial_noisy(const NoisySrc s, float* __restrict__ y_out, double* __restrict__ partial,
                                                                       long long per_sample) {
  const int n = blockIdx.y, b = blockIdx.x, B = gridDim.x;
  double sy = 0.0, sm = 0.0;
  for (long long i = (long long)b * kRedThreads + threadIdx.x; i < per_sample; i += (long long)B * kRedThreads) {
    float m;
    const float v = noisy_at(s, n, per_sample, i, m);
    if (y_out) y_out[(long long)n * per_sample + i] = v;
    sy += (double)v; sm += (double)m;
  }
  __shared__ double sh[2][kRedThreads / 32];
  sy = warp_sum(sy); sm = warp_sum(sm);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = sy; sh[1][threadIdx.x >> 5] = sm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < kRedThreads / 32; ++w) { a += sh[0][w]; c += sh[1][w]; }
    partial[((long long)n * 2 + 0) * B + b] = a;
    partial[((long long)n * 2 + 1) * B + b] = c;
  }
}

// yp = reflect_pad(mask * (noisy - mean)), mask_p = reflect_pad(mask), the noisy sample recomputed at the source index
__global__ void __launch_bounds__(256) k_center_pad_noisy(const NoisySrc s, const float* __restrict__ mean, float* __restrict__ yp,
                                                          float* __restrict__ mask_p, PadParams p) {
  const long long total = (long long)p.N * p.C * p.Fd * p.Fh * p.Fw;
  const long long per = (long long)p.C * p.D * p.H * p.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % p.Fw); long long r = i / p.Fw;
    int h = (int)(r % p.Fh); r /= p.Fh;
    int d = (int)(r % p.Fd); r /= p.Fd;           // r = n*C + c
    const int n = (int)(r / p.C), ch = (int)(r % p.C);
    const int sw = reflect_idx(w - p.pl, p.W), sh = reflect_idx(h - p.pt, p.H), sd = reflect_idx(d - p.pf, p.D);
    const long long idx = (((long long)ch * p.D + sd) * p.H + sh) * p.W + sw;
    float m;
    float v = __fsub_rn(noisy_at(s, n, per, idx, m), mean[n]);
    if (s.mode != kMaskNone) { v = __fmul_rn(m, v); mask_p[i] = m; }
    yp[i] = v;
  }
}

}  // namespace cdl
