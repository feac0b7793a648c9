// cdl_nle.cuh — blind noise-level estimate on the device: the reference's nle_mad
//
//     sigma_hat[n] = median(|HH * y[n]|) / 0.6745          (model/nle.py:17-27)
//
// HH = the diagonal detail filter of the 2-D bior4.4 analysis bank (model/wvlt.py:5-42: outer product of the
// decomposition high-pass with itself, both axes flipped), applied per channel with stride 2 and no padding
// (F.conv2d(y, hh, stride=2, groups=C)); the median runs over all C x Ho x Wo coefficients of a sample and is the LOWER
// median for an even count (torch.median).  sigma never leaves the device: the module takes it as a tensor.
//
// Kernels: one pass computes the coefficients (shared-memory tile, the 7 x 7 non-zero taps of the 10 x 10 filter),
// stores |.| and builds the first histogram; an exact radix select over the fp32 bit patterns (non-negative floats
// order like their bits) follows: 12 + 12 + 8 bits, one histogram pass over the stored magnitudes per level.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cdl {
namespace nle {

constexpr int kL = 10;                    // bior4.4 filter length
constexpr int kBins = 4096;               // histogram bins per level (12 bits; the last level uses 256)
constexpr int kTW = 32, kTH = 8;          // output tile of the coefficient pass

// pywt.Wavelet('bior4.4').dec_hi (PyWavelets 1.x table; CDF 9/7 high-pass = sqrt(2) x the published 7-tap filter), flipped:
// g[a] = dec_hi[9 - a].  The reference builds it from pywt at run time (model/wvlt.py:9; dependency unpinned and absent here).
__constant__ float c_g[kL] = {0.0f, 0.0f, -0.06453888262869706f, 0.04068941760916406f, 0.41809227322161724f,
                              -0.7884856164055829f, 0.41809227322161724f, 0.04068941760916406f, -0.06453888262869706f, 0.0f};

struct State { unsigned prefix; unsigned krem; unsigned count; unsigned pad; };

// |HH * y| -> mag (N, C*Ho*Wo) and level-0 histogram (key >> 20).  grid = (tiles_w * tiles_h, C, N)
__global__ void __launch_bounds__(256) k_nle_coeffs(const float* __restrict__ y, float* __restrict__ mag, unsigned* __restrict__ hist,
                                                    int C, int H, int W, int Ho, int Wo, int tiles_w) {
  __shared__ float tile[2 * kTH + kL - 2][2 * kTW + kL - 2 + 1];
  __shared__ unsigned sh[kBins / 2];                        // finite non-negative floats: key >> 20 < 2048
  const int n = blockIdx.z, c = blockIdx.y;
  const int ty = blockIdx.x / tiles_w, tx = blockIdx.x % tiles_w;
  const int i0 = ty * kTH, j0 = tx * kTW;
  for (int i = threadIdx.x; i < kBins / 2; i += 256) sh[i] = 0u;
  const float* src = y + ((size_t)n * C + c) * H * W;
  constexpr int RH = 2 * kTH + kL - 2, RW = 2 * kTW + kL - 2;
  for (int i = threadIdx.x; i < RH * RW; i += 256) {
    const int r = i / RW, q = i % RW;
    const int gy = 2 * i0 + r, gx = 2 * j0 + q;
    tile[r][q] = (gy < H && gx < W) ? __ldg(src + (size_t)gy * W + gx) : 0.0f;
  }
  __syncthreads();
  const int li = threadIdx.x / kTW, lj = threadIdx.x % kTW;
  const int oi = i0 + li, oj = j0 + lj;
  if (oi < Ho && oj < Wo) {
    float acc = 0.0f;
#pragma unroll
    for (int a = 2; a < 9; ++a)
#pragma unroll
      for (int b = 2; b < 9; ++b) acc = fmaf(tile[2 * li + a][2 * lj + b], __fmul_rn(c_g[a], c_g[b]), acc);
    const float m = fabsf(acc);
    mag[((size_t)n * C + c) * Ho * Wo + (size_t)oi * Wo + oj] = m;
    atomicAdd(&sh[__float_as_uint(m) >> 20], 1u);
  }
  __syncthreads();
  unsigned* hn = hist + (size_t)n * kBins;
  for (int i = threadIdx.x; i < kBins / 2; i += 256)
    if (sh[i]) atomicAdd(&hn[i], sh[i]);
}

// next-level histogram of the elements whose key matches the prefix found so far.  grid = (blocks, N)
// level 1: keys with (key >> 20) == prefix, bin = (key >> 8) & 0xfff;  level 2: (key >> 8) == prefix, bin = key & 0xff
__global__ void __launch_bounds__(256) k_nle_hist(const float* __restrict__ mag, const State* __restrict__ st, unsigned* __restrict__ hist,
                                                  long long per_sample, int level) {
  __shared__ unsigned sh[kBins];
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < kBins; i += 256) sh[i] = 0u;
  __syncthreads();
  const unsigned prefix = st[n].prefix;
  const int shift_p = level == 1 ? 20 : 8, shift_b = level == 1 ? 8 : 0;
  const unsigned mask = level == 1 ? 0xfffu : 0xffu;
  const float* m = mag + (size_t)n * per_sample;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < per_sample; i += (long long)gridDim.x * 256) {
    const unsigned key = __float_as_uint(__ldg(m + i));
    if ((key >> shift_p) == prefix) atomicAdd(&sh[(key >> shift_b) & mask], 1u);
  }
  __syncthreads();
  unsigned* hn = hist + (size_t)n * kBins;
  for (int i = threadIdx.x; i < kBins; i += 256)
    if (sh[i]) atomicAdd(&hn[i], sh[i]);
}

// one block per sample: the bin holding the element of rank krem, new prefix / rank; clears the histogram for the next
// level.  level 0 starts from rank (count - 1) / 2 (lower median); level 2 writes sigma_hat.
__global__ void __launch_bounds__(256) k_nle_scan(unsigned* __restrict__ hist, State* __restrict__ st, int level, unsigned count,
                                                  float* __restrict__ sigma_hat) {
  __shared__ unsigned part[256];
  __shared__ unsigned sel[2];
  const int n = blockIdx.x;
  unsigned* hn = hist + (size_t)n * kBins;
  const unsigned krem = level == 0 ? (count - 1u) / 2u : st[n].krem;
  constexpr int PER = kBins / 256;                           // 16 consecutive bins per thread
  unsigned loc[PER], s = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) { loc[i] = hn[threadIdx.x * PER + i]; s += loc[i]; hn[threadIdx.x * PER + i] = 0u; }
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned cum = 0; int t = 0;
    for (; t < 255; ++t) { if (cum + part[t] > krem) break; cum += part[t]; }
    sel[0] = (unsigned)t; sel[1] = cum;
  }
  __syncthreads();
  if (threadIdx.x == sel[0]) {
    unsigned cum = sel[1]; int b = 0;
    for (; b < PER - 1; ++b) { if (cum + loc[b] > krem) break; cum += loc[b]; }
    const unsigned bin = threadIdx.x * PER + b;
    const unsigned prev = level == 0 ? 0u : st[n].prefix;
    const unsigned prefix = level == 0 ? bin : (level == 1 ? ((prev << 12) | bin) : ((prev << 8) | bin));
    st[n].prefix = prefix; st[n].krem = krem - cum; st[n].count = count;
    if (level == 2) sigma_hat[n] = __fdiv_rn(__uint_as_float(prefix), 0.6745f);
  }
}

}  // namespace nle
}  // namespace cdl
