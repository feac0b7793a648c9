// cdl_prepost.cuh — pre_process / post_process of the reference (model/utils.py:5-33, 70-98):
// per-sample (masked) mean, centre + mask + reflect-pad to a multiple of the stride; crop + add mean.
#pragma once
#include "cdl_common.cuh"

namespace cdl {

constexpr int kRedThreads = 256;
constexpr int kRedBlocksPerSample = 64;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stage 1: partial[(n*2 + which)*B + b] = sum over a fixed slice (deterministic order), fp64 accumulation
__global__ void __launch_bounds__(kRedThreads) k_reduce_partial(const float* __restrict__ y, const float* __restrict__ mask,
                                                                 double* __restrict__ partial, long long per_sample) {
  const int n = blockIdx.y, b = blockIdx.x, B = gridDim.x;
  const float* yn = y + (long long)n * per_sample;
  const float* mn = mask ? mask + (long long)n * per_sample : nullptr;
  double sy = 0.0, sm = 0.0;
  const long long nvec = ((reinterpret_cast<uintptr_t>(yn) & 15) == 0 && (!mn || (reinterpret_cast<uintptr_t>(mn) & 15) == 0)) ? per_sample / 4 : 0;
  for (long long i = (long long)b * kRedThreads + threadIdx.x; i < nvec; i += (long long)B * kRedThreads) {
    float4 v = __ldg(reinterpret_cast<const float4*>(yn) + i);
    sy += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    if (mn) {
      float4 m = __ldg(reinterpret_cast<const float4*>(mn) + i);
      sm += ((double)m.x + (double)m.y) + ((double)m.z + (double)m.w);
    }
  }
  for (long long i = nvec * 4 + (long long)b * kRedThreads + threadIdx.x; i < per_sample; i += (long long)B * kRedThreads) {
    sy += (double)yn[i];
    if (mn) sm += (double)mn[i];
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

// stage 2: sums[2n] = sum(y_n), sums[2n+1] = sum(mask_n) or the element count
__global__ void k_reduce_final(const double* __restrict__ partial, double* __restrict__ sums, int B, int N, int has_mask, double count) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double a = 0.0, c = 0.0;
  for (int b = 0; b < B; ++b) { a += partial[((long long)n * 2 + 0) * B + b]; c += partial[((long long)n * 2 + 1) * B + b]; }
  sums[2 * n] = a;
  sums[2 * n + 1] = has_mask ? c : count;
}

// mean = fp32(sum) / fp32(denominator): x.sum()/mask.sum() resp. x.mean() (model/utils.py:10-13, 75-78)
__global__ void k_mean_from_sums(const double* __restrict__ sums, float* __restrict__ mean, int N) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  mean[n] = __fdiv_rn((float)sums[2 * n], (float)sums[2 * n + 1]);
}

__device__ __forceinline__ int reflect_idx(int i, int L) {   // F.pad(mode='reflect') index map
  if (i < 0) i = -i;
  if (i >= L) i = 2 * (L - 1) - i;
  return i;
}

struct PadParams {
  int N, C;
  int D, H, W;        // unpadded extents
  int Fd, Fh, Fw;     // padded extents
  int pf, pt, pl;     // front / top / left pad
};

// yp = reflect_pad(mask * (y - mean));  mask_p = reflect_pad(mask)   (model/utils.py:14-20, 79-85)
__global__ void __launch_bounds__(256) k_center_pad(const float* __restrict__ y, const float* __restrict__ mask, const float* __restrict__ mean,
                                                    float* __restrict__ yp, float* __restrict__ mask_p, PadParams p) {
  const long long total = (long long)p.N * p.C * p.Fd * p.Fh * p.Fw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % p.Fw); long long r = i / p.Fw;
    int h = (int)(r % p.Fh); r /= p.Fh;
    int d = (int)(r % p.Fd); r /= p.Fd;           // r = n*C + c
    int n = (int)(r / p.C);
    int sw = reflect_idx(w - p.pl, p.W), sh = reflect_idx(h - p.pt, p.H), sd = reflect_idx(d - p.pf, p.D);
    long long src = ((r * p.D + sd) * p.H + sh) * p.W + sw;
    float v = __fsub_rn(y[src], mean[n]);
    if (mask) {
      float m = mask[src];
      v = __fmul_rn(m, v);
      mask_p[i] = m;
    }
    yp[i] = v;
  }
}

// xhat = unpad(xphat) + mean   (model/utils.py:24-33, 89-98; the evident crop for every parity class)
__global__ void __launch_bounds__(256) k_unpad_add_mean(const float* __restrict__ xp, const float* __restrict__ mean, float* __restrict__ xhat, PadParams p) {
  const long long total = (long long)p.N * p.C * p.D * p.H * p.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % p.W); long long r = i / p.W;
    int h = (int)(r % p.H); r /= p.H;
    int d = (int)(r % p.D); r /= p.D;
    int n = (int)(r / p.C);
    long long src = ((r * p.Fd + d + p.pf) * p.Fh + h + p.pt) * p.Fw + w + p.pl;
    xhat[i] = __fadd_rn(xp[src], mean[n]);
  }
}

}  // namespace cdl
