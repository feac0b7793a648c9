// cdl_common.cuh — geometry and small device helpers shared by every kernel of libcdl_b200.
//
// Index model (all kernels): a coarse site q and a filter tap t touch the fine voxel
//     e = stride * q - o + t            (per axis),   valid iff 0 <= e < F
// with o = P/2 for the ordinary zero-padded operator of nn.Conv{2,3}d / nn.ConvTranspose{2,3}d
// (reference model/net.py:32-33,137-142) and o = P/2 - halo_front on the temporal axis of a slab.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cdl {

struct Geo {
  int N, C, M, K;
  int Fd, Fh, Fw;   // fine (padded) extents of yp / residual / xphat
  int Qd, Qh, Qw;   // coarse extents of z
  int Pd, Ph, Pw;   // filter extents
  int sd, s;        // stride along d (1 for 2D) and along h, w
  int od, oh, ow;   // origin offsets
  int ndim;
  __host__ __device__ long long fine_plane() const { return (long long)Fh * Fw; }
  __host__ __device__ long long fine_vol() const { return (long long)Fd * Fh * Fw; }
  __host__ __device__ long long coarse_vol() const { return (long long)Qd * Qh * Qw; }
  __host__ __device__ int taps() const { return Pd * Ph * Pw; }
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
// floor division / positive modulo for possibly negative numerators
__host__ __device__ constexpr int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
__host__ __device__ constexpr int pos_mod(int a, int b) { return a - floor_div(a, b) * b; }

// sign(x) * relu(|x| - t), the reference's ST (model/net.py:11-14), including t < 0 and x == 0.
__device__ __forceinline__ float soft_threshold(float v, float tau) {
  float a = fmaxf(fabsf(v) - tau, 0.0f);
  return (v > 0.0f) ? a : ((v < 0.0f) ? -a : 0.0f);
}

// tau = t0 + c * t1 with the reference's two roundings (no FMA contraction), model/net.py:85.
__device__ __forceinline__ float make_tau(float t0, float t1, float c) { return __fadd_rn(t0, __fmul_rn(c, t1)); }

}  // namespace cdl
