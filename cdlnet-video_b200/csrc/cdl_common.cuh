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

// ---- frame-recurrent CSR proximal operators (reference model/net.py:229-262), with the reference's operation order ----
__device__ __forceinline__ float sign_f(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }
// prox_CSR(u, zp, lambda, gamma) = ST(ST(u - zp - lambda*sign(zp), lambda*gamma) + zp + lambda*sign(zp), lambda)
__device__ __forceinline__ float prox_csr(float u, float zp, float lam, float gam) {
  const float ls = __fmul_rn(lam, sign_f(zp));
  const float inner = soft_threshold(__fsub_rn(__fsub_rn(u, zp), ls), __fmul_rn(lam, gam));
  return soft_threshold(__fadd_rn(__fadd_rn(inner, zp), ls), lam);
}
// prox_CSR_f2(u, zp, za, lambda, gamma1, gamma2) (model/net.py:244-262)
__device__ __forceinline__ float prox_csr_f2(float u, float zp, float za, float lam, float g1, float g2) {
  const float lg1 = __fmul_rn(lam, g1), lg2 = __fmul_rn(lam, g2);
  const float Ca = __fadd_rn(__fadd_rn(zp, __fmul_rn(lam, sign_f(zp))), __fmul_rn(lg2, sign_f(__fsub_rn(zp, za))));
  const float Cb = __fadd_rn(__fadd_rn(za, __fmul_rn(lam, sign_f(za))), __fmul_rn(lg1, sign_f(__fsub_rn(za, zp))));
  const float d = __fsub_rn(u, Ca);
  const float ls = __fmul_rn(lg1, sign_f(d));
  const float inner = soft_threshold(d, lg1);                      // gamma1 * lambda == lambda * gamma1
  const float mid = soft_threshold(__fadd_rn(__fsub_rn(inner, Cb), ls), lg2);
  return soft_threshold(__fsub_rn(__fadd_rn(mid, Cb), ls), lam);
}

// the proximal step of one analysis epilogue: plain ST, or a CSR variant against the neighbouring frames' codes
struct ProxArgs {
  const float* zprev;   // code of the previous frame (same layout as z) or nullptr
  const float* zafter;  // code of the next frame or nullptr
  const float* ga0;     // [M] gamma thresholds g[k,0,:] paired with zprev (g / g1), or with zafter when only zafter is given (g2)
  const float* ga1;     // [M] g[k,1,:]
  const float* gb0;     // [M] second pair (g2) when both neighbours are given
  const float* gb1;
};

}  // namespace cdl
