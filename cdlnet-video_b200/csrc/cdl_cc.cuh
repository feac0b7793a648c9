// cdl_cc.cuh — CUDA-core fp32 kernels for the two convolutions of the ISTA loop.
//
//   k_cc_analysis : u = A_k r (nn.Conv{2,3}d, reference model/net.py:85,87,200,205) fused with the
//                   update and soft-threshold  z <- ST(z - u, t0 + c*t1)  (model/net.py:11-14).
//                   z makes exactly one HBM round trip (read + write, in place).
//   k_cc_synthesis: x = B_k z (nn.ConvTranspose{2,3}d, model/net.py:87,90,205,210) fused with the
//                   residual  r = mask*x - yp.  Gather (output-stationary) form: no atomics, the
//                   accumulation order is fixed, results are run-to-run deterministic.
//
// These are the exact-arithmetic (fp32 FMA) kernels: the precision class of the reference's CPU path.
#pragma once
#include "cdl_common.cuh"

namespace cdl {

// ------------------------------------------------------------------------------------------------
// analysis + update + soft-threshold
// ------------------------------------------------------------------------------------------------
struct AnaParams {
  Geo g;
  const float* rin;   // (N,C,Fd,Fh,Fw) residual (or yp for the first iteration)
  float* z;           // (N,M,Qd,Qh,Qw), updated in place
  const float* wA;    // packed filters of this layer: [C][Pd][Ph][Pw][MPAD], m fastest
  const float* t0;    // [M] thresholds t[k,0,:]
  const float* t1;    // [M] thresholds t[k,1,:]
  const float* cvec;  // [N] sigma/255 per sample, or nullptr (c = 0)
  int first;          // 1: z <- ST(+u) (iteration 0), 0: z <- ST(z - u)
  ProxArgs prox;      // all null: plain soft threshold; else the CSR proximal operators (model/net.py:229-262)
  int TH, TWS;        // tile = TH coarse rows x TWS strips of 8 coarse sites; TH*TWS == 8 warps
  int tiles_h, tiles_w;
};

constexpr int kAnaThreads = 256;
constexpr int kAnaJB = 8;   // coarse sites per warp strip

template <int S, int PW>
__host__ __device__ constexpr int ana_winp() { return (((kAnaJB - 1) * S + PW) + 3) / 4 * 4; }

template <int S, int PW>
inline size_t ana_smem_bytes(const Geo& g, int TH, int TWS, int MBT) {
  int RH = S * (TH - 1) + g.Ph;
  int RWP = S * kAnaJB * (TWS - 1) + ana_winp<S, PW>();
  return (size_t)(round_up(g.C * g.Pd * RH * RWP, 4) + g.Ph * PW * 32 * MBT) * sizeof(float);
}

// Warp = one strip of 8 consecutive coarse sites along w; lane = subband group: lane handles
// m = lane + 32*i, i < MBT.  The r window of the strip is read from shared memory as a warp-wide
// broadcast, the filter taps as conflict-free consecutive words; MBT*8 accumulators per lane.
template <int S, int PW, int MBT>
__global__ void __launch_bounds__(kAnaThreads) k_cc_analysis(const AnaParams p) {
  constexpr int JB = kAnaJB;
  constexpr int WINP = ana_winp<S, PW>();
  constexpr int MPAD = 32 * MBT;
  const Geo& g = p.g;
  extern __shared__ __align__(16) float smem[];
  const int RH = S * (p.TH - 1) + g.Ph;
  const int RWP = S * JB * (p.TWS - 1) + WINP;
  float* rt = smem;                                           // [C][Pd][RH][RWP]
  float* wt = smem + round_up(g.C * g.Pd * RH * RWP, 4);      // [Ph][PW][MPAD]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_w = blockIdx.x % p.tiles_w, tile_h = blockIdx.x / p.tiles_w;
  const int qd = blockIdx.y, n = blockIdx.z;
  const int qh0 = tile_h * p.TH, qw0 = tile_w * p.TWS * JB;
  const int fd0 = g.sd * qd - g.od, fh0 = S * qh0 - g.oh, fw0 = S * qw0 - g.ow;

  // stage the residual halo tile, zero-filled outside the image (= the conv's zero padding)
  {
    const int total = g.C * g.Pd * RH * RWP;
    const float* src_n = p.rin + (long long)n * g.C * g.fine_vol();
    for (int i = tid; i < total; i += kAnaThreads) {
      int w = i % RWP, r1 = i / RWP;
      int h = r1 % RH, r2 = r1 / RH;
      int d = r2 % g.Pd, c = r2 / g.Pd;
      int gd = fd0 + d, gh = fh0 + h, gw = fw0 + w;
      float v = 0.0f;
      if (gd >= 0 && gd < g.Fd && gh >= 0 && gh < g.Fh && gw >= 0 && gw < g.Fw)
        v = __ldg(src_n + ((long long)(c * g.Fd + gd) * g.Fh + gh) * g.Fw + gw);
      rt[i] = v;
    }
  }

  const int wrow = warp / p.TWS, wstrip = warp % p.TWS;
  float acc[MBT][JB];
#pragma unroll
  for (int i = 0; i < MBT; ++i)
#pragma unroll
    for (int j = 0; j < JB; ++j) acc[i][j] = 0.0f;

  const int chunk = g.Ph * PW * MPAD;   // floats per (c, td) filter chunk
  for (int c = 0; c < g.C; ++c) {
    for (int td = 0; td < g.Pd; ++td) {
      __syncthreads();   // previous chunk fully consumed (first pass: nothing pending)
      {
        const float4* src = reinterpret_cast<const float4*>(p.wA + (long long)(c * g.Pd + td) * chunk);
        float4* dst = reinterpret_cast<float4*>(wt);
        for (int i = tid; i < chunk / 4; i += kAnaThreads) dst[i] = __ldg(src + i);
      }
      __syncthreads();   // chunk (and, first pass, the r tile) visible
      const float* rrow = rt + ((c * g.Pd + td) * RH + S * wrow) * RWP + S * JB * wstrip;
      for (int th = 0; th < g.Ph; ++th) {
        float rwin[WINP];
        const float4* r4 = reinterpret_cast<const float4*>(rrow + th * RWP);
#pragma unroll
        for (int v = 0; v < WINP / 4; ++v) {
          float4 q = r4[v];
          rwin[4 * v] = q.x; rwin[4 * v + 1] = q.y; rwin[4 * v + 2] = q.z; rwin[4 * v + 3] = q.w;
        }
        const float* wp = wt + th * PW * MPAD + lane;
#pragma unroll
        for (int tw = 0; tw < PW; ++tw) {
#pragma unroll
          for (int i = 0; i < MBT; ++i) {
            const float w = wp[tw * MPAD + 32 * i];
#pragma unroll
            for (int j = 0; j < JB; ++j) acc[i][j] = fmaf(w, rwin[j * S + tw], acc[i][j]);
          }
        }
      }
    }
  }

  // epilogue: z <- ST(z - u, tau)   (one read + one write of z per element)
  const int qh = qh0 + wrow;
  if (qh >= g.Qh) return;
  const float cval = p.cvec ? p.cvec[n] : 0.0f;
  const int qws = qw0 + JB * wstrip;
  if (qws >= g.Qw) return;
  const bool vec = ((g.Qw & 3) == 0) && (qws + JB <= g.Qw);
#pragma unroll
  for (int i = 0; i < MBT; ++i) {
    const int m = lane + 32 * i;
    if (m >= g.M) continue;
    const float tau = make_tau(p.t0[m], p.t1[m], cval);
    const long long zoff = (((long long)(n * g.M + m) * g.Qd + qd) * g.Qh + qh) * g.Qw + qws;
    float* zp = p.z + zoff;
    // CSR: one neighbour (prox_CSR with that neighbour's gamma pair) or both (prox_CSR_f2)
    const float* nb1 = p.prox.zprev ? p.prox.zprev : p.prox.zafter;
    const float* nb2 = (p.prox.zprev && p.prox.zafter) ? p.prox.zafter : nullptr;
    const float ga = nb1 ? make_tau(p.prox.ga0[m], p.prox.ga1[m], cval) : 0.0f;
    const float gb = nb2 ? make_tau(p.prox.gb0[m], p.prox.gb1[m], cval) : 0.0f;
    auto prox = [&](float v, int j) {
      if (!nb1) return soft_threshold(v, tau);
      if (!nb2) return prox_csr(v, __ldg(nb1 + zoff + j), tau, ga);
      return prox_csr_f2(v, __ldg(nb1 + zoff + j), __ldg(nb2 + zoff + j), tau, ga, gb);
    };
    if (vec) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (!p.first) { a = reinterpret_cast<const float4*>(zp)[0]; b = reinterpret_cast<const float4*>(zp)[1]; }
      float zin[JB] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float o[JB];
#pragma unroll
      for (int j = 0; j < JB; ++j) o[j] = prox(p.first ? acc[i][j] : __fsub_rn(zin[j], acc[i][j]), j);
      reinterpret_cast<float4*>(zp)[0] = make_float4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<float4*>(zp)[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
      for (int j = 0; j < JB; ++j) {
        if (qws + j < g.Qw) {
          float v = p.first ? acc[i][j] : __fsub_rn(zp[j], acc[i][j]);
          zp[j] = prox(v, j);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// synthesis (+ residual)
// ------------------------------------------------------------------------------------------------
struct SynParams {
  Geo g;
  const float* z;     // (N,M,Qd,Qh,Qw)
  const float* wB;    // packed filters of this layer: [M][C][Pd][Ph][PWP], tw fastest, zero padded
  const float* yp;    // (N,C,F) or nullptr
  const float* mask;  // (N,C,F) or nullptr
  float* out;         // (N,C,F)
  int residual;       // 1: out = mask*x - yp ; 0: out = x
  int TDc, THc, TWs;  // tile in cells (a cell = sd x s x s fine voxels); a thread owns one strip of 8 fine w
  int tiles_d, tiles_h, tiles_w;
  int lo_d, lo_h;     // lowest relative coarse offset touched per axis (<= 0)
  int ZD, ZH, ZWP;    // extents of the staged z halo tile (per subband)
  int MCH;            // subbands per shared-memory chunk
};

constexpr int kSynThreads = 256;
constexpr int kSynFW = 8;   // fine voxels along w per thread

template <int S, int PW>
struct SynW {
  static constexpr int OW = PW / 2;
  static constexpr int JB = kSynFW / S;
  static constexpr int LO = -((PW - 1 - OW) / S);
  static constexpr int HI = (kSynFW - 1 + OW) / S;
  static constexpr int ZSEG = HI - LO + 1;
  static constexpr int NV = (ZSEG + 3) / 4;
  static constexpr int PWP = (PW + 3) / 4 * 4;
};

// Thread = one strip of 8 fine voxels along w, for every phase (pd,ph) of its cell and every image
// channel; the subband loop is sequential in the thread, so no cross-lane reduction is needed.
// z is staged per MCH-subband chunk as a halo tile in shared memory (coalesced, zero-filled = the
// transposed conv's implicit border), filter taps are warp-wide broadcasts.
template <int S, int PW, int C, bool ND3>
__global__ void __launch_bounds__(kSynThreads) k_cc_synthesis(const SynParams p) {
  using W = SynW<S, PW>;
  constexpr int SD = ND3 ? S : 1;
  constexpr int FW = kSynFW, JB = W::JB, NV = W::NV, PWP = W::PWP, OW = W::OW, LO_W = W::LO;
  const Geo& g = p.g;
  extern __shared__ __align__(16) float smem[];
  const int zsz = p.ZD * p.ZH * p.ZWP;
  const int wsz = C * g.Pd * g.Ph * PWP;
  float* zt = smem;                              // [MCH][ZD][ZH][ZWP]
  float* wt = smem + round_up(p.MCH * zsz, 4);   // [MCH][C][Pd][Ph][PWP]

  const int tid = threadIdx.x;
  int bt = blockIdx.x;
  const int tile_w = bt % p.tiles_w; bt /= p.tiles_w;
  const int tile_h = bt % p.tiles_h; bt /= p.tiles_h;
  const int tile_d = bt;
  const int n = blockIdx.y;
  const int jd0 = tile_d * p.TDc, jh0 = tile_h * p.THc, jw0 = tile_w * p.TWs * JB;

  const int strip = tid % p.TWs;
  const int row = (tid / p.TWs) % p.THc;
  const int pl = tid / (p.TWs * p.THc);
  const bool active = pl < p.TDc;

  float acc[SD][S][C][FW];
#pragma unroll
  for (int a = 0; a < SD; ++a)
#pragma unroll
    for (int b = 0; b < S; ++b)
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < FW; ++j) acc[a][b][c][j] = 0.0f;

  const float* zn = p.z + (long long)n * g.M * g.coarse_vol();
  for (int mc = 0; mc < g.M; mc += p.MCH) {
    __syncthreads();
    {   // stage z halo tile of this subband chunk
      const int total = p.MCH * zsz;
      for (int i = tid; i < total; i += kSynThreads) {
        int col = i % p.ZWP, r1 = i / p.ZWP;
        int zh = r1 % p.ZH, r2 = r1 / p.ZH;
        int zd = r2 % p.ZD, mi = r2 / p.ZD;
        int qw = jw0 + LO_W + col, qh = jh0 + p.lo_h + zh, qd = jd0 + p.lo_d + zd, m = mc + mi;
        float v = 0.0f;
        if (m < g.M && qw >= 0 && qw < g.Qw && qh >= 0 && qh < g.Qh && qd >= 0 && qd < g.Qd)
          v = __ldg(zn + (((long long)m * g.Qd + qd) * g.Qh + qh) * g.Qw + qw);
        zt[i] = v;
      }
      const int wtotal = p.MCH * wsz;
      for (int i = tid; i < wtotal; i += kSynThreads) {
        int mi = i / wsz;
        wt[i] = (mc + mi < g.M) ? __ldg(p.wB + (long long)mc * wsz + i) : 0.0f;
      }
    }
    __syncthreads();
    if (!active) continue;
    for (int mi = 0; mi < p.MCH; ++mi) {
#pragma unroll
      for (int pd = 0; pd < SD; ++pd) {
        for (int td = pos_mod(pd + g.od, SD); td < g.Pd; td += SD) {
          const int zdi = pl + (pd + g.od - td) / SD - p.lo_d;   // numerator is an exact multiple of SD
#pragma unroll
          for (int ph = 0; ph < S; ++ph) {
            for (int th = pos_mod(ph + g.oh, S); th < g.Ph; th += S) {
              const int zhi = row + (ph + g.oh - th) / S - p.lo_h;
              const float4* zp = reinterpret_cast<const float4*>(zt + ((mi * p.ZD + zdi) * p.ZH + zhi) * p.ZWP + JB * strip);
              float zs[4 * NV];
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                float4 q = zp[v];
                zs[4 * v] = q.x; zs[4 * v + 1] = q.y; zs[4 * v + 2] = q.z; zs[4 * v + 3] = q.w;
              }
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const float4* wp = reinterpret_cast<const float4*>(wt + (((mi * C + c) * g.Pd + td) * g.Ph + th) * PWP);
                float w[PWP];
#pragma unroll
                for (int v = 0; v < PWP / 4; ++v) {
                  float4 q = wp[v];
                  w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
                }
#pragma unroll
                for (int j = 0; j < FW; ++j) {
#pragma unroll
                  for (int tw = 0; tw < PW; ++tw) {
                    if (pos_mod(j + OW - tw, S) == 0)
                      acc[pd][ph][c][j] = fmaf(w[tw], zs[floor_div(j + OW - tw, S) - LO_W], acc[pd][ph][c][j]);
                  }
                }
              }
            }
          }
        }
      }
    }
  }
  if (!active) return;

  // epilogue: residual (mask*x - yp, two roundings like the reference) or plain x
  const int jd = jd0 + pl, jh = jh0 + row;
  const int fw0 = S * jw0 + FW * strip;
  if (fw0 >= g.Fw) return;
  const bool vec = ((g.Fw & 3) == 0) && (fw0 + FW <= g.Fw);
#pragma unroll
  for (int pd = 0; pd < SD; ++pd) {
    const int fd = SD * jd + pd;
    if (fd >= g.Fd) continue;
#pragma unroll
    for (int ph = 0; ph < S; ++ph) {
      const int fh = S * jh + ph;
      if (fh >= g.Fh) continue;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const long long idx = (((long long)(n * C + c) * g.Fd + fd) * g.Fh + fh) * g.Fw + fw0;
        float o[FW];
#pragma unroll
        for (int j = 0; j < FW; ++j) o[j] = acc[pd][ph][c][j];
        if (vec) {
          if (p.residual) {
            float yv[FW], mv[FW];
            *reinterpret_cast<float4*>(yv) = *reinterpret_cast<const float4*>(p.yp + idx);
            *reinterpret_cast<float4*>(yv + 4) = *reinterpret_cast<const float4*>(p.yp + idx + 4);
            if (p.mask) {
              *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(p.mask + idx);
              *reinterpret_cast<float4*>(mv + 4) = *reinterpret_cast<const float4*>(p.mask + idx + 4);
#pragma unroll
              for (int j = 0; j < FW; ++j) o[j] = __fmul_rn(mv[j], o[j]);
            }
#pragma unroll
            for (int j = 0; j < FW; ++j) o[j] = __fsub_rn(o[j], yv[j]);
          }
          *reinterpret_cast<float4*>(p.out + idx) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(p.out + idx + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
          for (int j = 0; j < FW; ++j) {
            if (fw0 + j < g.Fw) {
              float x = o[j];
              if (p.residual) {
                if (p.mask) x = __fmul_rn(p.mask[idx + j], x);
                x = __fsub_rn(x, p.yp[idx + j]);
              }
              p.out[idx + j] = x;
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// filter repacking (cdl_set_weights)
// ------------------------------------------------------------------------------------------------
// analysis: (M,C,T) -> [C][T][MPAD], m fastest, zero padded
__global__ void k_pack_analysis(const float* __restrict__ w, float* __restrict__ out, int M, int C, int T, int MPAD) {
  long long total = (long long)C * T * MPAD;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int m = (int)(i % MPAD);
    long long ct = i / MPAD;
    int t = (int)(ct % T), c = (int)(ct / T);
    out[i] = (m < M) ? w[((long long)m * C + c) * T + t] : 0.0f;
  }
}
// synthesis: (M,C,Pd,Ph,Pw) -> [M][C][Pd][Ph][PWP], tw fastest, zero padded
__global__ void k_pack_synthesis(const float* __restrict__ w, float* __restrict__ out, int rows, int PW, int PWP) {
  long long total = (long long)rows * PWP;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int tw = (int)(i % PWP);
    long long r = i / PWP;
    out[i] = (tw < PW) ? w[r * PW + tw] : 0.0f;
  }
}

}  // namespace cdl
