// cdl_tc_synthesis.cuh — tcgen05 synthesis step for the video network (3D, P = 7x7x7, s = 2, C = 1):
//
//     out += B_k z        (reference model/net.py:205,210; nn.ConvTranspose3d, stride 2, output_padding 1)
//
// `out` is pre-initialised by the caller with -yp (so that it ends up holding the residual B z - yp) or
// with 0 (final D z).  Formulation: GEMM + col2im,
//     Pq[q, t] = sum_m z[q, m] * W[m, t]        M_gemm = 256 coarse sites / CTA pair, K = 176, N = 2 x 176
//     out[2q - 3 + t] += Pq[q, t]               (overlap-add of each site's 7x7x7 patch)
//
//   * cta_group::2, filters resident: each CTA keeps half of the bank (176 taps x 176 subbands, 124 KB).
//   * Every tcgen05.mma costs >= ~80 cycles whatever its N (measured, profiles/), so the taps are processed in
//     two passes of N = 176 (25 + 24 rows of 7 taps): 44 MMAs of 88 cycles per tile instead of 132 short ones.
//     The two accumulators (2 x 176 TMEM columns) double-buffer each other: the col2im of pass p overlaps the
//     MMAs of the next pass.
//   * A operand = the code tile: z is kept channels-last (N,Qd,Qh,Qw,176), so a producer thread (one per coarse
//     site = TMEM lane) reads its subbands with 256-bit loads, rounds to tf32 (RNE - the tensor core truncates)
//     and tcgen05.st's them into a 2-slot ring of 64 TMEM columns (3 K-chunks per pass; pass 1 re-reads the
//     tile from L1/L2).
//   * col2im: taps are ordered (th, td, tw).  A thread first combines its 7 tw-values with its w-neighbours
//     by warp shuffles (-> the 2 fine voxels of its own cell, plus 5 spill voxels per warp), then adds a
//     float2 to the CTA's fine tile in shared memory.  Warps (= h-rows of the tile) proceed in lock step
//     over th (named barrier between th groups), so no two warps ever touch the same row: no shared-memory
//     atomics.  A CTA pair sweeps consecutive coarse frames of one column, so the footprints of successive tiles
//     overlap in d inside shared memory (a ring of 9 fine planes): after each tile only the 2 planes that are final
//     are handed to the producer warps, which add them to `out` with red.global.add.v4.f32 (3.5x fewer global
//     reductions than flushing every 7-plane footprint) while the col2im warps start the next tile.
#pragma once
#include "cdl_common.cuh"
#include "cdl_tc_ptx.cuh"
#include "cdl_tc_analysis.cuh"

namespace cdl {
namespace tc {

constexpr int kKB = 176;                  // GEMM K of the synthesis (subbands, padded) = channels of the code layout
constexpr int kKBSteps = kKB / 8;         // 22
constexpr int kNBP = 176;                 // GEMM N per pass
constexpr int kRowsP0 = 25;               // (th,td) rows of 7 taps in pass 0 (pass 1: 24)
constexpr int kColDB = 0, kColAB = 2 * kNBP, kASlotB = 32;   // TMEM: D0 | D1 | A0..A3  (480 of 512)
constexpr int kASlotsB = 5;               // A ring depth (32 subbands = 4 K-steps per slot; 6 chunks per pass)
constexpr int kXD = 7, kXH = 13, kXW = 72;                   // fine footprint tile of one CTA (col 0 <-> fine w = 2*qw0 - 4)
constexpr int kXPlanes = 9;                                  // ring of fine d-planes: 7 being accumulated + 2 being flushed
constexpr int kXTile = kXPlanes * kXH * kXW;

struct SynTcParams {
  Geo g;
  const float* z;       // code in the internal quad-blocked layout (code_site_offset)
  float* out;           // (N,1,Fd,Fh,Fw), accumulated into
  const float* wpack;   // this layer: [2 ranks][2 passes][22 k-steps][11 groups][2][8][4]
  int tiles_w, tiles_h, ntiles;
  int seg, nseg, nunits; // a CTA pair sweeps `seg` consecutive coarse frames of one (n, h-tile, w-tile) column per unit
  int a_lo;             // 0: A = rna_tf32(z) ; 1: A = rna_tf32(z - rna_tf32(z))  (low part, used by the 3-term final synthesis)
  long long* dbg;
  int dbg_mode;         // development aid (results invalid): 512 = producers skip the code loads, 1024 = no L2 prefetch
};

constexpr size_t kSynSmemB = (size_t)2 * kKBSteps * (kNBP / 2) * 8 * sizeof(float);             // 123904
constexpr size_t kSynSmemX = (size_t)kXTile * sizeof(float);                                    // 9-plane footprint ring, 33696
constexpr size_t kSynSmemBytes = kSynSmemB + kSynSmemX + 256;

// filters (M,1,7,7,7) [index (m, td, th, tw)] -> B[pass][n = 7*row + tw, k = m], per-rank UMMA layout, tf32 RNE.
// Pass 0 holds rows 0..24, pass 1 rows 25..48 of the (th,td) row list (th-major); unused columns are zero.
__global__ void k_pack_tc_synthesis(const float* __restrict__ w, float* __restrict__ out, int M, int lo) {
  const int per_rank = 2 * kKBSteps * (kNBP / 2) * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per_rank; i += gridDim.x * blockDim.x) {
    const int rank = i / per_rank;
    int rem = i % per_rank;
    const int pass = rem / (kKBSteps * (kNBP / 2) * 8);
    rem %= kKBSteps * (kNBP / 2) * 8;
    const int ks = rem / ((kNBP / 2) * 8);
    rem %= (kNBP / 2) * 8;
    const int grp = rem / 64, kc = (rem / 32) % 2, r8 = (rem / 4) % 8, e = rem % 4;
    const int j = rank * (kNBP / 2) + grp * 8 + r8;            // column inside the pass accumulator
    const int m = ks * 8 + kc * 4 + e;
    const int row = pass * kRowsP0 + j / 7, tw = j % 7;
    float v = 0.0f;
    if (j < 7 * (pass == 0 ? kRowsP0 : 49 - kRowsP0) && m < M) {
      const int th = row / 7, td = row % 7;
      v = w[(size_t)m * kTaps + (td * 7 + th) * 7 + tw];
    }
    const float hi = ptx::to_tf32_rna(v);
    out[i] = lo ? ptx::to_tf32_rna(v - hi) : hi;                // lo: the part of W that tf32 rounding drops
  }
}

// out <- -yp (residual mode) before the scatter-add; plain float4 stream
__global__ void __launch_bounds__(256) k_neg_copy(const float* __restrict__ yp, float* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(yp) + i);
    reinterpret_cast<float4*>(out)[i] = make_float4(-v.x, -v.y, -v.z, -v.w);
  }
}

struct EdgeMasks { float lt31, lt30, gt0; };   // 1.0 / 0.0 lane masks: fma(x, mask, acc) adds x only where the neighbour exists

// NR consecutive (th,td) rows R0..R0+NR-1, all inside one th group, whose taps sit in u[7*(R-RC0) .. +6].
// Phase 1: w-direction reduction with shuffles (independent across rows), phase 2: all shared-memory loads,
// phase 3: adds and stores - so the NR read-modify-writes overlap instead of forming one dependent chain.
template <int R0, int NR, int RC0, int COLS>
__device__ __forceinline__ void rows_apply(const uint32_t (&u)[COLS], float* xs, int pbase, int hrow, int lane, const EdgeMasks& em) {
  constexpr int th = R0 / 7;
  const unsigned full = 0xffffffffu;
  float x0[NR], x1[NR], e1[NR], e2[NR], e3[NR], f5[NR], f6[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int o = 7 * (R0 + i - RC0);
    const float v0 = __uint_as_float(u[o]), v1 = __uint_as_float(u[o + 1]), v2 = __uint_as_float(u[o + 2]),
                v3 = __uint_as_float(u[o + 3]), v4 = __uint_as_float(u[o + 4]), v5 = __uint_as_float(u[o + 5]),
                v6 = __uint_as_float(u[o + 6]);
    const float a1 = __shfl_down_sync(full, v1, 1), a2 = __shfl_down_sync(full, v2, 1), a0 = __shfl_down_sync(full, v0, 2);
    const float b5 = __shfl_up_sync(full, v5, 1), b6 = __shfl_up_sync(full, v6, 1);
    const float n0 = __shfl_down_sync(full, v0, 1);            // lane 1's v0, used by lane 0 (spill to fine -1)
    x0[i] = fmaf(b5, em.gt0, fmaf(a1, em.lt31, v3));           // fine w = 2q   : taps 3 (own), 1 (q+1), 5 (q-1)
    x1[i] = fmaf(b6, em.gt0, fmaf(a0, em.lt30, fmaf(a2, em.lt31, v4)));   // fine w = 2q+1 : taps 4, 2 (q+1), 0 (q+2), 6 (q-1)
    e1[i] = v0; e2[i] = v1; e3[i] = v2 + n0; f5[i] = v5; f6[i] = v6;
  }
  float* rowp[NR];
  float2 cur[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int td = (R0 + i) % 7;
    const int ps = pbase + td - ((pbase + td >= kXPlanes) ? kXPlanes : 0);          // ring slot of fine plane 2*qd + td
    rowp[i] = xs + (ps * kXH + 2 * hrow + th) * kXW;
    cur[i] = *reinterpret_cast<const float2*>(rowp[i] + 4 + 2 * lane);
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    cur[i].x += x0[i]; cur[i].y += x1[i];
    *reinterpret_cast<float2*>(rowp[i] + 4 + 2 * lane) = cur[i];
  }
  // Spill voxels outside the warp's 64 own columns: lane 0 holds the left three (fine w = 2*qw0 - 3 .. -1 = tile
  // columns 1..3), lane 31 the right two (2*qw0 + 64, 65 = tile columns 68, 69).  One float4 read-modify-write at a
  // lane-dependent address serves both in a single divergent region (two lanes active) instead of two.
  if (lane == 0 || lane == 31) {
    const int off = lane ? 68 : 0;
    float4 q[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) q[i] = *reinterpret_cast<const float4*>(rowp[i] + off);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      q[i].x += lane ? f5[i] : 0.0f; q[i].y += lane ? f6[i] : e1[i];
      q[i].z += lane ? 0.0f : e2[i]; q[i].w += lane ? 0.0f : e3[i];
      *reinterpret_cast<float4*>(rowp[i] + off) = q[i];
    }
  }
}

// rows [R, REND) whose first row RC0 sits at u[0], split at th-group boundaries; warps run the th groups in lock
// step (named barrier 2) so that no two warps ever touch the same shared-memory row at the same time
template <int R, int REND, int RC0, int COLS>
__device__ __forceinline__ void rows_walk(const uint32_t (&u)[COLS], float* xs, int pbase, int hrow, int lane, const EdgeMasks& em) {
  if constexpr (R < REND) {
    constexpr int RB = ((R / 7) + 1) * 7 < REND ? ((R / 7) + 1) * 7 : REND;
    if constexpr (R % 7 == 0 && R > 0) ptx::named_bar_sync(2, 128);
    rows_apply<R, RB - R, RC0, COLS>(u, xs, pbase, hrow, lane, em);
    rows_walk<RB, REND, RC0, COLS>(u, xs, pbase, hrow, lane, em);
  }
}

struct SynTile { int n, qd, qh0, qw0, first, last; };
// j-th tile of CTA pair `pair`: unit u = pair + (j / seg) * npairs is a (n, d-segment, h-tile, w-tile) column of `seg`
// consecutive coarse frames, swept in order of qd
__device__ __forceinline__ SynTile syn_tile(const SynTcParams& p, int pair, int npairs, int j) {
  int u = pair + (j / p.seg) * npairs;
  const int k = j % p.seg;
  SynTile t;
  t.qw0 = (u % p.tiles_w) * kTW; u /= p.tiles_w;
  t.qh0 = (u % p.tiles_h) * 2 * kTH; u /= p.tiles_h;
  t.qd = (u % p.nseg) * p.seg + k;
  t.n = u / p.nseg;
  t.first = k == 0; t.last = k == p.seg - 1;
  return t;
}

// one accumulator pass (176 columns = rows [RFIRST, RLAST) of 7 taps) drained in three 64-column loads of 9/9/rest rows
template <int RFIRST, int RLAST>
__device__ __forceinline__ void syn_epilogue_pass(uint32_t dcol, float* xs, int pbase, int hrow, int lane, const EdgeMasks& em,
                                                  uint64_t* dempty_p, uint64_t* dfull_p, uint32_t parity, long long& tw, uint32_t rank) {
  using namespace ptx;
  CDL_TW(tw, mbar_wait(dfull_p, parity));
  tc_fence_after();
  {
    uint32_t u[64];
    tmem_ld64(dcol, u);
    tmem_wait_ld();
    rows_walk<RFIRST, RFIRST + 9, RFIRST, 64>(u, xs, pbase, hrow, lane, em);
  }
  {
    uint32_t u[64];
    tmem_ld64(dcol + 63, u);
    tmem_wait_ld();
    rows_walk<RFIRST + 9, RFIRST + 18, RFIRST + 9, 64>(u, xs, pbase, hrow, lane, em);
  }
  {
    uint32_t u[64];
    tmem_ld64(dcol + 126, u);                      // rows RFIRST+18 .. RLAST-1 (49 or 42 columns); the rest is ignored
    tmem_wait_ld();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { if (rank == 0) mbar_arrive(dempty_p); else mbar_arrive_cluster(dempty_p, 0); }   // accumulator free again
    rows_walk<RFIRST + 18, RLAST, RFIRST + 18, 64>(u, xs, pbase, hrow, lane, em);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_tc_synthesis(const SynTcParams p) {
  using namespace ptx;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sB = reinterpret_cast<float*>(smem_raw);
  float* sX = reinterpret_cast<float*>(smem_raw + kSynSmemB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kSynSmemB + kSynSmemX);
  uint64_t* wbar = bars + 0;
  uint64_t* afull = bars + 1;                  // [kASlotsB] (leader) producers of both CTAs -> MMA
  uint64_t* aempty = afull + kASlotsB;         // [kASlotsB] MMA commit (multicast) -> producers
  uint64_t* dfull = aempty + kASlotsB;         // [2] MMA commit (multicast) -> epilogue   (index = pass)
  uint64_t* dempty = dfull + 2;                // [2] (leader) epilogue warps of both CTAs -> MMA
  uint64_t* wready = dempty + 2;               //     (leader) the peer CTA's filters have landed
  uint64_t* xfull = wready + 1;                // [2] epilogue -> producers: footprint planes complete, flush them
  uint64_t* xfree = xfull + 2;                 // [2] producers -> epilogue: planes flushed and cleared
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfree + 2);

  const Geo& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  long long tw0 = 0, tw1 = 0, tw2 = 0, tw3 = 0, tw4 = 0, tw5 = 0;
  const long long tstart = clock64();

  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(wready, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&xfull[i], 4); mbar_init(&xfree[i], 8); }
    for (int i = 0; i < kASlotsB; ++i) { mbar_init(&afull[i], 16); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dempty[i], 8); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc<2>(tmem_slot, 512); tmem_relinquish<2>(); }
  for (int i = tid; i < kXTile; i += kThreads) sX[i] = 0.0f;
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(wbar, (uint32_t)kSynSmemB);
    const char* src = reinterpret_cast<const char*>(p.wpack) + (size_t)rank * kSynSmemB;
    const uint32_t piece = 30976;   // 123904 / 4
    for (int i = 0; i < 4; ++i) bulk_g2s(reinterpret_cast<char*>(sB) + i * piece, src + i * piece, piece, wbar);
  }
  tc_fence_before();
  cluster_sync_all();          // barriers initialised, TMEM allocated (filters may still be in flight)
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const int my_tiles = (pair < p.nunits) ? ((p.nunits - pair + npairs - 1) / npairs) * p.seg : 0;

  if (warp < 8) {
    // ============================== producers: code tile -> tf32 -> TMEM A ring; footprint flush ==============================
    // warp = 4*half + quad: TMEM lanes of tile row `quad`; `half` selects which half of every K-chunk this warp converts
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tbase + ((uint32_t)(quad * 32) << 16);
    // After tile j (coarse frame qd) the fine planes 2*qd-od and 2*qd-od+1 are final (all 7 of the footprint at the end
    // of a unit): out[fine] += plane, then clear its ring slot.
    auto flush_tile = [&](int j) {
      const SynTile t = syn_tile(p, pair, npairs, j);
      CDL_TW(tw1, mbar_wait(&xfull[j & 1], (j >> 1) & 1));
      const int fh0 = 2 * (t.qh0 + (int)rank * kTH) - 3, fw0 = 2 * t.qw0 - 4;
      float* on = p.out + (size_t)t.n * g.fine_vol();
      const int nrows = (t.last ? kXD : 2) * kXH;
      for (int r = warp; r < nrows; r += 8) {                    // one 72-float row per warp pass, 18 float4 per row
        const int h = r % kXH, pi = r / kXH;
        const int pr = 2 * t.qd + pi;                            // plane index along the sweep; fine d = pr - od
        const int gd = pr - g.od, gh = fh0 + h, gw = fw0 + 4 * lane;
        if (lane < kXW / 4) {
          float4* cell = reinterpret_cast<float4*>(sX + ((pr % kXPlanes) * kXH + h) * kXW + 4 * lane);
          const float4 v = *cell;
          *cell = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gd >= 0 && gd < g.Fd && gh >= 0 && gh < g.Fh && gw >= 0 && gw + 4 <= g.Fw)
            red_add_v4_f32(on + ((size_t)gd * g.Fh + gh) * g.Fw + gw, v);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&xfree[j & 1]);
    };
    const int a_lo = p.a_lo;
    auto cvt_a = [a_lo](float x) {
      const float hi = __uint_as_float(tf32_rna_bits(x));
      return a_lo ? __uint_as_float(tf32_rna_bits(x - hi)) : hi;
    };
    int it = 0;
    uint32_t gch = 0;
    float rg[6][16];
    for (; it < my_tiles; ++it) {
      const SynTile t = syn_tile(p, pair, npairs, it);
      const SynTile t2 = syn_tile(p, pair, npairs, it + 1 < my_tiles ? it + 1 : it);
      const int qh = t.qh0 + rank * kTH + quad, qw = t.qw0 + lane;
      const int valid = qh < g.Qh && qw < g.Qw && !(p.dbg_mode & 512);
      const float* zs = p.z + code_site_offset(((size_t)t.n * g.Qd + t.qd) * g.Qh + qh, g.Qw, qw);
      // bulk L2 prefetch two tiles ahead (the register refill below already requests tile it+1 during tile it): one
      // contiguous 22.5 KB burst per tile row keeps the DRAM reads sequential
      if (half == 0 && lane == 0) {
        for (int a = 1; a <= 1; ++a) {
          if (it + a >= my_tiles) break;
          const SynTile t3 = syn_tile(p, pair, npairs, it + a);
          const int qh3 = t3.qh0 + rank * kTH + quad;
          if (qh3 < g.Qh) {
            const int nq = (min(kTW, g.Qw - t3.qw0) + 7) & ~7;
            bulk_prefetch_l2(p.z + code_site_offset(((size_t)t3.n * g.Qd + t3.qd) * g.Qh + qh3, g.Qw, t3.qw0), (uint32_t)nq * kKB * 4);
          }
        }
      }
      // This thread's 88 subbands (its half of the six 32-subband K-chunks: 5 x 16 + 8) live in registers for the whole
      // tile: converted to tf32 once, stored to TMEM in both passes.  Right after a group's pass-1 store its registers
      // are reloaded with the NEXT tile's values, so every load has several chunk periods to land.
      auto load_group = [&](float (&dst)[16], const float* base, int c, int ok) {
        if (c < 5) {                                             // subbands 32c + 16*half + [0,16) = blocks 4c + 2*half, +1
          ldg256_pred(base + (4 * c + 2 * half) * kCodeBlk, *reinterpret_cast<float(*)[8]>(&dst[0]), ok);
          ldg256_pred(base + (4 * c + 2 * half + 1) * kCodeBlk, *reinterpret_cast<float(*)[8]>(&dst[8]), ok);
        } else {                                                 // subbands 160 + 8*half + [0,8) = block 20 + half
          ldg256_pred(base + (20 + half) * kCodeBlk, *reinterpret_cast<float(*)[8]>(&dst[0]), ok);
        }
      };
      if (it == 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) load_group(rg[c], zs, c, valid);
      }
      const float* zs2 = zs;                                     // the next tile's site (for the register refill)
      int valid2 = 0;
      if (it + 1 < my_tiles) {
        const int qh2 = t2.qh0 + rank * kTH + quad, qw2 = t2.qw0 + lane;
        valid2 = qh2 < g.Qh && qw2 < g.Qw && !(p.dbg_mode & 512);
        zs2 = p.z + code_site_offset(((size_t)t2.n * g.Qd + t2.qd) * g.Qh + qh2, g.Qw, qw2);
      }
#pragma unroll
      for (int pc = 0; pc < 12; ++pc, ++gch) {                   // 2 passes x 6 K-chunks of 32 subbands (the last holds 16)
        const int c = pc % 6;
        const uint32_t slot = gch % kASlotsB;
        const uint32_t acol = lane_addr + kColAB + slot * kASlotB;
        if (pc < 6) {                                            // first use of the group: round to tf32 in place
#pragma unroll
          for (int i = 0; i < 16; ++i) rg[c][i] = cvt_a(rg[c][i]);
        }
        CDL_TW(tw0, mbar_wait(&aempty[slot], ((gch / kASlotsB) & 1) ^ 1));
        tc_fence_after();
        if (c < 5) tmem_st16(acol + 16 * half, *reinterpret_cast<const uint32_t(*)[16]>(&rg[c][0]));
        else tmem_st8(acol + 8 * half, *reinterpret_cast<const uint32_t(*)[8]>(&rg[c][0]));
        CDL_TW(tw4, tmem_wait_st(); tc_fence_before(); __syncwarp(); if (lane == 0) { if (rank == 0) mbar_arrive(&afull[slot]); else mbar_arrive_cluster(&afull[slot], 0); });
        if (pc >= 6) load_group(rg[c], zs2, c, valid2);          // group is dead for this tile: refill it for the next one
      }
      CDL_TW(tw5, if (it > 0) flush_tile(it - 1));
    }
    if (it > 0) flush_tile(it - 1);                                // planes of the last tile
  } else if (warp < kMmaWarp) {
    // ============================== epilogue: col2im ==============================
    const int ew = warp - 8;
    const uint32_t lane_addr = tbase + ((uint32_t)(ew * 32) << 16);
    const EdgeMasks em = {lane < 31 ? 1.0f : 0.0f, lane < 30 ? 1.0f : 0.0f, lane > 0 ? 1.0f : 0.0f};
    for (int it = 0; it < my_tiles; ++it) {
      const SynTile t = syn_tile(p, pair, npairs, it);
      const int xb = it & 1;
      // the two planes this tile opens (2*qd+5, 2*qd+6) reuse the ring slots flushed after tile it-2 ...
      CDL_TW(tw1, mbar_wait(&xfree[xb], ((it >> 1) & 1) ^ 1));
      // ... and a new unit may start anywhere in the ring: wait for the complete flush that ended the previous unit
      if (t.first && it > 0) CDL_TW(tw1, mbar_wait(&xfree[(it - 1) & 1], ((it - 1) >> 1) & 1));
      const int pbase = (2 * t.qd) % kXPlanes;
      syn_epilogue_pass<0, kRowsP0>(lane_addr + kColDB, sX, pbase, ew, lane, em, &dempty[0], &dfull[0], it & 1, tw0, rank);
      syn_epilogue_pass<kRowsP0, 49>(lane_addr + kColDB + kNBP, sX, pbase, ew, lane, em, &dempty[1], &dfull[1], it & 1, tw0, rank);
      __syncwarp();
      if (lane == 0) mbar_arrive(&xfull[xb]);      // this warp's rows are in; 4 arrivals -> producers flush the final planes
      named_bar_sync(2, 128);                      // th lock step restarts with everyone at group 0
    }
  } else {
    // ============================== MMA issue (leader CTA, one thread) ==============================
    // (the warp-converged issue form of the analysis kernel measured SLOWER here: this kernel is bound by its col2im
    // warps, and a faster-spinning MMA warp only takes issue slots from the col2im warp it shares a scheduler with)
    if (rank == 1 && lane == 0) { mbar_wait(wbar, 0); mbar_arrive_cluster(wready, 0); }
    if (rank == 0 && lane == 0) {
      CDL_TW(tw2, mbar_wait(wbar, 0); mbar_wait_cluster(wready, 0));
      const uint64_t bdesc0 = make_smem_desc_kmajor_noswz(smem_u32(sB), 128, 256);
      constexpr uint32_t kBStep = ((kNBP / 2) * 32) >> 4;           // 16-byte units between k-steps of B (88 rows x 32 B)
      const uint32_t idesc = make_idesc_tf32(256, kNBP);
      int it = 0;
      uint32_t gch = 0;
      for (; it < my_tiles; ++it) {
        for (int pass = 0; pass < 2; ++pass) {
          CDL_TW(tw0, mbar_wait_cluster(&dempty[pass], (it & 1) ^ 1));
          tc_fence_after();
          const uint32_t dcol = tbase + kColDB + pass * kNBP;
          for (int c = 0; c < 6; ++c, ++gch) {
            const uint32_t slot = gch % kASlotsB;
            CDL_TW(tw1, mbar_wait_cluster(&afull[slot], (gch / kASlotsB) & 1));
            tc_fence_after();
            const uint32_t a0 = tbase + kColAB + slot * kASlotB;
            const uint64_t bd = bdesc0 + (uint64_t)(pass * kKBSteps + c * 4) * kBStep;
            if (c < 5) {
#pragma unroll
              for (int j = 0; j < 4; ++j) mma_tf32_ts<2>(dcol, a0 + j * 8, bd + (uint64_t)j * kBStep, idesc, (c | j) != 0);
            } else {
#pragma unroll
              for (int j = 0; j < 2; ++j) mma_tf32_ts<2>(dcol, a0 + j * 8, bd + (uint64_t)j * kBStep, idesc, 1);
            }
            mma_commit<2>(&aempty[slot]);             // A slot reusable once these MMAs have read it
          }
          mma_commit<2>(&dfull[pass]);                // this pass's accumulator is complete -> col2im (both CTAs)
        }
      }
    }
    __syncwarp();
  }
  if (p.dbg && lane == 0) {
    long long* d = p.dbg + ((size_t)blockIdx.x * 24 + warp) * 8;
    d[0] = clock64() - tstart; d[1] = tw0; d[2] = tw1; d[3] = tw2; d[4] = tw3; d[5] = tw4; d[6] = tw5;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc<2>(tbase, 512);
}

// ---- code layout conversion: internal quad-blocked channels-last <-> reference (N,M,Qd,Qh,Qw) ----
// One block = 32 consecutive sites of one row x 32 subbands, transposed through shared memory.  On the internal side
// that tile is 8 contiguous 512-byte pieces (4 w-blocks x 2 parities, 4 subband blocks each): one float4 per lane.
// grid = (rows * ceil(Qw/32), 176/32 rounded up, N); R = Qd*Qh rows per sample.
__device__ __forceinline__ void code_tile_piece(int warp, int lane, int& qloc, int& mloc) {
  const int wb = warp >> 1, par = warp & 1;                 // piece = (w-block, parity); lane = float4 index inside it
  qloc = 8 * wb + 2 * ((lane & 7) >> 1) + par;              // site inside the 32-site segment
  mloc = 8 * (lane >> 3) + 4 * (lane & 1);                  // first of 4 subbands inside the 32-subband slab
}
__global__ void __launch_bounds__(256) k_code_export(const float* __restrict__ zcl, float* __restrict__ z, int R, int Qw, int M) {
  __shared__ float t[32][33];                               // [site][subband]
  const int nseg = (Qw + 31) >> 5;
  const int row = blockIdx.x / nseg, qw0 = (blockIdx.x % nseg) * 32;
  const int m0 = blockIdx.y * 32, n = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int qloc, mloc;
  code_tile_piece(warp, lane, qloc, mloc);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (qw0 + qloc < Qw && m0 + mloc < kKB)
    v = *reinterpret_cast<const float4*>(zcl + code_site_offset((size_t)n * R + row, Qw, qw0 + qloc) + (size_t)((m0 + mloc) >> 3) * kCodeBlk + ((m0 + mloc) & 7));
  t[qloc][mloc] = v.x; t[qloc][mloc + 1] = v.y; t[qloc][mloc + 2] = v.z; t[qloc][mloc + 3] = v.w;
  __syncthreads();
  const size_t Q = (size_t)R * Qw;
  for (int r = warp; r < 32; r += 8) {
    const int m = m0 + r, qw = qw0 + lane;
    if (m < M && qw < Qw) z[((size_t)n * M + m) * Q + (size_t)row * Qw + qw] = t[lane][r];
  }
}
__global__ void __launch_bounds__(256) k_code_import(const float* __restrict__ z, float* __restrict__ zcl, int R, int Qw, int M) {
  __shared__ float t[32][33];
  const int nseg = (Qw + 31) >> 5;
  const int row = blockIdx.x / nseg, qw0 = (blockIdx.x % nseg) * 32;
  const int m0 = blockIdx.y * 32, n = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t Q = (size_t)R * Qw;
  for (int r = warp; r < 32; r += 8) {
    const int m = m0 + r, qw = qw0 + lane;
    t[lane][r] = (m < M && qw < Qw) ? z[((size_t)n * M + m) * Q + (size_t)row * Qw + qw] : 0.0f;
  }
  __syncthreads();
  int qloc, mloc;
  code_tile_piece(warp, lane, qloc, mloc);
  if (qw0 + qloc < Qw && m0 + mloc < kKB)
    *reinterpret_cast<float4*>(zcl + code_site_offset((size_t)n * R + row, Qw, qw0 + qloc) + (size_t)((m0 + mloc) >> 3) * kCodeBlk + ((m0 + mloc) & 7)) =
        make_float4(t[qloc][mloc], t[qloc][mloc + 1], t[qloc][mloc + 2], t[qloc][mloc + 3]);
}

}  // namespace tc
}  // namespace cdl
