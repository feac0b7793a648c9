// cdl_tc_synthesis.cuh — tcgen05 synthesis step for the video network (3D, P = 7x7x7, s = 2, C = 1):
//
//     out += B_k z        (reference model/net.py:205,210; nn.ConvTranspose3d, stride 2, output_padding 1)
//
// `out` is pre-initialised by the caller with -yp (so that it ends up holding the residual B z - yp) or
// with 0 (final D z).  Formulation: GEMM + col2im,
//     Pq[q, t] = sum_m z[m, q] * W[m, t]        M_gemm = 256 coarse sites / CTA pair, K = 176, N = 352
//     out[2q - 3 + t] += Pq[q, t]               (overlap-add of each site's 7x7x7 patch)
//
//   * cta_group::2, filters resident: each CTA keeps half of the bank (176 taps x 176 subbands, 124 KB).
//   * A operand = z^T tile: producer threads (one per coarse site = TMEM lane) read z[m, q] straight from
//     global memory (coalesced over the 32 sites of a warp), round to tf32 (RNE) and tcgen05.st it into
//     TMEM; two A buffers (2 x 176 columns) so loads of tile i+1 overlap the MMAs of tile i.
//   * N is processed in 6 chunks (5 x 64 + 32 taps); the accumulator is a 2-slot ring of 64 TMEM columns,
//     so the col2im epilogue of chunk c overlaps the MMAs of chunk c+1.
//   * col2im: taps are ordered (th, td, tw).  A thread first combines its 7 tw-values with its w-neighbours
//     by warp shuffles (-> the 2 fine voxels of its own cell, plus 5 spill voxels per warp), then adds a
//     float2 to the CTA's fine tile in shared memory.  Warps (= h-rows of the tile) proceed in lock step
//     over th (named barrier between th groups), so no two warps ever touch the same row: no shared-memory
//     atomics.  The finished 7 x 13 x 69 footprint is added to `out` with red.global.add (tiles overlap).
#pragma once
#include "cdl_common.cuh"
#include "cdl_tc_ptx.cuh"
#include "cdl_tc_analysis.cuh"

namespace cdl {
namespace tc {

constexpr int kKB = 176;                  // GEMM K of the synthesis (subbands, padded)
constexpr int kKBSteps = kKB / 8;         // 22
constexpr int kNB = 352;                  // GEMM N (343 taps padded)
constexpr int kNChunks = 6;               // 5 x 64 + 32
constexpr int kColA0 = 0, kColDB = 2 * kKB, kDSlot = 64;   // TMEM: A0 | A1 | D0 | D1  (480 of 512)
constexpr int kXD = 7, kXH = 13, kXW = 72;                 // fine footprint tile of one CTA (col 0 <-> fine w = 2*qw0 - 4)
constexpr int kXTile = kXD * kXH * kXW;

struct SynTcParams {
  Geo g;
  const float* z;       // (N,M,Qd,Qh,Qw)
  float* out;           // (N,1,Fd,Fh,Fw), accumulated into
  const float* wpack;   // this layer: [2 ranks][chunk][22 k-steps][rows/8][2][8][4]
  int tiles_w, tiles_h, ntiles;
};

__host__ __device__ constexpr int syn_chunk_rows(int nc) { return nc < 5 ? 32 : 16; }          // B rows per CTA in chunk nc
__host__ __device__ constexpr int syn_chunk_off(int nc) { return nc * (kKBSteps * 32 * 8); }   // float offset of chunk nc
constexpr size_t kSynSmemB = (size_t)(5 * 32 + 16) * kKB * sizeof(float);                      // 123904
constexpr size_t kSynSmemX = (size_t)kXTile * sizeof(float);                                   // 26208
constexpr size_t kSynSmemBytes = kSynSmemB + kSynSmemX + 256;

// filters (M,1,7,7,7) [index (m, td, th, tw)] -> B[n = t' = (th,td,tw), k = m], per-rank UMMA layout, tf32 RNE
__global__ void k_pack_tc_synthesis(const float* __restrict__ w, float* __restrict__ out, int M) {
  const int per_rank = (5 * 32 + 16) * kKB;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per_rank; i += gridDim.x * blockDim.x) {
    const int rank = i / per_rank;
    int rem = i % per_rank;
    int nc = rem / (kKBSteps * 32 * 8);
    if (nc > 5) nc = 5;
    rem -= syn_chunk_off(nc);
    const int rows = syn_chunk_rows(nc);
    const int ks = rem / (rows * 8);
    rem %= rows * 8;
    const int grp = rem / 64, kc = (rem / 32) % 2, r8 = (rem / 4) % 8, e = rem % 4;
    const int n = nc * 64 + rank * rows + grp * 8 + r8;        // GEMM N index = tap in (th,td,tw) order
    const int m = ks * 8 + kc * 4 + e;
    float v = 0.0f;
    if (n < kTaps && m < M) {
      const int th = n / 49, td = (n / 7) % 7, tw = n % 7;
      v = w[(size_t)m * kTaps + (td * 7 + th) * 7 + tw];
    }
    out[i] = ptx::to_tf32_rna(v);
  }
}

// out <- -yp (residual mode) before the scatter-add; plain float4 stream
__global__ void __launch_bounds__(256) k_neg_copy(const float* __restrict__ yp, float* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(yp) + i);
    reinterpret_cast<float4*>(out)[i] = make_float4(-v.x, -v.y, -v.z, -v.w);
  }
}

template <int R>
struct RowOps {
  // one (th,td) row of 7 tw-values -> this CTA's fine tile in shared memory
  static __device__ __forceinline__ void apply(const float (&v)[7], float* xs, int hrow, int lane) {
    constexpr int th = R / 7, td = R % 7;
    const unsigned full = 0xffffffffu;
    float a1 = __shfl_down_sync(full, v[1], 1), a2 = __shfl_down_sync(full, v[2], 1), a0 = __shfl_down_sync(full, v[0], 2);
    float b5 = __shfl_up_sync(full, v[5], 1), b6 = __shfl_up_sync(full, v[6], 1);
    float n0 = __shfl_down_sync(full, v[0], 1);                 // lane 1's v0, needed by lane 0 (spill to fine -1)
    float x0 = v[3], x1 = v[4];
    if (lane < 31) { x0 += a1; x1 += a2; }
    if (lane < 30) x1 += a0;
    if (lane > 0) { x0 += b5; x1 += b6; }
    float* row = xs + (td * kXH + 2 * hrow + th) * kXW;
    float2* cell = reinterpret_cast<float2*>(row + 4 + 2 * lane);
    float2 cur = *cell;
    cur.x += x0; cur.y += x1;
    *cell = cur;
    if (lane == 0) { row[1] += v[0]; row[2] += v[1]; row[3] += v[2] + n0; }
    if (lane == 31) { row[68] += v[5]; row[69] += v[6]; }
  }
};

__device__ __forceinline__ void syn_tile_coords(const SynTcParams& p, int tile, int& n, int& qd, int& qh0, int& qw0) {
  int tw = tile % p.tiles_w; tile /= p.tiles_w;
  int th = tile % p.tiles_h; tile /= p.tiles_h;
  qd = tile % p.g.Qd; n = tile / p.g.Qd;
  qh0 = th * 2 * kTH; qw0 = tw * kTW;
}

// compile-time walk over the columns of accumulator chunk NC: column J holds tap t' = 64*NC + J
template <int NC, int J, int COLS>
struct ChunkStep {
  static __device__ __forceinline__ void run(const uint32_t (&u)[COLS], float (&rowv)[7], float* xs, int hrow, int lane) {
    constexpr int t = NC * 64 + J;
    if constexpr (t < kTaps) {
      rowv[t % 7] = __uint_as_float(u[J]);
      if constexpr (t % 7 == 6) {
        // lock step over th: every warp finishes th-group g before any starts g+1
        if constexpr ((t / 7) % 7 == 0 && (t / 49) > 0) ptx::named_bar_sync(2, 128);
        RowOps<t / 7>::apply(rowv, xs, hrow, lane);
      }
    }
    if constexpr (J + 1 < COLS) ChunkStep<NC, J + 1, COLS>::run(u, rowv, xs, hrow, lane);
  }
};

template <int NC>
__device__ __forceinline__ void syn_epilogue_chunk(uint32_t taddr, float (&rowv)[7], float* xs, int hrow, int lane,
                                                   uint64_t* dempty_slot, uint64_t* dfull_slot, uint32_t parity) {
  using namespace ptx;
  constexpr int COLS = NC < 5 ? 64 : 32;
  mbar_wait(dfull_slot, parity);
  tc_fence_after();
  uint32_t u[COLS];
  if constexpr (COLS == 64) tmem_ld64(taddr, u); else tmem_ld32(taddr, u);
  tmem_wait_ld();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(dempty_slot, 0);           // accumulator slot is free again
  ChunkStep<NC, 0, COLS>::run(u, rowv, xs, hrow, lane);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_tc_synthesis(const SynTcParams p) {
  using namespace ptx;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sB = reinterpret_cast<float*>(smem_raw);
  float* sX = reinterpret_cast<float*>(smem_raw + kSynSmemB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kSynSmemB + kSynSmemX);
  uint64_t* wbar = bars + 0;
  uint64_t* afull = bars + 1;    // [2] (leader) producers of both CTAs -> MMA
  uint64_t* aempty = bars + 3;   // [2] MMA commit (multicast) -> producers
  uint64_t* dfull = bars + 5;    // [2] MMA commit (multicast) -> epilogue
  uint64_t* dempty = bars + 7;   // [2] (leader) epilogue warps of both CTAs -> MMA
  uint64_t* wready = bars + 9;   //     (leader) the peer CTA's filters have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const Geo& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(wready, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&afull[i], 16); mbar_init(&aempty[i], 1); mbar_init(&dfull[i], 1); mbar_init(&dempty[i], 8); }
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
  const size_t mstride = (size_t)g.coarse_vol();

  if (warp < 8) {
    // ============================== producers: z[m, q] -> tf32 -> TMEM A ==============================
    // warp = 4*half + quad: TMEM lanes of tile row `quad`, subbands [88*half, 88*half+88)
    const int quad = warp & 3, half = warp >> 2;
    const int m0 = half * (kKB / 2);
    const uint32_t lane_addr = tbase + ((uint32_t)(quad * 32) << 16);
    int it = 0;
    for (int tile = pair; tile < p.ntiles; tile += npairs, ++it) {
      const uint32_t ab = it & 1;
      int n, qd, qh0, qw0;
      syn_tile_coords(p, tile, n, qd, qh0, qw0);
      const int qh = qh0 + rank * kTH + quad, qw = qw0 + lane;
      const bool valid = qh < g.Qh && qw < g.Qw;
      const float* zq = p.z + (((size_t)n * g.M * g.Qd + qd) * g.Qh + qh) * g.Qw + qw + (size_t)m0 * mstride;
      // pull the next tile's rows of z towards L2 while this one is converted
      if (tile + npairs < p.ntiles) {
        int n2, qd2, qh02, qw02;
        syn_tile_coords(p, tile + npairs, n2, qd2, qh02, qw02);
        const int qh2 = qh02 + rank * kTH + quad;
        if (qh2 < g.Qh) {
          const float* z2 = p.z + (((size_t)n2 * g.M * g.Qd + qd2) * g.Qh + qh2) * g.Qw + qw02;
          for (int m = m0 + lane; m < m0 + kKB / 2 && m < g.M; m += 32) prefetch_l2(z2 + m * mstride);
        }
      }
      const uint32_t acol = lane_addr + kColA0 + ab * kKB + m0;
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint((valid && m0 + i < g.M) ? __ldg(zq + i * mstride) : 0.0f);
      mbar_wait(&aempty[ab], ((it >> 1) & 1) ^ 1);     // first batch is in flight while the MMAs still read this buffer
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(to_tf32_rna(__uint_as_float(v[i])));
      tmem_st32(acol, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint((valid && m0 + 32 + i < g.M) ? __ldg(zq + (32 + i) * mstride) : 0.0f);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(to_tf32_rna(__uint_as_float(v[i])));
      tmem_st32(acol + 32, v);
      {
        uint32_t w[24];
#pragma unroll
        for (int i = 0; i < 24; ++i) w[i] = __float_as_uint((valid && m0 + 64 + i < g.M) ? __ldg(zq + (64 + i) * mstride) : 0.0f);
#pragma unroll
        for (int i = 0; i < 24; ++i) w[i] = __float_as_uint(to_tf32_rna(__uint_as_float(w[i])));
        tmem_st16(acol + 64, *reinterpret_cast<const uint32_t(*)[16]>(&w[0]));
        tmem_st8(acol + 80, *reinterpret_cast<const uint32_t(*)[8]>(&w[16]));
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&afull[ab], 0);
    }
  } else if (warp < kMmaWarp) {
    // ============================== epilogue: col2im ==============================
    const int ew = warp - 8;
    const uint32_t lane_addr = tbase + ((uint32_t)(ew * 32) << 16);
    const int et = tid - 256;
    uint32_t gch = 0;
    for (int tile = pair; tile < p.ntiles; tile += npairs) {
      float rowv[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) rowv[i] = 0.0f;
      // six accumulator chunks, ring of 2 slots
#define CDL_CHUNK(NC) { const uint32_t s_ = gch & 1; \
        syn_epilogue_chunk<NC>(lane_addr + kColDB + s_ * kDSlot, rowv, sX, ew, lane, &dempty[s_], &dfull[s_], (gch >> 1) & 1); ++gch; }
      CDL_CHUNK(0) CDL_CHUNK(1) CDL_CHUNK(2) CDL_CHUNK(3) CDL_CHUNK(4) CDL_CHUNK(5)
#undef CDL_CHUNK
      named_bar_sync(2, 128);                    // footprint complete
      // flush: out[fine] += tile, then clear the tile for the next round
      int n, qd, qh0, qw0;
      syn_tile_coords(p, tile, n, qd, qh0, qw0);
      qh0 += rank * kTH;
      const int fd0 = 2 * qd - g.od, fh0 = 2 * qh0 - 3, fw0 = 2 * qw0 - 4;
      float* on = p.out + (size_t)n * g.fine_vol();
      for (int i = et; i < kXD * kXH * (kXW / 4); i += 128) {
        const int c4 = i % (kXW / 4), r = i / (kXW / 4);
        const int h = r % kXH, d = r / kXH;
        const int gd = fd0 + d, gh = fh0 + h, gw = fw0 + 4 * c4;
        float4* cell = reinterpret_cast<float4*>(sX + r * kXW + 4 * c4);
        const float4 v = *cell;
        *cell = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gd >= 0 && gd < g.Fd && gh >= 0 && gh < g.Fh && gw >= 0 && gw + 4 <= g.Fw) {
          float* dst = on + ((size_t)gd * g.Fh + gh) * g.Fw + gw;
          if (v.x != 0.f) red_add_f32(dst + 0, v.x);
          if (v.y != 0.f) red_add_f32(dst + 1, v.y);
          if (v.z != 0.f) red_add_f32(dst + 2, v.z);
          if (v.w != 0.f) red_add_f32(dst + 3, v.w);
        }
      }
      named_bar_sync(2, 128);                    // tile cleared before the next round's first add
    }
  } else {
    // ============================== MMA issue (leader CTA, one thread) ==============================
    if (rank == 1 && lane == 0) { mbar_wait(wbar, 0); mbar_arrive_cluster(wready, 0); }
    if (rank == 0 && lane == 0) {
      mbar_wait(wbar, 0);
      mbar_wait_cluster(wready, 0);
      const uint32_t sB_addr = smem_u32(sB);
      int it = 0;
      uint32_t gch = 0;
      for (int tile = pair; tile < p.ntiles; tile += npairs, ++it) {
        const uint32_t ab = it & 1;
        mbar_wait_cluster(&afull[ab], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t acol = tbase + kColA0 + ab * kKB;
        for (int nc = 0; nc < kNChunks; ++nc, ++gch) {
          const uint32_t s = gch & 1;
          mbar_wait_cluster(&dempty[s], ((gch >> 1) & 1) ^ 1);
          tc_fence_after();
          const int rows = syn_chunk_rows(nc);
          const uint32_t idesc = make_idesc_tf32(256, 2 * rows);
          const uint32_t boff = sB_addr + syn_chunk_off(nc) * 4;
          for (int ks = 0; ks < kKBSteps; ++ks) {
            const uint64_t bdesc = make_smem_desc_kmajor_noswz(boff + ks * rows * 32, 128, 256);
            mma_tf32_ts<2>(tbase + kColDB + s * kDSlot, acol + ks * 8, bdesc, idesc, ks > 0);
          }
          mma_commit<2>(&dfull[s]);
        }
        mma_commit<2>(&aempty[ab]);               // A buffer reusable once every chunk's MMAs have read it
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc<2>(tbase, 512);
}

}  // namespace tc
}  // namespace cdl
