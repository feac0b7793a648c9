// cdl_tc_analysis.cuh — tcgen05 analysis step for the video network (3D, P = 7x7x7, s = 2, C = 1):
//
//     z <- ST(z - A_k r, t0 + c*t1)            (reference model/net.py:200,205 + :11-14)
//
// as an implicit GEMM  U[q, m] = sum_t R[q, t] * W[m, t]  with  M_gemm = 256 coarse sites per CTA pair,
// N = 176 (169 subbands, zero padded), K = 392 (49 (td,th) rows x 8: the 7 w-taps plus the 8-byte-aligned window's
// first element, whose filter column is zero), kind::tf32, fp32 accumulation.
//
//   * cta_group::2: the two CTAs of a cluster own 128 coarse sites each (a 1 x 4 x 32 block) and share
//     the filter bank: each keeps HALF of it (88 subbands x 344 taps = 121 KB) resident in shared memory
//     for the whole launch, so the filters are read from L2 once per CTA, not once per tile.
//   * The A operand (im2col of the residual, C = 1) never exists in memory: producer threads (one per
//     coarse site = one TMEM lane) read their 7x7x7 window from a zero-filled halo tile of r in shared
//     memory, round to tf32 (RNE — tcgen05 would truncate) and write it straight into TMEM columns with
//     tcgen05.st; the MMA reads A from TMEM (".ts" form), B from shared memory.
//   * Accumulators are double buffered in TMEM (2 x 176 columns) so the epilogue of tile i (TMEM ->
//     registers -> z - u -> soft threshold -> z, one coalesced read + write of z) overlaps the MMAs of
//     tile i+1; the A operand is a 2-slot ring of 56-column chunks (7 MMAs each).
//   * Persistent: one CTA pair per SM pair, static round-robin over 256-site tiles.
//
// Warp roles (416 threads): warps 0-3 producers, warps 4-11 epilogue (two per TMEM lane quadrant, each owning
// half of the subbands and prefetching its z values into registers while the MMAs run), warp 12 MMA issue + TMEM alloc.
#pragma once
#include "cdl_common.cuh"
#include "cdl_tc_ptx.cuh"

namespace cdl {
namespace tc {

constexpr int kThreads = 416;
constexpr int kMmaWarp = 12;
constexpr int kNA = 176;                 // GEMM N of the analysis (subbands, padded)
constexpr int kNAH = kNA / 2;            // per-CTA half of the filter bank
constexpr int kP = 7, kTaps = 343;
constexpr int kKSteps = 49;              // one tf32 MMA K-step (8 columns) per (td,th) row: window element 0 (zero filter) + 7 taps
#ifndef CDL_ANA_ROWS
#define CDL_ANA_ROWS 7
#endif
constexpr int kChunkRows = CDL_ANA_ROWS;                         // (td,th) rows per A chunk = K-steps per chunk (8 columns each)
constexpr int kChunks = (49 + kChunkRows - 1) / kChunkRows;      // chunks per tile; the last one holds the remaining rows
constexpr int kRowsLast = 49 - kChunkRows * (kChunks - 1);
constexpr int kASlots = 160 / (8 * kChunkRows);                  // A ring depth: all TMEM columns left beside the two accumulators
constexpr int kTH = 4, kTW = 32;         // coarse tile per CTA: 1 x 4 x 32 (d,h,w); a pair stacks two along h
constexpr int kRD = 7, kRH = 2 * (kTH - 1) + kP, kRW = 72;   // residual halo tile (fine): 7 x 13 x 72 floats
constexpr int kRTile = kRD * kRH * kRW;  // 6552 floats = 26208 B (one TMA box)
constexpr int kRTilePad = 6560;          // buffer pitch: 26240 B, a multiple of 128 B (TMA destination alignment)
constexpr int kColD = 0, kColA = 2 * kNA, kASlot = 8 * kChunkRows;   // TMEM columns: D0 | D1 | A ring (<= 512)
static_assert(kASlots >= 2 && kColA + kASlots * kASlot <= 512, "A ring does not fit TMEM");

// store NC (multiple of 8) consecutive TMEM columns from registers
template <int NC>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const float* v) {
  using namespace ptx;
  if constexpr (NC >= 32) { tmem_st32(taddr, *reinterpret_cast<const uint32_t(*)[32]>(v)); tmem_st_cols<NC - 32>(taddr + 32, v + 32); }
  else if constexpr (NC >= 16) { tmem_st16(taddr, *reinterpret_cast<const uint32_t(*)[16]>(v)); tmem_st_cols<NC - 16>(taddr + 16, v + 16); }
  else if constexpr (NC >= 8) { tmem_st8(taddr, *reinterpret_cast<const uint32_t(*)[8]>(v)); tmem_st_cols<NC - 8>(taddr + 8, v + 8); }
}

struct AnaTcParams {
  Geo g;
  const float* rin;     // (N,1,Fd,Fh,Fw) residual (or yp for iteration 0), ALREADY rounded to tf32 (k_round_tf32)
  float* z;             // internal code layout: channels-last (N,Qd,Qh,Qw,176), updated in place
  const float* wpack;   // this layer: [2 ranks][43 k-steps][11 groups][2][8][4] tf32-rounded filters
  const float* t0;      // [M]
  const float* t1;      // [M]
  const float* cvec;    // [N] or nullptr
  int first;
  int tiles_w, tiles_h; // pair tiles along w (32 sites) and h (8 rows)
  int ntiles;           // N * Qd * tiles_h * tiles_w
  long long* dbg;       // optional [grid][16 warps][8] cycle counters
  int dbg_mode;         // development aid: 1 = producers only signal, 2 = epilogue skips the code traffic (results invalid)
};

constexpr size_t kAnaSmemB = (size_t)kKSteps * kNAH * 8 * sizeof(float);      // 121088
constexpr size_t kAnaSmemR = 2 * (size_t)kRTilePad * sizeof(float);           // 52480
constexpr size_t kAnaSmemT = 2 * kNA * sizeof(float);                         // thresholds t0 | t1
constexpr size_t kAnaSmemBytes = kAnaSmemB + kAnaSmemR + kAnaSmemT + 256;

// filters (M,1,7,7,7) -> per-rank K-major no-swizzle UMMA layout, tf32 RNE
__global__ void k_pack_tc_analysis(const float* __restrict__ w, float* __restrict__ out, int M) {
  const int total = 2 * kKSteps * kNAH * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int e = i % 4, r8 = (i / 4) % 8, kc = (i / 32) % 2, grp = (i / 64) % (kNAH / 8);
    int ks = (i / (kNAH * 8)) % kKSteps, rank = i / (kNAH * 8 * kKSteps);
    int m = rank * kNAH + grp * 8 + r8, j = kc * 4 + e;              // k-step ks = (td,th) row; column j: 0 = pad, 1..7 = tw 0..6
    float v = (m < M && j > 0) ? w[(size_t)m * kTaps + ks * 7 + (j - 1)] : 0.0f;
    out[i] = ptx::to_tf32_rna(v);
  }
}

// r -> tf32 (RNE) once per element, before the analysis: every element is used by ~43 windows and the tensor core
// would truncate.  Image-sized stream (4 MB per 16x256x256 clip), negligible next to the code traffic.
__global__ void __launch_bounds__(256) k_round_tf32(const float* __restrict__ src, float* __restrict__ dst, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    v.x = __uint_as_float(ptx::tf32_rna_bits(v.x)); v.y = __uint_as_float(ptx::tf32_rna_bits(v.y));
    v.z = __uint_as_float(ptx::tf32_rna_bits(v.z)); v.w = __uint_as_float(ptx::tf32_rna_bits(v.w));
    reinterpret_cast<float4*>(dst)[i] = v;
  }
}

__device__ __forceinline__ void ana_tile_coords(const AnaTcParams& p, int tile, int& n, int& qd, int& qh0, int& qw0) {
  int tw = tile % p.tiles_w; tile /= p.tiles_w;
  int th = tile % p.tiles_h; tile /= p.tiles_h;
  qd = tile % p.g.Qd; n = tile / p.g.Qd;
  qh0 = th * 2 * kTH; qw0 = tw * kTW;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_tc_analysis(const AnaTcParams p, const __grid_constant__ CUtensorMap rmap) {
  using namespace ptx;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sB = reinterpret_cast<float*>(smem_raw);
  float* sR = reinterpret_cast<float*>(smem_raw + kAnaSmemB);
  float* sT = reinterpret_cast<float*>(smem_raw + kAnaSmemB + kAnaSmemR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kAnaSmemB + kAnaSmemR + kAnaSmemT);
  uint64_t* wbar = bars + 0;
  uint64_t* afull = bars + 1;                  // [kASlots] (used in the leader CTA) producers of both CTAs -> MMA
  uint64_t* aempty = afull + kASlots;          // [kASlots] MMA commit (multicast) -> producers
  uint64_t* dfull = aempty + kASlots;          // [2]  MMA commit (multicast) -> epilogue
  uint64_t* dempty = dfull + 2;                // [2]  (leader) epilogue warps of both CTAs -> MMA
  uint64_t* wready = dempty + 2;               //      (leader) the peer CTA's filters have landed
  uint64_t* rfull = wready + 1;                // [2]  TMA: residual halo tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull + 2);

  const Geo& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  long long tw0 = 0, tw1 = 0, tw2 = 0, tw3 = 0, tw4 = 0, tw5 = 0;
  const long long tstart = clock64();

  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(wready, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&rfull[i], 1);
    for (int i = 0; i < kASlots; ++i) { mbar_init(&afull[i], 8); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dempty[i], 16); }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc<2>(tmem_slot, 512); tmem_relinquish<2>(); }
  for (int i = tid; i < kNA; i += kThreads) {
    sT[i] = (i < g.M) ? p.t0[i] : 0.0f;
    sT[kNA + i] = (i < g.M) ? p.t1[i] : 0.0f;
  }
  __syncthreads();
  if (tid == 0) {   // this CTA's half of the filter bank: 137984 B in 4 bulk copies
    mbar_expect_tx(wbar, (uint32_t)kAnaSmemB);
    const char* src = reinterpret_cast<const char*>(p.wpack) + (size_t)rank * kAnaSmemB;
    const uint32_t piece = (uint32_t)(kAnaSmemB / 4);   // 34496, multiple of 16
    for (int i = 0; i < 4; ++i) bulk_g2s(reinterpret_cast<char*>(sB) + i * piece, src + i * piece, piece, wbar);
  }
  tc_fence_before();
  cluster_sync_all();          // both CTAs: barriers initialised, TMEM allocated (filters may still be in flight)
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  if (warp < 4) {
    // ============================== producers: r tile -> im2col -> TMEM ==============================
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    // one thread per CTA drives the TMA: box 72 x 13 x 7 (w,h,d) of r with zero fill outside the clip = the conv's padding
    auto issue_tile_load = [&](int tile, int buf) {
      int n, qd, qh0, qw0;
      ana_tile_coords(p, tile, n, qd, qh0, qw0);
      qh0 += rank * kTH;
      mbar_expect_tx(&rfull[buf], kRTile * 4);
      tma_load_4d(sR + buf * kRTilePad, &rmap, 2 * qw0 - 4, 2 * qh0 - 3, 2 * qd - g.od, n, &rfull[buf]);   // column 0 <-> fine w = 2*qw0 - 4
    };
    int it = 0;
    uint32_t gchunk = 0;
    if (tid == 0) { tma_prefetch_desc(&rmap); if (pair < p.ntiles) issue_tile_load(pair, 0); }
    for (int tile = pair; tile < p.ntiles; tile += npairs, ++it) {
      const int buf = it & 1;
      CDL_TW(tw1, mbar_wait(&rfull[buf], (it >> 1) & 1); named_bar_sync(1, 128));   // tile `it` landed; everyone left tile it-1
      if (tid == 0 && tile + npairs < p.ntiles) issue_tile_load(tile + npairs, buf ^ 1);
      // this thread's coarse site: row `warp` of the CTA tile, column `lane`
      const float* rs = sR + buf * kRTilePad + (2 * warp) * kRW + 2 * lane;
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch, ++gchunk) {
        const uint32_t slot = gchunk % kASlots;
        // the 8-float window (fine w = 2q-4 .. 2q+3) of each (td,th) row goes to TMEM as is: 8 columns = one K-step
        float raw[8 * kChunkRows];
        const int nrows = (ch < kChunks - 1) ? kChunkRows : kRowsLast;
#pragma unroll
        for (int rr = 0; rr < kChunkRows; ++rr) {
          if (rr < nrows) {
            const int row = ch * kChunkRows + rr, td = row / kP, th = row % kP;
            const float* src = rs + (td * kRH + th) * kRW;
#pragma unroll
            if (!(p.dbg_mode & 1)) { for (int j = 0; j < 4; ++j) lds64(src + 2 * j, raw[8 * rr + 2 * j], raw[8 * rr + 2 * j + 1]); }
          }
        }
        CDL_TW(tw0, mbar_wait(&aempty[slot], ((gchunk / kASlots) & 1) ^ 1));
        tc_fence_after();
        const uint32_t acol = lane_addr + kColA + slot * kASlot;
        if (p.dbg_mode & 1) {
        } else if (ch < kChunks - 1) {
          tmem_st_cols<8 * kChunkRows>(acol, raw);
        } else {
          tmem_st_cols<8 * kRowsLast>(acol, raw);
        }
        CDL_TW(tw2, tmem_wait_st());
        CDL_TW(tw4, tc_fence_before(); __syncwarp(); if (lane == 0) { if (rank == 0) mbar_arrive(&afull[slot]); else mbar_arrive_cluster(&afull[slot], 0); });
      }
    }
  } else if (warp < kMmaWarp) {
    // ============================== epilogue: TMEM -> z update ==============================
    // warp = 4 + 4*half + quad: TMEM lanes [32*quad, 32*quad+32) (tile row `quad`), subbands [88*half, 88*half+88)
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int m0 = half * kNAH;
    const uint32_t lane_addr = tbase + ((uint32_t)(quad * 32) << 16);
    const uint32_t usign = p.first ? 0x80000000u : 0u;          // iteration 0: z_in = 0 and v = +u  (0 - (-u))
    float* sTau = sT;                                           // tau[m] of the sample the current tile belongs to
    int n_tau = -1;
    int it = 0;
    for (int tile = pair; tile < p.ntiles; tile += npairs, ++it) {
      const uint32_t ds = it & 1;
      int n, qd, qh0, qw0;
      ana_tile_coords(p, tile, n, qd, qh0, qw0);
      if (n != n_tau) {                                         // uniform over the 8 epilogue warps (same tile sequence)
        named_bar_sync(3, 256);
        const float cval = p.cvec ? p.cvec[n] : 0.0f;
        for (int i = tid - 128; i < kNA; i += 256) sTau[i] = (i < g.M) ? make_tau(p.t0[i], p.t1[i], cval) : 0.0f;
        named_bar_sync(3, 256);
        n_tau = n;
      }
      const int qh = qh0 + rank * kTH + quad, qw = qw0 + lane;
      const int valid = qh < g.Qh && qw < g.Qw && !(p.dbg_mode & 2);
      const int ld_ok = valid && !p.first;
      // this site's 88 subbands are contiguous (channels-last): 11 LDG.256 in, 11 STG.256 out
      float* zs = p.z + ((((size_t)n * g.Qd + qd) * g.Qh + qh) * g.Qw + qw) * kNA + m0;
      if (!p.first && half == 0 && lane == 0 && tile + npairs < p.ntiles) {   // next tile: this row's 32 sites x 704 B are contiguous
        int n2, qd2, qh02, qw02;
        ana_tile_coords(p, tile + npairs, n2, qd2, qh02, qw02);
        const int qh2 = qh02 + rank * kTH + quad;
        if (qh2 < g.Qh) {
          const int nq = min(kTW, g.Qw - qw02);
          bulk_prefetch_l2(p.z + ((((size_t)n2 * g.Qd + qd2) * g.Qh + qh2) * g.Qw + qw02) * kNA, (uint32_t)nq * kNA * 4);
        }
      }
      // all 88 code values of this site are requested BEFORE waiting for the accumulator (11 x LDG.256 in flight per
      // thread, 90 KB per SM): the DRAM latency hides behind the MMAs of this tile; results overwrite the registers
      float zr[kNAH];
#pragma unroll
      for (int b = 0; b < kNAH / 8; ++b) ldg256_pred(zs + 8 * b, *reinterpret_cast<float(*)[8]>(&zr[8 * b]), ld_ok);
      CDL_TW(tw0, mbar_wait(&dfull[ds], (it >> 1) & 1));
      tc_fence_after();
      const uint32_t dcol = lane_addr + kColD + ds * kNA + m0;
      uint32_t ua[8], ub[8];
      tmem_ld8(dcol, ua);
#pragma unroll
      for (int b = 0; b < kNAH / 8; ++b) {
        uint32_t (&u)[8] = (b & 1) ? ub : ua;
        uint32_t (&un)[8] = (b & 1) ? ua : ub;
        tmem_wait_ld();
        if (b + 1 < kNAH / 8) tmem_ld8(dcol + 8 * (b + 1), un);            // next 8 accumulator columns in flight
        const float4 t0 = *reinterpret_cast<const float4*>(sTau + m0 + 8 * b);
        const float4 t1 = *reinterpret_cast<const float4*>(sTau + m0 + 8 * b + 4);
        const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
          zr[8 * b + i] = soft_threshold(__fsub_rn(zr[8 * b + i], __uint_as_float(u[i] ^ usign)), tt[i]);
        stg256_pred(zs + 8 * b, *reinterpret_cast<const float(*)[8]>(&zr[8 * b]), valid);
      }
      tc_fence_before();                           // accumulator fully read: hand the TMEM slot back to the MMA warp
      __syncwarp();
      if (lane == 0) { if (rank == 0) mbar_arrive(&dempty[ds]); else mbar_arrive_cluster(&dempty[ds], 0); }
    }
  } else {
    // ============================== MMA issue (leader CTA, one thread) ==============================
    if (rank == 1 && lane == 0) { mbar_wait(wbar, 0); mbar_arrive_cluster(wready, 0); }
    if (rank == 0 && lane == 0) {
      CDL_TW(tw2, mbar_wait(wbar, 0); mbar_wait_cluster(wready, 0));
      const uint32_t idesc = make_idesc_tf32(256, kNA);
      const uint64_t bdesc0 = make_smem_desc_kmajor_noswz(smem_u32(sB), 128, 256);
      constexpr uint32_t kBStep = (kNAH * 32) >> 4;                 // 16-byte units between consecutive k-steps of B
      int it = 0;
      uint32_t gchunk = 0;
      for (int tile = pair; tile < p.ntiles; tile += npairs, ++it) {
        const uint32_t ds = it & 1;
        CDL_TW(tw0, mbar_wait_cluster(&dempty[ds], ((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t dcol = tbase + kColD + ds * kNA;
        for (int ch = 0; ch < kChunks; ++ch, ++gchunk) {
          const uint32_t slot = gchunk % kASlots;
          CDL_TW(tw1, mbar_wait_cluster(&afull[slot], (gchunk / kASlots) & 1));
          tc_fence_after();
          const uint32_t a0 = tbase + kColA + slot * kASlot;
          uint64_t bdesc = bdesc0 + (uint64_t)(ch * kChunkRows) * kBStep;   // descriptor start-address field advances by kBStep per k-step
          if (ch < kChunks - 1) {
#pragma unroll
            for (int j = 0; j < kChunkRows; ++j) mma_tf32_ts<2>(dcol, a0 + j * 8, bdesc + (uint64_t)j * kBStep, idesc, (ch | j) != 0);
          } else {
#pragma unroll
            for (int j = 0; j < kRowsLast; ++j) mma_tf32_ts<2>(dcol, a0 + j * 8, bdesc + (uint64_t)j * kBStep, idesc, 1);
          }
          mma_commit<2>(&aempty[slot]);             // A slot reusable once these MMAs have read it
        }
        mma_commit<2>(&dfull[ds]);                  // accumulator complete -> epilogue (both CTAs)
      }
    }
    __syncwarp();                                   // reconverge the MMA warp before the aligned cluster barrier
  }
  if (p.dbg && lane == 0) {
    long long* d = p.dbg + ((size_t)blockIdx.x * 24 + warp) * 8;
    d[0] = clock64() - tstart; d[1] = tw0; d[2] = tw1; d[3] = tw2; d[4] = tw3; d[5] = tw4; d[6] = tw5;
  }
  // teardown: everyone done (all MMAs were consumed by the epilogues before they exit)
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc<2>(tbase, 512);
}

}  // namespace tc
}  // namespace cdl
