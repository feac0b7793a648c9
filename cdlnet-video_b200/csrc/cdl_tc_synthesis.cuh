// cdl_tc_synthesis.cuh — tcgen05 synthesis step for the video network (3D, P = 7x7x7, s = 2, C = 1):
//
//     out += B_k z        (reference model/net.py:205,210; nn.ConvTranspose3d, stride 2, output_padding 1)
//
// `out` is pre-initialised by the caller with -yp (so that it ends up holding the residual B z - yp) or
// with 0 (final D z).  Formulation: GEMM + col2im,
//     Pq[q, t] = sum_m z[q, m] * W[m, t]        M_gemm = 2 x 128 coarse sites / CTA pair, K = 176, N = 2 x 176
//     out[2q - 3 + t] += Pq[q, t]               (overlap-add of each site's 7x7x7 patch)
//
// Round-2 design (the round-1 kernel fed the code through 8 producer warps and a 12-hand-off TMEM ring per tile and
// drained the accumulators with 4 lock-stepped col2im warps: 36 % tensor-pipe activity):
//   * NO producer warps.  The code is stored pre-biased in UMMA core-matrix order (cdl_tc_analysis.cuh), so the A
//     operand goes HBM -> TMA -> shared memory -> tensor core (SS form) without touching a register; the tensor core's
//     truncation of the pre-biased words IS the round-to-nearest tf32 rounding the parity bar needs.
//   * A tile = ONE coarse row x 128 sites per CTA (16 groups of 8 sites = 128 TMEM lanes); the two CTAs of a pair
//     (cta_group::2, M = 256) work on two consecutive tiles and share the resident filter bank (each keeps half: 124 KB).
//   * Both tap halves (25 + 24 rows of 7 taps, N = 176 each) accumulate in the same K loop, D0 | D1 = 352 TMEM columns:
//     every 16 KB A chunk (4 K-steps) is read from L2 once and used by 8 MMAs.  44 MMAs per tile issued by one
//     converged warp (3872 tensor-pipe cycles) from a ring of 3 slots (48 KB in flight), 6 waits + 6 commits per tile;
//     both CTAs' TMA loads complete on the LEADER's barrier (cta_group::2 form): no relay hop.
//   * col2im by 16 warps (4 per TMEM lane quadrant, a quarter of the 49 (th,td) rows each).  With a one-row tile every
//     (th,td) row lands on its own footprint row, so the warps never meet except on the 5 seam columns between
//     quadrants, which go to a small private spill array merged by the flush: no lock step, no named barriers, no
//     atomics inside the pass.  The w-direction overlap-add runs on warp shuffles (lane = site, permuted by the group order).
//   * A CTA sweeps consecutive coarse rows qh of one (n, qd, w-tile) column; the footprints of successive tiles overlap
//     in h inside a ring of 7 fine rows (x 7 fine frames x 264 columns): after each tile only the 2 rows that became
//     final are added to `out` (red.global.add.v4.f32) and cleared - by the same 16 warps, while the MMAs of the next
//     tile run.  The MMA and col2im phases of one accumulator set alternate (352 + 352 columns do not fit TMEM twice).
//   * Tiles are dealt to the CTAs as equal contiguous ranges of the (column, qh) sequence: perfect balance, long sweeps.
#pragma once
#include "cdl_common.cuh"
#include "cdl_tc_ptx.cuh"
#include "cdl_tc_analysis.cuh"

namespace cdl {
namespace tc {

constexpr int kKB = 176;                  // GEMM K of the synthesis (subbands, padded) = channels of the code layout
constexpr int kKBSteps = kKB / 8;         // 22
constexpr int kNBP = 176;                 // GEMM N per tap half
constexpr int kRowsP0 = 25;               // (th,td) rows of 7 taps in half 0 (half 1: 24)
constexpr int kSTileW = 128;              // coarse sites per tile: one row, 8 blocks of 16 = 16 groups of 8
constexpr int kSGroups = kSTileW / 8;
#ifndef CDL_SYN_KS
#define CDL_SYN_KS 4
#endif
#ifndef CDL_SYN_SLOTS
#define CDL_SYN_SLOTS 3
#endif
constexpr int kSChunkKS = CDL_SYN_KS;      // K-steps per A ring slot (the last chunk of a tile holds 2: the TMA box runs past the
                                          // 176 subbands and is zero-filled there)
constexpr int kSChunkK4 = 2 * kSChunkKS;  // 4-subband chunks per slot
constexpr int kSChunkFloats = kSGroups * kSChunkK4 * kCodeChunk;     // 4096 floats = 16 KB
constexpr int kSChunks = (kKBSteps + kSChunkKS - 1) / kSChunkKS;     // 6 chunks per tile
constexpr int kSSlots = CDL_SYN_SLOTS;    // A ring depth: 48 KB in flight.  Every slot is used exactly twice per tile, so slot AND
                                          // barrier parity of chunk c are compile-time (c % 3, c / 3).  Measured: a wait + commit pair
                                          // costs the issuing warp ~300 cycles, so it must be amortised over >= 8 MMAs (704 pipe cycles)
static_assert(kSChunks % kSSlots == 0 && (kSChunks / kSSlots) % 2 == 0, "every slot must be used an even number of times per tile (compile-time barrier parities)");
constexpr int kXW = 264;                  // footprint columns: col c <-> fine w = 2*qw0 - 4 + c (cols 1..261 used)
constexpr int kXPl = 7;                   // fine frames of one coarse frame's footprint
constexpr int kXRing = 7;                 // ring of fine rows: the 2 rows a tile finishes are flushed (by the same warps) before the
                                          // next tile's col2im starts, so they are free again for the 2 rows that tile opens
constexpr int kXTile = kXRing * kXPl * kXW;
constexpr int kXSpill = kXRing * kXPl * 4 * 8;   // per (ring row, frame, quadrant): 3 left + 2 right seam columns (padded to 8)
constexpr int kSynC2iWarps = 16;
constexpr int kSynMmaWarp = 16, kSynLoadWarp = 17;
constexpr int kSynThreads = 32 * 18;      // 576 (5 warps on two of the four sub-partitions: 96 registers per thread)
constexpr int kColD0 = 0, kColD1 = kNBP;  // TMEM: D0 | D1 (352 of 512 columns)

struct SynTcParams {
  Geo g;
  const float* z;       // code in the internal layout (L2 prefetch only; the operand itself comes through `zmap`)
  float* out;           // (N,1,Fd,Fh,Fw), accumulated into
  const float* wpack;   // this layer: [2 ranks][2 halves][22 k-steps][11 groups][2][8][4]
  int tiles_w;          // 128-site tiles per coarse row
  int sweep;            // tile order (see syn_tile)
  long long ntiles;     // N * Qd * tiles_w * Qh
  long long nrows;      // N * Qd * Qh rows of the code tensor (a tile coordinate >= nrows reads zeros)
  long long* dbg;
  int dbg_mode;         // development aid (results invalid): 64 = col2im skipped, 128 = flush skipped, 256 = no TMA data movement
};

constexpr size_t kSynSmemB = (size_t)2 * kKBSteps * (kNBP / 2) * 8 * sizeof(float);             // 123904
constexpr size_t kSynSmemX = ((size_t)kXTile * sizeof(float) + 127) / 128 * 128;                // 51840
constexpr size_t kSynSmemA = (size_t)kSSlots * kSChunkFloats * sizeof(float);                   // 49152
constexpr size_t kSynSmemS = (size_t)kXSpill * sizeof(float);                                   // 6272
constexpr size_t kSynSmemBytes = kSynSmemB + kSynSmemX + kSynSmemA + kSynSmemS + 512;   // + 48 mbarriers, TMEM slot

// filters (M,1,7,7,7) [index (m, td, th, tw)] -> B[half][n = 7*row + tw, k = m], per-rank UMMA layout, tf32 RNE.
// Half 0 holds rows 0..24, half 1 rows 25..48 of the (th,td) row list (th-major); unused columns are zero.
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
    const int j = rank * (kNBP / 2) + grp * 8 + r8;            // column inside the half's accumulator
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

// ---- tile sequence ----
struct SynTile { long long row; int n, qd, qh, qw0, valid, first, last; };
// Two orders, both dealing EQUAL CONTIGUOUS shares to the CTAs (a RUN is a maximal stretch of consecutive qh of one
// (n, qd, w-tile) column inside a share; the footprint ring carries over inside a run and is flushed completely at its end):
//   sweep = 0  tiles numbered column by column ((n, qd, w-tile), w-tile fastest), qh fastest inside a column; CTA `cta` owns
//              the range [cta*T/nctas, (cta+1)*T/nctas) of the whole launch.  Longest runs; used when a coarse frame has
//              too few tiles to keep every CTA busy on it (`out` then fits L2 anyway).
//   sweep = 1  the (w-tile, qh) tiles of ONE coarse frame are dealt to the CTAs once (share [cta*F/nctas, (cta+1)*F/nctas)),
//              and every CTA sweeps all (n, qd) frames over its share.  All CTAs work on the same coarse frame at the same
//              time, so the 3-4 scatter-adds a fine voxel receives from neighbouring coarse frames meet in L2 (in the
//              other order neighbouring frames are processed 4320 tiles apart on the 1080p clip: every red.global.add
//              was a DRAM read-modify-write, 15 of 59 GB per launch, profiles/r02o_ncu_kernels.md).
struct SynShare { long long a; long long count; int len; };
__device__ __forceinline__ SynShare syn_range(const SynTcParams& p, int cta, int nctas) {
  SynShare s;
  if (p.sweep) {
    const long long F = (long long)p.tiles_w * p.g.Qh;
    s.a = F * cta / nctas;
    s.len = (int)(F * (cta + 1) / nctas - s.a);
    s.count = (long long)s.len * p.g.N * p.g.Qd;
  } else {
    s.a = p.ntiles * cta / nctas;
    s.count = p.ntiles * (cta + 1) / nctas - s.a;
    s.len = 0;
  }
  return s;
}
__device__ __forceinline__ SynTile syn_tile(const SynTcParams& p, const SynShare& sh, int i) {
  SynTile t;
  t.valid = i < sh.count;
  if (!t.valid) { t.row = p.nrows; t.n = t.qd = t.qh = t.qw0 = 0; t.first = t.last = 0; return t; }
  if (p.sweep) {
    const int fi = i / sh.len, j = i - fi * sh.len;
    const long long s = sh.a + j;
    const int wt = (int)(s / p.g.Qh);
    t.qh = (int)(s - (long long)wt * p.g.Qh);
    t.qw0 = wt * kSTileW;
    t.n = fi / p.g.Qd; t.qd = fi - t.n * p.g.Qd;
    t.first = (j == 0) || t.qh == 0;
    t.last = (j == sh.len - 1) || t.qh == p.g.Qh - 1;
  } else {
    const long long tau = sh.a + i;
    long long col = tau / p.g.Qh;
    t.qh = (int)(tau - col * p.g.Qh);
    t.qw0 = (int)(col % p.tiles_w) * kSTileW; col /= p.tiles_w;
    t.qd = (int)(col % p.g.Qd);
    t.n = (int)(col / p.g.Qd);
    t.first = (i == 0) || t.qh == 0;
    t.last = (i == sh.count - 1) || t.qh == p.g.Qh - 1;
  }
  t.row = ((long long)t.n * p.g.Qd + t.qd) * p.g.Qh + t.qh;
  return t;
}

// ---- col2im of NR consecutive (th,td) rows whose taps sit in u[7*i .. 7*i+6] ----
// lane = site (permuted); x0 / x1 = the two fine voxels of the site's own cell after the w-direction overlap-add:
//   fine 2o   : taps 3 (own), 1 (o+1), 5 (o-1)          fine 2o+1 : taps 4 (own), 2 (o+1), 0 (o+2), 6 (o-1)
// A quadrant's 64 own columns are touched by its warps only.  What spills over the quadrant's ends (fine -3..-1 from
// site 0, fine 64, 65 from site 31) belongs to columns another quadrant's warps own: it goes to a small private spill
// array instead (per ring row, frame and quadrant) and is merged by the flush - plain read-modify-writes everywhere
// (fp32 shared-memory atomics are CAS loops on this architecture: measured 1.7x slower than the round-1 kernel).
struct C2iLane { int src1, src2, srcm; float m1, m2, mm; int o; };

template <int ROW0, int NR>
__device__ __forceinline__ void c2i_rows(const uint32_t (&u)[7 * NR], float* xs, float* ss, int pbase, int colbase, int q, const C2iLane& L) {
  const unsigned full = 0xffffffffu;
  float x0[NR], x1[NR], e3[NR];
  int cell[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int row = ROW0 + i, th = row / 7, td = row % 7;
    const float v0 = __uint_as_float(u[7 * i]), v1 = __uint_as_float(u[7 * i + 1]), v2 = __uint_as_float(u[7 * i + 2]),
                v3 = __uint_as_float(u[7 * i + 3]), v4 = __uint_as_float(u[7 * i + 4]), v5 = __uint_as_float(u[7 * i + 5]),
                v6 = __uint_as_float(u[7 * i + 6]);
    const float a1 = __shfl_sync(full, v1, L.src1), a2 = __shfl_sync(full, v2, L.src1), n0 = __shfl_sync(full, v0, L.src1);
    const float a0 = __shfl_sync(full, v0, L.src2);
    const float b5 = __shfl_sync(full, v5, L.srcm), b6 = __shfl_sync(full, v6, L.srcm);
    x0[i] = fmaf(b5, L.mm, fmaf(a1, L.m1, v3));
    x1[i] = fmaf(b6, L.mm, fmaf(a0, L.m2, fmaf(a2, L.m1, v4)));
    e3[i] = v2 + n0;
    int ps = pbase + th; if (ps >= kXRing) ps -= kXRing;          // ring slot of fine row 2*qh + th
    cell[i] = ps * kXPl + td;
  }
  float2 cur[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) cur[i] = *reinterpret_cast<const float2*>(xs + cell[i] * kXW + colbase);
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    cur[i].x += x0[i]; cur[i].y += x1[i];
    *reinterpret_cast<float2*>(xs + cell[i] * kXW + colbase) = cur[i];
  }
  if (L.o == 0 || L.o == 31) {                                     // one divergent region, two lanes: the seam columns
    const int off = L.o ? 4 : 0;                                   // [0..2] left spill (fine -3,-2,-1), [4..5] right spill (fine 64, 65)
    float4 sp[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) sp[i] = *reinterpret_cast<const float4*>(ss + (cell[i] * 4 + q) * 8 + off);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      sp[i].x += L.o ? __uint_as_float(u[7 * i + 5]) : __uint_as_float(u[7 * i]);
      sp[i].y += L.o ? __uint_as_float(u[7 * i + 6]) : __uint_as_float(u[7 * i + 1]);
      sp[i].z += L.o ? 0.0f : e3[i];
      *reinterpret_cast<float4*>(ss + (cell[i] * 4 + q) * 8 + off) = sp[i];
    }
  }
}

// rows [ROW0, ROW0 + NROWS) of the (th,td) row list, accumulator columns starting at `acol` (row ROW0's tap 0), drained
// four rows (one 32-column tcgen05.ld) at a time; the next load is in flight while the current rows are applied
// `released`: called once the last tcgen05.ld of the part has completed (the accumulator can go back to the MMA warp
// while the last rows are still being applied)
template <int ROW0, int NROWS, int DONE = 0, typename Rel>
__device__ __forceinline__ void c2i_part(uint32_t acol, uint32_t (&u)[32], float* xs, float* ss, int pbase, int colbase, int q, const C2iLane& L,
                                         Rel released) {
  using namespace ptx;
#ifndef CDL_SYN_EXP
#define CDL_SYN_EXP 0      // experiments (results invalid): 1 = TMEM reads only, 2 = everything but the TMEM reads
#endif
  if constexpr (DONE == 0 && CDL_SYN_EXP != 2) tmem_ld32(acol, u);
  constexpr int NR = (NROWS - DONE) < 4 ? (NROWS - DONE) : 4;
  tmem_wait_ld();
  uint32_t v[7 * NR];
#pragma unroll
  for (int i = 0; i < 7 * NR; ++i) v[i] = u[i];
  if constexpr (DONE + NR < NROWS) { if constexpr (CDL_SYN_EXP != 2) tmem_ld32(acol + 7 * (DONE + NR), u); }
  else released();
  if constexpr (CDL_SYN_EXP != 1) c2i_rows<ROW0 + DONE, NR>(v, xs, ss, pbase, colbase, q, L);
  else if (v[0] == 0x7fc12345u) xs[0] = 1.0f;        // keep the loads alive
  if constexpr (DONE + NR < NROWS) c2i_part<ROW0, NROWS, DONE + NR>(acol, u, xs, ss, pbase, colbase, q, L, released);
}

// LO = false: A = rna_tf32(z) (what the tensor core reads from the pre-biased words)
// LO = true : A = rna_tf32(z - rna_tf32(z)), formed in shared memory by the (then idle) col2im warps - the low part used
//             by the 3-term final dictionary synthesis
template <bool LO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSynThreads, 1) k_tc_synthesis(const SynTcParams p, const __grid_constant__ CUtensorMap zmap) {
  using namespace ptx;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sB = reinterpret_cast<float*>(smem_raw);
  float* sX = reinterpret_cast<float*>(smem_raw + kSynSmemB);
  float* sA = reinterpret_cast<float*>(smem_raw + kSynSmemB + kSynSmemX);
  float* sS = reinterpret_cast<float*>(smem_raw + kSynSmemB + kSynSmemX + kSynSmemA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kSynSmemB + kSynSmemX + kSynSmemA + kSynSmemS);
  uint64_t* wbar = bars + 0;
  uint64_t* wready = bars + 1;                 //      (leader) the peer CTA's filters have landed
  uint64_t* afull = bars + 2;                  // [3]  TMA: (leader) BOTH CTAs' A chunks landed -> MMA;  LO: this CTA's chunk landed
  uint64_t* aboth = afull + kSSlots;           // [3]  LO only: (leader) chunk transformed in place by the col2im warps of both CTAs -> MMA
  uint64_t* aempty = aboth + kSSlots;          // [3]  MMA commit (multicast) -> TMA warps
  uint64_t* dfull = aempty + kSSlots;          //      MMA commit (multicast) -> col2im of both CTAs
  uint64_t* dempty = dfull + 1;                //      (leader) col2im warps of both CTAs -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 1);

  const Geo& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  long long tw0 = 0, tw1 = 0, tw2 = 0;
  const long long tstart = clock64();

  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(wready, 1);
    for (int i = 0; i < kSSlots; ++i) {
      mbar_init(&afull[i], 1); mbar_init(&aboth[i], 2 * kSynC2iWarps); mbar_init(&aempty[i], 1);
    }
    mbar_init(dfull, 1);
    mbar_init(dempty, 2 * kSynC2iWarps);
    fence_mbar_init();
  }
  if (warp == kSynMmaWarp) { tmem_alloc<2>(tmem_slot, 512); tmem_relinquish<2>(); }
  for (int i = tid; i < kXTile; i += kSynThreads) sX[i] = 0.0f;
  for (int i = tid; i < kXSpill; i += kSynThreads) sS[i] = 0.0f;
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

  // this CTA's tile range and the pair's common number of rounds
  const SynShare share = syn_range(p, blockIdx.x, gridDim.x);
  const int rounds = (int)max(share.count, syn_range(p, blockIdx.x ^ 1, gridDim.x).count);

  if (warp < kSynC2iWarps) {
    // ============================== col2im + footprint flush ==============================
    const int q = warp & 3, part = warp >> 2;
    const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16);
    // lane -> site: TMEM lane 8*g' + i of the quadrant is row i of group g' = site 16*(g'>>1) + 2*i + (g'&1)
    C2iLane L;
    {
      const int gq = lane >> 3, i8 = lane & 7;
      L.o = 16 * (gq >> 1) + 2 * i8 + (gq & 1);
      auto lane_of = [](int o) { o &= 31; return 8 * (2 * (o >> 4) + (o & 1)) + ((o & 15) >> 1); };
      L.src1 = lane_of(L.o + 1); L.src2 = lane_of(L.o + 2); L.srcm = lane_of(L.o + 31);
      L.m1 = L.o < 31 ? 1.0f : 0.0f; L.m2 = L.o < 30 ? 1.0f : 0.0f; L.mm = L.o > 0 ? 1.0f : 0.0f;
    }
    const int colbase = 4 + 64 * q + 2 * L.o;
    for (int it = 0; it < rounds; ++it) {
      const SynTile t = syn_tile(p, share, it);
      if (LO) {
        // low part of the code, in place: word -> z = word - bias, hi = truncate(word) = rna(z), word' = (z - hi) + bias
#pragma unroll 1
        for (int c = 0; c < kSChunks; ++c) {
          const int slot = c % kSSlots;
          mbar_wait(&afull[slot], c / kSSlots);
          auto lo_word = [](float x) {
            const uint32_t b = __float_as_uint(x);
            if (b == 0u) return 0.0f;                                                   // TMA zero fill (no site): stays zero
            const float z = __uint_as_float(b - kCodeBias), hi = __uint_as_float(b & 0xffffe000u);
            return __uint_as_float(__float_as_uint(z - hi) + kCodeBias);
          };
#pragma unroll
          for (int h = 0; h * 32 * kSynC2iWarps < kSChunkFloats / 4; ++h) {             // 512 threads x 16 B per step
            const int idx = h * 32 * kSynC2iWarps + tid;
            if (idx < kSChunkFloats / 4) {
              float4* pa = reinterpret_cast<float4*>(sA + slot * kSChunkFloats) + idx;
              float4 w = *pa;
              w.x = lo_word(w.x); w.y = lo_word(w.y); w.z = lo_word(w.z); w.w = lo_word(w.w);
              *pa = w;
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_release(&aboth[slot], 0);                // 32 warps of the pair -> the leader's MMA warp
        }
      }
      CDL_TW(tw0, mbar_wait(dfull, it & 1));
      tc_fence_after();
      const int pbase = (2 * t.qh) % kXRing;
      auto release = [&]() {                                       // this warp's columns are in registers: accumulators free again
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (rank == 0) mbar_arrive(dempty); else mbar_arrive_cluster(dempty, 0); }
      };
#ifdef CDL_TC_PROFILE
      const long long tc0 = clock64();
#endif
      if (t.valid && !(p.dbg_mode & 64)) {
        uint32_t u[32];
        switch (part) {                                            // warp-uniform: a quarter of the 49 rows each
          case 0: c2i_part<0, 13>(lane_addr + kColD0, u, sX, sS, pbase, colbase, q, L, release); break;
          case 1: c2i_part<13, 12>(lane_addr + kColD0 + 7 * 13, u, sX, sS, pbase, colbase, q, L, release); break;
          case 2: c2i_part<25, 12>(lane_addr + kColD1, u, sX, sS, pbase, colbase, q, L, release); break;
          default: c2i_part<37, 12>(lane_addr + kColD1 + 7 * 12, u, sX, sS, pbase, colbase, q, L, release); break;
        }
      } else {
        release();
      }
#ifdef CDL_TC_PROFILE
      const long long tc1 = clock64();
      tw1 += tc1 - tc0;
#endif
      named_bar_sync(1, 32 * kSynC2iWarps);                        // every warp's rows are in the ring
      if (t.valid && !(p.dbg_mode & 128)) {
        // fine rows 2*qh and 2*qh+1 are final (all 7 at the end of a run): out += row, clear the ring slot.  Runs while
        // the MMAs of the next tile are in flight.
        const int nfl = t.last ? kXPl : 2;
        const int items = nfl * kXPl * (kXW / 4);
        float* on = p.out + (size_t)t.n * g.fine_vol();
        const int gw0 = 2 * t.qw0 - 4;
        for (int item = tid; item < items; item += 32 * kSynC2iWarps) {
          const int c4 = item % (kXW / 4), rest = item / (kXW / 4);
          const int td = rest % kXPl, pi = rest / kXPl;
          const int pr = 2 * t.qh + pi;                            // fine row along the sweep; fine h = pr - oh
          const int gd = 2 * t.qd + td - g.od, gh = pr - g.oh, gw = gw0 + 4 * c4;
          const int rc = (pr % kXRing) * kXPl + td;
          float4* cell = reinterpret_cast<float4*>(sX + rc * kXW + 4 * c4);
          float4 v = *cell;
          *cell = make_float4(0.f, 0.f, 0.f, 0.f);
          // the seam columns: quadrant qq's left spill lands on columns 64 qq + 1..3, its right spill on 64 qq + 68, 69
          if ((c4 & 15) == 0 && c4 < 64) {
            float4* sp = reinterpret_cast<float4*>(sS + (rc * 4 + (c4 >> 4)) * 8);
            const float4 l = *sp;
            *sp = make_float4(0.f, 0.f, 0.f, 0.f);
            v.y += l.x; v.z += l.y; v.w += l.z;
          } else if ((c4 & 15) == 1 && c4 >= 17) {
            float4* sp = reinterpret_cast<float4*>(sS + (rc * 4 + ((c4 - 17) >> 4)) * 8 + 4);
            const float4 r = *sp;
            *sp = make_float4(0.f, 0.f, 0.f, 0.f);
            v.x += r.x; v.y += r.y;
          }
          if (gd >= 0 && gd < g.Fd && gh >= 0 && gh < g.Fh && gw >= 0 && gw + 4 <= g.Fw)
            red_add_v4_f32(on + ((size_t)gd * g.Fh + gh) * g.Fw + gw, v);
        }
      }
      named_bar_sync(1, 32 * kSynC2iWarps);   // the flushed rows are clear before the next tile's col2im reuses their ring slots
#ifdef CDL_TC_PROFILE
      tw2 += clock64() - tc1;
#endif
    }
  } else if (warp == kSynMmaWarp) {
    // ============================== MMA issue (leader CTA; converged warp, elected lane) ==============================
    if (rank == 1 && lane == 0) { mbar_wait(wbar, 0); mbar_arrive_cluster(wready, 0); }
    if (rank == 0) {
      CDL_TW(tw2, mbar_wait(wbar, 0); mbar_wait_cluster(wready, 0));
      const uint32_t idesc = make_idesc_tf32(256, kNBP);
      const uint64_t bdesc0 = make_smem_desc_kmajor_noswz(smem_u32(sB), 128, 256);
      const uint64_t adesc0 = make_smem_desc_kmajor_noswz(smem_u32(sA), 128, kSChunkK4 * 128);
      constexpr uint32_t kBStep = ((kNBP / 2) * 32) >> 4;           // 16-byte units between k-steps of B (88 rows x 32 B)
      for (int it = 0; it < rounds; ++it) {
        CDL_TW(tw0, mbar_wait_cluster(dempty, (it & 1) ^ 1));       // col2im of the previous tile has drained D0 | D1
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kSChunks; ++c) {                        // slot, barrier parity and every descriptor: constants
          const int slot = c % kSSlots;
          CDL_TW(tw1, mbar_wait_cluster(LO ? &aboth[slot] : &afull[slot], c / kSSlots));
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < kSChunkKS; ++j) {
            const int ks = kSChunkKS * c + j;
            if (ks < kKBSteps) {
              const uint64_t ad = adesc0 + (uint64_t)((slot * kSChunkFloats * 4 + j * 256) >> 4);
              mma_tf32_ss_warp<2>(tbase + kColD0, ad, bdesc0 + (uint64_t)ks * kBStep, idesc, ks != 0);
              mma_tf32_ss_warp<2>(tbase + kColD1, ad, bdesc0 + (uint64_t)(kKBSteps + ks) * kBStep, idesc, ks != 0);
            }
          }
          mma_commit_warp<2>(&aempty[slot]);          // A slot reusable once these MMAs have read it (both CTAs)
        }
        mma_commit_warp<2>(dfull);                    // both accumulators complete -> col2im (both CTAs)
      }
    }
    __syncwarp();
  } else {
    // ============================== TMA: code tile -> A ring, 8 KB chunks ==============================
    if (lane == 0) {
      tma_prefetch_desc(&zmap);
      const uint64_t pol = l2_policy_evict_first();                 // the code streams through once: keep L2 for `out` (scatter-adds)
      const int G = code_groups_per_row(g.Qw);
      for (int it = 0; it < rounds; ++it) {
        const SynTile t = syn_tile(p, share, it);
        const int g0 = (t.qw0 >> 4) * 2;                             // first group of the tile in its row
        {                                                            // a tile's 16 groups are one contiguous run: one L2 prefetch
          const SynTile t2 = syn_tile(p, share, it + 1);            // (distance 2 / no evict-first hint: same kernel time, profiles/r02x_*)
          if (t2.valid) {
            const int g2 = (t2.qw0 >> 4) * 2, ng = min(kSGroups, G - g2);
            bulk_prefetch_l2_hint(p.z + ((size_t)t2.row * G + g2) * kCodeGroup, (uint32_t)ng * kCodeGroup * 4, pol);
          }
        }
#pragma unroll 1
        for (int c = 0; c < kSChunks; ++c) {
          const int slot = c % kSSlots;
          CDL_TW(tw0, mbar_wait(&aempty[slot], (c / kSSlots) ^ 1));
          if (LO) {
            // the chunk is transformed in place before the MMA may read it: completion on this CTA's own barrier, then
            // the col2im warps of both CTAs arrive on the leader's aboth
            mbar_expect_tx(&afull[slot], kSChunkFloats * 4);
            tma_load_3d_hint(sA + slot * kSChunkFloats, &zmap, c * kSChunkK4 * kCodeChunk, g0, (int)t.row, &afull[slot], pol);
          } else {
            // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of both
            if (p.dbg_mode & 256) { if (rank == 0) mbar_arrive(&afull[slot]); continue; }
            if (rank == 0) mbar_expect_tx(&afull[slot], 2 * kSChunkFloats * 4);
            tma_load_3d_2cta(sA + slot * kSChunkFloats, &zmap, c * kSChunkK4 * kCodeChunk, g0, (int)t.row, &afull[slot], pol);
          }
        }
      }
    }
    __syncwarp();
  }
  if (p.dbg && lane == 0) {
    long long* d = p.dbg + ((size_t)blockIdx.x * 24 + warp) * 8;
    d[0] = clock64() - tstart; d[1] = tw0; d[2] = tw1; d[3] = tw2;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == kSynMmaWarp) tmem_dealloc<2>(tbase, 512);
}

// ---- code layout conversion: internal (groups of 8 same-parity sites, pre-biased) <-> reference (N,M,Qd,Qh,Qw) ----
// One block = 32 consecutive sites of one row (4 groups) x 32 subbands, transposed through shared memory.  On the
// internal side a thread moves one float4 = 4 subbands of one site; 8 lanes fill a 128-byte chunk.
// grid = (rows * ceil(Qw/32), 176/32 rounded up, N); R = Qd*Qh rows per sample.
__device__ __forceinline__ void code_tile_piece(int t, int& grp, int& k4l, int& i8, int& qloc) {
  grp = t >> 6; k4l = (t >> 3) & 7; i8 = t & 7;
  qloc = 16 * (grp >> 1) + 2 * i8 + (grp & 1);             // site inside the 32-site segment
}
__global__ void __launch_bounds__(256) k_code_export(const float* __restrict__ zcl, float* __restrict__ z, int R, int Qw, int M) {
  __shared__ float t[32][33];                               // [site][subband]
  const int nseg = (Qw + 31) >> 5;
  const int row = blockIdx.x / nseg, qw0 = (blockIdx.x % nseg) * 32;
  const int m0 = blockIdx.y * 32, n = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = code_groups_per_row(Qw);
  int grp, k4l, i8, qloc;
  code_tile_piece(threadIdx.x, grp, k4l, i8, qloc);
  const int gidx = (qw0 >> 4) * 2 + grp, k4 = (m0 >> 2) + k4l;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gidx < G && k4 < kCodeK4) {
    const float4 w = *reinterpret_cast<const float4*>(zcl + (((size_t)n * R + row) * G + gidx) * kCodeGroup + (size_t)k4 * kCodeChunk + i8 * 4);
    v = make_float4(code_dec(w.x), code_dec(w.y), code_dec(w.z), code_dec(w.w));
  }
  t[qloc][4 * k4l] = v.x; t[qloc][4 * k4l + 1] = v.y; t[qloc][4 * k4l + 2] = v.z; t[qloc][4 * k4l + 3] = v.w;
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
  const int G = code_groups_per_row(Qw);
  const size_t Q = (size_t)R * Qw;
  for (int r = warp; r < 32; r += 8) {
    const int m = m0 + r, qw = qw0 + lane;
    t[lane][r] = (m < M && qw < Qw) ? z[((size_t)n * M + m) * Q + (size_t)row * Qw + qw] : 0.0f;     // layout padding = 0
  }
  __syncthreads();
  int grp, k4l, i8, qloc;
  code_tile_piece(threadIdx.x, grp, k4l, i8, qloc);
  const int gidx = (qw0 >> 4) * 2 + grp, k4 = (m0 >> 2) + k4l;
  if (gidx < G && k4 < kCodeK4)
    *reinterpret_cast<float4*>(zcl + (((size_t)n * R + row) * G + gidx) * kCodeGroup + (size_t)k4 * kCodeChunk + i8 * 4) =
        make_float4(code_enc(t[qloc][4 * k4l]), code_enc(t[qloc][4 * k4l + 1]), code_enc(t[qloc][4 * k4l + 2]), code_enc(t[qloc][4 * k4l + 3]));
}

}  // namespace tc
}  // namespace cdl
