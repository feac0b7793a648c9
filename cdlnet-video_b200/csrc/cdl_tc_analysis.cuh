// cdl_tc_analysis.cuh — tcgen05 analysis step for the video network (3D, P = 7x7x7, s = 2, C = 1):
//
//     z <- ST(z - A_k r, t0 + c*t1)            (reference model/net.py:200,205 + :11-14)
//
// as an implicit GEMM  U[q, m] = sum_t R[q, t] * W[m, t]  with  M_gemm = 256 coarse sites per CTA pair,
// N = 176 (169 subbands, zero padded), K = 392 (49 (td,th) rows x 8: the 7 w-taps plus the 8-byte-aligned window's
// first element, whose filter column is zero), kind::tf32, fp32 accumulation.
//
//   * cta_group::2: the two CTAs of a cluster share the filter bank: each keeps HALF of it (88 subbands x 392
//     taps = 135 KB) resident in shared memory for the whole launch.
//   * The A operand (im2col of the residual, C = 1) never exists anywhere: the MMA reads it straight from a TMA-loaded
//     halo tile of r through an OVERLAPPING shared-memory descriptor.  With stride 2 the 8-float windows of sites
//     q and q+2 start 16 bytes apart, so 8 same-parity sites of one row form a legal K-major core matrix (8 rows
//     x 16 B, contiguous) and the second half of the window is the same memory 16 B further on (LBO = 16).  A CTA
//     therefore takes the sites of ONE w-parity (rank 0 even, rank 1 odd) of a 16 x 16 pair tile: 16 rows x 8
//     sites = 128 TMEM lanes, one 144-byte row window per 8-row group (SBO = 144).  Rows are stored h-parity major
//     (TMA element stride 2 along h), which makes the group stride constant for every (td,th).
//     No producer warps, no LDS/tcgen05.st traffic, no A ring: one warp waits for the tile and issues 49 MMAs.
//   * r is rounded to tf32 (RNE) once per element by k_round_tf32 (tcgen05 truncates its operands).
//   * Accumulators are double buffered in TMEM (2 x 176 columns): the epilogue of tile i (TMEM -> registers ->
//     z - u -> soft threshold -> z, one coalesced read + write of z) overlaps the MMAs of tile i+1.
//   * Persistent: one CTA pair per SM pair, static round-robin over 256-site tiles.
//
// Warp roles (576 threads): warps 0-15 epilogue (four per TMEM lane quadrant, each owning 5 or 6 of the 22 subband
// blocks and prefetching the next tile's z values into registers), warp 16 MMA issue + TMEM alloc, warp 17 TMA loads.
#pragma once
#include "cdl_common.cuh"
#include "cdl_tc_ptx.cuh"

namespace cdl {
namespace tc {

constexpr int kThreads = 416;            // synthesis kernel
constexpr int kMmaWarp = 12;             // synthesis kernel
constexpr int kNA = 176;                 // GEMM N of the analysis (subbands, padded)
constexpr int kNAH = kNA / 2;            // per-CTA half of the filter bank
constexpr int kP = 7, kTaps = 343;
constexpr int kKSteps = 49;              // one tf32 MMA K-step (8 columns) per (td,th) row: window element 0 (zero filter) + 7 taps
constexpr int kTH = 4, kTW = 32;         // synthesis: coarse tile per CTA 1 x 4 x 32 (d,h,w); a pair stacks two along h

// ---- analysis geometry ----
constexpr int kAEpiWarps = 16;            // four per TMEM lane quadrant: 6 / 6 / 5 / 5 of the 22 subband blocks
constexpr int kAThreads = 32 * (kAEpiWarps + 2);
constexpr int kAMmaWarp = kAEpiWarps, kALoadWarp = kAEpiWarps + 1;
constexpr int kATile = 16;               // pair tile: 16 x 16 coarse sites (h x w); each CTA takes one w-parity (16 rows x 8 sites)
constexpr int kARW = 36;                 // floats per row window: 8 sites x 4 + the 4-float tail of the last site (144 B)
constexpr int kARows = kATile + 3;       // rows per h-parity: coarse row + th/2, th/2 <= 3
constexpr int kARD = 7;
constexpr int kABox = kARW * kARows * kARD;          // 4788 floats = 19152 B per h-parity box
constexpr int kABoxPitch = 4800;                     // 19200 B: multiple of 128 B (TMA destination alignment)
constexpr int kABuf = 2 * kABoxPitch;                // one tile: both h-parities
constexpr int kColD = 0;                             // TMEM columns: D0 | D1

// ---- internal code layout of the tensor-core path: UMMA core matrices of 8 same-parity sites, pre-biased values ----
// Along w, a row of sites is cut into blocks of 16; the 8 even and the 8 odd sites of a block form one GROUP each
// (qw = 16 b + 2 i + p  ->  group 2 b + p, row i of the group).  A group stores its 176 subbands as 44 chunks of
// [8 sites][4 subbands] = 128 bytes: exactly one K-major UMMA core matrix.  Consequences:
//   * synthesis: a TMA box (c chunks x 16 groups) lands in shared memory as a legal no-swizzle K-major A operand
//     (LBO = 128 B between the two chunks of a K-step, SBO = c * 128 B between groups) - the code goes from HBM to the
//     tensor core without passing through a register (no producer warps);
//   * analysis: a CTA owns the sites of one w-parity (overlapping-descriptor im2col), so the 8 lanes of a tile row hold
//     the 8 rows of one group: a 128-bit access per lane fills a whole 128-byte line.
// Values are stored PRE-BIASED: word = bits(z) + 0x1000.  tcgen05 kind::tf32 truncates its fp32 operand words to 19 bits,
// and truncate(bits + 0x1000) == round-to-nearest(ties away)(z): the tensor core reads rna_tf32(z) straight from memory,
// while the analysis epilogue recovers the exact fp32 z with one integer subtraction.  (Storing z itself rounded to tf32
// costs 1.1-1.5e-4 on xhat; storing it unrounded and letting the tensor core truncate biases B z by 2^-11.)
// Offset (floats) of subband 0 of site (row, qw), row = (n*Qd + qd)*Qh + qh; subband m sits (m >> 2) * 32 + (m & 3) further on.
constexpr int kCodeChunk = 32;                      // floats per chunk: [8 sites][4 subbands]
constexpr int kCodeK4 = kNA / 4;                    // 44 chunks per group
constexpr int kCodeGroup = kCodeK4 * kCodeChunk;    // 1408 floats per group of 8 same-parity sites
constexpr uint32_t kCodeBias = 0x1000u;             // half a tf32 ulp, in fp32 mantissa units
__host__ __device__ __forceinline__ int code_groups_per_row(int Qw) { return 2 * ((Qw + 15) >> 4); }
__host__ __device__ __forceinline__ size_t code_site_offset(size_t row, int Qw, int qw) {
  return (row * (size_t)code_groups_per_row(Qw) + (size_t)(2 * (qw >> 4) + (qw & 1))) * kCodeGroup + (size_t)(((qw & 15) >> 1) * 4);
}
__host__ __device__ __forceinline__ size_t code_floats(size_t rows, int Qw) { return rows * (size_t)code_groups_per_row(Qw) * kCodeGroup; }
__device__ __forceinline__ float code_enc(float z) { return __uint_as_float(__float_as_uint(z) + kCodeBias); }
__device__ __forceinline__ float code_dec(float w) { return __uint_as_float(__float_as_uint(w) - kCodeBias); }

struct AnaTcParams {
  Geo g;
  float* z;             // internal code layout (see code_site_offset), pre-biased words, updated in place
  const float* wpack;   // this layer: [2 ranks][49 k-steps][11 groups][2][8][4] tf32-rounded filters
  const float* t0;      // [M]
  const float* t1;      // [M]
  const float* cvec;    // [N] or nullptr
  int first;
  int tiles_w, tiles_h; // pair tiles along w and h (16 x 16 sites)
  int ntiles;           // N * Qd * tiles_h * tiles_w
  long long* dbg;       // optional [grid][24 warps][8] cycle counters
  int dbg_mode;         // development aid (results invalid): 2 = epilogue skips the code traffic, 16 = no TMA loads, 32 = no MMAs
};

constexpr size_t kAnaSmemB = (size_t)kKSteps * kNAH * 8 * sizeof(float);      // 137984
constexpr size_t kAnaSmemR = 2 * (size_t)kABuf * sizeof(float);               // 76800
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
// would truncate.  Two copies are written: dst0 in r's own layout (read by the even-site CTA) and dst1 with every row
// shifted right by two floats (row pitch Fw + 4, dst1[x] = r[x - 2]) for the odd-site CTA, whose windows start at
// fine w = 2*q - 4 with q odd: TMA needs the box start 16-byte aligned in global memory.
// Image-sized streams (4 MB per 16x256x256 clip), negligible next to the code traffic.
// Inside cdl_forward the pass also re-arms the residual buffer for the next synthesis (reset[i] = -yp[i]; reset aliases
// src, which is dead once read) - one launch and one image pass less per iteration than a separate k_neg_copy.
// Temporal slabs (cdl_analysis_step_halo): on the Pd - s seam frames a rank shares with a neighbour both hold PARTIAL
// sums of B z; the neighbour's partial (received into prev / next, (N, ov, Fh, Fw) each) is added here, in the pass that
// reads r anyway.  In residual mode both partials carry -yp, so yp is added back once.
struct HaloFuse {
  const float* prev;      // partial sums of the previous rank on my first `ov` frames (or nullptr)
  const float* next;      // partial sums of the next rank on my last `ov` frames (or nullptr)
  const float* yp;        // non-null: add yp back on the seam frames (residual mode)
  int ov, Fd;
  long long frame4;       // float4s per frame: Fh * Fw / 4
};

template <bool HALO>
__global__ void __launch_bounds__(256) k_round_tf32(const float* src, float* __restrict__ dst0, float* __restrict__ dst1,
                                                    int Fw4, long long n4, const float* __restrict__ yp, float* reset, const HaloFuse h) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(src)[i];
    if (HALO) {
      const long long fr = i / h.frame4, within = i - fr * h.frame4;
      const int f = (int)(fr % h.Fd);
      const long long n = fr / h.Fd;
      const float* add = nullptr;
      if (h.prev && f < h.ov) add = h.prev + ((n * h.ov + f) * h.frame4 + within) * 4;
      else if (h.next && f >= h.Fd - h.ov) add = h.next + ((n * h.ov + (f - (h.Fd - h.ov))) * h.frame4 + within) * 4;
      if (add) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(add));
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        if (h.yp) {
          const float4 y = __ldg(reinterpret_cast<const float4*>(h.yp) + i);
          v.x += y.x; v.y += y.y; v.z += y.z; v.w += y.w;
        }
      }
    }
    if (reset) {
      const float4 y = __ldg(reinterpret_cast<const float4*>(yp) + i);
      reinterpret_cast<float4*>(reset)[i] = make_float4(-y.x, -y.y, -y.z, -y.w);
    }
    v.x = __uint_as_float(ptx::tf32_rna_bits(v.x)); v.y = __uint_as_float(ptx::tf32_rna_bits(v.y));
    v.z = __uint_as_float(ptx::tf32_rna_bits(v.z)); v.w = __uint_as_float(ptx::tf32_rna_bits(v.w));
    reinterpret_cast<float4*>(dst0)[i] = v;
    const long long row = i / Fw4;
    const int j = (int)(i - row * Fw4);
    float* d = dst1 + row * (4LL * Fw4 + 4) + 4 * j;
    if (j == 0) *reinterpret_cast<float2*>(d) = make_float2(0.0f, 0.0f);
    *reinterpret_cast<float2*>(d + 2) = make_float2(v.x, v.y);
    *reinterpret_cast<float2*>(d + 4) = make_float2(v.z, v.w);
    if (j == Fw4 - 1) *reinterpret_cast<float2*>(d + 6) = make_float2(0.0f, 0.0f);
  }
}

// r[seam frames] += neighbour's partial (+ yp): the stand-alone form of the fusion above, for the final D z and for
// the kernel families whose analysis step has no rounding pass.  One thread per float4 of the 2 * ov seam frames.
__global__ void __launch_bounds__(256) k_halo_add(float* __restrict__ r, const HaloFuse h, int N) {
  const long long per = (long long)h.ov * h.frame4;                  // float4s per sample and seam
  const long long total = 2 * per * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int side = (int)(i / (per * N));                            // 0: head seam (prev), 1: tail seam (next)
    const long long j = i - (long long)side * per * N;
    const float* src = side ? h.next : h.prev;
    if (!src) continue;
    const long long n = j / per, w = j - n * per;                     // w = frame-in-seam * frame4 + within
    const long long f0 = side ? (long long)(h.Fd - h.ov) : 0;
    const long long idx = (n * h.Fd + f0) * h.frame4 + w;
    float4 v = reinterpret_cast<float4*>(r)[idx];
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + j);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    if (h.yp) {
      const float4 y = __ldg(reinterpret_cast<const float4*>(h.yp) + idx);
      v.x += y.x; v.y += y.y; v.z += y.z; v.w += y.w;
    }
    reinterpret_cast<float4*>(r)[idx] = v;
  }
}

// Tile order: w fastest, then the coarse FRAME, then the 16-row band, then the sample.  A fine frame of r is needed by
// 3.5 coarse frames; with the frame index running inside a band the 74 pairs sweep a 38-row strip of r through all
// frames (reuse distance tiles_w tiles = a few MB, an L2 hit) instead of re-reading every frame from DRAM 3.5 times
// (measured on the 1080p clip with frames outermost: 14 GB of 102 GB per launch, profiles/r02o_ncu_kernels.md).
__device__ __forceinline__ void ana_tile_coords(const AnaTcParams& p, int tile, int& n, int& qd, int& qh0, int& qw0) {
  int tw = tile % p.tiles_w; tile /= p.tiles_w;
  qd = tile % p.g.Qd; tile /= p.g.Qd;
  int th = tile % p.tiles_h; n = tile / p.tiles_h;
  qh0 = th * kATile; qw0 = tw * kATile;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kAThreads, 1) k_tc_analysis(const AnaTcParams p, const __grid_constant__ CUtensorMap rmap0,
                                                                                            const __grid_constant__ CUtensorMap rmap1) {
  using namespace ptx;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sB = reinterpret_cast<float*>(smem_raw);
  float* sR = reinterpret_cast<float*>(smem_raw + kAnaSmemB);
  float* sT = reinterpret_cast<float*>(smem_raw + kAnaSmemB + kAnaSmemR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kAnaSmemB + kAnaSmemR + kAnaSmemT);
  uint64_t* wbar = bars + 0;
  uint64_t* wready = bars + 1;                 //      (leader) the peer CTA's filters have landed
  uint64_t* rfull = bars + 2;                  // [2]  TMA: this CTA's residual tile landed
  uint64_t* rboth = rfull + 2;                 // [2]  (leader) both CTAs' tiles landed -> MMA
  uint64_t* rempty = rboth + 2;                // [2]  MMA commit (multicast) -> loaders
  uint64_t* dfull = rempty + 2;                // [2]  MMA commit (multicast) -> epilogue
  uint64_t* dempty = dfull + 2;                // [2]  (leader) epilogue warps of both CTAs -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

  const Geo& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  long long tw0 = 0, tw1 = 0, tw2 = 0;
  const long long tstart = clock64();

  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(wready, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&rfull[i], 1); mbar_init(&rboth[i], 2); mbar_init(&rempty[i], 1);
      mbar_init(&dfull[i], 1); mbar_init(&dempty[i], 2 * kAEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kAMmaWarp) { tmem_alloc<2>(tmem_slot, 512); tmem_relinquish<2>(); }
  for (int i = tid; i < kNA; i += kAThreads) {
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

  if (warp == kALoadWarp) {
    // ============================== TMA: residual halo tile, one w-parity, h-parity major ==============================
    if (lane == 0) {
      const CUtensorMap* rmap = rank ? &rmap1 : &rmap0;     // odd sites read the copy shifted by two floats
      tma_prefetch_desc(rmap);
      // (An L2 tensor prefetch of the residual boxes two tiles ahead was measured: the MMA warp's wait for r tiles fell from
      //  15.5 % to 10.7 %, but the duplicated L2 traffic cost more - 0.641 vs 0.597 ms per launch on 16 clips, 22.6 vs 17.8 ms
      //  on the 1080p clip, profiles/r02x_*.  Not kept.)
      int it = 0;
      for (int tile = pair; tile < p.ntiles; tile += npairs, ++it) {
        const int buf = it & 1;
        int n, qd, qh0, qw0;
        ana_tile_coords(p, tile, n, qd, qh0, qw0);
        CDL_TW(tw0, mbar_wait(&rempty[buf], ((it >> 1) & 1) ^ 1));       // the MMAs of tile it-2 have read this buffer
        if (p.dbg_mode & 16) { mbar_arrive(&rfull[buf]); } else {
        mbar_expect_tx(&rfull[buf], 2 * kABox * 4);
        // window column 0 <-> fine w = 2*q - 4; box rows are fine h = f0, f0+2, ...: parity f0 & 1, half-row index f0 >> 1
        const int w0 = rank ? 2 * qw0 : 2 * qw0 - 4, d0 = 2 * qd - g.od;      // rank 1: r[2*qw0 - 2] sits at x = 2*qw0
        const int fe = 2 * qh0 - 3, fo = 2 * qh0 - 2;                     // first fine row for even / odd th
        tma_load_5d(sR + buf * kABuf, rmap, w0, fe >> 1, fe & 1, d0, n, &rfull[buf]);
        tma_load_5d(sR + buf * kABuf + kABoxPitch, rmap, w0, fo >> 1, fo & 1, d0, n, &rfull[buf]);
        }
        CDL_TW(tw1, mbar_wait(&rfull[buf], (it >> 1) & 1));
        if (rank == 0) mbar_arrive(&rboth[buf]); else mbar_arrive_cluster(&rboth[buf], 0);
      }
    }
    __syncwarp();
  } else if (warp < kAEpiWarps) {
    // ============================== epilogue: TMEM -> z update ==============================
    // warp = quad + 4*part: TMEM lanes [32*quad, 32*quad+32) = tile rows 4*quad..4*quad+3 x 8 sites of this CTA's
    // w-parity; subband blocks [b0, b0+nb) of the 22 (parts 0,1: 6 blocks, parts 2,3: 5 blocks)
    const int quad = warp & 3, part = warp >> 2;
    const int nb = part < 2 ? 6 : 5, b0 = part < 2 ? 6 * part : 12 + 5 * (part - 2);
    const int m0 = 8 * b0;
    constexpr int kMaxB = 6;
    const int r4 = lane >> 3, i8 = lane & 7;
    const uint32_t lane_addr = tbase + ((uint32_t)(quad * 32) << 16);
    const uint32_t usign = p.first ? 0x80000000u : 0u;          // iteration 0: z_in = 0 and v = +u  (0 - (-u))
    float* sTau = sT;                                           // tau[m] of the sample the current tile belongs to
    int n_tau = -1;
    // Rolling register prefetch: this thread's (up to) 48 code values of a site live in zr[]; as soon as block b of tile i has been
    // stored, block b of tile i+1 is requested into the same registers, so every load has a whole tile time
    // (~4 us) to arrive - DRAM latency never shows up in the epilogue, at no extra register cost.
    auto site = [&](int tile, float*& zs, int& n, int& valid) {
      int qd, qh0, qw0;
      ana_tile_coords(p, tile, n, qd, qh0, qw0);
      const int qh = qh0 + 4 * quad + r4, qw = qw0 + 2 * i8 + (int)rank;
      // bit 0: the row exists (loads and stores happen); bit 1: the site exists (else it is layout padding: stored as 0)
      valid = (qh < g.Qh && !(p.dbg_mode & 2)) ? (1 | (qw < g.Qw ? 2 : 0)) : 0;
      // this thread's 4-subband chunks, 128 B apart; the 8 lanes of a tile row (one group) fill each 128-byte line
      zs = p.z + code_site_offset(((size_t)n * g.Qd + qd) * g.Qh + min(qh, g.Qh - 1), g.Qw, qw) + 2 * b0 * kCodeChunk;
    };
    // ... and ahead of the register loads one bulk L2 prefetch per tile row (the two groups of a 16-site block are
    // 2 x 5632 B contiguous) turns the DRAM reads into large sequential bursts
    auto prefetch_rows = [&](int tile) {
      if (p.first || part != 0 || i8 != 0 || (r4 & 1) != (int)rank || tile >= p.ntiles) return;
      int n2, qd2, qh02, qw02;
      ana_tile_coords(p, tile, n2, qd2, qh02, qw02);
      const int qh2 = qh02 + 4 * quad + r4;
      if (qh2 >= g.Qh) return;
      bulk_prefetch_l2(p.z + code_site_offset(((size_t)n2 * g.Qd + qd2) * g.Qh + qh2, g.Qw, qw02), 2u * kCodeGroup * 4);   // both parities' groups
    };
    float zr[8 * kMaxB];
    float* zs; int n, valid;
#ifndef CDL_ANA_PF
#define CDL_ANA_PF 2          // bulk L2 prefetch distance in tiles (3+ thrashes L2: measured 4.3 -> 5.1 ms)
#endif
#pragma unroll
    for (int a = 1; a < CDL_ANA_PF; ++a) prefetch_rows(pair + a * npairs);
    const uint64_t pol = l2_policy_evict_first();                   // code lines are dead after use: evict-first loads and stores
    auto load_block = [&](const float* base, int b, int ok) {      // 8 subbands = two chunks of 4, pre-biased words
      ldg128_pred_hint(base + 2 * kCodeChunk * b, *reinterpret_cast<float(*)[4]>(&zr[8 * b]), ok, pol);
      ldg128_pred_hint(base + 2 * kCodeChunk * b + kCodeChunk, *reinterpret_cast<float(*)[4]>(&zr[8 * b + 4]), ok, pol);
    };
    if (pair < p.ntiles) {
      site(pair, zs, n, valid);
#pragma unroll
      for (int b = 0; b < kMaxB; ++b) load_block(zs, b, (valid & 1) && !p.first && b < nb);
    }
    int it = 0;
    for (int tile = pair; tile < p.ntiles; tile += npairs, ++it) {
      const uint32_t ds = it & 1;
      if (n != n_tau) {                                         // uniform over the 8 epilogue warps (same tile sequence)
        named_bar_sync(3, 32 * kAEpiWarps);
        const float cval = p.cvec ? p.cvec[n] : 0.0f;
        for (int i = tid; i < kNA; i += 32 * kAEpiWarps) sTau[i] = (i < g.M) ? make_tau(p.t0[i], p.t1[i], cval) : 0.0f;
        named_bar_sync(3, 32 * kAEpiWarps);
        n_tau = n;
      }
      prefetch_rows(tile + CDL_ANA_PF * npairs);
      float* zs2 = zs; int n2 = n, valid2 = 0;
      if (tile + npairs < p.ntiles) site(tile + npairs, zs2, n2, valid2);
      const int ld2 = (valid2 & 1) && !p.first;
      const int st_ok = valid & 1;
      const uint32_t site_mask = (valid & 2) ? 0xffffffffu : 0u;   // layout padding beyond Qw holds (pre-biased) zeros
      CDL_TW(tw0, mbar_wait(&dfull[ds], (it >> 1) & 1));
      tc_fence_after();
      const uint32_t dcol = lane_addr + kColD + ds * kNA + m0;
      uint32_t ua[8], ub[8];
      tmem_ld8(dcol, ua);
#pragma unroll
      for (int b = 0; b < kMaxB; ++b) {
        if (b < nb) {                                          // warp-uniform
        uint32_t (&u)[8] = (b & 1) ? ub : ua;
        uint32_t (&un)[8] = (b & 1) ? ua : ub;
        tmem_wait_ld();
        if (b + 1 < nb) tmem_ld8(dcol + 8 * (b + 1), un);                  // next 8 accumulator columns in flight
        const float4 t0 = *reinterpret_cast<const float4*>(sTau + m0 + 8 * b);
        const float4 t1 = *reinterpret_cast<const float4*>(sTau + m0 + 8 * b + 4);
        const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float zin = p.first ? 0.0f : code_dec(zr[8 * b + i]);
          const float zo = soft_threshold(__fsub_rn(zin, __uint_as_float(u[i] ^ usign)), tt[i]);
          zr[8 * b + i] = code_enc(__uint_as_float(__float_as_uint(zo) & site_mask));
        }
        stg128_pred_hint(zs + 2 * kCodeChunk * b, *reinterpret_cast<const float(*)[4]>(&zr[8 * b]), st_ok, pol);
        stg128_pred_hint(zs + 2 * kCodeChunk * b + kCodeChunk, *reinterpret_cast<const float(*)[4]>(&zr[8 * b + 4]), st_ok, pol);
        load_block(zs2, b, ld2);                                 // same registers: tile i+1, block b
        }
      }
      tc_fence_before();                           // accumulator fully read: hand the TMEM slot back to the MMA warp
      __syncwarp();
      if (lane == 0) { if (rank == 0) mbar_arrive(&dempty[ds]); else mbar_arrive_cluster(&dempty[ds], 0); }
      zs = zs2; n = n2; valid = valid2;
    }
  } else {
    // ============================== MMA issue (leader CTA) ==============================
    // The whole warp runs this loop converged and one elected lane issues (mma_tf32_ss_warp): the issue cost per MMA
    // stays far below the tensor pipe's 88 cycles, so the warp runs ahead of the pipe (its queue holds ~18 MMAs) and
    // the commits / barrier waits cost nothing.
    if (rank == 1 && lane == 0) { mbar_wait(wbar, 0); mbar_arrive_cluster(wready, 0); }
    if (rank == 0) {
      CDL_TW(tw2, mbar_wait(wbar, 0); mbar_wait_cluster(wready, 0));
      const uint32_t idesc = make_idesc_tf32(256, kNA);
      const uint64_t bdesc0 = make_smem_desc_kmajor_noswz(smem_u32(sB), 128, 256);
      // A: K-major, no swizzle, 8-site core matrices; second K half 16 B further on (LBO), next row group 144 B on (SBO)
      const uint64_t adesc0 = make_smem_desc_kmajor_noswz(smem_u32(sR), 16, kARW * 4);
      constexpr uint32_t kBStep = (kNAH * 32) >> 4;                 // 16-byte units between consecutive k-steps of B
      int it = 0;
      for (int tile = pair; tile < p.ntiles;) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {                               // unrolled over the two buffers: constant operands
          if (tile < p.ntiles) {
            CDL_TW(tw0, mbar_wait_cluster(&dempty[u], ((it >> 1) & 1) ^ 1));
            CDL_TW(tw1, mbar_wait_cluster(&rboth[u], (it >> 1) & 1));
            tc_fence_after();
            const uint32_t dcol = tbase + kColD + u * kNA;
#pragma unroll
            for (int ks = 0; ks < kKSteps; ++ks) {
              const int td = ks / kP, th = ks % kP;
              const uint32_t aoff = (uint32_t)(u * kABuf + (th & 1) * kABoxPitch + (td * kARows + (th >> 1)) * kARW) * 4;
              if (!(p.dbg_mode & 32)) mma_tf32_ss_warp<2>(dcol, adesc0 + (uint64_t)(aoff >> 4), bdesc0 + (uint64_t)ks * kBStep, idesc, ks != 0);
            }
            mma_commit_warp<2>(&rempty[u]);         // r buffer reusable once these MMAs have read it (both CTAs)
            mma_commit_warp<2>(&dfull[u]);          // accumulator complete -> epilogue (both CTAs)
            tile += npairs; ++it;
          }
        }
      }
    }
    __syncwarp();                                   // reconverge the MMA warp before the aligned cluster barrier
  }
  if (p.dbg && lane == 0) {
    long long* d = p.dbg + ((size_t)blockIdx.x * 24 + warp) * 8;
    d[0] = clock64() - tstart; d[1] = tw0; d[2] = tw1; d[3] = tw2;
  }
  // teardown: everyone done (all MMAs were consumed by the epilogues before they exit)
  tc_fence_before();
  cluster_sync_all();
  if (warp == kAMmaWarp) tmem_dealloc<2>(tbase, 512);
}

}  // namespace tc
}  // namespace cdl
