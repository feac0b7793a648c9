// cdl_tc_synthesis_h.cuh — tcgen05 synthesis step for the video network, "tap half per CTA" form:
//
//     out += B_k z        (reference model/net.py:205,210; nn.ConvTranspose3d, stride 2, output_padding 1)
//
// Same GEMM + col2im formulation, code layout, TMA box, lane permutation and footprint ring as cdl_tc_synthesis.cuh,
// but the work of a CTA pair is split along N instead of along M:
//   * the 49 (td,th) rows of 7 taps are listed td-major (= the filter's own memory order); CTA `half` of a pair owns rows
//     [25*half, 25*half + 25 - half) = N = 176 accumulator columns, for EVERY tile of the pair's tile range.  Its filter
//     half (176 x 176 x 4 B = 124 KB) stays resident in shared memory for the whole launch.
//   * plain cta_group::1 MMAs (M = 128 sites, N = 176, K = 8 per instruction; 22 per tile, 1936 tensor-pipe cycles): the
//     per-SM tensor throughput is the same as in the cta_group::2 form, but a tile's accumulator is 176 TMEM columns
//     instead of 352, so TWO fit: the MMAs of tile i+1 run while the 16 col2im warps drain tile i.  (Measured on the
//     paired form, profiles/r02o_ncu_kernels.md: MMA and drain phases alternate, 10.2 k cycles per tile pair against
//     3.9 k cycles of tensor-pipe work; the MMA warp waits 46 % of the time for the accumulator.)
//   * the footprint ring of a CTA holds only the 4 fine frames its rows reach (td 0..3 / td 3..6): 29.6 KB instead of
//     51.8 KB, which pays for a 4-slot A ring (64 KB in flight).
//   * both CTAs of a pair read the same code tiles (the second read is an L2 hit); there is no cluster, no cross-CTA
//     barrier and no cta_group::2 instruction in this kernel.
#pragma once
#include "cdl_common.cuh"
#include "cdl_tc_ptx.cuh"
#include "cdl_tc_analysis.cuh"
#include "cdl_tc_synthesis.cuh"

namespace cdl {
namespace tc {
namespace h {

constexpr int kRows0 = 25;                // rows of half 0 (td-major list: td 0..2 complete + td = 3, th 0..3); half 1: 24
#ifndef CDL_SYNH_KS
#define CDL_SYNH_KS 4                     // K-steps per A ring slot (4 KB each)
#endif
#ifndef CDL_SYNH_SLOTS
#define CDL_SYNH_SLOTS 4                  // A ring depth (power of two)
#endif
constexpr int kChKS = CDL_SYNH_KS;
constexpr int kChK4 = 2 * kChKS;          // 4-subband chunks per slot
constexpr int kChFloats = kSGroups * kChK4 * kCodeChunk;
constexpr int kChunks = (kKBSteps + kChKS - 1) / kChKS;      // chunks per tile (the last one may run past the 176 subbands: zero-filled)
constexpr int kSlots = CDL_SYNH_SLOTS;
static_assert((kSlots & (kSlots - 1)) == 0, "ring depth must be a power of two");
static_assert((2 * kChunks) % kSlots == 0, "two tiles must use every slot equally often (compile-time slots in the MMA loop)");
constexpr int kUses2 = 2 * kChunks / kSlots;                  // uses of every slot per two-tile round
constexpr int kSlotShift = kSlots == 1 ? 0 : (kSlots == 2 ? 1 : (kSlots == 4 ? 2 : 3));
constexpr int kPl = 4;                    // fine frames per CTA footprint: td 0..3 (half 0) / td 3..6 (half 1)
constexpr int kXTileH = kXRing * kPl * kXW;
constexpr int kXSpillH = kXRing * kPl * 4 * 8;
constexpr int kThreadsH = 32 * 18;        // 16 col2im warps + MMA warp + TMA warp

constexpr size_t kSmemB = (size_t)kKBSteps * kNBP * 8 * sizeof(float);                      // 123904
constexpr size_t kSmemX = ((size_t)kXTileH * sizeof(float) + 127) / 128 * 128;              // 29568
constexpr size_t kSmemA = (size_t)kSlots * kChFloats * sizeof(float);                       // 65536
constexpr size_t kSmemS = (size_t)kXSpillH * sizeof(float);                                 // 3584
constexpr size_t kSmemBytesH = kSmemB + kSmemX + kSmemA + kSmemS + 512;

// filters (M,1,7,7,7) [index (m, td, th, tw)] -> B[half][n = 7*local_row + tw, k = m] in the UMMA K-major no-swizzle
// layout [22 k-steps][22 column groups][2][8][4], tf32 RNE.  Row list td-major: the column j of half h is filter element
// 175*h + j of subband m.  Unused columns are zero.
__global__ void k_pack_tc_synthesis_h(const float* __restrict__ w, float* __restrict__ out, int M, int lo) {
  const int per_half = kKBSteps * kNBP * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per_half; i += gridDim.x * blockDim.x) {
    const int half = i / per_half;
    int rem = i % per_half;
    const int ks = rem / (kNBP * 8);
    rem %= kNBP * 8;
    const int grp = rem / 64, kc = (rem / 32) % 2, r8 = (rem / 4) % 8, e = rem % 4;
    const int j = grp * 8 + r8;                                  // accumulator column
    const int m = ks * 8 + kc * 4 + e;
    float v = 0.0f;
    if (j < 7 * (half == 0 ? kRows0 : 49 - kRows0) && m < M) v = w[(size_t)m * kTaps + 7 * kRows0 * half + j];
    const float hi = ptx::to_tf32_rna(v);
    out[i] = lo ? ptx::to_tf32_rna(v - hi) : hi;
  }
}

// col2im of NR consecutive rows of the td-major list (see c2i_rows in cdl_tc_synthesis.cuh for the lane algebra)
template <int ROW0, int NR>
__device__ __forceinline__ void c2i_rows_h(const uint32_t (&u)[7 * NR], float* xs, float* ss, int pbase, int colbase, int q, const C2iLane& L) {
  constexpr int HALF = ROW0 >= kRows0 ? 1 : 0;
  const unsigned full = 0xffffffffu;
  float x0[NR], x1[NR], e3[NR];
  int cell[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int row = ROW0 + i, td = row / 7, th = row % 7;
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
    cell[i] = ps * kPl + (td - 3 * HALF);
  }
  float2 cur[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) cur[i] = *reinterpret_cast<const float2*>(xs + cell[i] * kXW + colbase);
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    cur[i].x += x0[i]; cur[i].y += x1[i];
    *reinterpret_cast<float2*>(xs + cell[i] * kXW + colbase) = cur[i];
  }
  if (L.o == 0 || L.o == 31) {                                     // the seam columns between lane quadrants
    const int off = L.o ? 4 : 0;
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

// rows [ROW0, ROW0 + NROWS) of the td-major list; `acol` = TMEM address of row ROW0's tap 0.  Four rows per 32-column
// tcgen05.ld, the next load in flight while the current rows are applied; `released` runs once the last load has landed.
template <int ROW0, int NROWS, int DONE = 0, typename Rel>
__device__ __forceinline__ void c2i_part_h(uint32_t acol, uint32_t (&u)[32], float* xs, float* ss, int pbase, int colbase, int q, const C2iLane& L,
                                           Rel released) {
  using namespace ptx;
  if constexpr (DONE == 0) tmem_ld32(acol, u);
  constexpr int NR = (NROWS - DONE) < 4 ? (NROWS - DONE) : 4;
  tmem_wait_ld();
  uint32_t v[7 * NR];
#pragma unroll
  for (int i = 0; i < 7 * NR; ++i) v[i] = u[i];
  if constexpr (DONE + NR < NROWS) tmem_ld32(acol + 7 * (DONE + NR), u);
  else released();
  c2i_rows_h<ROW0 + DONE, NR>(v, xs, ss, pbase, colbase, q, L);
  if constexpr (DONE + NR < NROWS) c2i_part_h<ROW0, NROWS, DONE + NR>(acol, u, xs, ss, pbase, colbase, q, L, released);
}

// tile sequence of a PAIR: the same numbering as syn_tile (columns (n, qd, w-tile), qh fastest), range [t0, t1)
template <bool LO>
__global__ void __launch_bounds__(kThreadsH, 1) k_tc_synthesis_h(const SynTcParams p, const __grid_constant__ CUtensorMap zmap) {
  using namespace ptx;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sB = reinterpret_cast<float*>(smem_raw);
  float* sX = reinterpret_cast<float*>(smem_raw + kSmemB);
  float* sA = reinterpret_cast<float*>(smem_raw + kSmemB + kSmemX);
  float* sS = reinterpret_cast<float*>(smem_raw + kSmemB + kSmemX + kSmemA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kSmemB + kSmemX + kSmemA + kSmemS);
  uint64_t* wbar = bars + 0;
  uint64_t* afull = bars + 1;                  // [4]  TMA: chunk landed (-> MMA; LO: -> col2im warps, which transform it in place)
  uint64_t* aboth = afull + kSlots;            // [4]  LO only: chunk transformed -> MMA
  uint64_t* aempty = aboth + kSlots;           // [4]  MMA commit -> TMA warp
  uint64_t* dfull = aempty + kSlots;           // [2]  MMA commit -> col2im
  uint64_t* dempty = dfull + 2;                // [2]  col2im warps -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

  const Geo& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.x & 1, pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  long long tw0 = 0, tw1 = 0, tw2 = 0;
  const long long tstart = clock64();

  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int i = 0; i < kSlots; ++i) { mbar_init(&afull[i], 1); mbar_init(&aboth[i], kSynC2iWarps); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dempty[i], kSynC2iWarps); }
    fence_mbar_init();
  }
  if (warp == kSynMmaWarp) { tmem_alloc<1>(tmem_slot, 512); tmem_relinquish<1>(); }
  for (int i = tid; i < kXTileH; i += kThreadsH) sX[i] = 0.0f;
  for (int i = tid; i < kXSpillH; i += kThreadsH) sS[i] = 0.0f;
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(wbar, (uint32_t)kSmemB);
    const char* src = reinterpret_cast<const char*>(p.wpack) + (size_t)half * kSmemB;
    const uint32_t piece = 30976;   // 123904 / 4
    for (int i = 0; i < 4; ++i) bulk_g2s(reinterpret_cast<char*>(sB) + i * piece, src + i * piece, piece, wbar);
  }
  tc_fence_before();
  __syncthreads();             // barriers initialised, TMEM allocated (filters may still be in flight)
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  long long t0, t1;
  syn_range(p, pair, npairs, t0, t1);
  const int ntl = (int)(t1 - t0);

  if (warp < kSynC2iWarps) {
    // ============================== col2im + footprint flush ==============================
    const int q = warp & 3, part = warp >> 2;
    const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16);
    C2iLane L;
    {
      const int gq = lane >> 3, i8 = lane & 7;
      L.o = 16 * (gq >> 1) + 2 * i8 + (gq & 1);
      auto lane_of = [](int o) { o &= 31; return 8 * (2 * (o >> 4) + (o & 1)) + ((o & 15) >> 1); };
      L.src1 = lane_of(L.o + 1); L.src2 = lane_of(L.o + 2); L.srcm = lane_of(L.o + 31);
      L.m1 = L.o < 31 ? 1.0f : 0.0f; L.m2 = L.o < 30 ? 1.0f : 0.0f; L.mm = L.o > 0 ? 1.0f : 0.0f;
    }
    const int colbase = 4 + 64 * q + 2 * L.o;
    const int sel = 4 * half + part;                               // warp-uniform
    for (int it = 0; it < ((p.dbg_mode & 1024) ? 0 : ntl); ++it) {
      const SynTile t = syn_tile(p, t0, t1, it);
      const int b = it & 1;
      if (LO) {
        // low part of the code, in place: word -> z = word - bias, hi = truncate(word) = rna(z), word' = (z - hi) + bias
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          const int gc = kChunks * it + c, slot = gc & (kSlots - 1);
          mbar_wait(&afull[slot], (gc >> kSlotShift) & 1);
          auto lo_word = [](float x) {
            const uint32_t bb = __float_as_uint(x);
            if (bb == 0u) return 0.0f;                                                  // TMA zero fill (no site): stays zero
            const float z = __uint_as_float(bb - kCodeBias), hi = __uint_as_float(bb & 0xffffe000u);
            return __uint_as_float(__float_as_uint(z - hi) + kCodeBias);
          };
#pragma unroll
          for (int hh = 0; hh < kChFloats / 4 / (32 * kSynC2iWarps); ++hh) {            // 512 threads x 16 B per step
            float4* pa = reinterpret_cast<float4*>(sA + slot * kChFloats) + hh * 32 * kSynC2iWarps + tid;
            float4 w = *pa;
            w.x = lo_word(w.x); w.y = lo_word(w.y); w.z = lo_word(w.z); w.w = lo_word(w.w);
            *pa = w;
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&aboth[slot]);
        }
      }
      CDL_TW(tw0, mbar_wait(&dfull[b], (it >> 1) & 1));
      tc_fence_after();
      const int pbase = (2 * t.qh) % kXRing;
      auto release = [&]() {                                       // this warp's columns are in registers: accumulator b is free again
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dempty[b]);
      };
#ifdef CDL_TC_PROFILE
      const long long tc0 = clock64();
#endif
      if (!(p.dbg_mode & 64)) {
        uint32_t u[32];
        const uint32_t d = lane_addr + (uint32_t)(b * kNBP);
        switch (sel) {                                             // warp-uniform: a quarter of the half's rows each
          case 0: c2i_part_h<0, 7>(d, u, sX, sS, pbase, colbase, q, L, release); break;
          case 1: c2i_part_h<7, 6>(d + 7 * 7, u, sX, sS, pbase, colbase, q, L, release); break;
          case 2: c2i_part_h<13, 6>(d + 7 * 13, u, sX, sS, pbase, colbase, q, L, release); break;
          case 3: c2i_part_h<19, 6>(d + 7 * 19, u, sX, sS, pbase, colbase, q, L, release); break;
          case 4: c2i_part_h<25, 6>(d, u, sX, sS, pbase, colbase, q, L, release); break;
          case 5: c2i_part_h<31, 6>(d + 7 * 6, u, sX, sS, pbase, colbase, q, L, release); break;
          case 6: c2i_part_h<37, 6>(d + 7 * 12, u, sX, sS, pbase, colbase, q, L, release); break;
          default: c2i_part_h<43, 6>(d + 7 * 18, u, sX, sS, pbase, colbase, q, L, release); break;
        }
      } else {
        release();
      }
#ifdef CDL_TC_PROFILE
      const long long tc1 = clock64();
      tw1 += tc1 - tc0;
#endif
      named_bar_sync(1, 32 * kSynC2iWarps);                        // every warp's rows are in the ring
      if (!(p.dbg_mode & 128)) {
        // fine rows 2*qh and 2*qh+1 are final (all 7 at the end of a run): out += row, clear the ring slot
        const int nfl = t.last ? kXRing : 2;
        const int items = nfl * kPl * (kXW / 4);
        float* on = p.out + (size_t)t.n * g.fine_vol();
        const int gw0 = 2 * t.qw0 - 4;
        for (int item = tid; item < items; item += 32 * kSynC2iWarps) {
          const int c4 = item % (kXW / 4), rest = item / (kXW / 4);
          const int tl = rest % kPl, pi = rest / kPl;
          const int pr = 2 * t.qh + pi;                            // fine row along the sweep; fine h = pr - oh
          const int gd = 2 * t.qd + tl + 3 * half - g.od, gh = pr - g.oh, gw = gw0 + 4 * c4;
          const int rc = (pr % kXRing) * kPl + tl;
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
    // ============================== MMA issue (converged warp, elected lane) ==============================
    CDL_TW(tw2, mbar_wait(wbar, 0));
    const uint32_t idesc = make_idesc_tf32(128, kNBP);
    const uint64_t bdesc0 = make_smem_desc_kmajor_noswz(smem_u32(sB), 128, 256);
    const uint64_t adesc0 = make_smem_desc_kmajor_noswz(smem_u32(sA), 128, kChK4 * 128);
    constexpr uint32_t kBStep = (kNBP * 32) >> 4;                   // 16-byte units between k-steps of B (176 rows x 32 B)
    const bool no_ring = (p.dbg_mode & 512) != 0, no_hand = (p.dbg_mode & 1024) != 0;   // development aids (results invalid)
    for (int it = 0; it < ntl; it += 2) {                           // two tiles per round: slots, buffers and descriptors are constants
      const int rnd = it >> 1;                                      // 2 * kChunks chunks per round = kUses2 uses of every slot
#pragma unroll
      for (int u2 = 0; u2 < 2; ++u2) {
        if (it + u2 < ntl) {
          if (!no_hand) CDL_TW(tw0, mbar_wait(&dempty[u2], (rnd & 1) ^ 1));   // col2im of tile it + u2 - 2 has drained this accumulator
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            const int lc = kChunks * u2 + c, slot = lc & (kSlots - 1);
            if (!no_ring) CDL_TW(tw1, mbar_wait(LO ? &aboth[slot] : &afull[slot], (rnd * kUses2 + (lc >> kSlotShift)) & 1));
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < kChKS; ++j) {
              const int ks = kChKS * c + j;
              if (ks < kKBSteps) {
                const uint64_t ad = adesc0 + (uint64_t)((slot * kChFloats * 4 + j * 256) >> 4);
                mma_tf32_ss_warp<1>(tbase + u2 * kNBP, ad, bdesc0 + (uint64_t)ks * kBStep, idesc, ks != 0);
              }
            }
            if (!no_ring) mma_commit_warp<1>(&aempty[slot]);        // A slot reusable once these MMAs have read it
          }
          if (!no_hand) mma_commit_warp<1>(&dfull[u2]);             // accumulator complete -> col2im
        }
      }
    }
    if (no_hand) { mma_commit_warp<1>(&dfull[0]); mbar_wait(&dfull[0], 0); }   // drain before the teardown
    __syncwarp();
  } else {
    // ============================== TMA: code tile -> A ring, 16 KB chunks ==============================
    if (lane == 0) {
      tma_prefetch_desc(&zmap);
      const int G = code_groups_per_row(g.Qw);
      for (int it = 0; it < ntl; ++it) {
        const SynTile t = syn_tile(p, t0, t1, it);
        const int g0 = (t.qw0 >> 4) * 2;                             // first group of the tile in its row
        if (half == 0) {                                             // the next tile's 16 groups are one contiguous run: one L2 prefetch per pair
          const SynTile t2 = syn_tile(p, t0, t1, it + 1);
          if (t2.valid) {
            const int g2 = (t2.qw0 >> 4) * 2, ng = min(kSGroups, G - g2);
            bulk_prefetch_l2(p.z + ((size_t)t2.row * G + g2) * kCodeGroup, (uint32_t)ng * kCodeGroup * 4);
          }
        }
#pragma unroll 1
        if (p.dbg_mode & 512) continue;
        for (int c = 0; c < kChunks; ++c) {
          const int gc = kChunks * it + c, slot = gc & (kSlots - 1);
          CDL_TW(tw0, mbar_wait(&aempty[slot], ((gc >> kSlotShift) & 1) ^ 1));
          if (p.dbg_mode & 256) { mbar_arrive(&afull[slot]); continue; }
          mbar_expect_tx(&afull[slot], kChFloats * 4);
          tma_load_3d(sA + slot * kChFloats, &zmap, c * kChK4 * kCodeChunk, g0, (int)t.row, &afull[slot]);
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
  __syncthreads();
  if (warp == kSynMmaWarp) tmem_dealloc<1>(tbase, 512);
}

}  // namespace h
}  // namespace tc
}  // namespace cdl
