// cdl_tc2_synthesis.cuh — tcgen05 residual synthesis for the 2-D networks with stride 1 (P = 7x7, s = 1, C <= 3,
// M <= 64: BASELINE configs 1b, 3 and 4):
//
//     out += [mask *] B_k z        (reference model/net.py:87; nn.ConvTranspose2d, stride 1, padding 3; GDLNet :670)
//
// `out` is pre-initialised by the caller with -yp, so that it ends up holding the residual mask * B z - yp.
// Formulation: GEMM + col2im, as in the video kernel (cdl_tc_synthesis.cuh), cta_group::1:
//     Pq[q, (c,th,tw)] = sum_m z[q, m] * W[m, c, th, tw]      M_gemm = 128 sites (4 rows x 32), K = M (<= 64), N = 176
//     out[c, h + th - 3, w + tw - 3] += Pq[q, (c,th,tw)]      (overlap-add of each site's 7x7xC patch)
//
//   * A operand = the code tile, TS form.  z stays in the reference's planar layout (N,M,H,W): a producer thread is
//     one site = one TMEM lane, so for a fixed subband the 32 lanes of a warp read 128 contiguous bytes.  Values are
//     rounded to tf32 (RNE - the tensor core truncates) and tcgen05.st'ed into a 2-slot ring of 64 TMEM columns;
//     the next tile's values are requested before the current ones are stored.
//   * B operand = the filter bank, resident: [K/8 steps][176 rows (c,th,tw8)][8 subbands] = 45 KB; column
//     (c*7 + th)*8 + tw of the accumulator, tw = 7 is padding (every (c,th) row is one tcgen05.ld of 8 columns).
//   * Accumulators double buffered: 2 x 176 + 2 x 64 = 480 of 512 TMEM columns.
//   * col2im (4 warps = the 4 tile rows), WRITE-ONCE PRIVATE FOOTPRINTS (round 2; the round-1 kernel walked the (c,th)
//     rows in lock step with a named barrier per step and read-modify-wrote one shared footprint: 2.67 ms per launch on
//     config 4, this one 2.17 ms).  Warp r owns a private buffer priv[r][(c,th)][40]: row (c,th) of it receives exactly
//     one value per column and tile (columns 0..31 from the lanes' own sums, 32..37 from the spill sums of lanes 0..5),
//     so it is WRITTEN, not accumulated: no read-modify-write, no clearing, no ordering between warps.  Inside a warp
//     the 7 tw values are combined across lanes with rotate-shuffles; the accumulator is drained in 32-column
//     tcgen05.ld's (four (c,th) rows each) with the next load in flight.
//   * The overlap-add moves into the flush: out[c, h0-3+y, w0-3+x] += sum_{r = max(0,y-6)}^{min(3,y)} priv[r][c, y-r][x]
//     (at most 4 terms, times the mask in JDD flush mode), executed by the eight PRODUCER warps one tile behind on a
//     double-buffered priv (2 x 13.4 KB), with red.global.add.
//   * JDD mask: by default `out` starts from zero and one image pass (k_mask_residual) forms mask * out - yp after
//     the scatter-add; with CDL_TC2D_MASKPASS=0 `out` starts from -yp and the flush multiplies by the mask.
//   * Only the RESIDUAL synthesis runs here; the final dictionary synthesis D z (its rounding would land directly on
//     xhat) stays on the exact fp32 CUDA-core kernel.
//
// Warp roles (416 threads): warps 0-7 producers + flush (two per TMEM lane quadrant, half of the subbands each),
// warps 8-11 col2im, warp 12 MMA issue + TMEM alloc.
#pragma once
#include <type_traits>
#include "cdl_common.cuh"
#include "cdl_tc_ptx.cuh"
#include "cdl_tc2_analysis.cuh"

namespace cdl {
namespace tc2 {

constexpr int kSTH = 4, kSTW = 32;          // CTA tile: rows x sites = 128 TMEM lanes
constexpr int kSN = 176;                    // GEMM N: 21 (c,th) rows x 8 (7 taps + pad), padded to 22 rows
constexpr int kSColD = 0, kSColA = 2 * kSN; // TMEM: D0 | D1 | A0 | A1
constexpr int kFY = kSTH + kP - 1;          // 10 footprint rows   (row 0 <-> image row h0 - 3)
constexpr int kFX = kSTW + kP - 1;          // 38 footprint columns (col 0 <-> image column w0 - 3)
constexpr int kFPitch = 40;
constexpr int kSProdWarps = 8, kSColWarps = 4;
constexpr int kSMmaWarp = kSProdWarps + kSColWarps;
constexpr int kSThreads = 32 * (kSProdWarps + kSColWarps + 1);   // 416

struct Syn2Params {
  int N, C, M, H, W;
  int Kg;               // GEMM K: M rounded up to a multiple of 16 (<= kNMax)
  const float* z;       // (N, M, H, W)
  float* out;           // (N, C, H, W), accumulated into
  const float* mask;    // (N, C, H, W) or nullptr: multiplies B z
  int lo_code;          // 1: A = tf32(z - tf32(z)), the low part of the code (second term of the 3-term final D z)
  const float* wpack;   // this layer: [Kg/8 k-steps][22 groups][2][8][4] tf32-rounded filters
  int tiles_w, tiles_h, ntiles;
};

__host__ __device__ inline uint32_t syn_b_bytes(int Kg) { return (uint32_t)(Kg / 8) * kSN * 32u; }

// filters (M,C,7,7) [ConvTranspose2d weight (in = M, out = C, th, tw)] -> B[n = (c*7 + th)*8 + tw, k = m], UMMA K-major
// lo != 0: the part of W that tf32 rounding drops, tf32(W - tf32(W)) (3-term final dictionary synthesis)
__global__ void k_pack_tc2_synthesis(const float* __restrict__ w, float* __restrict__ out, int M, int C, int Kg, int lo = 0) {
  const int total = (Kg / 8) * kSN * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i % 4, r8 = (i / 4) % 8, kc = (i / 32) % 2, grp = (i / 64) % (kSN / 8), ks = i / (kSN * 8);
    const int n = grp * 8 + r8, m = ks * 8 + kc * 4 + e;
    const int row = n >> 3, tw = n & 7, c = row / kP, th = row % kP;
    const float v = (row < kP * C && tw < kP && m < M) ? w[(((size_t)m * C + c) * kP + th) * kP + tw] : 0.0f;
    const float hi = ptx::to_tf32_rna(v);
    out[i] = lo ? ptx::to_tf32_rna(v - hi) : hi;
  }
}

// JDD mode, optional: out <- mask * out - yp as one image pass after the scatter-add (out then starts from zero and the
// footprint flush needs no mask loads)
__global__ void __launch_bounds__(256) k_mask_residual(float* out, const float* __restrict__ mask, const float* __restrict__ yp, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(out)[i];
    const float4 m = __ldg(reinterpret_cast<const float4*>(mask) + i);
    const float4 y = __ldg(reinterpret_cast<const float4*>(yp) + i);
    reinterpret_cast<float4*>(out)[i] = make_float4(__fsub_rn(__fmul_rn(m.x, v.x), y.x), __fsub_rn(__fmul_rn(m.y, v.y), y.y),
                                                    __fsub_rn(__fmul_rn(m.z, v.z), y.z), __fsub_rn(__fmul_rn(m.w, v.w), y.w));
  }
}

__device__ __forceinline__ void syn_tile_coords(const Syn2Params& p, int tile, int& n, int& h0, int& w0) {
  const int tw = tile % p.tiles_w; tile /= p.tiles_w;
  const int th = tile % p.tiles_h;
  n = tile / p.tiles_h;
  h0 = th * kSTH; w0 = tw * kSTW;
}

constexpr int kPrivRows = kMaxC * kP;                       // 21 (c,th) rows per warp
constexpr int kPrivWarp = kPrivRows * kFPitch;              // floats per warp
constexpr int kPrivBuf = kSColWarps * kPrivWarp;            // floats per buffer (4 warps)

__host__ __device__ inline uint32_t syn_smem_bytes(int Kg) {
  return align128(syn_b_bytes(Kg)) + (uint32_t)(2 * kPrivBuf * 4) + 128u;
}

__global__ void __launch_bounds__(kSThreads, 1) k_tc2_synthesis(const Syn2Params p) {
  using namespace ptx;
  using tc2::mbar_wait;                        // bounded wait (traps instead of hanging)
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t bbytes = syn_b_bytes(p.Kg);
  float* sB = reinterpret_cast<float*>(smem_raw);
  float* sP = reinterpret_cast<float*>(smem_raw + align128(bbytes));                  // [2][4 warps][21][40]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + align128(bbytes) + 2 * kPrivBuf * 4);
  uint64_t* wbar = bars + 0;
  uint64_t* afull = bars + 1;                  // [2] producers -> MMA
  uint64_t* aempty = afull + 2;                // [2] MMA commit -> producers
  uint64_t* dfull = aempty + 2;                // [2] MMA commit -> col2im
  uint64_t* dempty = dfull + 2;                // [2] col2im -> MMA
  uint64_t* xfull = dempty + 2;                // [2] col2im -> producers: private footprints of a tile written
  uint64_t* xfree = xfull + 2;                 // [2] producers -> col2im: flushed, buffer reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfree + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stride = gridDim.x;
  const int ksteps = p.Kg >> 3;
  const int nrows = kP * p.C;                  // (c,th) rows in use

  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&afull[i], kSProdWarps); mbar_init(&aempty[i], 1);
      mbar_init(&dfull[i], 1); mbar_init(&dempty[i], kSColWarps);
      mbar_init(&xfull[i], kSColWarps); mbar_init(&xfree[i], kSProdWarps);
    }
    fence_mbar_init();
  }
  if (warp == kSMmaWarp) { tmem_alloc<1>(tmem_slot, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(wbar, bbytes);
    for (uint32_t o = 0; o < bbytes; o += 16384u) {
      const uint32_t n = (bbytes - o < 16384u) ? (bbytes - o) : 16384u;
      bulk_g2s(reinterpret_cast<char*>(sB) + o, reinterpret_cast<const char*>(p.wpack) + o, n, wbar);
    }
  }

  if (warp < kSProdWarps) {
    // ============================== producers: code tile -> tf32 -> TMEM A ring; overlap-add + flush ==============================
    const int quad = warp & 3, half = warp >> 2;
    const int nh = p.Kg >> 1;
    const int m0 = half * nh;
    const uint32_t lane_addr = tbase + ((uint32_t)(quad * 32) << 16);
    const size_t plane = (size_t)p.H * p.W;
    float rg[32];
    // Tile coordinates are carried incrementally (tile += stride): the div / mod decode per use was 10 % of the kernel's
    // instructions in an issue-bound kernel (ncu source page, profiles/r02ab_lines_tc2_synthesis.txt).
    struct TileC { int n, th, tw; };
    int sw, sh, sn;
    { int t = stride; sw = t % p.tiles_w; t /= p.tiles_w; sh = t % p.tiles_h; sn = t / p.tiles_h; }
    auto decode = [&](int tile) { TileC c; c.tw = tile % p.tiles_w; tile /= p.tiles_w; c.th = tile % p.tiles_h; c.n = tile / p.tiles_h; return c; };
    auto advance = [&](TileC c) {
      c.tw += sw; int cy = c.tw >= p.tiles_w; c.tw -= cy ? p.tiles_w : 0;
      c.th += sh + cy; cy = c.th >= p.tiles_h; c.th -= cy ? p.tiles_h : 0;
      c.n += sn + cy;
      return c;
    };
    const int nvalid = max(0, min(nh, p.M - m0));                      // subbands this warp half really has
    auto load_tile = [&](int tile, const TileC& c) {
      const int h0 = c.th * kSTH, w0 = c.tw * kSTW;
      const int ok = (tile < p.ntiles) && (h0 + quad < p.H) && (w0 + lane < p.W);
      const int n = ok ? c.n : 0;
      const char* base = reinterpret_cast<const char*>(p.z + (((size_t)n * p.M + m0) * p.H + (ok ? h0 + quad : 0)) * p.W + (ok ? w0 + lane : 0));
      const size_t pitch = plane * 4;
      if (nvalid == 32 && __all_sync(0xffffffffu, ok)) {                // interior tile, full subband half: unpredicated loads
#pragma unroll
        for (int j = 0; j < 32; ++j) { rg[j] = ldg_f32_stream(base); base += pitch; }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) rg[j] = ldg_f32_pred(base + (size_t)j * pitch, ok && j < nvalid);
      }
    };
    // flush of one tile: out[c, h0-3+y, w0-3+x] += sum_r priv[r][c*7 + (y-r)][x]  over the rows r with 0 <= y-r <= 6.
    // Thread = one footprint column x of one channel c and one half of the 10 footprint rows: the (x, c, half) decode is done
    // once per kernel, the row validity of the <= 4 terms is known at compile time, the output address advances by one image
    // row.  (The element-per-thread form with a div / mod decode per output was 20 % of the kernel's instructions.)
    const int fx = tid % kFPitch, fslot = tid / kFPitch;                // 40 columns (38 used) x 6 slots = 3 channels x 2 halves
    const int fc = fslot >> 1, fhalf = fslot & 1;
    const bool factive = fx < kFX && fc < p.C && fslot < 2 * kMaxC;
    auto flush_rows = [&](auto HALF, const float* pv, int n, int h0, int w0) {
      constexpr int Y0 = decltype(HALF)::value * (kFY / 2);
      const int gw = w0 - kP / 2 + fx;
      if (gw < 0 || gw >= p.W) return;
      const float* pc = pv + fc * kP * kFPitch + fx;
      const int gh0 = h0 - kP / 2 + Y0;                                 // image row of this thread's first footprint row
      const long long o = (((long long)n * p.C + fc) * p.H + gh0) * p.W + gw;      // may point above the image: only dereferenced in range
      float* po = p.out + o;
      const float* pm = p.mask ? p.mask + o : nullptr;
#pragma unroll
      for (int yy = 0; yy < kFY / 2; ++yy) {
        const int y = Y0 + yy;
        float v = 0.0f;
#pragma unroll
        for (int r = 0; r < kSTH; ++r) {
          const int th = y - r;
          if (th >= 0 && th < kP) v += pc[r * kPrivWarp + th * kFPitch];
        }
        const int gh = gh0 + yy;
        if (gh >= 0 && gh < p.H) {
          if (pm) v *= __ldg(pm);
          red_add_f32(po, v);
        }
        po += p.W;
        if (pm) pm += p.W;
      }
    };
    auto flush_tile = [&](int j, const TileC& c) {
      const int xb = j & 1;
      mbar_wait(&xfull[xb], (j >> 1) & 1);
      const float* pv = sP + xb * kPrivBuf;
      if (factive) {
        if (fhalf == 0) flush_rows(std::integral_constant<int, 0>{}, pv, c.n, c.th * kSTH, c.tw * kSTW);
        else flush_rows(std::integral_constant<int, 1>{}, pv, c.n, c.th * kSTH, c.tw * kSTW);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&xfree[xb]);
    };
    TileC ccur = decode(blockIdx.x), cprev = ccur, cnext;
    load_tile(blockIdx.x, ccur);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += stride, ++it) {
      const int b = it & 1, u = it >> 1;
      uint32_t rt[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) rt[j] = tf32_rna_bits(rg[j]);
      if (p.lo_code) {
#pragma unroll
        for (int j = 0; j < 32; ++j) rt[j] = tf32_rna_bits(__fsub_rn(rg[j], __uint_as_float(rt[j])));
      }
      cnext = advance(ccur);
      load_tile(tile + stride, cnext);
      mbar_wait(&aempty[b], (u & 1) ^ 1);
      tc_fence_after();
      const uint32_t acol = lane_addr + (uint32_t)(kSColA + b * kNMax + m0);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (8 * q < nh) tmem_st8(acol + 8 * q, *reinterpret_cast<const uint32_t(*)[8]>(&rt[8 * q]));
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull[b]);
      if (it > 0) flush_tile(it - 1, cprev);     // one tile behind: its col2im ran while this tile's A was produced
      cprev = ccur; ccur = cnext;
    }
    if (it > 0) flush_tile(it - 1, cprev);
  } else if (warp < kSMmaWarp) {
    // ============================== col2im: accumulator -> write-once private footprint ==============================
    const int r = warp - kSProdWarps;
    const uint32_t lane_addr = tbase + ((uint32_t)(r * 32) << 16);
    // one (c,th) row: combine the 7 w-taps across lanes (lane L receives tap tw of lane (L - tw) mod 32: its own column for
    // L >= tw, the spill column 32 + L otherwise) and WRITE columns L and 32 + L of the private row
    // (the lane >= tw test as two 0/1 weights: two FFMAs per tap and no select - exact, the weights are 0 and 1)
    float mo[kP], ms[kP];
#pragma unroll
    for (int tw = 1; tw < kP; ++tw) { mo[tw] = lane >= tw ? 1.0f : 0.0f; ms[tw] = 1.0f - mo[tw]; }
    auto put_row = [&](const uint32_t* v, float* row) {
      float own = __uint_as_float(v[0]), spill = 0.0f;
#pragma unroll
      for (int tw = 1; tw < kP; ++tw) {
        const float w = __shfl_sync(0xffffffffu, __uint_as_float(v[tw]), (lane - tw) & 31);
        own = fmaf(w, mo[tw], own); spill = fmaf(w, ms[tw], spill);
      }
      row[lane] = own;
      if (lane < kP - 1) row[32 + lane] = spill;
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += stride, ++it) {
      const int b = it & 1, u = it >> 1;
      mbar_wait(&xfree[b], (u & 1) ^ 1);         // the flush of tile it-2 has read this private buffer
      mbar_wait(&dfull[b], u & 1);
      tc_fence_after();
      const uint32_t dcol = lane_addr + (uint32_t)(kSColD + b * kSN);
      float* priv = sP + b * kPrivBuf + r * kPrivWarp;
      uint32_t ua[32], ub[32];
      tmem_ld32(dcol, ua);                       // rows 0..3
#pragma unroll
      for (int g = 0; g < 5; ++g) {              // groups of four (c,th) rows = 32 accumulator columns
        uint32_t (&cur)[32] = (g & 1) ? ub : ua;
        uint32_t (&nxt)[32] = (g & 1) ? ua : ub;
        if (4 * g < nrows) {                     // warp-uniform
          tmem_wait_ld();
          if (g < 4) { if (4 * (g + 1) < nrows) tmem_ld32(dcol + 32 * (g + 1), nxt); }
          else if (20 < nrows) tmem_ld8(dcol + 160, *reinterpret_cast<uint32_t(*)[8]>(&nxt[0]));   // row 20
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (4 * g + q < nrows) put_row(&cur[8 * q], priv + (4 * g + q) * kFPitch);
        }
      }
      if (20 < nrows) { tmem_wait_ld(); put_row(&ub[0], priv + 20 * kFPitch); }
      tc_fence_before();                         // accumulator fully read, private footprint written
      __syncwarp();
      if (lane == 0) { mbar_arrive(&dempty[b]); mbar_arrive(&xfull[b]); }
    }
  } else {
    // ============================== MMA issue: whole warp converged, one elected lane issues ==============================
    mbar_wait(wbar, 0);
    const uint32_t idesc = make_idesc_tf32(128, kSN);
    const uint64_t bdesc0 = make_smem_desc_kmajor_noswz(smem_u32(sB), 128, 256);
    constexpr uint32_t kBStep = (kSN * 32) >> 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += stride, ++it) {
      const int b = it & 1, u = it >> 1;
      mbar_wait(&dempty[b], (u & 1) ^ 1);
      mbar_wait(&afull[b], u & 1);
      tc_fence_after();
      const uint32_t dcol = tbase + (uint32_t)(kSColD + b * kSN);
      const uint32_t acol = tbase + (uint32_t)(kSColA + b * kNMax);
      for (int j = 0; j < ksteps; ++j)
        mma_tf32_ts_warp<1>(dcol, acol + 8 * j, bdesc0 + (uint64_t)j * kBStep, idesc, j != 0);
      mma_commit_warp<1>(&aempty[b]);
      mma_commit_warp<1>(&dfull[b]);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSMmaWarp) tmem_dealloc<1>(tbase, 512);
}

}  // namespace tc2
}  // namespace cdl
