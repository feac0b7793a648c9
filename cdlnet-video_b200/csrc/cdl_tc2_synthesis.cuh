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
//   * col2im (4 warps = the 4 tile rows): the (c,th) rows are walked in lock step - at step t warp r handles
//     th = (t + r) mod 7, which lands on footprint row r + th = t + 2r (- 7 when wrapped): distinct for the four
//     warps (tests/test_tc2_operand_cpu.py), a named barrier between steps - no shared-memory atomics.  Inside a warp
//     the 7 tw values are combined across lanes with rotate-shuffles, so that every footprint column is
//     read-modify-written by exactly one lane (own column L, lanes 0..5 also the spill column 32 + L).
//     The finished 10 x 38 x C footprint is added to `out` with red.global.add (times the mask in JDD mode) and cleared.
//   * JDD mask: by default `out` starts from zero and one image pass (k_mask_residual) forms mask * out - yp after
//     the scatter-add; with CDL_TC2D_MASKPASS=0 `out` starts from -yp and the flush multiplies by the mask (slower:
//     the mask loads sit on the flush's critical path, 10.0 vs 5.5 ms per launch on config 3).
//   * Measured (B200): config 4 (64 x 3 x 512^2, M = 64) 2.67 ms per launch, config 3 (32 x 3 x 1024^2, mask) 5.46 ms
//     (fp32 CUDA-core kernel: 15.2 / 30.4 ms); bit-exact against it on exactly representable data.  The col2im warps
//     (tcgen05.ld + 6 shuffles + lock-step barrier per (c,th) row), not the tensor pipe or HBM, set the pace.
//   * Only the RESIDUAL synthesis runs here; the final dictionary synthesis D z (its rounding would land directly on
//     xhat) stays on the exact fp32 CUDA-core kernel.
//
// Warp roles (416 threads): warps 0-7 producers (two per TMEM lane quadrant, half of the subbands each), warps 8-11
// col2im + flush, warp 12 MMA issue + TMEM alloc.
#pragma once
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
  const float* wpack;   // this layer: [Kg/8 k-steps][22 groups][2][8][4] tf32-rounded filters
  int tiles_w, tiles_h, ntiles;
};

__host__ __device__ inline uint32_t syn_b_bytes(int Kg) { return (uint32_t)(Kg / 8) * kSN * 32u; }
__host__ __device__ inline uint32_t syn_smem_bytes(int Kg) {
  return align128(syn_b_bytes(Kg)) + (uint32_t)(kMaxC * kFY * kFPitch * 4) + 128u;
}

// filters (M,C,7,7) [ConvTranspose2d weight (in = M, out = C, th, tw)] -> B[n = (c*7 + th)*8 + tw, k = m], UMMA K-major
__global__ void k_pack_tc2_synthesis(const float* __restrict__ w, float* __restrict__ out, int M, int C, int Kg) {
  const int total = (Kg / 8) * kSN * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i % 4, r8 = (i / 4) % 8, kc = (i / 32) % 2, grp = (i / 64) % (kSN / 8), ks = i / (kSN * 8);
    const int n = grp * 8 + r8, m = ks * 8 + kc * 4 + e;
    const int row = n >> 3, tw = n & 7, c = row / kP, th = row % kP;
    const float v = (row < kP * C && tw < kP && m < M) ? w[(((size_t)m * C + c) * kP + th) * kP + tw] : 0.0f;
    out[i] = ptx::to_tf32_rna(v);
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

__global__ void __launch_bounds__(kSThreads, 1) k_tc2_synthesis(const Syn2Params p) {
  using namespace ptx;
  using tc2::mbar_wait;                        // bounded wait (traps instead of hanging)
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t bbytes = syn_b_bytes(p.Kg);
  float* sB = reinterpret_cast<float*>(smem_raw);
  float* sX = reinterpret_cast<float*>(smem_raw + align128(bbytes));                  // [C][10][40] footprint
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + align128(bbytes) + kMaxC * kFY * kFPitch * 4);
  uint64_t* wbar = bars + 0;
  uint64_t* afull = bars + 1;                  // [2] producers -> MMA
  uint64_t* aempty = afull + 2;                // [2] MMA commit -> producers
  uint64_t* dfull = aempty + 2;                // [2] MMA commit -> col2im
  uint64_t* dempty = dfull + 2;                // [2] col2im -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stride = gridDim.x;
  const int ksteps = p.Kg >> 3;

  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&afull[i], kSProdWarps); mbar_init(&aempty[i], 1);
      mbar_init(&dfull[i], 1); mbar_init(&dempty[i], kSColWarps);
    }
    fence_mbar_init();
  }
  if (warp == kSMmaWarp) { tmem_alloc<1>(tmem_slot, 512); tmem_relinquish<1>(); }
  for (int i = tid; i < kMaxC * kFY * kFPitch; i += kSThreads) sX[i] = 0.0f;
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
    // ============================== producers: code tile -> tf32 -> TMEM A ring ==============================
    // warp = quad + 4*half: TMEM lanes of tile row `quad`; subbands [half*Kg/2, (half+1)*Kg/2), 8 per tcgen05.st
    const int quad = warp & 3, half = warp >> 2;
    const int nh = p.Kg >> 1;                                   // 8, 16, 24 or 32 subbands per thread
    const int m0 = half * nh;
    const uint32_t lane_addr = tbase + ((uint32_t)(quad * 32) << 16);
    const size_t plane = (size_t)p.H * p.W;
    float rg[32];
    auto load_tile = [&](int tile) {
      int n = 0, h0 = 0, w0 = 0, ok = 0;
      if (tile < p.ntiles) {
        syn_tile_coords(p, tile, n, h0, w0);
        ok = (h0 + quad < p.H) && (w0 + lane < p.W);
      }
      const char* base = reinterpret_cast<const char*>(p.z + (((size_t)n * p.M + m0) * p.H + (h0 + quad)) * p.W + (w0 + lane));
#pragma unroll
      for (int j = 0; j < 32; ++j)
        rg[j] = ldg_f32_pred(base + (size_t)j * plane * 4, ok && j < nh && m0 + j < p.M);
    };
    load_tile(blockIdx.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += stride, ++it) {
      const int b = it & 1, u = it >> 1;
      uint32_t rt[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) rt[j] = tf32_rna_bits(rg[j]);
      load_tile(tile + stride);                                 // the next tile's values are in flight during the hand-off
      mbar_wait(&aempty[b], (u & 1) ^ 1);                      // the MMAs of tile it-2 have read this slot
      tc_fence_after();
      const uint32_t acol = lane_addr + (uint32_t)(kSColA + b * kNMax + m0);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (8 * q < nh) tmem_st8(acol + 8 * q, *reinterpret_cast<const uint32_t(*)[8]>(&rt[8 * q]));   // warp-uniform
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull[b]);
    }
  } else if (warp < kSMmaWarp) {
    // ============================== col2im + flush ==============================
    const int r = warp - kSProdWarps;                           // tile row = TMEM lane quadrant (warp & 3 == r)
    const int ct = tid - 32 * kSProdWarps;                      // 0..127
    const uint32_t lane_addr = tbase + ((uint32_t)(r * 32) << 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += stride, ++it) {
      const int b = it & 1, u = it >> 1;
      int n, h0, w0;
      syn_tile_coords(p, tile, n, h0, w0);
      mbar_wait(&dfull[b], u & 1);
      tc_fence_after();
      const uint32_t dcol = lane_addr + (uint32_t)(kSColD + b * kSN);
      for (int c = 0; c < p.C; ++c) {
#pragma unroll
        for (int t = 0; t < kP; ++t) {
          const int th = (t + r) % kP;
          uint32_t v[8];
          tmem_ld8(dcol + (uint32_t)((c * kP + th) * 8), v);
          tmem_wait_ld();
          if (c == p.C - 1 && t == kP - 1) {     // accumulator fully read: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dempty[b]);
          }
          named_bar_sync(2, 32 * kSColWarps);    // lock step: the four warps are on four different footprint rows
          // Column x of the footprint row collects tap tw of lane x - tw.  A lane must not read-modify-write columns its
          // neighbours also touch (no ordering between lanes), so the 7 taps are first combined across lanes with
          // rotate-shuffles: lane L receives tap tw of lane (L - tw) mod 32 - for L >= tw that is a contribution to its
          // own column L, for L < tw it comes from lane 32 + L - tw and belongs to the spill column 32 + L.
          float own = __uint_as_float(v[0]), spill = 0.0f;
#pragma unroll
          for (int tw = 1; tw < kP; ++tw) {
            const float w = __shfl_sync(0xffffffffu, __uint_as_float(v[tw]), (lane - tw) & 31);
            if (lane >= tw) own += w; else spill += w;
          }
          float* row = sX + (c * kFY + r + th) * kFPitch;
          row[lane] += own;                                   // every column is touched by exactly one lane
          if (lane < kP - 1) row[32 + lane] += spill;
        }
      }
      named_bar_sync(2, 32 * kSColWarps);        // footprint complete
      const int nf = p.C * kFY * kFX;
      for (int i = ct; i < nf; i += 32 * kSColWarps) {
        const int x = i % kFX, y = (i / kFX) % kFY, c = i / (kFX * kFY);
        float* cell = sX + (c * kFY + y) * kFPitch + x;
        float v = *cell;
        *cell = 0.0f;
        const int gh = h0 - kP / 2 + y, gw = w0 - kP / 2 + x;
        if (gh >= 0 && gh < p.H && gw >= 0 && gw < p.W) {
          const size_t o = (((size_t)n * p.C + c) * p.H + gh) * p.W + gw;
          if (p.mask) v *= __ldg(p.mask + o);
          red_add_f32(p.out + o, v);
        }
      }
      // (the next tile's first accumulation step starts with the same named barrier: the cleared footprint is visible)
    }
  } else {
    // ============================== MMA issue: whole warp converged, one elected lane issues ==============================
    mbar_wait(wbar, 0);
    const uint32_t idesc = make_idesc_tf32(128, kSN);
    const uint64_t bdesc0 = make_smem_desc_kmajor_noswz(smem_u32(sB), 128, 256);
    constexpr uint32_t kBStep = (kSN * 32) >> 4;                 // 16-byte units between consecutive k-steps of B
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
      mma_commit_warp<1>(&aempty[b]);            // A slot reusable once these MMAs have read it
      mma_commit_warp<1>(&dfull[b]);             // accumulator complete -> col2im
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSMmaWarp) tmem_dealloc<1>(tbase, 512);
}

}  // namespace tc2
}  // namespace cdl
