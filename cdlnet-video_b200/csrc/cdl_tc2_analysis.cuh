// cdl_tc2_analysis.cuh — tcgen05 analysis step for the 2-D networks with stride 1 (CDLNet / JDD_CDLNet / GDLNet,
// P = 7x7, s = 1, C <= 3, M <= 64: BASELINE configs 1b, 3 and 4):
//
//     z <- ST(z - A_k r, t0 + c*t1)            (reference model/net.py:85,87 + :11-14; GDLNet :668,670)
//
// Default for tf32/auto plans of this geometry (CDL_TC2D=0 at plan creation falls back to the exact fp32 CUDA-core
// kernels of cdl_cc.cuh; fine width must be a multiple of 4).  Bit-exact against the fp32 kernel on exactly
// representable data, 2-6e-5 on xhat against the oracle / fp32 family (tests/test_tc2_gpu.py, profiles/r01k_*);
// the operand construction is also modelled on the CPU (tests/test_tc2_operand_cpu.py).
// Measured (B200, config 3: 32 x 3 x 1024^2, M = 64): 3.38 ms per launch = 5.08 TB/s of code traffic = 78 % of the
// measured HBM copy bandwidth (fp32 CUDA-core kernel: 20.8 ms).
//
// Implicit GEMM  U[q, m] = sum_{c,th,j} R[q, (c,th,j)] * W[m, (c,th,j)],  kind::tf32, fp32 accumulation in TMEM,
// cta_group::1 (the filter bank is 7*C*64*32 B <= 43 KB, no need to split it over a CTA pair):
//   * K = 7*C steps of 8: one (c,th) filter row per step = window element 0 (zero filter column) + the 7 w-taps;
//     N = M rounded up to 16.
//   * The im2col operand is never built.  With stride 1 the 8-float windows of sites w and w+4 start 16 bytes apart,
//     so 8 sites of EQUAL RESIDUE w mod 4 of one image row form a legal K-major core matrix (8 rows x 16 B,
//     contiguous) and the second K half is the same memory 16 B further on: the MMA reads A straight from a row
//     window in shared memory through an overlapping descriptor (LBO = 16 B, SBO = 144 B), exactly the construction
//     of the video kernel (cdl_tc_analysis.cuh) with four residues in place of two parities.
//   * A descriptor start must be 16-byte aligned and a TMA box must start 16-byte aligned in global memory (measured,
//     profiles/r01_tcgen05_mma_cost.log), so the residues 1..3 cannot be addressed in place: two "shifter" warps copy
//     the TMA-staged halo tile (40 floats x 22 rows x C) into FOUR operand copies, copy rho shifted left by rho floats,
//     rounding to tf32 (RNE; the tensor core would truncate) on the way - the rounding pre-pass of the video path
//     costs nothing extra here and r is read from HBM once, unrounded.
//   * CTA tile = 16 rows x 32 sites; residue rho owns accumulator rho: TMEM lane = (row, i), site w = w0 + rho + 4 i.
//     4 accumulators x 64 columns, double buffered = all 512 TMEM columns.  84 MMAs (C = 3) of N = 64 per tile
//     (~95 cycles each in SS form) = 8 k cycles against 262 KB of code traffic per tile (11 k cycles at the SM's share
//     of HBM): HBM-bound by design.
//   * Epilogue: a thread owns FOUR CONSECUTIVE sites (w0 + 4 i + rho, rho = 0..3) of one row, so the reference's own
//     planar layout z (N,M,H,W) is already perfectly coalesced: one 128-bit load and store per subband, 8 lanes = one
//     128-byte line, 4 lines per warp instruction.  No internal code layout, no import/export, and the CUDA-core
//     synthesis keeps working on the same buffer.  Loads run one 8-subband block ahead (across tile boundaries).
//   * Persistent: one CTA per SM, static round-robin over tiles.
//
// Warp roles (384 threads): warps 0-7 epilogue (two per TMEM lane quadrant, half of the subbands each), warp 8 MMA
// issue + TMEM alloc, warp 9 TMA loads, warps 10-11 shift/round the staged tile into the four operand copies.
#pragma once
#include "cdl_common.cuh"
#include "cdl_tc_ptx.cuh"

namespace cdl {
namespace tc2 {

constexpr int kTH = 16, kTW = 32;          // CTA tile: rows x sites
constexpr int kP = 7;
constexpr int kRW = 36;                    // floats per operand row window: 8 sites x 4 + the 4-float tail of the last site (144 B)
constexpr int kRows = kTH + kP - 1;        // 22 image rows per tile (halo 3 + 3)
constexpr int kSW = 40;                    // floats per staged row: kRW + 3 shifts, rounded up to a 16-byte multiple (TMA box)
constexpr int kMaxC = 3;
constexpr int kNMax = 64;                  // TMEM columns per accumulator
constexpr int kEpiWarps = 8, kShiftWarps = 2;
constexpr int kMmaWarp = kEpiWarps, kLoadWarp = kEpiWarps + 1, kShiftWarp0 = kEpiWarps + 2;
constexpr int kThreads = 32 * (kEpiWarps + 2 + kShiftWarps);   // 384

__host__ __device__ constexpr uint32_t align128(uint32_t v) { return (v + 127u) / 128u * 128u; }

// dynamic shared-memory layout (bytes), a function of C and the GEMM N only
struct SmemLayout {
  uint32_t b;            // filters: [7*C k-steps][Ng/8 groups][2 k-chunks][8 rows][4] floats
  uint32_t op;           // operand copies: [2 buffers][4 residues][C*22 rows][36] floats
  uint32_t stage;        // TMA staging: [2 buffers][C*22 rows][40] floats
  uint32_t tau;          // t0 | t1, kNMax floats each
  uint32_t bars;
  uint32_t total;
  uint32_t copy_pitch, buf_pitch, stage_pitch, b_bytes, stage_bytes;
};
__host__ __device__ inline SmemLayout smem_layout(int C, int Ng) {
  SmemLayout L;
  L.b = 0;
  L.b_bytes = (uint32_t)(kP * C * Ng * 32);
  L.copy_pitch = (uint32_t)(C * kRows * kRW * 4);              // multiple of 16 (22*36*4 = 198*16)
  L.buf_pitch = 4 * L.copy_pitch;
  L.op = align128(L.b_bytes);
  L.stage_bytes = (uint32_t)(C * kRows * kSW * 4);
  L.stage_pitch = align128(L.stage_bytes);                     // TMA destination: 128-byte aligned
  L.stage = align128(L.op + 2 * L.buf_pitch);
  L.tau = L.stage + 2 * L.stage_pitch;
  L.bars = L.tau + 2 * kNMax * 4;
  L.total = L.bars + 128;
  return L;
}

struct Ana2Params {
  int N, C, M, H, W;
  int Ng;               // GEMM N: M rounded up to a multiple of 16 (<= kNMax)
  float* z;             // (N, M, H, W), updated in place
  const float* wpack;   // this layer's filters in the layout of SmemLayout::b, tf32-rounded
  const float* t0;      // [M]
  const float* t1;      // [M]
  const float* cvec;    // [N] or nullptr
  int first;
  int tiles_w, tiles_h, ntiles;
};

// filters (M,C,7,7) -> K-major no-swizzle UMMA layout per k-step (c,th); column j: 0 = pad, 1..7 = tw 0..6; tf32 RNE
__global__ void k_pack_tc2_analysis(const float* __restrict__ w, float* __restrict__ out, int M, int C, int Ng) {
  const int total = kP * C * Ng * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i % 4, r8 = (i / 4) % 8, kc = (i / 32) % 2, grp = (i / 64) % (Ng / 8), ks = i / (Ng * 8);
    const int m = grp * 8 + r8, j = kc * 4 + e, c = ks / kP, th = ks % kP;
    const float v = (m < M && j > 0) ? w[(((size_t)m * C + c) * kP + th) * kP + (j - 1)] : 0.0f;
    out[i] = ptx::to_tf32_rna(v);
  }
}

__device__ __forceinline__ void ldg128_pred(const float* p, float (&v)[4], int ok) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t"
               "mov.b32 %0, 0; mov.b32 %1, 0; mov.b32 %2, 0; mov.b32 %3, 0;\n\t"
               "@p ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "r"(ok));
}
__device__ __forceinline__ void stg128_pred(float* p, const float (&v)[4], int ok) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p st.global.v4.f32 [%4], {%0,%1,%2,%3};\n\t}"
               ::"f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "l"(p), "r"(ok) : "memory");
}

// mbarrier wait that cannot hang the device: a hand-off that does not arrive within ~2 s traps (reported by the next
// CUDA call as a launch failure instead of a hung GPU).  The clock is only read every 1024 failed polls.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}

__device__ __forceinline__ void tile_coords(const Ana2Params& p, int tile, int& n, int& h0, int& w0) {
  const int tw = tile % p.tiles_w; tile /= p.tiles_w;
  const int th = tile % p.tiles_h;
  n = tile / p.tiles_h;
  h0 = th * kTH; w0 = tw * kTW;
}

__global__ void __launch_bounds__(kThreads, 1) k_tc2_analysis(const Ana2Params p, const __grid_constant__ CUtensorMap rmap) {
  using namespace ptx;
  using tc2::mbar_wait;                        // the bounded wait above, not ptx::mbar_wait
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const SmemLayout L = smem_layout(p.C, p.Ng);
  float* sB = reinterpret_cast<float*>(smem_raw + L.b);
  uint8_t* sOp = smem_raw + L.op;
  uint8_t* sStage = smem_raw + L.stage;
  float* sT = reinterpret_cast<float*>(smem_raw + L.tau);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
  uint64_t* wbar = bars + 0;                   //      filters landed
  uint64_t* sfull = bars + 1;                  // [2]  TMA: staged tile landed
  uint64_t* sempty = sfull + 2;                // [2]  shifters: staged tile consumed
  uint64_t* ofull = sempty + 2;                // [2]  shifters: operand copies written -> MMA
  uint64_t* oempty = ofull + 2;                // [2]  MMA commit: operand copies read
  uint64_t* dfull = oempty + 2;                // [2]  MMA commit: accumulators complete -> epilogue
  uint64_t* dempty = dfull + 2;                // [2]  epilogue warps: accumulators read -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stride = gridDim.x;

  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sfull[i], 1); mbar_init(&sempty[i], kShiftWarps);
      mbar_init(&ofull[i], kShiftWarps); mbar_init(&oempty[i], 1);
      mbar_init(&dfull[i], 1); mbar_init(&dempty[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc<1>(tmem_slot, 512); tmem_relinquish<1>(); }
  for (int i = tid; i < kNMax; i += kThreads) {
    sT[i] = (i < p.M) ? p.t0[i] : 0.0f;
    sT[kNMax + i] = (i < p.M) ? p.t1[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  if (tid == 0) {                              // the filter bank, in pieces of <= 16 KB
    mbar_expect_tx(wbar, L.b_bytes);
    for (uint32_t o = 0; o < L.b_bytes; o += 16384u) {
      const uint32_t n = (L.b_bytes - o < 16384u) ? (L.b_bytes - o) : 16384u;
      bulk_g2s(reinterpret_cast<char*>(sB) + o, reinterpret_cast<const char*>(p.wpack) + o, n, wbar);
    }
  }

  if (warp == kLoadWarp) {
    // ============================== TMA: halo tile of r, all channels, unrounded ==============================
    if (lane == 0) {
      tma_prefetch_desc(&rmap);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += stride, ++it) {
        const int b = it & 1, u = it >> 1;
        int n, h0, w0;
        tile_coords(p, tile, n, h0, w0);
        mbar_wait(&sempty[b], (u & 1) ^ 1);                           // the shifters have consumed tile it-2
        mbar_expect_tx(&sfull[b], L.stage_bytes);
        // staged column x <-> image column w0 - 4 + x (16-byte aligned start: w0 is a multiple of 32); rows h0-3 ..
        // h0+18; out-of-range elements arrive as zero = the convolution's zero padding
        tma_load_4d(sStage + b * L.stage_pitch, &rmap, w0 - 4, h0 - (kP / 2), 0, n, &sfull[b]);
      }
    }
    __syncwarp();
  } else if (warp >= kShiftWarp0) {
    // ============================== staged tile -> four shifted, tf32-rounded operand copies ==============================
    // copy rho, row y, float k  =  tf32(staged[y][k + rho]),  k < 36:   site i's window starts at float 4 i of copy rho,
    // i.e. at image column w0 - 4 + rho + 4 i = w - 4 for the site w = w0 + rho + 4 i.
    const int st = tid - 32 * kShiftWarp0;
    const int nelem = p.C * kRows * kSW;
    const int cp = (int)(L.copy_pitch >> 2);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += stride, ++it) {
      const int b = it & 1, u = it >> 1;
      mbar_wait(&sfull[b], u & 1);
      mbar_wait(&oempty[b], (u & 1) ^ 1);                             // the MMAs of tile it-2 have read these copies
      tc_fence_after();
      const float* sg = reinterpret_cast<const float*>(sStage + b * L.stage_pitch);
      float* op = reinterpret_cast<float*>(sOp + b * L.buf_pitch);
      for (int idx = st; idx < nelem; idx += 32 * kShiftWarps) {
        const int row = idx / kSW, x = idx - row * kSW;
        if (x < kRW + 3) {
          const float v = __uint_as_float(tf32_rna_bits(sg[idx]));
          float* o = op + row * kRW + x;
#pragma unroll
          for (int rho = 0; rho < 4; ++rho) {
            const int k = x - rho;
            if (k >= 0 && k < kRW) o[rho * cp - rho] = v;
          }
        }
      }
      fence_async_smem();                        // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) { mbar_arrive(&ofull[b]); mbar_arrive(&sempty[b]); }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issue: whole warp converged, one elected lane issues ==============================
    mbar_wait(wbar, 0);
    const uint32_t idesc = make_idesc_tf32(128, p.Ng);
    const uint64_t bdesc0 = make_smem_desc_kmajor_noswz(smem_u32(sB), 128, 256);
    // A: K-major, no swizzle; rows of a core matrix = 8 sites 16 B apart, second K half 16 B further on (LBO), next
    // 8-row group = next image row, 144 B on (SBO)
    const uint64_t adesc0 = make_smem_desc_kmajor_noswz(smem_u32(sOp), 16, kRW * 4);
    const uint32_t bstep = (uint32_t)(p.Ng * 32) >> 4;                 // 16-byte units between consecutive k-steps of B
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += stride, ++it) {
      const int b = it & 1, u = it >> 1;
      mbar_wait(&dempty[b], (u & 1) ^ 1);
      mbar_wait(&ofull[b], u & 1);
      tc_fence_after();
      for (int rho = 0; rho < 4; ++rho) {
        const uint32_t dcol = tbase + (uint32_t)((b * 4 + rho) * kNMax);
        const uint32_t abase = (uint32_t)b * L.buf_pitch + (uint32_t)rho * L.copy_pitch;
        for (int c = 0; c < p.C; ++c) {
#pragma unroll
          for (int th = 0; th < kP; ++th) {
            const int ks = c * kP + th;
            const uint32_t aoff = abase + (uint32_t)((c * kRows + th) * kRW * 4);
            mma_tf32_ss_warp<1>(dcol, adesc0 + (uint64_t)(aoff >> 4), bdesc0 + (uint64_t)ks * bstep, idesc, ks != 0);
          }
        }
      }
      mma_commit_warp<1>(&oempty[b]);            // operand copies reusable once these MMAs have read them
      mma_commit_warp<1>(&dfull[b]);             // accumulators complete -> epilogue
    }
    __syncwarp();
  } else {
    // ============================== epilogue: TMEM -> z update ==============================
    // warp = quad + 4*part: TMEM lanes [32*quad, 32*quad+32) = tile rows 4*quad .. 4*quad+3 x 8 lanes i; this thread's
    // four sites are w0 + 4 i + rho; subbands [part*Ng/2, (part+1)*Ng/2) in blocks of 8.
    const int quad = warp & 3, part = warp >> 2;
    const int nb = p.Ng >> 4;                                          // 8-subband blocks per warp and tile
    const int mpart = part * (p.Ng >> 1);
    const int hrow = 4 * quad + (lane >> 3), i8 = lane & 7;
    const uint32_t lane_addr = tbase + ((uint32_t)(quad * 32) << 16);
    const uint32_t usign = p.first ? 0x80000000u : 0u;                 // iteration 0: z_in = 0 and v = +u  (0 - (-u))
    const size_t plane = (size_t)p.H * p.W;
    const int my_tiles = (p.ntiles - (int)blockIdx.x + stride - 1) / stride;
    const int total = my_tiles * nb;                                   // flattened (tile, block) steps of this warp

    struct Step { float* zp; int valid, m, n, it, blk; };
    auto locate = [&](int s, Step& o) {
      o.it = s / nb; o.blk = s - o.it * nb;
      o.m = mpart + 8 * o.blk;
      o.valid = 0; o.zp = p.z; o.n = 0;
      const int tile = (int)blockIdx.x + o.it * stride;
      if (s >= total || tile >= p.ntiles) return;
      int h0, w0;
      tile_coords(p, tile, o.n, h0, w0);
      const int h = h0 + hrow, w = w0 + 4 * i8;
      o.valid = (h < p.H && w < p.W);                                  // W is a multiple of 4: a float4 is all in or all out
      o.zp = p.z + (((size_t)o.n * p.M + o.m) * p.H + h) * p.W + w;
    };
    auto load = [&](const Step& s, float (&zr)[8][4]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ldg128_pred(s.zp + j * plane, zr[j], s.valid && !p.first && (s.m + j < p.M));
    };
    auto process = [&](const Step& s, float (&zr)[8][4]) {
      const int b = s.it & 1, u = s.it >> 1;
      if (s.blk == 0) { mbar_wait(&dfull[b], u & 1); tc_fence_after(); }          // warp-uniform
      uint32_t acc[4][8];
#pragma unroll
      for (int rho = 0; rho < 4; ++rho) tmem_ld8(lane_addr + (uint32_t)((b * 4 + rho) * kNMax + s.m), acc[rho]);
      tmem_wait_ld();
      if (s.blk == nb - 1) {                     // accumulators fully read: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dempty[b]);
      }
      const float cval = p.cvec ? p.cvec[s.n] : 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float tau = make_tau(sT[s.m + j], sT[kNMax + s.m + j], cval);
        float o[4];
#pragma unroll
        for (int rho = 0; rho < 4; ++rho)
          o[rho] = soft_threshold(__fsub_rn(zr[j][rho], __uint_as_float(acc[rho][j] ^ usign)), tau);
        stg128_pred(s.zp + j * plane, o, s.valid && (s.m + j < p.M));
      }
    };

    float za[8][4], zb[8][4];
    Step sa, sb;
    locate(0, sa);
    load(sa, za);
    for (int s = 0; s < total; s += 2) {         // loads run one block ahead of the accumulator reads, across tiles
      locate(s + 1, sb);
      load(sb, zb);
      process(sa, za);
      if (s + 1 < total) {
        locate(s + 2, sa);
        load(sa, za);
        process(sb, zb);
      }
    }
  }
  // teardown: every MMA has been consumed by the epilogues before they leave their loop
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<1>(tbase, 512);
}

}  // namespace tc2
}  // namespace cdl
