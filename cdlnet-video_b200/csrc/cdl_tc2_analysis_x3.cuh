// cdl_tc2_analysis_x3.cuh — "accurate mode" CANDIDATE of k_tc2_analysis: 3-term (hi/lo) tf32 analysis, OPT-IN with
// CDL_TC2D_ANA=3 at plan creation.  NOT YET RUN ON HARDWARE (written after the round's GPU budget was spent).
//
//     u = A r  ~=  tf32(r) tf32(A) + tf32(r - tf32(r)) tf32(A) + tf32(r) tf32(A - tf32(A))
//
// Why: the arithmetic model (scripts/tc2_emulate_cpu.py) shows that the single-pass tf32 ANALYSIS carries the
// 2-D family's error at the named configurations (config 4: 4.9e-5 as built, 2.6e-5 with a 3-term analysis = the same
// as an exact one), and that the error grows with code density (1.6e-4 predicted on a deliberately hot GDLNet vector).
// The kernel body below IS k_tc2_analysis (generated from it; loader, epilogue, barriers and the overlapping-descriptor
// operand are unchanged) with three differences:
//   * the shifter warps write EIGHT operand copies per tile: copies 0..3 = tf32(r) as before, copies 4..7 = the
//     rounding remainder tf32(r - tf32(r)) - into ONE operand buffer (8 x 9.5 KB; a second one does not fit next to
//     two filter banks), so the shift of tile t+1 waits for the MMAs of tile t;
//   * the filter bank has a second half, tf32(A - tf32(A)) (k_pack_tc2_analysis_x3);
//   * three MMAs per K-step accumulate into the same TMEM columns: 252 MMAs per tile (C = 3) = ~24 k cycles against
//     11 k cycles of code traffic: this mode is tensor-pipe-bound, ~2.2x the time of the default analysis.
//   (The N = 128 variant of DESIGN.md 9 - hi and lo filters side by side in N, 2x instead of 3x - needs half-tile
//    accumulators and a different epilogue; this one keeps the validated epilogue.)
#pragma once
#include "cdl_tc2_analysis.cuh"

namespace cdl {
namespace tc2 {

__host__ __device__ inline SmemLayout smem_layout_x3(int C, int Ng) {
  SmemLayout L;
  L.b = 0;
  L.b_bytes = 2u * (uint32_t)(kP * C * Ng * 32);                // hi bank | lo bank
  L.copy_pitch = (uint32_t)(C * kRows * kRW * 4);
  L.buf_pitch = 8 * L.copy_pitch;                              // hi copies 0..3 | lo copies 4..7, ONE buffer
  L.op = align128(L.b_bytes);
  L.stage_bytes = (uint32_t)(C * kRows * kSW * 4);
  L.stage_pitch = align128(L.stage_bytes);
  L.stage = align128(L.op + L.buf_pitch);
  L.tau = L.stage + 2 * L.stage_pitch;
  L.bars = L.tau + 2 * kNMax * 4;
  L.total = L.bars + 128;
  return L;
}

// filters (M,C,7,7) -> two banks in the layout of k_pack_tc2_analysis: tf32(W) and tf32(W - tf32(W))
__global__ void k_pack_tc2_analysis_x3(const float* __restrict__ w, float* __restrict__ out, int M, int C, int Ng) {
  const int bank = kP * C * Ng * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * bank; i += gridDim.x * blockDim.x) {
    const int lo = i / bank, q = i % bank;
    const int e = q % 4, r8 = (q / 4) % 8, kc = (q / 32) % 2, grp = (q / 64) % (Ng / 8), ks = q / (Ng * 8);
    const int m = grp * 8 + r8, j = kc * 4 + e, c = ks / kP, th = ks % kP;
    const float v = (m < M && j > 0) ? w[(((size_t)m * C + c) * kP + th) * kP + (j - 1)] : 0.0f;
    const float hi = ptx::to_tf32_rna(v);
    out[i] = lo ? ptx::to_tf32_rna(v - hi) : hi;
  }
}

__global__ void __launch_bounds__(kThreads, 1) k_tc2_analysis_x3(const Ana2Params p, const __grid_constant__ CUtensorMap rmap) {
  using namespace ptx;
  using tc2::mbar_wait;                        // the bounded wait above, not ptx::mbar_wait
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const SmemLayout L = smem_layout_x3(p.C, p.Ng);
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
      mbar_wait(&oempty[0], (it & 1) ^ 1);                            // ONE operand buffer: the MMAs of tile it-1 have read it
      tc_fence_after();
      const float* sg = reinterpret_cast<const float*>(sStage + b * L.stage_pitch);
      float* op = reinterpret_cast<float*>(sOp);
      for (int idx = st; idx < nelem; idx += 32 * kShiftWarps) {
        const int row = idx / kSW, x = idx - row * kSW;
        if (x < kRW + 3) {
          const float raw = sg[idx];
          const float v = __uint_as_float(tf32_rna_bits(raw));                       // hi part: copies 0..3
          const float vl = __uint_as_float(tf32_rna_bits(__fsub_rn(raw, v)));        // lo part: copies 4..7
          float* o = op + row * kRW + x;
#pragma unroll
          for (int rho = 0; rho < 4; ++rho) {
            const int k = x - rho;
            if (k >= 0 && k < kRW) { o[rho * cp - rho] = v; o[(4 + rho) * cp - rho] = vl; }
          }
        }
      }
      fence_async_smem();                        // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) { mbar_arrive(&ofull[0]); mbar_arrive(&sempty[b]); }
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
    const uint64_t bdesc_lo = bdesc0 + (uint64_t)((L.b_bytes >> 1) >> 4);   // second bank: tf32(W - tf32(W))
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += stride, ++it) {
      const int b = it & 1, u = it >> 1;
      mbar_wait(&dempty[b], (u & 1) ^ 1);
      mbar_wait(&ofull[0], it & 1);
      tc_fence_after();
      for (int rho = 0; rho < 4; ++rho) {
        const uint32_t dcol = tbase + (uint32_t)((b * 4 + rho) * kNMax);
        const uint32_t abase = (uint32_t)rho * L.copy_pitch;
        for (int c = 0; c < p.C; ++c) {
#pragma unroll
          for (int th = 0; th < kP; ++th) {
            const int ks = c * kP + th;
            const uint32_t aoff = abase + (uint32_t)((c * kRows + th) * kRW * 4);
            // u = r_hi W_hi + r_lo W_hi + r_hi W_lo  (the dropped r_lo W_lo term is ~2^-22 relative)
            mma_tf32_ss_warp<1>(dcol, adesc0 + (uint64_t)(aoff >> 4), bdesc0 + (uint64_t)ks * bstep, idesc, ks != 0);
            mma_tf32_ss_warp<1>(dcol, adesc0 + (uint64_t)((aoff + 4 * L.copy_pitch) >> 4), bdesc0 + (uint64_t)ks * bstep, idesc, 1);
            mma_tf32_ss_warp<1>(dcol, adesc0 + (uint64_t)(aoff >> 4), bdesc_lo + (uint64_t)ks * bstep, idesc, 1);
          }
        }
      }
      mma_commit_warp<1>(&oempty[0]);            // operand copies reusable once these MMAs have read them
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
