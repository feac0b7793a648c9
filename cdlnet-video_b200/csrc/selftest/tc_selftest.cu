// tc_selftest.cu — bring-up checks for the tcgen05 primitives used by the tensor-core kernels.
// Each case computes D[128*CG x N] = A[128*CG x K] * B[N x K]^T (tf32, fp32 accumulate) with operands
// that are exactly representable, so any mismatch is a layout / descriptor error, not rounding.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o tc_selftest tc_selftest.cu
//   run  : ./tc_selftest <cg:1|2> <ts:0|1> <N> <KS> [rep probe commit_every overlap nclusters]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../cdl_tc_ptx.cuh"

using namespace cdl::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

// B packed per K-step j (8 k-values): [j][n/8][k/4][n%8][k%4]  (SBO = 256 B between 8-row groups, LBO = 128 B between k-chunks)
// For CG == 2 each CTA holds rows [rank*N/2, (rank+1)*N/2) of B, same packing with N/2 rows.
template <int CG, bool TS>
__global__ void __launch_bounds__(128) k_gemm(const float* __restrict__ A, const float* __restrict__ Bp, float* __restrict__ D, int N, int K, int rep, long long* cyc, int cgroup, int ovl) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[8];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  const int KS = K / 8;
  const int NL = N / CG;                       // B rows held by this CTA
  float* sB = reinterpret_cast<float*>(smem);                       // KS * NL * 8 floats
  float* sA = sB + (size_t)KS * NL * 8;                             // KS * 128 * 8 floats (SS only)

  if (warp == 0) { tmem_alloc<CG>(&tmem_base_s, 512); tmem_relinquish<CG>(); }
  if (tid == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bar2[i], 1); fence_mbar_init(); }
  // B: this CTA's rows
  for (int i = tid; i < KS * NL * 8; i += 128) {
    int j = i / (NL * 8), rem = i % (NL * 8);
    sB[i] = Bp[(size_t)j * N * 8 + (size_t)rank * NL * 8 + rem];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const int row = rank * 128 + tid;            // global A row of this thread == TMEM lane tid
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  const int ACOL = 256;                        // A operand columns in TMEM (TS)
  if (TS) {
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t v[8];
      for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(A[(size_t)row * K + k0 + i]);
      tmem_st8(lane_addr + ACOL + k0, v);
    }
    tmem_wait_st();
  } else {
    if (ovl) {      // overlapping-window source: A[m][8j+k] = S[rank][j][(m/8)*36 + 4*(m%8) + k]; A carries S in its first floats
      for (int i = tid; i < KS * 1024; i += 128) sA[i] = A[(size_t)2 * 128 * K + (size_t)rank * KS * 1024 + i];
    } else
    for (int k = 0; k < K; ++k) {
      int j = k / 8, kk = k % 8;
      sA[(size_t)j * 1024 + (tid / 8) * 64 + (kk / 4) * 32 + (tid % 8) * 4 + (kk % 4)] = A[(size_t)row * K + k];
    }
  }
  fence_async_smem();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

#ifdef SELFTEST_WARP_ISSUE
  if (rank == 0 && warp == 0) {      // converged warp, elected lane issues
#ifdef SELFTEST_CONST
#ifndef SELFTEST_KS
#define SELFTEST_KS 7        // K-steps (distinct operand pairs) the constant-operand loop cycles through; run with the same KS argument
#endif
    if (tbase != 0) __trap();
    const uint32_t idesc = make_idesc_tf32(128 * CG, 176);
    const uint64_t bd0 = make_smem_desc_kmajor_noswz(smem_u32(smem), 128, 256);
    const long long t0 = clock64();
    for (int r = 0; r < rep; ++r) {
      if (TS) {
#pragma unroll
        for (int j = 0; j < SELFTEST_KS; ++j)
          mma_tf32_ts_warp<CG>(0, ACOL + j * 8, bd0 + (uint64_t)((j * (176 / CG) * 32) >> 4), idesc, 1);
      } else {
        const uint64_t ad0 = make_smem_desc_kmajor_noswz(smem_u32(smem) + SELFTEST_KS * (176 / CG) * 32, 16, 144);
#pragma unroll
        for (int j = 0; j < SELFTEST_KS; ++j)
          mma_tf32_ss_warp<CG>(0, ad0 + (uint64_t)((j * 4096) >> 4), bd0 + (uint64_t)((j * (176 / CG) * 32) >> 4), idesc, 1);
      }
      if (cgroup > 0) mma_commit_warp<CG>(&bar2[r & 7]);
    }
#else
    const uint32_t idesc = make_idesc_tf32(128 * CG, N);
    const uint64_t bd0 = make_smem_desc_kmajor_noswz(smem_u32(sB), 128, 256);
    const long long t0 = clock64();
    for (int r = 0; r < rep; ++r)
    for (int j = 0; j < KS; ++j) {
      mma_tf32_ts_warp<CG>(tbase, tbase + ACOL + j * 8, bd0 + (uint64_t)((j * NL * 32) >> 4), idesc, j > 0);
      if (cgroup > 0 && (((r * KS + j + 1) & (cgroup - 1)) == 0)) mma_commit_warp<CG>(&bar2[(r + j) & 7]);
    }
#endif
    const long long t1 = clock64();
    mma_commit_warp<CG>(&bar);
    const long long t2 = clock64();
    mbar_wait(&bar, 0);
    if (cyc && tid == 0) { cyc[0] = clock64() - t0; cyc[1] = t1 - t0; cyc[2] = t2 - t1; }
  }
#else
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc_tf32(128 * CG, N);
    const long long t0 = clock64();
    for (int r = 0; r < rep; ++r)
    for (int j = 0; j < KS; ++j) {
      uint64_t bdesc = make_smem_desc_kmajor_noswz(smem_u32(sB) + j * NL * 32, 128, 256);
      if (TS) mma_tf32_ts<CG>(tbase, tbase + ACOL + j * 8, bdesc, idesc, j > 0);
      else {
        uint64_t adesc = ovl ? make_smem_desc_kmajor_noswz(smem_u32(sA) + j * 4096, 16, 144) : make_smem_desc_kmajor_noswz(smem_u32(sA) + j * 4096, 128, 256);
        mma_tf32_ss<CG>(tbase, adesc, bdesc, idesc, j > 0);
      }
      if (cgroup > 0 && (((r * KS + j + 1) & (cgroup - 1)) == 0)) mma_commit<CG>(&bar2[(r + j) & 7]);   // commit every `cgroup` (power of 2) MMAs, nobody waits
    }
    const long long t1 = clock64();
    mma_commit<CG>(&bar);
    const long long t2 = clock64();
    mbar_wait(&bar, 0);
    if (cyc) { cyc[0] = clock64() - t0; cyc[1] = t1 - t0; cyc[2] = t2 - t1; }
  }
#endif
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    tmem_ld8(lane_addr + c0, v);
    tmem_wait_ld();
    for (int i = 0; i < 8; ++i) D[(size_t)row * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) tmem_dealloc<CG>(tbase, 512);
}

int main(int argc, char** argv) {
  int cg = argc > 1 ? atoi(argv[1]) : 1, ts = argc > 2 ? atoi(argv[2]) : 0, N = argc > 3 ? atoi(argv[3]) : 176, KS = argc > 4 ? atoi(argv[4]) : 7;
  int rep = argc > 5 ? atoi(argv[5]) : 1, probe = argc > 6 ? atoi(argv[6]) : 0, cgroup = argc > 7 ? atoi(argv[7]) : 0, ovl = argc > 8 ? atoi(argv[8]) : 0;
  const int nclusters = argc > 9 ? atoi(argv[9]) : 1;   // > 1: the same GEMM on that many clusters at once (whole-chip MMA rate; every cluster writes the same D)
  const int M = 128 * cg, K = KS * 8;
  std::vector<float> A((size_t)M * K + (size_t)2 * 128 * K + 2 * KS * 1024), B((size_t)N * K), Bp((size_t)N * K), D((size_t)M * N, -1.f), R((size_t)M * N);
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int)((s >> 20) % 9) - 4; };
  for (auto& v : A) v = rnd() / 4.0f;
  if (ovl) {
    float* S = A.data() + (size_t)2 * 128 * K;
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < K; ++k) {
        int rk = m / 128, ml = m % 128, j = k / 8, kk = k % 8;
        A[(size_t)m * K + k] = S[(size_t)rk * KS * 1024 + j * 1024 + (ml / 8) * 36 + 4 * (ml % 8) + kk];
      }
  }
  for (auto& v : B) v = rnd() / 8.0f;
  if (probe) {
    // Operand-rounding probe: D[m][n] = A[m][0] exactly as the tensor core reads it (B picks k = 0; everything else 0).
    // A[m][0] carries all 13 sub-tf32 mantissa bits in varied patterns, both signs, several exponents.
    for (auto& v : A) v = 0.0f;
    for (auto& v : B) v = 0.0f;
    for (int n = 0; n < N; ++n) B[(size_t)n * K] = 1.0f;
    for (int m = 0; m < M; ++m) {
      uint32_t bits = 0x3f800000u + ((uint32_t)(m % 7) << 23) + ((uint32_t)m * 0x9E3779B1u & 0x007fffffu);
      if (m & 1) bits |= 0x80000000u;
      if (m == 2) bits = 0x3f800000u | 0x1fffu;          // 1 + (2^13 - 1) ulp: truncation -> 1, nearest -> 1 + 2^-10
      if (m == 3) bits = 0x3f800000u | 0x1000u;          // exactly half way
      if (m == 4) bits = 0x00001000u;                    // the encoding of zero used by the pre-biased code layout
      memcpy(&A[(size_t)m * K], &bits, 4);
    }
  }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      int j = k / 8, kk = k % 8;
      // full-N packing; a CTA of a pair takes the contiguous row range [rank*N/2, ...) of each K-step block
      Bp[(size_t)j * N * 8 + (size_t)(n / 8) * 64 + (kk / 4) * 32 + (n % 8) * 4 + (kk % 4)] = B[(size_t)n * K + k];
    }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      R[(size_t)m * N + n] = (float)acc;
    }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bp.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
  size_t smem = ((size_t)KS * (N / cg) * 8 + (size_t)KS * 1024) * 4 + 1024;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cg * nclusters); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  long long* dcyc; CK(cudaMalloc(&dcyc, 32)); CK(cudaMemset(dcyc, 0, 32));
  void (*fn)(const float*, const float*, float*, int, int, int, long long*, int, int) =
      cg == 1 ? (ts ? k_gemm<1, true> : k_gemm<1, false>) : (ts ? k_gemm<2, true> : k_gemm<2, false>);
  CK(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaLaunchKernelEx(&cfg, fn, (const float*)dA, (const float*)dB, dD, N, K, rep, dcyc, cgroup, ovl));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  long long hcs[4]; CK(cudaMemcpy(hcs, dcyc, 32, cudaMemcpyDeviceToHost));
  long long hc = hcs[0];
  printf("issue timing cg=%d ts=%d N=%d: %d MMAs issued in %lld cycles (%.1f each), commit %lld, all complete after %lld\n", cg, ts, N, rep * KS, hcs[1], (double)hcs[1] / (rep * KS), hcs[2], hcs[0]);
  if (rep > 1) printf("timing cg=%d ts=%d N=%d KS=%d rep=%d commit-every=%d : %.1f cycles per MMA\n", cg, ts, N, KS, rep, cgroup, (double)hc / ((double)rep * KS));
  if (probe) {
    long ntrunc = 0, nrna = 0, nother = 0;
    for (int m = 0; m < M; ++m) {
      uint32_t a, d;
      memcpy(&a, &A[(size_t)m * K], 4); memcpy(&d, &D[(size_t)m * N], 4);
      const uint32_t tr = a & 0xffffe000u, rn = (a + 0x1000u) & 0xffffe000u;
      if (d == tr) ++ntrunc;
      if (d == rn) ++nrna;
      if (d != tr && d != rn) { ++nother; if (nother < 5) printf("  row %d: a=%08x d=%08x trunc=%08x rna=%08x\n", m, a, d, tr, rn); }
    }
    printf("operand probe ts=%d cg=%d: %d rows, D == truncate(A) on %ld, D == rna(A) on %ld, neither on %ld  => the tensor core %s its fp32 operand words\n",
           ts, cg, M, ntrunc, nrna, nother, (ntrunc == M) ? "TRUNCATES" : (nrna == M ? "ROUNDS (rna)" : "does something else with"));
    return ntrunc == M ? 0 : 3;
  }
  double maxerr = 0; long bad = 0;
  for (size_t i = 0; i < D.size(); ++i) { double e = fabs((double)D[i] - R[i]); if (e > maxerr) maxerr = e; if (e > 1e-5) ++bad; }
  printf("selftest cg=%d ts=%d M=%d N=%d K=%d : max|err|=%.3e bad=%ld/%zu  %s\n", cg, ts, M, N, K, maxerr, bad, D.size(), bad ? "FAIL" : "PASS");
  if (bad) {
    int shown = 0;
    for (int m = 0; m < M && shown < 6; ++m)
      for (int n = 0; n < N && shown < 6; ++n)
        if (fabs(D[(size_t)m * N + n] - R[(size_t)m * N + n]) > 1e-5) { printf("  D[%d][%d]=%g ref=%g\n", m, n, D[(size_t)m * N + n], R[(size_t)m * N + n]); ++shown; }
  }
  return bad ? 1 : 0;
}
