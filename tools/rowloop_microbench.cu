// Microbenchmark of the FP64-engine row loop (bn52.cuh, mmd_rows_core) in isolation: which resource bounds it?
// Build variants with -DSCB_EXP_NO_SHFL / -DSCB_EXP_NO_CHAIN / -DSCB_EXP_NO_IADD and -DBLOCKS_PER_SM=n, run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-D...] tools/rowloop_microbench.cu -o /tmp/rl && /tmp/rl
// Prints one JSON line: cycles per warp-row per scheduler and the implied FP64-pipe utilisation.  No correctness claim
// (operands are arbitrary 52-bit integers; the experiment switches change the arithmetic).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../protocols/secure_comparison_b200/csrc/bn52.cuh"

#ifndef BLOCKS_PER_SM
#define BLOCKS_PER_SM 3
#endif
#ifndef TPI_
#define TPI_ 4
#endif
#ifndef PAIR_
#define PAIR_ false
#endif
using namespace scb;
constexpr int L = 10, TPI = TPI_, K = L * TPI, GPW = 32 / TPI;

__global__ void __launch_bounds__(128, BLOCKS_PER_SM) k_rows(const double* in, double* out, int passes, int flags) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, t = lane % TPI;
  double* area = sm + (threadIdx.x >> 5) * ((2 * K + 2) * GPW) + lane / TPI;
  uint64_t* sq = reinterpret_cast<uint64_t*>(area + (K + 1) * GPW);
  double av[L], nv[L], r[L];
  for (int i = 0; i < L; i++) {
    av[i] = in[(t * L + i) % 64];
    nv[i] = in[(t * L + i + 7) % 64];
    r[i] = av[i];
  }
  for (int j = t; j < K + 1; j += TPI) area[j * GPW] = in[j % 64];
  for (int j = t; j < K; j += TPI) sq[j * GPW] = 0ull;
  __syncwarp();
  const double np = in[5];
  for (int p = 0; p < passes; p++) {
    uint64_t col[L], hi;
    mmd_rows_core<L, TPI, PAIR_>(r, nv, np, area, sq, flags, col, hi);
    resolve_carries<L, TPI>(col, hi, r);
    stage_digits52<L, TPI>(area, r);
  }
  double acc = 0;
  for (int i = 0; i < L; i++) acc += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv) {
  const int passes = argc > 1 ? atoi(argv[1]) : 2000;
  const int flags = argc > 2 ? atoi(argv[2]) : 0;
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double hin[64];
  for (int i = 0; i < 64; i++) hin[i] = static_cast<double>((0x9e3779b97f4aull * (i + 1)) & ((1ull << 52) - 1));
  double *in, *out;
  cudaMalloc(&in, sizeof(hin));
  cudaMalloc(&out, sizeof(double) * 128 * BLOCKS_PER_SM * nsm);
  cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
  const size_t smem = sizeof(double) * 4 * (2 * K + 2) * GPW;
  const int grid = BLOCKS_PER_SM * nsm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_rows<<<grid, 128, smem>>>(in, out, passes / 10, flags);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_rows<<<grid, 128, smem>>>(in, out, passes, flags);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rows, 128, smem);
  const double warp_rows_per_sched = static_cast<double>(passes) * K * BLOCKS_PER_SM;   // 4 warps per block = 1 per scheduler per block
  const double cycles = ms * 1e-3 * 1.965e9;
  printf("{\"variant\":\"%s\",\"tpi\":%d,\"pair\":%d,\"flags\":%d,\"blocks_per_sm\":%d,\"occupancy\":%d,\"ms\":%.3f,\"cycles_per_warp_row_per_scheduler\":%.1f,"
         "\"fp64_pipe_util\":%.3f,\"err\":\"%s\"}\n",
#if defined(SCB_EXP_NO_SHFL)
         "no_shfl"
#elif defined(SCB_EXP_NO_CHAIN)
         "no_chain"
#elif defined(SCB_EXP_NO_IADD)
         "no_iadd"
#else
         "baseline"
#endif
         , TPI, PAIR_ ? 1 : 0, flags, BLOCKS_PER_SM, occ, ms, cycles / warp_rows_per_sched, 128.0 / (cycles / warp_rows_per_sched), cudaGetErrorString(e));
  return 0;
}
