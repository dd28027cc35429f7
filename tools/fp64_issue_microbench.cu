// FP64-pipe issue microbenchmark for B200 (sm_100a): how many issue slots does a DFMA cost when it is mixed with
// integer work?  Variants per loop iteration (8 independent chains each):
//   mode 0: 16 DFMA                      mode 1: 16 DFMA + 16 IADD3 (64-bit adds = 2 x 8... counted as instructions)
//   mode 2: 16 DFMA + 32 IADD3           mode 3: 16 DFMA + 16 IMAD.WIDE     mode 4: 8 DFMA + 8 DADD + 16 IADD3 (the kernel's mix)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue_microbench fp64_issue_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) bench(uint32_t* out, const uint32_t* in) {
  double d[8], a[8];
  uint32_t x[16];
  uint64_t w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d[i] = (double)in[(threadIdx.x + i) & 255];
    a[i] = 1.0 + 1e-9 * (double)in[(threadIdx.x + 3 * i) & 255];
    w[i] = in[(threadIdx.x + 5 * i) & 255];
  }
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = in[(threadIdx.x + 7 * i) & 255];
  double t = 1.0000001;
  uint32_t u = in[threadIdx.x & 255] | 1u;
#pragma unroll 2
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = __fma_rz(a[i], t, d[i]);
    if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 8; i++) d[i] = __dsub_rn(d[i], t);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) d[i] = __fma_rz(a[i], -t, d[i]);
    }
    if (MODE == 1 || MODE == 2 || MODE == 4) {
#pragma unroll
      for (int i = 0; i < 16; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(u));
    }
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(u + 1));
    }
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; i++) w[i] = (uint64_t)x[i] * u + w[i];
#pragma unroll
      for (int i = 0; i < 8; i++) w[i] = (uint64_t)x[i + 8] * u + w[i];
    }
    t += 1e-7;
    u += 2;
  }
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= (uint64_t)d[i] ^ w[i];
#pragma unroll
  for (int i = 0; i < 16; i++) acc ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)acc ^ (uint32_t)(acc >> 32);
}

template <int MODE>
int run(const char* name, int bps, int nsm, uint32_t* out, uint32_t* in, double fp64_per_iter, double other_per_iter) {
  int grid = bps * nsm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  bench<MODE><<<grid, 256>>>(out, in);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < 5; r++) bench<MODE><<<grid, 256>>>(out, in);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
  double iters = (double)grid * 256 * ITERS;
  double clk = ms * 1e-3 * 1.965e9;
  printf("{\"bench\":\"%s\",\"blocks_per_sm\":%d,\"ms\":%.4f,\"fp64_per_clk_per_sm\":%.2f,\"other_per_clk_per_sm\":%.2f,\"warp_instr_per_clk_per_smsp\":%.3f}\n", name, bps, ms,
         iters * fp64_per_iter / clk / nsm, iters * other_per_iter / clk / nsm, iters * (fp64_per_iter + other_per_iter) / 32 / clk / nsm / 4);
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  uint32_t *out, *in;
  CK(cudaMalloc(&out, sizeof(uint32_t) * 8 * 256 * nsm)); CK(cudaMalloc(&in, 4096));
  uint32_t hin[1024]; for (int i = 0; i < 1024; i++) hin[i] = 0x9e3779b9u * (i + 1);
  CK(cudaMemcpy(in, hin, 4096, cudaMemcpyHostToDevice));
  for (int bps : {2, 4}) {
    run<0>("16 dfma", bps, nsm, out, in, 16, 0);
    run<1>("16 dfma + 16 iadd", bps, nsm, out, in, 16, 16);
    run<2>("16 dfma + 32 iadd", bps, nsm, out, in, 16, 32);
    run<3>("16 dfma + 16 imad.wide", bps, nsm, out, in, 16, 16);
    run<4>("8 dfma + 8 dadd + 16 iadd", bps, nsm, out, in, 16, 16);
  }
  return 0;
}
