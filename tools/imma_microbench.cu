// imma_microbench.cu -- how fast is the legacy warp-level INT8 tensor path (mma.sync.m16n8k32.s32.u8.u8) on sm_100a, and does
// it run next to a DFMA stream?  Evidence for DESIGN.md section 8 ("where the next factor is"): the reduction half of a
// Montgomery product is a contraction with a matrix that is fixed for the whole batch (the Toeplitz matrix of the modulus
// in 8-bit digits), so it could leave the FP64 pipe -- IF a tensor path that is programmable from registers is fast enough.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/rl_imma tools/imma_microbench.cu && tools/rl_imma
//
// Prints one JSON line per mode: multiply-adds (IMMA) and DFMAs per clock and SM.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

template <int NACC>
__device__ __forceinline__ void imma_step(int (&c)[NACC][4], const unsigned (&a)[4], const unsigned (&b)[2]) {
#pragma unroll
  for (int j = 0; j < NACC; j++) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
}

// mode 0: IMMA only, 1: DFMA only, 2: both in every warp (8 IMMA + 16 DFMA per iteration)
template <int MODE>
__global__ void __launch_bounds__(256) k_mix(int iters, int* out, double* dout, double seed) {
  constexpr int NACC = 8, NF = 16;
  unsigned a[4], b[2];
  int c[NACC][4];
  double f[NF];
  const unsigned tid = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = tid * 2654435761u + i;
  b[0] = tid ^ 0x9e3779b9u;
  b[1] = tid * 40503u;
#pragma unroll
  for (int j = 0; j < NACC; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) c[j][i] = 0;
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = seed + i;
  const double m = 1.0 + seed * 1e-9;
  for (int it = 0; it < iters; it++) {
    if (MODE != 1) imma_step<NACC>(c, a, b);
    if (MODE != 0) {
#pragma unroll
      for (int i = 0; i < NF; i++) f[i] = __fma_rz(f[i], m, seed);
    }
  }
  int s = 0;
#pragma unroll
  for (int j = 0; j < NACC; j++) s += c[j][0] ^ c[j][1] ^ c[j][2] ^ c[j][3];
  double fs = 0;
#pragma unroll
  for (int i = 0; i < NF; i++) fs += f[i];
  out[tid] = s;
  dout[tid] = fs;
}

template <int MODE>
static void run(const char* name, int sms, double ghz, int blocks_per_sm) {
  const int iters = 20000, threads = 256;
  const int grid = sms * blocks_per_sm;
  int* out;
  double* dout;
  cudaMalloc(&out, sizeof(int) * grid * threads);
  cudaMalloc(&dout, sizeof(double) * grid * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_mix<MODE><<<grid, threads>>>(1000, out, dout, 1.5);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k_mix<MODE><<<grid, threads>>>(iters, out, dout, 1.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double warps = static_cast<double>(grid) * threads / 32;
  const double clk = best * 1e-3 * ghz * 1e9;
  const double imma_mac = (MODE != 1) ? warps * iters * 8.0 * 4096.0 : 0.0;
  const double dfma = (MODE != 0) ? warps * iters * 16.0 * 32.0 : 0.0;
  printf("{\"mode\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"imma_mac_per_clk_per_sm\": %.1f, \"imma_tmac_per_s\": %.2f, "
         "\"dfma_per_clk_per_sm\": %.2f, \"err\": \"%s\"}\n",
         name, blocks_per_sm, best, imma_mac / clk / sms, imma_mac / (best * 1e-3) / 1e12, dfma / clk / sms, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(dout);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const double ghz = 1.965;   // B200 SM clock under load on this pool (bench.py samples it)
  for (int bps : {1, 2, 4}) {
    run<0>("imma_only", sms, ghz, bps);
    run<1>("dfma_only", sms, ghz, bps);
    run<2>("imma_and_dfma_same_warp", sms, ghz, bps);
  }
  return 0;
}
