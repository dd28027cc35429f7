// Second IMAD-pipe microbenchmark (B200, sm_100a): the first one let ptxas hoist the loop-invariant
// product out of the "carry-less IMAD.WIDE" loop, so that line measured IADD3.  Here the multiplier
// changes every iteration and the SASS mix of every variant is checked with tools/sass_loops.py.
//   mode 0: IMAD.WIDE.U32 Rd, Ra, Rb, Rc   (64-bit accumulate in the instruction)
//   mode 1: IMAD.WIDE.U32 Rd, Ra, Rb, RZ + IADD3 / IADD3.X (two products folded by one 3-input add pair)
//   mode 2: IMAD (32-bit low product, accumulate)
//   mode 3: IMAD.HI.U32 + IMAD (separate high / low halves)
//   mode 4: DFMA
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) bench(uint32_t* out, const uint32_t* in) {
  uint32_t a[8];
  uint64_t acc[8];
  double dacc[8], da[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = in[(threadIdx.x + i) & 255] | 1u;
    acc[i] = in[(threadIdx.x + 9 * i) & 255];
    dacc[i] = (double)acc[i];
    da[i] = (double)a[i] * 1e-9;
  }
  uint32_t t = in[threadIdx.x & 255] | 3u, t2 = t * 7u;
  double dt = 1.0000001;
  if (MODE == 1) {
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        uint64_t p, q;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[i]), "r"(t));
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(q) : "r"(a[i]), "r"(t2));
        acc[i] = acc[i] + p + q;
      }
      t += 0x9e3779b9u;
      t2 ^= t;
    }
  } else
#pragma unroll 2
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = (uint64_t)a[i] * t + acc[i];
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = (uint64_t)a[i] * t2 + acc[i];
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        uint64_t p, q;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[i]), "r"(t));
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(q) : "r"(a[i]), "r"(t2));
        acc[i] = acc[i] + p + q;
      }
    } else if (MODE == 2) {
      uint32_t* r = reinterpret_cast<uint32_t*>(acc);
#pragma unroll
      for (int i = 0; i < 16; i++) r[i] = a[i & 7] * (i < 8 ? t : t2) + r[i];
    } else if (MODE == 3) {
      uint32_t* r = reinterpret_cast<uint32_t*>(acc);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        r[2 * i] = a[i] * t + r[2 * i];
        r[2 * i + 1] = __umulhi(a[i], t2) + r[2 * i + 1];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) dacc[i] = fma(da[i], dt, dacc[i]);
#pragma unroll
      for (int i = 0; i < 8; i++) dacc[i] = fma(da[i], -dt, dacc[i]);
      dt += 1e-7;
    }
    t += 0x9e3779b9u;
    t2 ^= t;
  }
  uint64_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x ^= acc[i] ^ (uint64_t)dacc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)x ^ (uint32_t)(x >> 32);
}

template <int MODE>
int run(const char* name, int blocks_per_sm, int nsm, uint32_t* out, uint32_t* in, double ops_per_iter) {
  int grid = blocks_per_sm * nsm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  bench<MODE><<<grid, 256>>>(out, in);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < 5; r++) bench<MODE><<<grid, 256>>>(out, in);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 5;
  double ops = (double)grid * 256 * ITERS * ops_per_iter;
  printf("{\"bench\":\"%s\",\"blocks_per_sm\":%d,\"ms\":%.4f,\"tera_ops_per_s\":%.3f,\"ops_per_clk_per_sm_at_1965MHz\":%.2f}\n", name, blocks_per_sm, ms,
         ops / (ms * 1e-3) / 1e12, ops / (ms * 1e-3) / nsm / 1.965e9);
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  uint32_t *out, *in;
  CK(cudaMalloc(&out, sizeof(uint32_t) * 8 * 256 * nsm)); CK(cudaMalloc(&in, 4096));
  uint32_t hin[1024]; for (int i = 0; i < 1024; i++) hin[i] = 0x9e3779b9u * (i + 1);
  CK(cudaMemcpy(in, hin, 4096, cudaMemcpyHostToDevice));
  for (int bps : {2, 4, 8}) {
    run<0>("imad_wide_accumulate (MAC)", bps, nsm, out, in, 16);
    run<1>("imad_wide_rz+iadd3 (MAC)", bps, nsm, out, in, 16);
    run<2>("imad_lo32 (instr)", bps, nsm, out, in, 16);
    run<3>("imad_hi+imad_lo (instr)", bps, nsm, out, in, 16);
    run<4>("dfma (instr)", bps, nsm, out, in, 16);
  }
  return 0;
}
