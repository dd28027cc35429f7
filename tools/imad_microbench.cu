// IMAD-pipe microbenchmark for B200 (sm_100a): measures the sustained rate of the
// 32x32->64 multiply-accumulate forms the Montgomery kernels are built from, so the
// integer roofline denominator (MAC/clk/SM) is a measured number, not a planning figure.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_microbench imad_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

// mode 0: independent mad.wide.u32 (IMAD.WIDE.U32, no carry), 8 accumulators
// mode 1: two carry chains of 8 (mad.lo.cc/madc.hi.cc -> IMAD.WIDE.U32.X), 16 MAC per iter
// mode 2: mad.lo.u32 (plain 32-bit IMAD), 8 accumulators
// mode 3: mode 1 + one shfl + 4 iadd3 per 16 MAC (co-issue test)
// mode 4: mode 1 + one LDS + one shfl + 4 add per 16 MAC
template <int MODE>
__global__ void __launch_bounds__(256) bench(uint32_t* out, const uint32_t* in, unsigned long long* cycles) {
  __shared__ uint32_t sm[1024];
  uint32_t a[8], r[18];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = in[(threadIdx.x + i) & 255] | 1u;
#pragma unroll
  for (int i = 0; i < 18; i++) r[i] = in[(threadIdx.x + 7 * i) & 255];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i & 255];
  uint32_t t = in[threadIdx.x & 255] | 3u;
  uint32_t x0 = 1, x1 = 2, x2 = 3, x3 = 4;
  __syncthreads();
  unsigned long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {
      unsigned long long* r64 = reinterpret_cast<unsigned long long*>(r);
#pragma unroll
      for (int rep = 0; rep < 2; rep++)
#pragma unroll
        for (int i = 0; i < 8; i++)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r64[i]) : "r"(a[i]), "r"(t));
    } else if (MODE == 2) {
#pragma unroll
      for (int rep = 0; rep < 2; rep++)
#pragma unroll
        for (int i = 0; i < 8; i++)
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a[i]), "r"(t));
    } else {
      // chain E on r[0..7] (4 MACs) x2 and chain O on r[8..15]: emulate two independent 8-MAC carry chains
      asm volatile(
          "mad.lo.cc.u32 %0, %8, %16, %0;\n\t madc.hi.cc.u32 %1, %8, %16, %1;\n\t"
          "madc.lo.cc.u32 %2, %9, %16, %2;\n\t madc.hi.cc.u32 %3, %9, %16, %3;\n\t"
          "madc.lo.cc.u32 %4, %10, %16, %4;\n\t madc.hi.cc.u32 %5, %10, %16, %5;\n\t"
          "madc.lo.cc.u32 %6, %11, %16, %6;\n\t madc.hi.cc.u32 %7, %11, %16, %7;\n\t"
          "madc.lo.cc.u32 %0, %12, %16, %0;\n\t madc.hi.cc.u32 %1, %12, %16, %1;\n\t"
          "madc.lo.cc.u32 %2, %13, %16, %2;\n\t madc.hi.cc.u32 %3, %13, %16, %3;\n\t"
          "madc.lo.cc.u32 %4, %14, %16, %4;\n\t madc.hi.cc.u32 %5, %14, %16, %5;\n\t"
          "madc.lo.cc.u32 %6, %15, %16, %6;\n\t madc.hi.u32 %7, %15, %16, %7;\n\t"
          : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(t));
      asm volatile(
          "mad.lo.cc.u32 %0, %8, %16, %0;\n\t madc.hi.cc.u32 %1, %8, %16, %1;\n\t"
          "madc.lo.cc.u32 %2, %9, %16, %2;\n\t madc.hi.cc.u32 %3, %9, %16, %3;\n\t"
          "madc.lo.cc.u32 %4, %10, %16, %4;\n\t madc.hi.cc.u32 %5, %10, %16, %5;\n\t"
          "madc.lo.cc.u32 %6, %11, %16, %6;\n\t madc.hi.cc.u32 %7, %11, %16, %7;\n\t"
          "madc.lo.cc.u32 %0, %12, %16, %0;\n\t madc.hi.cc.u32 %1, %12, %16, %1;\n\t"
          "madc.lo.cc.u32 %2, %13, %16, %2;\n\t madc.hi.cc.u32 %3, %13, %16, %3;\n\t"
          "madc.lo.cc.u32 %4, %14, %16, %4;\n\t madc.hi.cc.u32 %5, %14, %16, %5;\n\t"
          "madc.lo.cc.u32 %6, %15, %16, %6;\n\t madc.hi.u32 %7, %15, %16, %7;\n\t"
          : "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
          : "r"(a[1]), "r"(a[0]), "r"(a[3]), "r"(a[2]), "r"(a[5]), "r"(a[4]), "r"(a[7]), "r"(a[6]), "r"(t));
      if (MODE >= 3) {
        uint32_t s = __shfl_sync(0xffffffffu, r[0], (it & 3) * 8);
        if (MODE == 4) s += sm[(threadIdx.x * 4 + it) & 1023];
        asm volatile("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, %3, 0;"
                     : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3) : "r"(s));
        t ^= x3;
      }
    }
  }
  unsigned long long t1 = clock64();
  uint32_t acc = x0 ^ x1 ^ x2 ^ x3;
#pragma unroll
  for (int i = 0; i < 18; i++) acc ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run(const char* name, int blocks_per_sm, int threads, int nsm, uint32_t* out, uint32_t* in, unsigned long long* cyc, double mac_per_iter) {
  int grid = blocks_per_sm * nsm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  bench<MODE><<<grid, threads>>>(out, in, cyc);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  bench<MODE><<<grid, threads>>>(out, in, cyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  static unsigned long long h[4096];
  CK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < grid; i++) avg += (double)h[i]; avg /= grid;
  double macs_sm = (double)blocks_per_sm * threads * ITERS * mac_per_iter;   // per SM
  printf("{\"bench\":\"%s\",\"blocks_per_sm\":%d,\"threads\":%d,\"mac_per_clk_per_sm\":%.2f,\"ms\":%.4f,\"tmac_per_s\":%.3f,\"eff_mhz\":%.0f}\n",
         name, blocks_per_sm, threads, macs_sm / avg, ms, macs_sm * nsm / (ms * 1e-3) / 1e12, avg / (ms * 1e-3) / 1e6);
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, nsm, p.clockRate);
  uint32_t *out, *in; unsigned long long* cyc;
  CK(cudaMalloc(&out, sizeof(uint32_t) * 4096 * 1024)); CK(cudaMalloc(&in, 4096)); CK(cudaMalloc(&cyc, 8 * 4096));
  uint32_t hin[1024]; for (int i = 0; i < 1024; i++) hin[i] = 0x9e3779b9u * (i + 1);
  CK(cudaMemcpy(in, hin, 4096, cudaMemcpyHostToDevice));
  int cfgs[][2] = {{1, 128}, {1, 256}, {2, 256}, {4, 256}, {8, 256}};
  for (auto& c : cfgs) {
    run<0>("imad_wide_indep", c[0], c[1], nsm, out, in, cyc, 16);
    run<1>("imad_wide_x_carry_2chains", c[0], c[1], nsm, out, in, cyc, 16);
    run<2>("imad_lo32", c[0], c[1], nsm, out, in, cyc, 16);
    run<3>("carry_chains+shfl+4add", c[0], c[1], nsm, out, in, cyc, 16);
    run<4>("carry_chains+lds+shfl+4add", c[0], c[1], nsm, out, in, cyc, 16);
  }
  return 0;
}
