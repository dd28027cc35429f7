// scb200_peaks.cu -- measured issue-rate ceilings of the two multiplier pipes the engine uses, taken on the GPU at hand.
//
// bench.py's roofline is an integer-multiply roofline (BASELINE north star); neither pipe rate is in MEASURED_PEAKS.json
// (it holds HBM and bf16 figures), so the library measures them itself: a register-resident stream of accumulating
// IMAD.WIDE.U32 (32x32+64 -> 64, the multiplier of the IMAD engine, bn29.cuh) and a stream of DFMA (the 52x52-bit
// multiplier of the FP64 engine, bn52.cuh), 8 independent chains per thread, multiplier changing every iteration so that
// nothing is hoisted.  Same kernels as tools/imad_microbench2.cu (modes 0 and 4), behind the C ABI.
#include "../../../include/scb200.h"

#include <cstdint>
#include <cuda_runtime.h>

#include "host_ctx.hpp"

namespace {

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k_peak(uint32_t* out, const uint32_t* in) {
  uint32_t a[8];
  uint64_t acc[8];
  double dacc[8], da[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = in[(threadIdx.x + i) & 255] | 1u;
    acc[i] = in[(threadIdx.x + 9 * i) & 255];
    dacc[i] = static_cast<double>(acc[i]);
    da[i] = static_cast<double>(a[i]) * 1e-9;
  }
  uint32_t t = in[threadIdx.x & 255] | 3u, t2 = t * 7u;
  double dt = 1.0000001;
#pragma unroll 2
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = static_cast<uint64_t>(a[i]) * t + acc[i];
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = static_cast<uint64_t>(a[i]) * t2 + acc[i];
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
  for (int i = 0; i < 8; i++) x ^= acc[i] ^ static_cast<uint64_t>(dacc[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = static_cast<uint32_t>(x) ^ static_cast<uint32_t>(x >> 32);
}

}  // namespace

extern "C" int scb_measure_peak(scb_ctx* ctx, int which, double* ops_per_second) {
  if (!ctx || !ops_per_second || (which != SCB_PEAK_IMAD_WIDE && which != SCB_PEAK_DFMA)) return scb::fail(ctx, SCB_EINVAL, "scb_measure_peak: bad argument");
  cudaSetDevice(ctx->device);
  const int blocks_per_sm = 8, reps = 5;
  const int grid = blocks_per_sm * ctx->num_sms;
  void* buf = nullptr;
  int rc = ctx->scratch(3, sizeof(uint32_t) * (static_cast<size_t>(grid) * 256 + 1024), &buf);
  if (rc) return rc;
  uint32_t* out = static_cast<uint32_t*>(buf);
  uint32_t* in = out + static_cast<size_t>(grid) * 256;
  uint32_t hin[1024];
  for (int i = 0; i < 1024; i++) hin[i] = 0x9e3779b9u * (i + 1);
  cudaError_t e = cudaMemcpyAsync(in, hin, sizeof(hin), cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return scb::cuda_fail(ctx, e, "H2D copy");
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return scb::fail(ctx, SCB_ECUDA, "cudaEventCreate");
  auto launch = [&]() {
    if (which == SCB_PEAK_IMAD_WIDE)
      k_peak<0><<<grid, 256, 0, ctx->stream>>>(out, in);
    else
      k_peak<1><<<grid, 256, 0, ctx->stream>>>(out, in);
  };
  launch();                                               // warm-up
  cudaEventRecord(e0, ctx->stream);
  for (int r = 0; r < reps; r++) launch();
  cudaEventRecord(e1, ctx->stream);
  e = cudaStreamSynchronize(ctx->stream);
  float ms = 0.f;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx->launches += reps + 1;
  if (e != cudaSuccess || ms <= 0.f) return scb::cuda_fail(ctx, e, "scb_measure_peak");
  const double ops = static_cast<double>(grid) * 256.0 * ITERS * 16.0 * reps;
  *ops_per_second = ops / (ms * 1e-3);
  return SCB_OK;
}
