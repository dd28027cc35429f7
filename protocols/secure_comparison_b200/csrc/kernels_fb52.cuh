// kernels_fb52.cuh -- fixed-base window tables and the fixed-base product on the FP64 engine (bn52.cuh): the 52-bit twins of
// k_fb_bases / k_fb_fill / k_fb16_fill / group_fixed_base_mul in kernels_generic.cuh (same layout, doubles instead of uint32).
#pragma once
#include "kernels_f64.cuh"

namespace scb {

template <int L, int TPI>
__device__ __forceinline__ void copy52(double* dst, const double* src) {
  const int t = (threadIdx.x & 31) % TPI;
#pragma unroll
  for (int i = 0; i < L; i++) dst[t * L + i] = src[t * L + i];
}

// words (canonical residue, kw words) -> Montgomery form at dst
template <int L, int TPI>
__device__ __forceinline__ void words_to_mont52(const Modulus52Dev* __restrict__ m, double* sm_group, double* dst, const uint32_t* __restrict__ words,
                                                int kw) {
  const int t = (threadIdx.x & 31) % TPI;
#pragma unroll
  for (int i = 0; i < L; i++) dst[t * L + i] = u52_to_double(digit52_from_words(words, kw, t * L + i));
  mmd<L, TPI>(m, sm_group, dst, dst, m->r2);
}

// -------------------------------------------------------------------------------------------------
// Fixed-base tables: T[j][v] = base^(v * 256^j) in Montgomery form, j < nwin, v < 256; then 16-bit windows from pairs.
// -------------------------------------------------------------------------------------------------
template <int L, int TPI>
__global__ void __launch_bounds__(32) k52_fb_bases(const Modulus52Dev* __restrict__ m, const uint32_t* __restrict__ base, int k32, double* table, int nwin) {
  extern __shared__ double smem52[];
  constexpr int K = Geo52<L, TPI>::K;
  const int grp = (threadIdx.x & 31) / TPI;
  double* sm_group = smem52 + grp;
  // every group of the warp computes the same thing into its own scratch copy; group 0's copy is the table itself
  double* T = table + (grp == 0 ? 0 : static_cast<size_t>(nwin) * 256 * K + static_cast<size_t>(grp - 1) * K * 2);
  double* cur = (grp == 0) ? T + K : T;
  words_to_mont52<L, TPI>(m, sm_group, cur, base, k32);
  for (int j = 1; j < nwin; j++) {
    double* nxt = (grp == 0) ? table + (static_cast<size_t>(j) * 256 + 1) * K : (cur == T ? T + K : T);
    mmd<L, TPI>(m, sm_group, nxt, cur, cur);
    for (int s = 1; s < 8; s++) mmd<L, TPI>(m, sm_group, nxt, nxt, nxt);
    cur = nxt;
  }
}

template <int L, int TPI>
__global__ void __launch_bounds__(BLOCK_THREADS, (L <= 10 ? 3 : 2)) k52_fb_fill(const Modulus52Dev* __restrict__ m, double* table, int nwin, double* dump) {
  extern __shared__ double smem52[];
  using GI = GroupIndex52<L, TPI>;
  constexpr int K = Geo52<L, TPI>::K;
  double* sm_group = GI::warp_smem(smem52) + (threadIdx.x & 31) / TPI;
  const size_t j = GI::slot();
  const bool active = j < static_cast<size_t>(nwin);
  // inactive groups recompute window 0 into a private dump area (keeps the warp converged)
  double* T = active ? table + j * 256 * K : dump + (j - nwin) * 2 * K;
  const double* base1 = active ? T + K : table + K;
  if (active) copy52<L, TPI>(T, m->one);
  const double* prev = base1;
  for (int v = 2; v < 256; v++) {
    double* dst = active ? T + static_cast<size_t>(v) * K : T + (v & 1) * K;
    mmd<L, TPI>(m, sm_group, dst, prev, base1);
    prev = dst;
  }
}

template <int L, int TPI>
__global__ void __launch_bounds__(BLOCK_THREADS, (L <= 10 ? 3 : 2)) k52_fb16_fill(const Modulus52Dev* __restrict__ m, const double* table8, double* table16, size_t M,
                                                                  double* dump, unsigned long long* queue) {
  extern __shared__ double smem52[];
  using GI = GroupIndex52<L, TPI>;
  constexpr int K = Geo52<L, TPI>::K;
  double* sm_group = GI::warp_smem(smem52) + (threadIdx.x & 31) / TPI;
  const size_t slot = GI::slot();
  size_t idx;
  bool active;
  while (GI::next(queue, M, idx, active)) {
    const size_t j = idx >> 16, v = idx & 0xffff;
    double* dst = active ? table16 + idx * K : dump + slot * K;
    mmd<L, TPI>(m, sm_group, dst, table8 + ((2 * j) * 256 + (v & 255)) * K, table8 + ((2 * j + 1) * 256 + (v >> 8)) * K);
  }
}

// acc (Montgomery, lazy) *= base^exp using the window table (wbits = 8 or 16).  e: this group's exponent words.
// (A register-resident variant of this chain and of the squaring chain of step_4f -- one more inlined copy of the row loop in
// the fused kernel -- measured SLOWER: 110 ms against 98 ms for k52_step4_main at B = 65,536; the memory-to-memory form stays.)
template <int L, int TPI>
__device__ __forceinline__ void group_fixed_base_mul52(const Modulus52Dev* __restrict__ m, double* sm_group, double* acc, const double* table, int nwin,
                                                       const uint32_t* __restrict__ e, int ewords, int wbits) {
  constexpr int K = Geo52<L, TPI>::K;
  for (int j = 0; j < nwin; j++) {
    const uint32_t v = get_bits(e, ewords, j * wbits, wbits);
    mmd<L, TPI>(m, sm_group, acc, acc, table + ((static_cast<size_t>(j) << wbits) + v) * K);
  }
}

}  // namespace scb
