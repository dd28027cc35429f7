// bn52.cuh -- the second big-integer multiplier of the engine: Montgomery arithmetic on the FP64 pipe.
//
// Why.  The integer engine (bn29.cuh) keeps the IMAD.WIDE pipe ~87 % busy, which is all that pipe gives: 32
// 32x32-bit multiply-adds per clock per SM.  The FP64 pipe of a B200 SM retires 64 DFMA per clock and is idle.
// A DFMA on exact integers < 2^52 is a 52x52-bit multiplier: with the classic anchor trick (Emmart, Zheng, Weems,
// ARITH 2018)
//      hi = fma_rz(a, b, 2^104)                  = 2^104 + floor(a b / 2^52) 2^52      (ulp of [2^104, 2^105) is 2^52)
//      lo = fma_rz(a, b, (2^104 + 2^52) - hi)    = 2^52  + (a b mod 2^52)              (exact)
// both halves of the 104-bit product appear, exactly, in the MANTISSA bits of two doubles, and mantissa bit patterns
// add as 64-bit integers (IADD3 on the ALU pipe; the known exponent constants are subtracted once per column).
// Three FP64 operations (DFMA, DADD, DFMA) buy 52 x 52 bits of product where one IMAD.WIDE buys 29 x 29 in the
// carry-free radix-2^29 scheme: per clock and SM that is (64 / 3) * 2704 = 57.7 k bit^2 against 32 * 841 = 26.9 k.
//
// Representation.  K = L * TPI digits of 52 bits, stored as doubles holding exact integers in [0, 2^52); TPI adjacent
// lanes of a warp own L consecutive digits each; Montgomery radix R = 2^(52 K) > 8 n.  Values are lazy in [0, 2n) but
// their DIGITS are always fully normalised (< 2^52) -- the anchor trick needs a b < 2^104 -- so every product ends with an
// exact carry resolution across the lanes.  Nothing here is approximate: all arithmetic is on integers that the
// floating-point formats represent exactly, and the CUDA intrinsics fix the rounding mode of every operation.
//
// Row loop (CIOS by rows, like bn29.cuh): per row j and per lane
//      AL[i] += lo(a_i b_j), AH[i] += hi(a_i b_j)            L times   (AL/AH: 64-bit integer accumulators of column i)
//      q = lo(AL[0] * n') from lane 0, one double SHFL
//      AL[i] += lo(n_i q),   AH[i] += hi(n_i q)              L times
//      AL[1] += AH[0] + (AL[0] >> 52); the low 52 bits of AL[0] (zero in lane 0) go to the lane below as its new top column
// with the column shift done by register renaming (row loop unrolled L times).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "bn29.cuh"

namespace scb {

constexpr int MAXK52 = 160;                               // digits of the largest supported modulus (8192 bits)
constexpr uint64_t MASK52 = (1ull << 52) - 1ull;
constexpr uint64_t BITS_2P52 = 0x4330000000000000ull;     // bit pattern of 2^52
constexpr uint64_t BITS_2P104 = 0x4670000000000000ull;    // bit pattern of 2^104
#define SCB_2P52 4503599627370496.0
#define SCB_2P104 20282409603651670423947251286016.0

// Montgomery constants of one modulus for the FP64 engine (global memory; written once by scb_modulus_load).
struct Modulus52Dev {
  int K;                 // digits (L*TPI of the kernel configuration), 0 = this modulus has no FP64 configuration
  int k32;               // 32-bit words of the modulus at the C ABI
  double np;             // -n^-1 mod 2^52
  double n[MAXK52];      // digits of n
  double one[MAXK52];    // R mod n
  double r2[MAXK52];     // R^2 mod n
  double r3[MAXK52];     // R^3 mod n
  double unit[MAXK52];   // the integer 1
};

template <int L, int TPI>
struct Geo52 {
  static_assert(32 % TPI == 0, "TPI must divide the warp");
  static constexpr int K = L * TPI;
  static constexpr int GPW = 32 / TPI;
  static constexpr int WARP_SMEM_DOUBLES = (K + 1) * GPW;   // +1 row: the row loop prefetches b[j+1]
};

__device__ __forceinline__ double u52_to_double(uint64_t v) { return __longlong_as_double(static_cast<long long>(v | BITS_2P52)) - SCB_2P52; }
__device__ __forceinline__ uint64_t double_to_u52(double d) {
  return static_cast<uint64_t>(__double_as_longlong(d + SCB_2P52)) & MASK52;
}
__device__ __forceinline__ uint64_t raw(double d) { return static_cast<uint64_t>(__double_as_longlong(d)); }

// Exact carry resolution: column values col[p] (true integers, < 2^63, weight 2^(52 (tL + p))) -> digits < 2^52.
// All 32 lanes of the warp must call this together.
template <int L, int TPI>
__device__ __forceinline__ void resolve_carries(uint64_t (&col)[L], uint64_t hi_of_last, double (&r)[L]) {
  static_assert(L >= 2, "resolve_carries: a lane holds at least two digits");
  const int t = (threadIdx.x & 31) % TPI;
  // within the lane: d[p] = low 52 bits, carry on; `hi_of_last` joins the carry leaving the lane (weight 2^(52 L))
  uint64_t c = 0;
#pragma unroll
  for (int p = 0; p < L; p++) {
    c += col[p];
    col[p] = c & MASK52;
    c >>= 52;
  }
  c += hi_of_last;
  // the lane's carry-out (up to ~2^58) enters the next lane at digits 0 and 1
  uint64_t cin = __shfl_up_sync(FULL, c, 1, TPI);
  if (t == 0) cin = 0;
  uint64_t s = col[0] + (cin & MASK52);
  col[0] = s & MASK52;
  s = col[1] + (s >> 52) + (cin >> 52);
  col[1] = s & MASK52;
  uint64_t k = s >> 52;
  // What is left is a carry of 0 or 1 out of digit 1: it moves on only through digits that are all ones (about 2^-46 per
  // lane and product on random data), so the common case ends here; the loop ripples it through the rest of the lane, then
  // lane by lane from digit 0 (warp-uniform trip count; driven by test_fp64_engine_carry_resolution_rare_paths).
  // Measured at B = 65,536 against an unconditional second sweep: 1.0-1.5 % on every exponentiation kernel; a carry-save
  // first step (all columns split at once, depth independent of L) measured 0.5-2 % slower than this form.  (The plain
  // 1024-bit launches of 3 x 65,536 elements are the exception when timed ALONE -- 54.2 ms with the full sweep, 56.6 ms with
  // this form -- but inside a pass, where the other CRT chain runs next to them, this form wins: 988.6 against 990.5 ms per
  // pass; profiles/r02_resolve_plain.log.)
  int from = 2;
  while (__any_sync(FULL, k != 0)) {
#pragma unroll
    for (int p = 0; p < L; p++) {
      if (p >= from) {
        s = col[p] + k;
        col[p] = s & MASK52;
        k = s >> 52;
      }
    }
    uint64_t kin = __shfl_up_sync(FULL, k, 1, TPI);
    if (t == 0) kin = 0;
    k = kin;
    from = 0;
  }
#pragma unroll
  for (int p = 0; p < L; p++) r[p] = u52_to_double(col[p]);
}

// ---- the row loop -------------------------------------------------------------------------------------------------
constexpr int PASS_SAVE = 1;     // pass 1 of the pair arithmetic: its quotient digits are what pass 2 reads (every PAIR pass stores them)
constexpr int PASS_INJECT = 2;   // start the accumulator from R - m, m = the stored quotient digits (pair arithmetic, pass 2)

// The row loop of one Montgomery product on register operands, returning the UNRESOLVED columns:
//   sum_p col[p] 2^(52 p) + hi_last 2^(52 L)  (per lane)  =  this lane's share of (a * b [+ R - m] + q n) / R
//   av[L], nv[L] : this lane's digits of a and of the modulus; b is ALREADY staged in shared memory (digit j at
//   sm_stream[j * GPW], made visible with __syncwarp() by stage_digits52(); the pad row K may hold anything).
//   PAIR: the two hooks of the square-modulus arithmetic (kernels_f64_pair.cuh); flags is warp-uniform.
// All 32 lanes of the warp must call this together.
//
// What bounds it (tools/rowloop_microbench.cu, profiles/r02_rowloop_microbench.jsonl): the FP64 pipe's issue rate for this
// instruction mix, not latency -- 211-215 cycles per warp-row and scheduler whether 1, 2, 3 or 4 blocks are resident per SM.
// Breaking the serial quotient chain on purpose buys 9 %, dropping the shuffles 4 %, replacing the 64-bit adds nothing; a
// software-pipelined variant of this loop (a_0 b_{j+1} one row early, products of the previous row's quotient digit inside
// the next chain, accumulation delayed by one product pair) measured the same 215 cycles, and a fused sweep that runs both
// passes of a pair squaring at once (two accumulator sets, 244 registers, a 40 KB unrolled body) measured 243 cycles per
// row-equivalent against 228 for two separate passes (profiles/r02_rowloop_microbench_dual.jsonl): the plain form stays.
template <int L, int TPI, bool PAIR>
__device__ __forceinline__ void mmd_rows_core(const double (&av)[L], const double (&nv)[L], const double np, const double* __restrict__ sm_stream,
                                              uint64_t* __restrict__ sm_q, const int flags, uint64_t (&col)[L], uint64_t& hi_last) {
  constexpr int GPW = Geo52<L, TPI>::GPW;
  const int t = (threadIdx.x & 31) % TPI;
  const bool top = t == TPI - 1;
  const bool low_lane = t == 0;
  const bool inject = PAIR && (flags & PASS_INJECT) != 0;
  const double C1 = SCB_2P104, C12 = SCB_2P104 + SCB_2P52;
  // Exponent bookkeeping: every raw hi carries BITS_2P104 and every raw lo BITS_2P52 on top of its mantissa.  A column
  // receives exactly 2 of each per row while it is in this lane, so the total is subtracted up front (mod 2^64): the
  // column at position i of the first row lives i + 1 rows, a column entering at the top lives L rows.
  uint64_t AL[L], AH[L];
#pragma unroll
  for (int i = 0; i < L; i++) {
    AL[i] = 0ull - 2ull * (i + 1) * BITS_2P52;
    AH[i] = 0ull - 2ull * (i + 1) * BITS_2P104;
  }
  if (PAIR && inject) {
    // R - m = complement(m) + 1 enters as the accumulator's initial value: digit j of the complement in column j (it is
    // summed into the quotient digit of row j like every product that lands there).  L conversions per lane and pass
    // instead of one load + add in every row.
    __syncwarp();
#pragma unroll
    for (int i = 0; i < L; i++) AL[i] += double_to_u52(__longlong_as_double(static_cast<long long>(sm_q[(t * L + i) * GPW]))) ^ MASK52;
    if (low_lane) AL[0] += 1ull;
    __syncwarp();
  }
  const double* bp = sm_stream;
  uint64_t* qp = sm_q;
  double bj = bp[0];
#pragma unroll 1
  for (int blk = 0; blk < TPI; blk++) {
#pragma unroll
    for (int rr = 0; rr < L; rr++) {
      // logical column i of this row lives in AL/AH[(i + rr) % L]
      bp += GPW;
      const double bnext = bp[0];                  // the last row reads the pad row
      {
        const double hi = __fma_rz(av[0], bj, C1);
        const double lo = __fma_rz(av[0], bj, __dsub_rn(C12, hi));
        AH[rr % L] += raw(hi);
        AL[rr % L] += raw(lo);
      }
      // quotient digit: q = (column 0 mod 2^52) * n' mod 2^52 (the outstanding exponent constants have zero low bits)
      const double v0 = u52_to_double(AL[rr % L] & MASK52);
      const double qh = __fma_rz(v0, np, C1);
      double q = __fma_rz(v0, np, __dsub_rn(C1, qh));
#ifndef SCB_EXP_NO_SHFL
      q = __shfl_sync(FULL, q, 0, TPI);
#endif
#ifdef SCB_EXP_NO_CHAIN
      q = bnext;                                   /* experiment (tools/rowloop_microbench.cu): no serial quotient chain */
#endif
      if (PAIR) {
        // every pass stores its quotient digits (as the doubles they are; nobody reads what passes 2 and 3 leave), so the
        // loop has no flag-dependent branch: with `if (save)` ptxas put real branches between the rows and the launch of
        // the B randomizers went from 333 to 344 ms
        if (low_lane) qp[0] = raw(q);
        qp += GPW;
      }
#pragma unroll
      for (int i = 1; i < L; i++) {
        const double hi = __fma_rz(av[i], bj, C1);
        const double lo = __fma_rz(av[i], bj, __dsub_rn(C12, hi));
#ifndef SCB_EXP_NO_IADD
        AH[(i + rr) % L] += raw(hi);
        AL[(i + rr) % L] += raw(lo);
#else
        AL[(i + rr) % L] ^= raw(hi) ^ raw(lo);     /* experiment: logic ops instead of the 64-bit adds */
#endif
      }
#pragma unroll
      for (int i = 0; i < L; i++) {
        const double hi = __fma_rz(nv[i], q, C1);
        const double lo = __fma_rz(nv[i], q, __dsub_rn(C12, hi));
#ifndef SCB_EXP_NO_IADD
        AH[(i + rr) % L] += raw(hi);
        AL[(i + rr) % L] += raw(lo);
#else
        AL[(i + rr) % L] ^= raw(hi) ^ raw(lo);
#endif
      }
      // column 0 is complete (all its exponent constants are cancelled): push its carry up, hand its low part down
      const uint64_t t0 = AL[rr % L];
      AL[(1 + rr) % L] += AH[rr % L] + (t0 >> 52);
      const uint64_t low = t0 & MASK52;             // 0 in lane 0 (Montgomery)
#ifndef SCB_EXP_NO_SHFL
      const uint64_t in = __shfl_down_sync(FULL, low, 1, TPI);
#else
      const uint64_t in = low;
#endif
      AL[rr % L] = (top ? 0ull : in) - 2ull * L * BITS_2P52;    // becomes logical column L-1 of the next row
      AH[rr % L] = 0ull - 2ull * L * BITS_2P104;
      bj = bnext;
    }
  }
  // after TPI*L rows the renaming is back to the identity.  The column at final position p lived only L-1-p rows:
  // give back the exponent constants of the p+1 rows it did not see.
#pragma unroll
  for (int p = 0; p < L; p++) {
    AL[p] += 2ull * (p + 1) * BITS_2P52;
    AH[p] += 2ull * (p + 1) * BITS_2P104;
  }
  // fold the high accumulators one column up
  col[0] = AL[0];
#pragma unroll
  for (int p = 1; p < L; p++) col[p] = AL[p] + AH[p - 1];
  hi_last = AH[L - 1];
}

// r = a * b * R^-1 mod n (value < 2n, digits < 2^52); r may be the same array as av.
template <int L, int TPI>
__device__ __forceinline__ void mmd_rows(const double (&av)[L], const double (&nv)[L], const double np, const double* __restrict__ sm_group,
                                         double (&r)[L]) {
  uint64_t col[L], hi_last;
  mmd_rows_core<L, TPI, false>(av, nv, np, sm_group, nullptr, 0, col, hi_last);
  resolve_carries<L, TPI>(col, hi_last, r);
}

// Stage this lane's L digits into the group's shared-memory operand area and make them visible.
template <int L, int TPI>
__device__ __forceinline__ void stage_digits52(double* sm_group, const double (&bv)[L]) {
  constexpr int GPW = Geo52<L, TPI>::GPW;
  const int t = (threadIdx.x & 31) % TPI;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < L; i++) sm_group[(t * L + i) * GPW] = bv[i];
  __syncwarp();
}

// memory-to-memory product for the few places outside the exponentiation loop (conversions): dst = a * b * R^-1 mod n
template <int L, int TPI>
__device__ __noinline__ void mmd(const Modulus52Dev* __restrict__ m, double* sm_group, double* dst, const double* a, const double* b) {
  const int t = (threadIdx.x & 31) % TPI;
  double av[L], nv[L];
  {
    double bv[L];
#pragma unroll
    for (int i = 0; i < L; i++) bv[i] = b[t * L + i];
#pragma unroll
    for (int i = 0; i < L; i++) av[i] = a[t * L + i];
    stage_digits52<L, TPI>(sm_group, bv);
  }
#pragma unroll
  for (int i = 0; i < L; i++) nv[i] = m->n[t * L + i];
  double r[L];
  mmd_rows<L, TPI>(av, nv, m->np, sm_group, r);
#pragma unroll
  for (int i = 0; i < L; i++) dst[t * L + i] = r[i];
}

// dst = a + b with digits re-normalised (values < 4n stay far below R)
template <int L, int TPI>
__device__ __forceinline__ void add52(double* dst, const double* a, const double* b) {
  const int t = (threadIdx.x & 31) % TPI;
  uint64_t col[L];
  double r[L];
#pragma unroll
  for (int i = 0; i < L; i++) col[i] = double_to_u52(a[t * L + i]) + double_to_u52(b[t * L + i]);
  resolve_carries<L, TPI>(col, 0ull, r);
#pragma unroll
  for (int i = 0; i < L; i++) dst[t * L + i] = r[i];
}

// digit d (52 bits) of the number stored in words[0..kw)
__device__ __forceinline__ uint64_t digit52_from_words(const uint32_t* __restrict__ words, int kw, int d) {
  const int bit = d * 52;
  const int w = bit >> 5, sh = bit & 31;
  const uint64_t w0 = (w < kw) ? __ldg(words + w) : 0u;
  const uint64_t w1 = (w + 1 < kw) ? __ldg(words + w + 1) : 0u;
  const uint64_t w2 = (w + 2 < kw) ? __ldg(words + w + 2) : 0u;
  uint64_t v = (w0 | (w1 << 32)) >> sh;
  if (sh) v |= w2 << (64 - sh);
  return v & MASK52;
}

}  // namespace scb
