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
  static constexpr int WARP_SMEM_DOUBLES = (K + 2) * GPW;   // +2 zero rows: the row loop multiplies by b[j+1] and loads b[j+2] one row ahead
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
  uint64_t k = (s >> 52) + (cin >> 52);
#pragma unroll
  for (int p = 1; p < L; p++) {
    s = col[p] + k;
    col[p] = s & MASK52;
    k = s >> 52;
  }
  // second-order carries (0 or 1 leaving a lane): needs another round only when a lane's digits are all ones
  while (__any_sync(FULL, k != 0)) {
    uint64_t kin = __shfl_up_sync(FULL, k, 1, TPI);
    if (t == 0) kin = 0;
    k = kin;
#pragma unroll
    for (int p = 0; p < L; p++) {
      s = col[p] + k;
      col[p] = s & MASK52;
      k = s >> 52;
    }
  }
#pragma unroll
  for (int p = 0; p < L; p++) r[p] = u52_to_double(col[p]);
}

// ---- the row loop -------------------------------------------------------------------------------------------------
// One exact 52 x 52 -> 104-bit product in the mantissas of two doubles, added as raw bit patterns into a column pair.
#define SCB_PROD_HI(x, y) __fma_rz((x), (y), C1)
#define SCB_PROD_LO(x, y, hi) __fma_rz((x), (y), __dsub_rn(C12, (hi)))

constexpr int PASS_SAVE = 1;     // store the complemented quotient digits of this pass (pair arithmetic, pass 1)
constexpr int PASS_INJECT = 2;   // add the stored complement (+1) into the low column of every row (pair arithmetic, pass 2)

// The row loop of one Montgomery product on register operands, returning the UNRESOLVED columns:
//   sum_p col[p] 2^(52 p) + hi_last 2^(52 L)  (per lane)  =  this lane's share of (a * b [+ R - m] + q n) / R
//   av[L], nv[L] : this lane's digits of a and of the modulus; b is ALREADY staged in shared memory (digit j at
//   sm_stream[j * GPW], made visible with __syncwarp() by stage_digits52()); rows K and K+1 of the area must hold 0.0
//   (zero_pad_rows52(), once per kernel) because the loop multiplies by b[j+1] and loads b[j+2] one row ahead.
//   PAIR: the two hooks of the square-modulus arithmetic (kernels_f64_pair.cuh); flags is warp-uniform.
//
// Schedule (model with the same data flow and constants: tests/mmd52_model.py).  The quotient digit is a serial chain
// through every row -- column 0 -> q = lo(col0 * n') -> SHFL -> n_0 q, n_1 q -> carry into column 1 = the next column 0 --
// about a dozen dependent operations, and the two passes of a warp over an unrolled row used to run it with nothing else
// in flight (ncu: "wait" stalls on exactly those instructions, FP64 pipe 47-62 % busy).  Here the chain is interleaved, in
// SOURCE order (ptxas keeps it), with the work that does not depend on it: the products n_i q of the PREVIOUS row for
// i >= 2, the products a_i b_j of this row, and a_0 b_{j+1}, which is computed one row early so that column 1 is complete
// -- and the next chain can start -- as soon as n_0 q and n_1 q have landed.
// All 32 lanes of the warp must call this together.
template <int L, int TPI, bool PAIR>
__device__ __forceinline__ void mmd_rows_core(const double (&av)[L], const double (&nv)[L], const double np, const double* __restrict__ sm_stream,
                                              uint64_t* __restrict__ sm_q, const int flags, uint64_t (&col)[L], uint64_t& hi_last) {
  static_assert(L >= 6, "the interleaved schedule places products 1..5 by hand");
  constexpr int GPW = Geo52<L, TPI>::GPW;
  const int t = (threadIdx.x & 31) % TPI;
  const bool top = t == TPI - 1;
  const bool low_lane = t == 0;
  const bool save = PAIR && (flags & PASS_SAVE) != 0, inject = PAIR && (flags & PASS_INJECT) != 0;
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
  const double* bp = sm_stream;
  const uint64_t* qin = sm_q;
  uint64_t* qout = sm_q;
  double b0 = bp[0], b1 = bp[GPW];
  {
    // prologue: a_0 b_0 (every later row finds its a_0 b_j already in place), first complement digit + 1
    const double hi = SCB_PROD_HI(av[0], b0);
    const double lo = SCB_PROD_LO(av[0], b0, hi);
    AH[0] += raw(hi);
    AL[0] += raw(lo);
    if (PAIR && inject && low_lane) AL[0] += qin[0] + 1ull;
  }
  double qprev = 0.0;                                     // quotient digit of the previous row (its n_i q for i >= 2 are still owed)
#pragma unroll 1
  for (int blk = 0; blk < TPI; blk++) {
#pragma unroll
    for (int rr = 0; rr < L; rr++) {
      // logical column i of this row lives in AL/AH[(i + rr) % L]
#define COL(i) (((i) + rr) % L)
#define OWED(i) (((i) + rr - 1 + L) % L)                  /* column i of the PREVIOUS row */
      bp += GPW;
      const double b2 = bp[GPW];                           // b[j+2]
      uint64_t cnext = 0ull;
      if (PAIR && inject) cnext = qin[GPW];                // complement digit j+1 (row K of the area holds 0)
      qin += GPW;
      // ---- chain A, step 1: column 0 -> double
      const double v0 = u52_to_double(AL[COL(0)] & MASK52);
      // owed from the previous row: n_i qprev, i = 2, 3 (none in the first row of a block: they were paid before the back-edge)
      if (rr > 0) {
        const double h2 = SCB_PROD_HI(nv[2], qprev), h3 = SCB_PROD_HI(nv[3], qprev);
        const double l2 = SCB_PROD_LO(nv[2], qprev, h2), l3 = SCB_PROD_LO(nv[3], qprev, h3);
        AH[OWED(2)] += raw(h2);
        AL[OWED(2)] += raw(l2);
        AH[OWED(3)] += raw(h3);
        AL[OWED(3)] += raw(l3);
      }
      // ---- chain A, step 2
      const double qh = __fma_rz(v0, np, C1);
      if (rr > 0) {
#pragma unroll
        for (int i = 4; i < L; i += 2) {
          const double ha = SCB_PROD_HI(nv[i], qprev), hb = (i + 1 < L) ? SCB_PROD_HI(nv[(i + 1) % L], qprev) : 0.0;
          const double la = SCB_PROD_LO(nv[i], qprev, ha);
          AH[OWED(i)] += raw(ha);
          AL[OWED(i)] += raw(la);
          if (i + 1 < L) {
            const double lb = SCB_PROD_LO(nv[(i + 1) % L], qprev, hb);
            AH[OWED(i + 1)] += raw(hb);
            AL[OWED(i + 1)] += raw(lb);
          }
        }
      }
      // ---- chain A, step 3
      const double qs = __dsub_rn(C1, qh);
      {
        // a_1 b_j, a_2 b_j
        const double h1 = SCB_PROD_HI(av[1], b0), h2 = SCB_PROD_HI(av[2], b0);
        const double l1 = SCB_PROD_LO(av[1], b0, h1), l2 = SCB_PROD_LO(av[2], b0, h2);
        AH[COL(1)] += raw(h1);
        AL[COL(1)] += raw(l1);
        AH[COL(2)] += raw(h2);
        AL[COL(2)] += raw(l2);
      }
      // ---- chain A, step 4: q = lo(v0 * n')
      double q = __fma_rz(v0, np, qs);
      {
        // a_3 b_j, a_4 b_j, and a_0 b_{j+1} one row early (column 1 is column 0 of the next row)
        const double h3 = SCB_PROD_HI(av[3], b0), h4 = SCB_PROD_HI(av[4], b0), hn = SCB_PROD_HI(av[0], b1);
        const double l3 = SCB_PROD_LO(av[3], b0, h3), l4 = SCB_PROD_LO(av[4], b0, h4), ln = SCB_PROD_LO(av[0], b1, hn);
        AH[COL(3)] += raw(h3);
        AL[COL(3)] += raw(l3);
        AH[COL(4)] += raw(h4);
        AL[COL(4)] += raw(l4);
        AH[COL(1)] += raw(hn);
        AL[COL(1)] += raw(ln);
      }
      // ---- chain A, step 5: broadcast from the lane that owns column 0 of the number
      q = __shfl_sync(FULL, q, 0, TPI);
      if (PAIR && inject && low_lane) AL[COL(1)] += cnext;
#pragma unroll
      for (int i = 5; i < L; i += 2) {
        const double ha = SCB_PROD_HI(av[i], b0), hb = (i + 1 < L) ? SCB_PROD_HI(av[(i + 1) % L], b0) : 0.0;
        const double la = SCB_PROD_LO(av[i], b0, ha);
        AH[COL(i)] += raw(ha);
        AL[COL(i)] += raw(la);
        if (i + 1 < L) {
          const double lb = SCB_PROD_LO(av[(i + 1) % L], b0, hb);
          AH[COL(i + 1)] += raw(hb);
          AL[COL(i + 1)] += raw(lb);
        }
      }
      if (PAIR && save) {
        if (low_lane) qout[0] = double_to_u52(q) ^ MASK52;
      }
      qout += GPW;
      // ---- chain B: n_0 q, n_1 q complete columns 0 and 1
      {
        const double h0 = SCB_PROD_HI(nv[0], q), h1 = SCB_PROD_HI(nv[1], q);
        const double l0 = SCB_PROD_LO(nv[0], q, h0), l1 = SCB_PROD_LO(nv[1], q, h1);
        AH[COL(0)] += raw(h0);
        AL[COL(0)] += raw(l0);
        AH[COL(1)] += raw(h1);
        AL[COL(1)] += raw(l1);
      }
      // column 0 is complete (all its exponent constants are cancelled): push its carry up, hand its low part down
      const uint64_t t0 = AL[COL(0)];
      AL[COL(1)] += AH[COL(0)] + (t0 >> 52);
      const uint64_t low = t0 & MASK52;             // 0 in lane 0 (Montgomery)
      const uint64_t in = __shfl_down_sync(FULL, low, 1, TPI);
      AL[COL(0)] = (top ? 0ull : in) - 2ull * L * BITS_2P52;    // becomes logical column L-1 of the next row
      AH[COL(0)] = 0ull - 2ull * L * BITS_2P104;
      if (rr == L - 1) {
        // last row of the block: pay n_i q, i >= 2, before the back-edge
#pragma unroll
        for (int i = 2; i < L; i += 2) {
          const double ha = SCB_PROD_HI(nv[i], q), hb = (i + 1 < L) ? SCB_PROD_HI(nv[(i + 1) % L], q) : 0.0;
          const double la = SCB_PROD_LO(nv[i], q, ha);
          AH[COL(i)] += raw(ha);
          AL[COL(i)] += raw(la);
          if (i + 1 < L) {
            const double lb = SCB_PROD_LO(nv[(i + 1) % L], q, hb);
            AH[COL(i + 1)] += raw(hb);
            AL[COL(i + 1)] += raw(lb);
          }
        }
      }
      qprev = q;
      b0 = b1;
      b1 = b2;
#undef COL
#undef OWED
    }
  }
  // after TPI*L rows the renaming is back to the identity.  The column at final position p lived only L-1-p rows: give
  // back the exponent constants of the p+1 rows it did not see; position 0 also took the (zero) product a_0 b_K early.
#pragma unroll
  for (int p = 0; p < L; p++) {
    AL[p] += 2ull * (p + 1) * BITS_2P52 - (p == 0 ? BITS_2P52 : 0ull);
    AH[p] += 2ull * (p + 1) * BITS_2P104 - (p == 0 ? BITS_2P104 : 0ull);
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

// rows K and K+1 of a staging area must read as 0.0 (see mmd_rows_core); stage_digits52 never touches them
template <int L, int TPI>
__device__ __forceinline__ void zero_pad_rows52(double* sm_area) {
  constexpr int GPW = Geo52<L, TPI>::GPW;
  if ((threadIdx.x & 31) % TPI == 0) {
    sm_area[Geo52<L, TPI>::K * GPW] = 0.0;
    sm_area[(Geo52<L, TPI>::K + 1) * GPW] = 0.0;
  }
  __syncwarp();
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
