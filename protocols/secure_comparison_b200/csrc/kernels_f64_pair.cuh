// kernels_f64_pair.cuh -- exponentiation modulo a SQUARE n^2 on the FP64 engine, in base-n ("pair") form.
//
// Every heavy exponentiation of the protocol has a square modulus: rho^N mod N^2 (Initiator's randomizer,
// initiator.py:109), b^p mod p^2 and c^(p-1) mod p^2 (the KeyHolder's CRT halves of randomize / decrypt,
// keyholder.py:126-128, 195).  A residue x mod n^2 is held as a pair (a, b), x = a + b n, both halves K-digit numbers
// (K = digits of the ROOT n) kept lazily below 4n, and the Montgomery product for the root's radix R = 2^(52 K)
//
//      (a, b) (x) (c, d) = (a + b n)(c + d n) R^-1  mod n^2
//
// is three K-digit Montgomery passes -- two for a squaring -- instead of one 2K-digit product (= four K-digit ones):
// with pass 1  a c + m n = t R  (m: the K quotient digits of that reduction)
//
//      x y      = a c + (a d + b c) n  =  t R + (a d + b c - m) n                         (mod n^2)
//      x y R^-1 = t + [ REDC(a d + (R - m)) + REDC(b c) + (n - 1) ] n                     (mod n^2)
//
// R - m is injected digit by digit as the complement of m (plus one), so no pass sees a negative number; the "- 1"
// that compensates it is folded into "+ (n - 1)".  Squaring: t = REDC(a a), b' = REDC(a (2b) + (R - m)) + (n - 1).
// Digit products per squaring: 4 K^2 against 8 K^2 for the plain Montgomery product on the 2K-digit modulus n^2.
// Model and bounds (a < 2n, b < 4n closed under both operations when R >= 16 n): tests/pairmul_model.py.
//
// Data flow of the kernel is k_modexp52's: one group of TPI lanes per exponentiation, running value and modulus in
// registers, window table in the group's global scratch, dynamic work queue; inputs and outputs are the 32-bit-word
// arrays of the C ABI and outputs are canonical residues in [0, n^2).
#pragma once
#include "kernels_f64.cuh"

namespace scb {

// Per-square-modulus constants (global memory; written by link_square_modulus() on the host): pairs in base n.
struct PairConsts52Dev {
  int K;                 // digits of the root
  int w;                 // piece width of the input conversion: bits(n) - 1, so that 2^w <= n
  double oneA[MAXK52], oneB[MAXK52];   // R mod n^2            (Montgomery form of 1)
  double r2A[MAXK52], r2B[MAXK52];     // R^2 mod n^2          (x -> x R)
  double sA[MAXK52], sB[MAXK52];       // 2^w R mod n^2        (Montgomery form of the Horner radix)
};

template <int L, int TPI>
struct GeoPair {
  static constexpr int K = L * TPI;
  static constexpr int GPW = 32 / TPI;
  static constexpr int S2_OFF = (K + 1) * GPW;           // second streamed-operand area (each: K digits + the prefetch pad row)
  static constexpr int Q_OFF = 2 * (K + 1) * GPW;        // quotient digits of pass 1 (K doubles, stored as their bit patterns)
  static constexpr int WARP_SMEM_WORDS64 = (3 * K + 2) * GPW;
};

// Measured and not kept (profiles/r02u_dual.log): a window product in TWO passes, the second one accumulating a d + b c +
// (R - m) with a single reduction (5 K^2 digit products instead of 6 K^2; tests/pairmul_model.py, mul_dual, bounds hold).
// Bit-identical, no spills -- and 20 % SLOWER on both launches (333 -> 399 ms, 91 -> 111 ms) although only a fifth of the
// passes changed.  The likely cause is code size: a second unrolled row loop (28 KB next to the 20 KB of the first) in a
// main loop whose hot code is already ~35-40 KB; k52_step4_main showed the same sensitivity in the other direction
// (kernels_dgk52.cuh, group_modexp52_mem).  Not confirmed with an instruction-cache counter.
//
// The row loop with the two hooks of the pair arithmetic: bn52.cuh, mmd_rows_core<L, TPI, true>.
template <int L, int TPI>
__device__ __forceinline__ void mmd_rows_pair(const double (&av)[L], const double (&nv)[L], const double np, const double* __restrict__ sm_stream,
                                              uint64_t* __restrict__ sm_q, const int flags, uint64_t (&col)[L], uint64_t& hi_last) {
  mmd_rows_core<L, TPI, true>(av, nv, np, sm_stream, sm_q, flags, col, hi_last);
}

// col += the lane's digits of (n - 1)   (n is odd: only digit 0 of the whole number changes)
template <int L, int TPI>
__device__ __forceinline__ void add_n_minus_1(uint64_t (&col)[L], const double (&nv)[L]) {
#pragma unroll
  for (int p = 0; p < L; p++) col[p] += double_to_u52(nv[p]);
  if ((threadIdx.x & 31) % TPI == 0) col[0] -= 1ull;
}

template <int L, int TPI>
__device__ __forceinline__ void add_digits(uint64_t (&col)[L], const double (&d)[L]) {
#pragma unroll
  for (int p = 0; p < L; p++) col[p] += double_to_u52(d[p]);
}

// Memory-to-memory pair product for the cold paths (conversions): (dA, dB) = (xA, xB) (x) (yA, yB).  Arrays are K-digit
// slots (each lane touches its own digits); dst may alias x.  One copy of the row loop per kernel (not inlined).
template <int L, int TPI>
__device__ __noinline__ void pair_mul_mem(const Modulus52Dev* __restrict__ m, double* sm_group, double* dA, double* dB, const double* xA,
                                          const double* xB, const double* yA, const double* yB) {
  using GP = GeoPair<L, TPI>;
  const int t = (threadIdx.x & 31) % TPI;
  double* s1 = sm_group;
  double* s2 = sm_group + GP::S2_OFF;
  uint64_t* sq = reinterpret_cast<uint64_t*>(sm_group + GP::Q_OFF);
  double av[L], nv[L], tn[L], s1v[L];
  {
    double cv[L];
#pragma unroll
    for (int i = 0; i < L; i++) cv[i] = yA[t * L + i];
    stage_digits52<L, TPI>(s1, cv);
#pragma unroll
    for (int i = 0; i < L; i++) cv[i] = yB[t * L + i];
    stage_digits52<L, TPI>(s2, cv);
  }
#pragma unroll
  for (int i = 0; i < L; i++) nv[i] = m->n[t * L + i];
#pragma unroll
  for (int i = 0; i < L; i++) av[i] = xA[t * L + i];
  const double np = m->np;
#pragma unroll 1
  for (int pass = 0; pass < 3; pass++) {
    if (pass == 2) {
#pragma unroll
      for (int i = 0; i < L; i++) av[i] = xB[t * L + i];
    }
    uint64_t col[L], hi;
    mmd_rows_pair<L, TPI>(av, nv, np, pass == 1 ? s2 : s1, sq, pass == 0 ? PASS_SAVE : (pass == 1 ? PASS_INJECT : 0), col, hi);
    if (pass == 0) {
      resolve_carries<L, TPI>(col, hi, tn);
    } else if (pass == 1) {
      add_n_minus_1<L, TPI>(col, nv);
      resolve_carries<L, TPI>(col, hi, s1v);
    } else {
      add_digits<L, TPI>(col, s1v);
      resolve_carries<L, TPI>(col, hi, s1v);
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < L; i++) dA[t * L + i] = tn[i];
#pragma unroll
  for (int i = 0; i < L; i++) dB[t * L + i] = s1v[i];
}

// 52-bit digit d of the `width`-bit piece starting at bit `pos` of words[0..kw)
__device__ __forceinline__ uint64_t piece_digit52(const uint32_t* __restrict__ words, int kw, int pos, int width, int d) {
  const int left = width - 52 * d;
  if (left <= 0) return 0ull;
  const int bit = pos + 52 * d;
  const int w = bit >> 5, sh = bit & 31;
  const uint64_t w0 = (w < kw) ? __ldg(words + w) : 0u;
  const uint64_t w1 = (w + 1 < kw) ? __ldg(words + w + 1) : 0u;
  const uint64_t w2 = (w + 2 < kw) ? __ldg(words + w + 2) : 0u;
  uint64_t v = (w0 | (w1 << 32)) >> sh;
  if (sh) v |= w2 << (64 - sh);
  v &= MASK52;
  if (left < 52) v &= (1ull << left) - 1ull;
  return v;
}

// words (kw of them, any value) -> Montgomery pair of (value mod n^2) at (dA, dB): Horner over w-bit pieces, each < 2^w <= n
template <int L, int TPI>
__device__ __forceinline__ void load_mont_pair(const Modulus52Dev* __restrict__ m, const PairConsts52Dev* __restrict__ pc, double* sm_group,
                                               double* dA, double* dB, const uint32_t* __restrict__ words, int kw) {
  const int t = (threadIdx.x & 31) % TPI;
  const int w = pc->w;
  const int npieces = (32 * kw + w - 1) / w;             // launch-uniform
#pragma unroll
  for (int i = 0; i < L; i++) {
    dA[t * L + i] = u52_to_double(piece_digit52(words, kw, (npieces - 1) * w, w, t * L + i));
    dB[t * L + i] = 0.0;
  }
  for (int p = npieces - 2; p >= 0; p--) {
    pair_mul_mem<L, TPI>(m, sm_group, dA, dB, dA, dB, pc->sA, pc->sB);          // * 2^w (plain domain)
    uint64_t col[L];
    double r[L];
#pragma unroll
    for (int i = 0; i < L; i++) col[i] = double_to_u52(dA[t * L + i]) + piece_digit52(words, kw, p * w, w, t * L + i);
    resolve_carries<L, TPI>(col, 0ull, r);
#pragma unroll
    for (int i = 0; i < L; i++) dA[t * L + i] = r[i];
  }
  pair_mul_mem<L, TPI>(m, sm_group, dA, dB, dA, dB, pc->r2A, pc->r2B);
}

// Montgomery pair at (xA, xB) -> canonical integer a + b n in [0, n^2) -> kout words.  The final reductions and the
// K x K product b n run serially in lane 0 of the group over the group's global scratch (once per OUTPUT value:
// ~K^2 64-bit multiplies against the ~10^7 lane-instructions of the exponentiation).  tmp: 4K uint64 of scratch.
template <int L, int TPI>
__device__ __noinline__ void pair_store(const Modulus52Dev* __restrict__ m, double* sm_group, double* xA, double* xB, uint64_t* tmp,
                                        uint32_t* __restrict__ words, int kout, bool do_store) {
  constexpr int K = Geo52<L, TPI>::K;
  const int t = (threadIdx.x & 31) % TPI;
  pair_mul_mem<L, TPI>(m, sm_group, xA, xB, xA, xB, m->unit, m->unit + K);   // (1, 0): unit[K..) is zero (MAXK52 >= 2K for every root)
  __syncwarp();
  uint64_t* a = tmp;            // K digits
  uint64_t* b = tmp + K;        // K digits
  uint64_t* o = tmp + 2 * K;    // 2K output digits
  if (t == 0) {
    for (int j = 0; j < K; j++) {
      a[j] = double_to_u52(xA[j]);
      b[j] = double_to_u52(xB[j]);
    }
    // a mod n (a <= n), carry into b; then b mod n (b <= 3n + 1)
    for (int it = 0; it < 6; it++) {
      uint64_t* v = it < 2 ? a : b;
      int cmp = 0;
      for (int j = K - 1; j >= 0; j--) {
        const uint64_t y = double_to_u52(m->n[j]);
        if (v[j] != y) { cmp = v[j] > y ? 1 : -1; break; }
      }
      if (cmp < 0) continue;
      uint64_t borrow = 0;
      for (int j = 0; j < K; j++) {
        const uint64_t y = double_to_u52(m->n[j]) + borrow;
        borrow = v[j] < y ? 1ull : 0ull;
        v[j] = (v[j] - y) & MASK52;
      }
      if (it < 2) {                                         // a wrapped once: b += 1
        uint64_t c = 1;
        for (int j = 0; j < K && c; j++) {
          const uint64_t s = b[j] + c;
          b[j] = s & MASK52;
          c = s >> 52;
        }
      }
    }
    // o = a + b * n   (operand scanning, 104-bit products split at bit 52)
    for (int j = 0; j < K; j++) {
      o[j] = a[j];
      o[K + j] = 0;
    }
    for (int i = 0; i < K; i++) {
      const uint64_t bi = b[i];
      uint64_t carry = 0;
      for (int j = 0; j < K; j++) {
        const uint64_t nj = double_to_u52(m->n[j]);
        const uint64_t lo64 = bi * nj, hi64 = __umul64hi(bi, nj);
        const uint64_t plo = lo64 & MASK52, phi = (hi64 << 12) | (lo64 >> 52);
        const uint64_t s = o[i + j] + plo + carry;
        o[i + j] = s & MASK52;
        carry = (s >> 52) + phi;
      }
      for (int j = i + K; j < 2 * K && carry; j++) {
        const uint64_t s = o[j] + carry;
        o[j] = s & MASK52;
        carry = s >> 52;
      }
    }
  }
  __syncwarp();
  if (do_store) {
    for (int w = t; w < kout; w += TPI) {
      const int bit = w * 32;
      const int d = bit / 52, r = bit - d * 52;
      uint64_t x = 0;
      if (d < 2 * K) x = o[d] >> r;
      if (d + 1 < 2 * K && r > 20) x |= o[d + 1] << (52 - r);
      words[w] = static_cast<uint32_t>(x);
    }
  }
  __syncwarp();
}

template <int L, int TPI>
__host__ __device__ constexpr size_t modexp52p_scratch_doubles(int wbits) {
  // X pair + table of pairs + 4K uint64 for pair_store
  return static_cast<size_t>(Geo52<L, TPI>::K) * (2 + 4 + 2 * (static_cast<size_t>(1) << wbits));
}

// (outA, outB) = x~^e for ONE group.  Shared exponent with a sliding-window program (prog != nullptr; format: kernels_f64.cuh)
// or fixed windows over per-element exponent words.  One software loop, one inlined copy of the row loop: every
// multiplication is a short sequence of passes (squaring: 2, product: 3) and all control flow is launch-uniform --
// only the table INDEX of a window product is per-group data.
template <int L, int TPI>
__device__ __forceinline__ void group_modexp52_pair(const Modulus52Dev* __restrict__ m, const PairConsts52Dev* __restrict__ pc, double* sm_group,
                                                    double* outA, double* outB, const double* xA, const double* xB,
                                                    const uint32_t* __restrict__ e, int ewords, int ebits, int wbits,
                                                    const uint16_t* __restrict__ prog, int nprog, int pw, double* table) {
  using GP = GeoPair<L, TPI>;
  constexpr int K = Geo52<L, TPI>::K;
  const int t = (threadIdx.x & 31) % TPI;
  double* s1 = sm_group;
  double* s2 = sm_group + GP::S2_OFF;
  uint64_t* sq = reinterpret_cast<uint64_t*>(sm_group + GP::Q_OFF);
  const bool sliding = prog != nullptr;
  const int nwin = sliding ? 1 : (ebits + wbits - 1) / wbits;
  if (nwin == 0) {
#pragma unroll
    for (int i = 0; i < L; i++) {
      outA[t * L + i] = pc->oneA[t * L + i];
      outB[t * L + i] = pc->oneB[t * L + i];
    }
    return;
  }
  auto tabA = [&](uint32_t idx) -> double* { return table + static_cast<size_t>(idx) * 2 * K + t * L; };
  auto tabB = [&](uint32_t idx) -> double* { return table + static_cast<size_t>(idx) * 2 * K + K + t * L; };
  double ra[L], rb[L], nv[L], tn[L], s1v[L];
#pragma unroll
  for (int i = 0; i < L; i++) nv[i] = m->n[t * L + i];
  const double np = m->np;
#pragma unroll
  for (int i = 0; i < L; i++) {
    ra[i] = xA[t * L + i];
    rb[i] = xB[t * L + i];
  }
  // table layout -- sliding: table[i] = x~^(2i+1), i < T = 2^(pw-1), built by multiplying with the staged x~^2;
  //                 fixed:   table[v] = x~^v, v < 2^wbits (table[0] = 1~), built by multiplying with the staged x~.
  int first_free;                                        // next table slot the build fills
  int build_left;                                        // products of the table build still to run
  if (sliding) {
#pragma unroll
    for (int i = 0; i < L; i++) {
      tabA(0)[i] = ra[i];
      tabB(0)[i] = rb[i];
    }
    first_free = 1;
    build_left = (1 << (pw - 1)) - 1;
  } else {
#pragma unroll
    for (int i = 0; i < L; i++) {
      tabA(0)[i] = pc->oneA[t * L + i];
      tabB(0)[i] = pc->oneB[t * L + i];
      tabA(1)[i] = ra[i];
      tabB(1)[i] = rb[i];
    }
    first_free = 2;
    build_left = (1 << wbits) - 2;
    stage_digits52<L, TPI>(s1, ra);
    stage_digits52<L, TPI>(s2, rb);
  }
  // phases: 0 = the squaring x~ -> x~^2 that feeds the sliding table build, 1 = table build, 2 = main loop
  int phase = sliding ? 0 : 1;
  const uint32_t idx0 = sliding ? __ldg(prog) : get_bits(e, ewords, (nwin - 1) * wbits, wbits);
  if (phase == 1 && build_left == 0) {
    // fixed windows of one bit: the table is {1~, x~}
#pragma unroll
    for (int i = 0; i < L; i++) {
      ra[i] = tabA(idx0)[i];
      rb[i] = tabB(idx0)[i];
    }
    phase = 2;
  }
  // main-loop cursors
  int pe = 1, sq_left = 0;                               // sliding: program cursor / squarings left before the pending product
  uint32_t cur_idx = 0xFF;
  bool need_mul = false;
  int wi = nwin - 2, pos = 0;                            // fixed: window index / squarings done in the current window
  int pass = 0;
  bool op_square = false;
#pragma unroll 1
  for (;;) {
    if (pass == 0) {
      // ---- fetch the next multiplication and stage its streamed operands
      if (phase == 0) {
        op_square = true;
      } else if (phase == 1) {
        op_square = false;                                 // multiplier already staged in s1 / s2
      } else {
        bool is_mul;
        uint32_t idx = 0;
        if (sliding) {
          if (sq_left == 0 && !need_mul) {
            if (pe >= nprog) break;
            const uint32_t en = __ldg(prog + pe++);
            sq_left = static_cast<int>(en >> 8);
            cur_idx = en & 0xFFu;
            need_mul = cur_idx != 0xFFu;
            if (sq_left == 0 && !need_mul) continue;
          }
          if (sq_left > 0) {
            is_mul = false;
            sq_left--;
          } else {
            is_mul = true;
            idx = cur_idx;
            need_mul = false;
          }
        } else {
          if (wi < 0) break;
          if (pos < wbits) {
            is_mul = false;
            pos++;
          } else {
            is_mul = true;
            idx = get_bits(e, ewords, wi * wbits, wbits);
            pos = 0;
            wi--;
          }
        }
        op_square = !is_mul;
        if (is_mul) {
          double cv[L];
#pragma unroll
          for (int i = 0; i < L; i++) cv[i] = tabA(idx)[i];
          stage_digits52<L, TPI>(s1, cv);
#pragma unroll
          for (int i = 0; i < L; i++) cv[i] = tabB(idx)[i];
          stage_digits52<L, TPI>(s2, cv);
        }
      }
      if (op_square) stage_digits52<L, TPI>(s1, ra);
    } else if (pass == 1 && op_square) {
      // stream 2b for the cross term of the squaring: the digits of b are normalised, so doubling is a one-bit shift across the
      // digits (the bit leaving a lane's top digit enters the next lane; 2b < 8n < R, nothing leaves the number)
      uint64_t d[L];
      double b2[L];
#pragma unroll
      for (int i = 0; i < L; i++) d[i] = double_to_u52(rb[i]);
      uint32_t carry_in = __shfl_up_sync(FULL, static_cast<uint32_t>(d[L - 1] >> 51), 1, TPI);
      if (t == 0) carry_in = 0u;
#pragma unroll
      for (int i = L - 1; i >= 1; i--) d[i] = ((d[i] << 1) & MASK52) | (d[i - 1] >> 51);
      d[0] = ((d[0] << 1) & MASK52) | carry_in;
#pragma unroll
      for (int i = 0; i < L; i++) b2[i] = u52_to_double(d[i]);
      stage_digits52<L, TPI>(s1, b2);
    }
    uint64_t col[L], hi;
    mmd_rows_pair<L, TPI>(ra, nv, np, (pass == 1 && !op_square) ? s2 : s1, sq, pass == 0 ? PASS_SAVE : (pass == 1 ? PASS_INJECT : 0), col, hi);
    bool done = false;
    if (pass == 0) {
      resolve_carries<L, TPI>(col, hi, tn);
      pass = 1;
    } else if (pass == 1) {
      add_n_minus_1<L, TPI>(col, nv);
      if (op_square) {
        resolve_carries<L, TPI>(col, hi, rb);
        done = true;
      } else {
        resolve_carries<L, TPI>(col, hi, s1v);
#pragma unroll
        for (int i = 0; i < L; i++) ra[i] = rb[i];         // pass 2 multiplies b by the staged c
        pass = 2;
      }
    } else {
      add_digits<L, TPI>(col, s1v);
      resolve_carries<L, TPI>(col, hi, rb);
      done = true;
    }
    if (done) {
#pragma unroll
      for (int i = 0; i < L; i++) ra[i] = tn[i];
      pass = 0;
      if (phase == 0) {
        // x~^2 stays staged for the whole table build; the running value goes back to x~
        stage_digits52<L, TPI>(s1, ra);
        stage_digits52<L, TPI>(s2, rb);
#pragma unroll
        for (int i = 0; i < L; i++) {
          ra[i] = tabA(0)[i];
          rb[i] = tabB(0)[i];
        }
        phase = 1;
      } else if (phase == 1) {
#pragma unroll
        for (int i = 0; i < L; i++) {
          tabA(first_free)[i] = ra[i];
          tabB(first_free)[i] = rb[i];
        }
        first_free++;
        build_left--;
      }
      if (phase == 1 && build_left == 0) {
#pragma unroll
        for (int i = 0; i < L; i++) {
          ra[i] = tabA(idx0)[i];
          rb[i] = tabB(idx0)[i];
        }
        phase = 2;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < L; i++) {
    outA[t * L + i] = ra[i];
    outB[t * L + i] = rb[i];
  }
}

// out[i] = base[i]^exp[i] mod n^2 (n = the root modulus `m`); base: kbase words (any value, reduced on load)
// Two blocks per SM (255 registers), not three: the pair loop keeps two more L-digit values alive than k_modexp52, and at the
// 168-register cap of three blocks ptxas not only spilled them but, starved of temporaries, scheduled the row loop with its
// FP64 results consumed 1-6 instructions after they were issued (tools/sass_deps.py); the row loop itself does not care about
// occupancy (profiles/r02_rowloop_microbench.jsonl).  Measured at B = 65,536: 371 -> 340 ms (10x4), 420 -> 388 ms (10x2).
#ifndef SCB_PAIR_BLOCKS_PER_SM
#define SCB_PAIR_BLOCKS_PER_SM 2
#endif
template <int L, int TPI>
__global__ void __launch_bounds__(BLOCK_THREADS, SCB_PAIR_BLOCKS_PER_SM)
k_modexp52p(const Modulus52Dev* __restrict__ m, const PairConsts52Dev* __restrict__ pc, const uint32_t* __restrict__ base, int kbase,
            const uint32_t* __restrict__ exp, int ewords, size_t exp_stride, int ebits, int wbits, double* scratch, uint32_t* __restrict__ out, int kout,
            size_t B, unsigned long long* queue, const uint16_t* __restrict__ prog, int nprog, int pw) {
  extern __shared__ double smem52[];
  using GI = GroupIndex52<L, TPI>;
  using GP = GeoPair<L, TPI>;
  constexpr int K = Geo52<L, TPI>::K;
  double* sm_group = smem52 + (threadIdx.x >> 5) * GP::WARP_SMEM_WORDS64 + (threadIdx.x & 31) / TPI;
  const size_t slot = GI::slot();
  double* XA = scratch + slot * modexp52p_scratch_doubles<L, TPI>(wbits);
  double* XB = XA + K;
  uint64_t* TMP = reinterpret_cast<uint64_t*>(XB + K);
  double* table = XB + K + 4 * K;
  size_t inst;
  bool active;
  while (GI::next(queue, B, inst, active)) {
    load_mont_pair<L, TPI>(m, pc, sm_group, XA, XB, base + inst * kbase, kbase);
    __syncwarp();
    group_modexp52_pair<L, TPI>(m, pc, sm_group, XA, XB, XA, XB, exp + inst * exp_stride, ewords, ebits, wbits, prog, nprog, pw, table);
    __syncwarp();
    pair_store<L, TPI>(m, sm_group, XA, XB, TMP, out + inst * kout, kout, active);
  }
}

}  // namespace scb
