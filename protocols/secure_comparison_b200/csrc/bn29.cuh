// bn29.cuh -- cooperative big-integer Montgomery arithmetic for sm_100a (B200).
//
// Representation.  A residue mod an odd modulus n is held as K = L*TPI digits of W = 29 bits
// ("radix 2^29").  A group of TPI adjacent lanes of one warp works on one big integer; lane t of the
// group owns digits [t*L, (t+1)*L).  The Montgomery radix R = 2^(29*K) is chosen with R > 4n (in fact
// R/n >= 2^20 for every supported size), so every intermediate can stay "lazy" in [0, 2n) with digits
// <= 2^29, and only final outputs are reduced to the canonical residue in [0, n) (what GMP / CPython
// return in the reference).
//
// Why 29-bit digits.  On B200 the accumulating IMAD.WIDE.U32 Rd, Ra, Rb, Rc (32x32+64 -> 64) issues on the fmaheavy
// pipe only, 32 lanes/clk/SM nominal (measured 29.6: 8.6 TMAC/s, profiles/r01_imad_microbench2.jsonl and
// scb_measure_peak); the carry-chained form IMAD.WIDE.U32.X (what mad.lo.cc/madc.hi.cc compile to, and what a
// radix-2^32 CIOS needs) reaches 27.5/clk/SM, IMAD.HI + IMAD.LO pairs 18.  With 29-bit digits a product is < 2^58, and a
// 64-bit column accumulator can absorb the 2*L <= 62 products it sees while it belongs to one lane, so the whole
// inner loop is carry-less IMAD.WIDE.U32 with one 64-bit shift/add per row.
//
// One Montgomery product r = a*b*R^-1 mod n (CIOS by rows, one digit b_j per row):
//   col[i] += a[i]*b_j                (L IMAD.WIDE per lane; b_j is a shared-memory broadcast)
//   q = (col[0] * n') mod 2^29        (lane 0 of the group, broadcast with one SHFL)
//   col[i] += n[i]*q                  (L IMAD.WIDE)
//   the lane's lowest column is now finished: its high part is added to the lane's next column and
//   its low 29 bits go to the lane below (one SHFL) where they become that lane's new top column.
// The columns are never moved: the row loop is unrolled L times and logical column i of row rr lives
// in register pair acc[(i + rr) % L], so the per-row column shift is pure renaming.
//
// mm() is a memory-to-memory operation (operands and result are K-digit arrays in global scratch,
// each lane touching only its own slice) and is deliberately NOT inlined: its unrolled body is
// ~14 KB of SASS per (L, TPI) and every kernel in the library calls the same copy.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace scb {

constexpr int W = 29;
constexpr uint32_t DMASK = (1u << W) - 1u;
constexpr unsigned FULL = 0xffffffffu;
constexpr int MAXK = 288;   // digits of the largest supported modulus (8192-bit N^2 of Paillier-4096)
constexpr int INV_MAXL = 280;   // signed 30-bit limbs of the widest operand of the modular-inverse kernel

// Montgomery constants of one modulus (global memory; written once by scb_modulus_load).
struct ModulusDev {
  int K;            // digits (L*TPI of the kernel configuration that serves this modulus)
  int k32;          // 32-bit words of the modulus at the C ABI
  int bits;         // bit length of n
  uint32_t np;      // -n^-1 mod 2^29
  uint32_t n[MAXK];     // digits of n
  uint32_t one[MAXK];   // R mod n      (Montgomery form of 1)
  uint32_t r2[MAXK];    // R^2 mod n    (x -> x*R)
  uint32_t r3[MAXK];    // R^3 mod n
  uint32_t unit[MAXK];  // the integer 1 (x*R -> x)
  // modular-inverse kernel (k_modinv): n as 30-bit limbs, n^-1 mod 2^30, limb count, outer-iteration cap
  int len30;
  int inv_cap;
  uint32_t ninv30;
  int32_t n30[INV_MAXL];
};

template <int L, int TPI>
struct Geo {
  static_assert(32 % TPI == 0, "TPI must divide the warp");
  static_assert(L >= 3 && L <= 31, "a 64-bit column must absorb 2L products of 58 bits");
  static constexpr int K = L * TPI;
  static constexpr int GPW = 32 / TPI;                 // groups (big integers) per warp
  static constexpr int WARP_SMEM_WORDS = (K + 1) * GPW;   // +1 row: the row loop prefetches b[j+1]
};

// What a lane needs to know about its group.
template <int L, int TPI>
struct Grp {
  const ModulusDev* m;
  uint32_t* sm;      // staging area of the group inside the warp's shared memory: digit j at sm[j*GPW]
  int t;             // lane index inside the group
};

template <int L, int TPI>
__device__ __forceinline__ Grp<L, TPI> grp_init(const ModulusDev* m, uint32_t* warp_smem) {
  const int lane = threadIdx.x & 31;
  Grp<L, TPI> g;
  g.m = m;
  g.t = lane % TPI;
  g.sm = warp_smem + lane / TPI;
  return g;
}

__device__ __forceinline__ uint64_t madw(uint32_t a, uint32_t b, uint64_t c) {
  // C expression on purpose: NVVM emits the accumulating IMAD.WIDE.U32 Rd, Ra, Rb, Rc.  The same
  // thing as inline PTX `mad.wide.u32` makes ptxas 12.9 split each column into two
  // IMAD.WIDE(.., RZ) + IADD3 + IADD3.X, which costs 40% more issue slots.
  return static_cast<uint64_t>(a) * b + c;
}

// Stage the K digits at `src` (global) into the group's shared-memory operand area.
template <int L, int TPI>
__device__ __forceinline__ void stage_from(const Grp<L, TPI>& g, const uint32_t* src) {
  constexpr int GPW = Geo<L, TPI>::GPW;
  uint32_t tmp[L];
#pragma unroll
  for (int i = 0; i < L; i++) tmp[i] = src[g.t * L + i];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < L; i++) g.sm[(g.t * L + i) * GPW] = tmp[i];
  __syncwarp();
}

// dst = a * b * R^-1 mod n, lazy.  a, b, dst: K-digit arrays of THIS group in global memory (dst may
// alias a and/or b).  Bounds: a*b < 4 n R / 2^18 or so is plenty -- in practice a, b are lazy values
// (< 2n, or sums of a few of them) with digits <= 2^29 + 4; the result is < 2n with digits <= 2^29.
// All 32 lanes of the warp must call this together.
// The row loop of one Montgomery product on register operands: r = a * b * R^-1 mod n (lazy), with
//   av[L], nv[L] : this lane's digits of a and of the modulus (registers),
//   b            : ALREADY staged in the group's shared-memory area (digit j at sm_group[j * GPW]) and made visible
//                  with __syncwarp() by the caller (stage_digits()); the pad row K may hold anything,
//   r[L]         : this lane's digits of the result (may be the same array as av: it is written after the last row).
// All 32 lanes of the warp must call this together (full-mask shuffles).
template <int L, int TPI>
__device__ __forceinline__ void mm_rows(const uint32_t (&av)[L], const uint32_t (&nv)[L], const uint32_t np, const uint32_t* __restrict__ sm_group,
                                        uint32_t (&r)[L]) {
  constexpr int GPW = Geo<L, TPI>::GPW;
  const int t = (threadIdx.x & 31) % TPI;
  const bool top = t == TPI - 1;
  uint64_t acc[L];
#pragma unroll
  for (int i = 0; i < L; i++) acc[i] = 0;
  const uint32_t* bp = sm_group;
  uint32_t bj = bp[0];
#pragma unroll 1
  for (int blk = 0; blk < TPI; blk++) {
#pragma unroll
    for (int rr = 0; rr < L; rr++) {
      // logical column i of this row lives in acc[(i + rr) % L]
      bp += GPW;
      const uint32_t bnext = bp[0];               // the last row reads the pad row
      acc[rr % L] = madw(av[0], bj, acc[rr % L]);
      uint32_t q = (static_cast<uint32_t>(acc[rr % L]) * np) & DMASK;
      q = __shfl_sync(FULL, q, 0, TPI);
#pragma unroll
      for (int i = 1; i < L; i++) acc[(i + rr) % L] = madw(av[i], bj, acc[(i + rr) % L]);
      const uint64_t t0 = madw(nv[0], q, acc[rr % L]);
#pragma unroll
      for (int i = 1; i < L; i++) acc[(i + rr) % L] = madw(nv[i], q, acc[(i + rr) % L]);
      acc[(1 + rr) % L] += t0 >> W;
      const uint32_t lo = static_cast<uint32_t>(t0) & DMASK;   // 0 in lane 0 (Montgomery)
      const uint32_t in = __shfl_down_sync(FULL, lo, 1, TPI);
      acc[rr % L] = top ? 0u : in;                // becomes logical column L-1 of the next row
      bj = bnext;
    }
  }
  // after TPI*L rows the renaming is back to the identity: logical column i is acc[i].
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < L; i++) {
    c += acc[i];
    r[i] = static_cast<uint32_t>(c) & DMASK;
    c >>= W;
  }
  uint32_t c_lo = __shfl_up_sync(FULL, static_cast<uint32_t>(c), 1, TPI);
  uint32_t c_hi = __shfl_up_sync(FULL, static_cast<uint32_t>(c >> 32), 1, TPI);
  if (t == 0) { c_lo = 0; c_hi = 0; }
  uint64_t c2 = ((static_cast<uint64_t>(c_hi) << 32) | c_lo) + r[0];
  r[0] = static_cast<uint32_t>(c2) & DMASK;
  c2 = (c2 >> W) + r[1];
  r[1] = static_cast<uint32_t>(c2) & DMASK;
  r[2] += static_cast<uint32_t>(c2 >> W);        // <= 2^29: stays a legal lazy digit
}

// Stage this lane's L digits (registers) into the group's shared-memory operand area and make them visible.
template <int L, int TPI>
__device__ __forceinline__ void stage_digits(uint32_t* sm_group, const uint32_t (&bv)[L]) {
  constexpr int GPW = Geo<L, TPI>::GPW;
  const int t = (threadIdx.x & 31) % TPI;
  __syncwarp();                                   // every lane is done reading the previous operand
#pragma unroll
  for (int i = 0; i < L; i++) sm_group[(t * L + i) * GPW] = bv[i];
  __syncwarp();
}

// dst = a * b * R^-1 mod n, lazy.  a, b, dst: K-digit arrays of THIS group in global memory (dst may
// alias a and/or b).  Bounds: a*b < 4 n R / 2^18 or so is plenty -- in practice a, b are lazy values
// (< 2n, or sums of a few of them) with digits <= 2^29 + 4; the result is < 2n with digits <= 2^29.
// All 32 lanes of the warp must call this together.
template <int L, int TPI>
__device__ __noinline__ void mm(const ModulusDev* __restrict__ m, uint32_t* sm_group, uint32_t* dst, const uint32_t* a,
                                const uint32_t* b) {
  const int t = (threadIdx.x & 31) % TPI;
  uint32_t av[L], nv[L];
  {
    uint32_t bv[L];
#pragma unroll
    for (int i = 0; i < L; i++) bv[i] = b[t * L + i];
#pragma unroll
    for (int i = 0; i < L; i++) av[i] = a[t * L + i];
    stage_digits<L, TPI>(sm_group, bv);
  }
#pragma unroll
  for (int i = 0; i < L; i++) nv[i] = m->n[t * L + i];
  uint32_t r[L];
  mm_rows<L, TPI>(av, nv, m->np, sm_group, r);
#pragma unroll
  for (int i = 0; i < L; i++) dst[t * L + i] = r[i];
}

// dst = a + b (lazy).  Sums of a few lazy values stay far below R (R/n >= 2^20).
template <int L, int TPI>
__device__ __forceinline__ void lazy_add(const Grp<L, TPI>& g, uint32_t* dst, const uint32_t* a, const uint32_t* b) {
  uint32_t c = 0, r[L];
#pragma unroll
  for (int i = 0; i < L; i++) {
    const uint32_t s = a[g.t * L + i] + b[g.t * L + i] + c;   // < 2^31
    r[i] = s & DMASK;
    c = s >> W;
  }
  uint32_t cin = __shfl_up_sync(FULL, c, 1, TPI);
  if (g.t == 0) cin = 0;
  r[0] += cin;                                    // <= 2^29 + 3: legal lazy digit
#pragma unroll
  for (int i = 0; i < L; i++) dst[g.t * L + i] = r[i];
}

template <int L, int TPI>
__device__ __forceinline__ void copy_digits(const Grp<L, TPI>& g, uint32_t* dst, const uint32_t* src) {
#pragma unroll
  for (int i = 0; i < L; i++) dst[g.t * L + i] = src[g.t * L + i];
}

// ---------------------------------------------------------------------------------------------
// radix conversion at the kernel boundary (the C ABI speaks little-endian 32-bit words)
// ---------------------------------------------------------------------------------------------
// digit d of the number stored in words[0..k32)
__device__ __forceinline__ uint32_t digit_from_words(const uint32_t* __restrict__ words, int k32, int d) {
  const int bit = d * W;
  const int w = bit >> 5, sh = bit & 31;
  const uint32_t lo = (w < k32) ? __ldg(words + w) : 0u;
  const uint32_t hi = (w + 1 < k32) ? __ldg(words + w + 1) : 0u;
  return __funnelshift_r(lo, hi, sh) & DMASK;
}

// words (read-only input of this launch) -> K digits at dst (any value < 2^(29K))
template <int L, int TPI>
__device__ __noinline__ void words_to_digits(const Grp<L, TPI> g, uint32_t* dst, const uint32_t* __restrict__ words, int k32) {
#pragma unroll
  for (int i = 0; i < L; i++) dst[g.t * L + i] = digit_from_words(words, k32, g.t * L + i);
}

// Canonicalise the value staged in the group's shared-memory area in place: exact digits in
// [0, 2^29), then subtract n while >= n (at most `max_sub` times).  Serial in lane 0 of the group;
// used once per OUTPUT value, never in inner loops.
template <int L, int TPI>
__device__ __noinline__ void canonicalize_staged(const Grp<L, TPI> g, int max_sub = 2) {
  constexpr int K = Geo<L, TPI>::K;
  constexpr int GPW = Geo<L, TPI>::GPW;
  __syncwarp();
  if (g.t == 0) {
    uint32_t* s = g.sm;
    const uint32_t* nd = g.m->n;
    uint64_t c = 0;
    for (int j = 0; j < K; j++) {
      c += s[j * GPW];
      s[j * GPW] = static_cast<uint32_t>(c) & DMASK;
      c >>= W;
    }
    for (int it = 0; it < max_sub; it++) {
      int cmp = 0;
      for (int j = K - 1; j >= 0; j--) {
        const uint32_t x = s[j * GPW], y = nd[j];
        if (x != y) { cmp = x > y ? 1 : -1; break; }
      }
      if (cmp < 0) break;
      uint32_t borrow = 0;
      for (int j = 0; j < K; j++) {
        const uint32_t x = s[j * GPW], y = nd[j] + borrow;
        borrow = x < y ? 1u : 0u;
        s[j * GPW] = (x - y) & DMASK;
      }
    }
  }
  __syncwarp();
}

// word w (32 bits) of the canonical staged value
template <int L, int TPI>
__device__ __forceinline__ uint32_t word_from_staged(const Grp<L, TPI>& g, int w) {
  constexpr int K = Geo<L, TPI>::K;
  constexpr int GPW = Geo<L, TPI>::GPW;
  const int bit = w * 32;
  const int d = bit / W, r = bit - d * W;
  uint64_t v = 0;
  if (d < K) v = static_cast<uint64_t>(g.sm[d * GPW]) >> r;
  if (d + 1 < K) v |= static_cast<uint64_t>(g.sm[(d + 1) * GPW]) << (W - r);
  if (d + 2 < K) v |= static_cast<uint64_t>(g.sm[(d + 2) * GPW]) << (2 * W - r);
  return static_cast<uint32_t>(v);
}

// src (lazy ORDINARY value, i.e. already out of Montgomery form) -> canonical -> k32 words.
template <int L, int TPI>
__device__ __noinline__ void canonical_store(const Grp<L, TPI> g, const uint32_t* src, uint32_t* __restrict__ words, int k32,
                                                bool do_store) {
  stage_from(g, src);
  canonicalize_staged(g);
  if (do_store)
    for (int w = g.t; w < k32; w += TPI) words[w] = word_from_staged(g, w);
  __syncwarp();
}

// true when the canonical staged value equals the small constant v (all lanes of the group agree)
template <int L, int TPI>
__device__ __noinline__ bool staged_equals_small(const Grp<L, TPI> g, uint32_t v) {
  constexpr int GPW = Geo<L, TPI>::GPW;
  bool ok = true;
#pragma unroll
  for (int i = 0; i < L; i++) {
    const int d = g.t * L + i;
    ok = ok && (g.sm[d * GPW] == (d == 0 ? v : 0u));
  }
  // AND over the TPI lanes of the group
  for (int off = 1; off < TPI; off <<= 1) ok = __shfl_xor_sync(FULL, ok ? 1 : 0, off, TPI) && ok;
  return ok;
}

}  // namespace scb
