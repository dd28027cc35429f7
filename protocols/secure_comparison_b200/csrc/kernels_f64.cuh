// kernels_f64.cuh -- batched modular exponentiation on the FP64 engine (bn52.cuh).
//
// Same structure as k_modexp / group_modexp of the integer engine: one group of TPI lanes per exponentiation, the
// running value and the modulus stay in registers across the whole exponentiation, fixed-window table in the
// group's global scratch, dynamic work queue.  Inputs and outputs are the 32-bit-word arrays of the C ABI; outputs
// are canonical residues in [0, n).
#pragma once
#include "bn52.cuh"
#include "kernels_generic.cuh"

namespace scb {

template <int L, int TPI>
struct GroupIndex52 {
  static constexpr int GPW = Geo52<L, TPI>::GPW;
  static constexpr int GROUPS_PER_BLOCK = GPW * WARPS_PER_BLOCK;
  __device__ static __forceinline__ size_t slot() {
    return (static_cast<size_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5)) * GPW + (threadIdx.x & 31) / TPI;
  }
  __device__ static __forceinline__ bool next(unsigned long long* queue, size_t B, size_t& inst, bool& active) {
    unsigned long long base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(queue, static_cast<unsigned long long>(GPW));
    base = __shfl_sync(FULL, base, 0);
    if (base >= B) return false;
    inst = base + (threadIdx.x & 31) / TPI;
    active = inst < B;
    if (!active) inst = B - 1;
    return true;
  }
  __device__ static __forceinline__ double* warp_smem(double* smem) { return smem + (threadIdx.x >> 5) * Geo52<L, TPI>::WARP_SMEM_DOUBLES; }
};

template <int L, int TPI>
__host__ __device__ constexpr size_t modexp52_scratch_doubles(int wbits) {
  return static_cast<size_t>(Geo52<L, TPI>::K) * (3 + (static_cast<size_t>(1) << wbits));
}

// words (kw of them, value < 2^(104 K)) -> Montgomery form at dst (digits < 2^52, value < 4n); tmp: another K-digit slot
template <int L, int TPI>
__device__ __forceinline__ void load_mont52(const Modulus52Dev* __restrict__ m, double* sm_group, double* dst, double* tmp,
                                            const uint32_t* __restrict__ words, int kw) {
  constexpr int K = Geo52<L, TPI>::K;
  const int t = (threadIdx.x & 31) % TPI;
#pragma unroll
  for (int i = 0; i < L; i++) dst[t * L + i] = u52_to_double(digit52_from_words(words, kw, t * L + i));
  mmd<L, TPI>(m, sm_group, dst, dst, m->r2);
  if (32 * kw > 52 * K) {                       // launch-uniform: double-width operand x = lo + hi R
#pragma unroll
    for (int i = 0; i < L; i++) tmp[t * L + i] = u52_to_double(digit52_from_words(words, kw, K + t * L + i));
    mmd<L, TPI>(m, sm_group, tmp, tmp, m->r3);
    add52<L, TPI>(dst, dst, tmp);
  }
}

// src (ordinary value, digits < 2^52, value <= a few n) -> canonical residue -> kout words.  Serial in lane 0 of the
// group over the shared-memory staging area (once per OUTPUT value).
template <int L, int TPI>
__device__ __noinline__ void canonical_store52(const Modulus52Dev* __restrict__ m, double* sm_group, const double* src, uint32_t* __restrict__ words,
                                               int kout, bool do_store) {
  constexpr int K = Geo52<L, TPI>::K;
  constexpr int GPW = Geo52<L, TPI>::GPW;
  const int t = (threadIdx.x & 31) % TPI;
  double v[L];
#pragma unroll
  for (int i = 0; i < L; i++) v[i] = src[t * L + i];
  stage_digits52<L, TPI>(sm_group, v);
  if (t == 0) {
    for (int it = 0; it < 3; it++) {
      int cmp = 0;
      for (int j = K - 1; j >= 0; j--) {
        const double x = sm_group[j * GPW], y = m->n[j];
        if (x != y) { cmp = x > y ? 1 : -1; break; }
      }
      if (cmp < 0) break;
      uint64_t borrow = 0;
      for (int j = 0; j < K; j++) {
        const uint64_t x = double_to_u52(sm_group[j * GPW]), y = double_to_u52(m->n[j]) + borrow;
        borrow = x < y ? 1ull : 0ull;
        sm_group[j * GPW] = u52_to_double((x - y) & MASK52);
      }
    }
  }
  __syncwarp();
  if (do_store) {
    for (int w = t; w < kout; w += TPI) {
      const int bit = w * 32;
      const int d = bit / 52, r = bit - d * 52;
      uint64_t x = 0;
      if (d < K) x = double_to_u52(sm_group[d * GPW]) >> r;
      if (d + 1 < K && r > 20) x |= double_to_u52(sm_group[(d + 1) * GPW]) << (52 - r);
      words[w] = static_cast<uint32_t>(x);
    }
  }
  __syncwarp();
}

// acc = x~^e (Montgomery form) for ONE group; xm, acc: K-digit slots, table: (1 << wbits) slots (global scratch).
template <int L, int TPI>
__device__ __forceinline__ void group_modexp52(const Modulus52Dev* __restrict__ m, double* sm_group, double* acc, const double* xm,
                                               const uint32_t* __restrict__ e, int ewords, int ebits, int wbits, double* table) {
  constexpr int K = Geo52<L, TPI>::K;
  const int t = (threadIdx.x & 31) % TPI;
  const int nwin = (ebits + wbits - 1) / wbits;
  if (nwin == 0) {
#pragma unroll
    for (int i = 0; i < L; i++) acc[t * L + i] = m->one[t * L + i];
    return;
  }
  double r[L], nv[L];
#pragma unroll
  for (int i = 0; i < L; i++) nv[i] = m->n[t * L + i];
  const double np = m->np;
#pragma unroll
  for (int i = 0; i < L; i++) table[t * L + i] = m->one[t * L + i];
#pragma unroll
  for (int i = 0; i < L; i++) r[i] = xm[t * L + i];
#pragma unroll
  for (int i = 0; i < L; i++) table[K + t * L + i] = r[i];
  stage_digits52<L, TPI>(sm_group, r);
  const int ntab = (1 << wbits) - 2;
  const int nmain = (nwin - 1) * (wbits + 1);
  int pos = 0, wi = nwin - 2;                              // position inside the current window group / index of its window (no divisions in the loop)
#pragma unroll 1
  for (int step = 0; step < ntab + nmain; step++) {
    if (step >= ntab) {
      if (step == ntab) {
        const uint32_t win = get_bits(e, ewords, (nwin - 1) * wbits, wbits);
#pragma unroll
        for (int i = 0; i < L; i++) r[i] = table[static_cast<size_t>(win) * K + t * L + i];
      }
      if (pos < wbits) {
        stage_digits52<L, TPI>(sm_group, r);
      } else {
        const uint32_t win = get_bits(e, ewords, wi * wbits, wbits);
        double bv[L];
#pragma unroll
        for (int i = 0; i < L; i++) bv[i] = table[static_cast<size_t>(win) * K + t * L + i];
        stage_digits52<L, TPI>(sm_group, bv);
      }
    }
    mmd_rows<L, TPI>(r, nv, np, sm_group, r);
    if (step >= ntab) {
      if (pos == wbits) {
        pos = 0;
        wi--;
      } else {
        pos++;
      }
    }
    if (step < ntab) {
#pragma unroll
      for (int i = 0; i < L; i++) table[static_cast<size_t>(step + 2) * K + t * L + i] = r[i];
    }
  }
  if (nmain == 0) {
    const uint32_t win = get_bits(e, ewords, 0, wbits);
#pragma unroll
    for (int i = 0; i < L; i++) r[i] = table[static_cast<size_t>(win) * K + t * L + i];
  }
#pragma unroll
  for (int i = 0; i < L; i++) acc[t * L + i] = r[i];
}

// acc = x~^e for a SHARED exponent compiled on the host into a sliding-window program (scb_ctx::register_exponent):
//   prog[0]            = table index of the leading window
//   prog[i], i >= 1    = (number of squarings << 8) | table index of the odd window value to multiply by (0xFF = none)
// table[i] = x~^(2i+1), i < 2^(pw-1): one squaring (x~^2, staged once) and 2^(pw-1) - 1 products.  About 4 % fewer
// products than the fixed-window loop for 1024..2048-bit exponents (GMP's mpz_powm, which the reference reaches through
// gmpy2.powmod, is a sliding-window method as well).  The program is the same for the whole launch, so every branch here
// is launch-uniform; one inlined copy of the row loop serves the three phases.
template <int L, int TPI>
__device__ __forceinline__ void group_modexp52_prog(const Modulus52Dev* __restrict__ m, double* sm_group, double* acc, const double* xm,
                                                    const uint16_t* __restrict__ prog, int nprog, int pw, double* table) {
  constexpr int K = Geo52<L, TPI>::K;
  const int t = (threadIdx.x & 31) % TPI;
  const int T = 1 << (pw - 1);
  double r[L], nv[L];
#pragma unroll
  for (int i = 0; i < L; i++) nv[i] = m->n[t * L + i];
  const double np = m->np;
#pragma unroll
  for (int i = 0; i < L; i++) r[i] = xm[t * L + i];
#pragma unroll
  for (int i = 0; i < L; i++) table[t * L + i] = r[i];
  const uint32_t idx0 = __ldg(prog);
  int phase = 0, ti = 1, pe = 1, sq_left = 0;
  uint32_t cur_idx = 0xFF;
  bool need_mul = false;
#pragma unroll 1
  for (;;) {
    if (phase == 0) {
      stage_digits52<L, TPI>(sm_group, r);                       // x~ * x~
    } else if (phase == 2) {
      if (sq_left == 0 && !need_mul) {
        if (pe >= nprog) break;
        const uint32_t e = __ldg(prog + pe++);
        sq_left = static_cast<int>(e >> 8);
        cur_idx = e & 0xFFu;
        need_mul = cur_idx != 0xFFu;
        if (sq_left == 0 && !need_mul) continue;
      }
      if (sq_left > 0) {
        stage_digits52<L, TPI>(sm_group, r);                     // squaring
        sq_left--;
      } else {
        double bv[L];
#pragma unroll
        for (int i = 0; i < L; i++) bv[i] = table[static_cast<size_t>(cur_idx) * K + t * L + i];
        stage_digits52<L, TPI>(sm_group, bv);                    // window product
        need_mul = false;
      }
    }                                                            // phase 1: x~^2 is still staged
    mmd_rows<L, TPI>(r, nv, np, sm_group, r);
    if (phase == 0) {
      stage_digits52<L, TPI>(sm_group, r);                       // x~^2 stays staged for the whole table build
#pragma unroll
      for (int i = 0; i < L; i++) r[i] = table[t * L + i];       // back to x~
      phase = 1;
    } else if (phase == 1) {
#pragma unroll
      for (int i = 0; i < L; i++) table[static_cast<size_t>(ti) * K + t * L + i] = r[i];
      ti++;
    }
    if (phase == 1 && ti == T) {
#pragma unroll
      for (int i = 0; i < L; i++) r[i] = table[static_cast<size_t>(idx0) * K + t * L + i];
      phase = 2;
    }
  }
#pragma unroll
  for (int i = 0; i < L; i++) acc[t * L + i] = r[i];
}

// Test hook for the carry resolution of the FP64 engine (its second-order ripple loop has probability ~2^-52 per
// product on random data, so the parity tests never reach it): cols [G][K] column values, hi [G][TPI] the extra
// high accumulator of each lane, out [G][K] the normalised 52-bit digits.  One warp per GPW numbers.
template <int L, int TPI>
__global__ void k_debug_resolve52(const uint64_t* __restrict__ cols, const uint64_t* __restrict__ hi, uint64_t* __restrict__ out, size_t G) {
  constexpr int K = Geo52<L, TPI>::K;
  constexpr int GPW = Geo52<L, TPI>::GPW;
  const int lane = threadIdx.x & 31, t = lane % TPI;
  const size_t rounds = (G + GPW - 1) / GPW;
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t g = rd * GPW + lane / TPI;
    const bool active = g < G;
    if (!active) g = G - 1;
    uint64_t col[L];
    double r[L];
#pragma unroll
    for (int i = 0; i < L; i++) col[i] = cols[g * K + t * L + i];
    resolve_carries<L, TPI>(col, hi[g * TPI + t], r);
    if (active) {
#pragma unroll
      for (int i = 0; i < L; i++) out[g * K + t * L + i] = double_to_u52(r[i]);
    }
  }
}

// out[i] = (mul_in[i] *) base[i]^exp[i] mod n
// registers: L = 10 needs ~150 (3 blocks of 128 threads per SM: measured 3 % faster than 4 blocks with a few spills),
// L = 15 needs ~230 (2 blocks per SM)
// (2 blocks / 255 registers: the 1024-bit launches unchanged, the is-zero launch 50.9 -> 53.3 ms: profiles/r02t_bounds.log)
template <int L, int TPI>
__global__ void __launch_bounds__(BLOCK_THREADS, (L <= 10 ? 3 : 2))
k_modexp52(const Modulus52Dev* __restrict__ m, const uint32_t* __restrict__ base, int kbase, const uint32_t* __restrict__ exp, int ewords,
           size_t exp_stride, int ebits, int wbits, double* scratch, const uint32_t* __restrict__ mul_in, uint32_t* __restrict__ out, int kout,
           size_t B, unsigned long long* queue, const uint16_t* __restrict__ prog, int nprog, int pw) {
  extern __shared__ double smem52[];
  using GI = GroupIndex52<L, TPI>;
  constexpr int K = Geo52<L, TPI>::K;
  double* sm_group = GI::warp_smem(smem52) + (threadIdx.x & 31) / TPI;
  const size_t slot = GI::slot();
  double* X = scratch + slot * modexp52_scratch_doubles<L, TPI>(wbits);
  double* ACC = X + K;
  double* TMP = ACC + K;
  double* table = TMP + K;
  size_t inst;
  bool active;
  while (GI::next(queue, B, inst, active)) {
    load_mont52<L, TPI>(m, sm_group, X, TMP, base + inst * kbase, kbase);
    if (prog != nullptr)
      group_modexp52_prog<L, TPI>(m, sm_group, ACC, X, prog, nprog, pw, table);
    else
      group_modexp52<L, TPI>(m, sm_group, ACC, X, exp + inst * exp_stride, ewords, ebits, wbits, table);
    if (mul_in != nullptr) {
      load_mont52<L, TPI>(m, sm_group, TMP, X, mul_in + inst * kout, kout);
      mmd<L, TPI>(m, sm_group, ACC, ACC, TMP);
    }
    mmd<L, TPI>(m, sm_group, TMP, ACC, m->unit);               // leave Montgomery form
    canonical_store52<L, TPI>(m, sm_group, TMP, out + inst * kout, kout, active);
  }
}

}  // namespace scb
