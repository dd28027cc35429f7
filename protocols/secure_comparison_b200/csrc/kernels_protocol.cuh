// kernels_protocol.cuh -- the fused secure-comparison phases.
//
// Two kinds of kernels:
//   * "prep" kernels: one THREAD per comparison, plain radix-2^32 word arithmetic for the small
//     plaintext bookkeeping (bit vectors alpha / alpha~ / beta, the comparisons with (N-1)/2, the
//     shifts z >> l, the exact division of Paillier's L function).
//   * "group" kernels: one GROUP of TPI lanes per comparison (or per ciphertext), sequences of mm().
#pragma once
#include "kernels_generic.cuh"

namespace scb {

constexpr int MAXW = INV_MAXW;          // words of the widest operand at the ABI (8192 bits + slack)
constexpr int MAX_L = 64;               // comparison bit length l <= 64 (bit vectors travel as uint64)

// Protocol-level key material (device copy; built by scb_paillier_set_* / scb_dgk_set_*).
struct KeysDev {
  // ---- Paillier ----
  int kN;                       // words of N; ciphertexts have 2*kN words
  uint32_t Nw[MAXW];            // N
  uint32_t halfw[MAXW];         // (N-1)/2
  uint32_t ninv32;              // N^-1 mod 2^32 (exact division by N)
  uint32_t NR2[MAXK];           // N*R^2 mod N^2, digits of the N^2 configuration: mm(m, NR2) = m*N*R
  // ---- DGK ----
  int kn;                       // words of n
  int l;                        // comparison bit length the DGK key was made for (u > 2^(l+2))
  int u_bits;                   // bit length of u: blinding exponents r_i < u
  uint32_t g_m[MAXK];           // g~ (Montgomery form mod n)
  uint32_t c_sa[2][2][MAXK];    // [s == -1][alpha_i] : (enc(s) * enc(alpha_i))~   (initiator.py:459-476)
  // ---- DGK secret part ----
  int kp;                       // words of p
  // ---- Paillier secret part, CRT form (KeyHolder only; scb_paillier_set_private_crt) ----
  int kP;                       // words of the Paillier primes p, q (= kN / 2)
  uint32_t Pw[2][MAXW / 4 + 2]; // p, q
  uint32_t Pinv32[2];           // p^-1, q^-1 mod 2^32 (exact division in L_p, L_q)
  uint32_t crt_enc[2][MAXK];    // (q^2 (q^-2 mod p^2)) R^2, (p^2 (p^-2 mod q^2)) R^2 mod N^2   (digits of the N^2 configuration)
  uint32_t crt_dec[2][MAXK];    // (h_p q (q^-1 mod p)) R^2, (h_q p (p^-1 mod q)) R^2 mod N        (digits of the N configuration)
};

// ------------------------------------------------------------------------------------------------
// multiword helpers (thread-private, radix 2^32)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mw_less(const uint32_t* __restrict__ x, const uint32_t* __restrict__ y, int k) {
  for (int i = k - 1; i >= 0; i--) {
    const uint32_t a = x[i], b = y[i];
    if (a != b) return a < b;
  }
  return false;
}
__device__ __forceinline__ uint64_t mw_low64(const uint32_t* __restrict__ x, int k) {
  return static_cast<uint64_t>(x[0]) | (k > 1 ? static_cast<uint64_t>(x[1]) << 32 : 0ull);
}
__device__ __forceinline__ uint64_t mask_l(int l) { return l >= 64 ? ~0ull : ((1ull << l) - 1ull); }

// out (kout words) = ((x + y) >> l) + plus   with x, y of k words (y nullable = 0), plus a small constant.
__device__ __forceinline__ void mw_add_shift(uint32_t* __restrict__ out, int kout, const uint32_t* __restrict__ x,
                                             const uint32_t* __restrict__ y, int k, int l, uint32_t plus) {
  uint32_t s[MAXW / 2 + 4];
  uint32_t carry = 0;
  for (int i = 0; i < k; i++) {
    const uint64_t t = static_cast<uint64_t>(x[i]) + (y ? y[i] : 0u) + carry;
    s[i] = static_cast<uint32_t>(t);
    carry = static_cast<uint32_t>(t >> 32);
  }
  s[k] = carry;
  const int ws = l >> 5, bs = l & 31;
  uint64_t add = plus;
  for (int j = 0; j < kout; j++) {
    const int idx = j + ws;
    const uint32_t lo = idx <= k ? s[idx] : 0u;
    const uint32_t hi = idx + 1 <= k ? s[idx + 1] : 0u;
    const uint32_t w = bs ? ((lo >> bs) | (hi << (32 - bs))) : lo;
    const uint64_t t = static_cast<uint64_t>(w) + add;
    out[j] = static_cast<uint32_t>(t);
    add = t >> 32;
  }
}

// ------------------------------------------------------------------------------------------------
// prep kernels
// ------------------------------------------------------------------------------------------------
// Initiator bookkeeping for steps 3, 4c, 4e (initiator.py:260-291, 373):
//   alpha = r mod 2^l, alpha~ = (r - N) mod 2^l, r_small = (r < (N-1)/2)
static __global__ void k_prep_initiator(const KeysDev* __restrict__ keys, const uint32_t* __restrict__ r, int l, uint64_t* __restrict__ alpha,
                                 uint64_t* __restrict__ alpha_tilde, uint8_t* __restrict__ r_small, uint8_t* __restrict__ r_bad, size_t B) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int k = keys->kN;
  const uint32_t* ri = r + i * k;
  const uint64_t mask = mask_l(l);
  const uint64_t rl = mw_low64(ri, k), nl = mw_low64(keys->Nw, k);
  alpha[i] = rl & mask;
  alpha_tilde[i] = (rl - nl) & mask;
  r_small[i] = mw_less(ri, keys->halfw, k) ? 1 : 0;
  if (r_bad) r_bad[i] = mw_less(ri, keys->Nw, k) ? 0 : 1;     // step_4c asserts 0 <= r < N (initiator.py:286-288)
}

// KeyHolder bookkeeping for steps 2, 4a (keyholder.py:196, 213): beta = z mod 2^l, d = (z < (N-1)/2)
static __global__ void k_prep_keyholder_bits(const KeysDev* __restrict__ keys, const uint32_t* __restrict__ z, int l, uint64_t* __restrict__ beta,
                                      uint8_t* __restrict__ dflag, size_t B) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int k = keys->kN;
  beta[i] = mw_low64(z + i * k, k) & mask_l(l);
  dflag[i] = mw_less(z + i * k, keys->halfw, k) ? 1 : 0;
}

// Paillier L function: q = (x - 1) / N exactly (x = 1 mod N), by Hensel division in radix 2^32.
// x: [B][2kN] words, q: [B][kN] words.
static __global__ void __launch_bounds__(64) k_paillier_L(const KeysDev* __restrict__ keys, const uint32_t* __restrict__ x, uint32_t* __restrict__ q, size_t B) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B) return;
  const int k = keys->kN;
  uint32_t t[MAXW / 2 + 2];
  const uint32_t* xi = x + idx * 2 * k;
  // t = (x - 1) mod 2^(32k)
  uint32_t borrow = 1;
  for (int i = 0; i < k; i++) {
    const uint32_t v = xi[i];
    t[i] = v - borrow;
    borrow = (v < borrow) ? 1u : 0u;
  }
  const uint32_t ninv = keys->ninv32;
  for (int j = 0; j < k; j++) {
    const uint32_t qj = t[j] * ninv;
    q[idx * k + j] = qj;
    // t -= qj * N << (32 j)   (only the low k words matter)
    uint64_t carry = 0;
    for (int i = 0; i + j < k; i++) {
      const uint64_t prod = static_cast<uint64_t>(qj) * keys->Nw[i] + carry;
      const uint32_t lo = static_cast<uint32_t>(prod);
      carry = prod >> 32;
      const uint32_t old = t[i + j];
      t[i + j] = old - lo;
      carry += (old < lo) ? 1u : 0u;
    }
  }
}

// KeyHolder bookkeeping for steps 4j, 5 (keyholder.py:249-253, 274-287):
//   delta_b = OR_e flags[e];  m1 = z >> l;  m2 = (z < (N-1)/2) ? (z + N) >> l : m1;  m3 = delta_b
// plaintext words have kN words each, laid out [B][3][kN].
static __global__ void k_prep_keyholder_step5(const KeysDev* __restrict__ keys, const uint32_t* __restrict__ z, const uint8_t* __restrict__ flags,
                                       int nflags, const uint8_t* __restrict__ delta_b_in, int l, uint8_t* __restrict__ delta_b,
                                       uint32_t* __restrict__ plain, size_t B) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int k = keys->kN;
  uint8_t db = 0;
  if (flags)
    for (int e = 0; e < nflags; e++) db |= flags[i * nflags + e];
  else
    db = delta_b_in[i] ? 1 : 0;
  db = db ? 1 : 0;
  if (delta_b) delta_b[i] = db;
  const uint32_t* zi = z + i * k;
  uint32_t* m1 = plain + i * 3 * k;
  uint32_t* m2 = m1 + k;
  uint32_t* m3 = m2 + k;
  mw_add_shift(m1, k, zi, nullptr, k, l, 0);
  if (mw_less(zi, keys->halfw, k))
    mw_add_shift(m2, k, zi, keys->Nw, k, l, 0);
  else
    for (int j = 0; j < k; j++) m2[j] = m1[j];
  m3[0] = db;
  for (int j = 1; j < k; j++) m3[j] = 0;
}

// Initiator bookkeeping for the fused steps 6+7 (initiator.py:518-564):
//   result = zeta * (1 + e*N) * V   with   e = (N - (r >> l) - (1 - delta_a)) mod N
//   V = delta_b_enc^-1 (delta_a = 1)  or  delta_b_enc (delta_a = 0)
// because (1 + aN)^-1 = 1 - aN and (1 - aN)(1 - N) = 1 - (a+1)N  (mod N^2).
static __global__ void k_prep_step67(const KeysDev* __restrict__ keys, const uint32_t* __restrict__ r, const uint8_t* __restrict__ delta_a, int l,
                              uint32_t* __restrict__ e, uint8_t* __restrict__ r_small, size_t B) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B) return;
  const int k = keys->kN;
  const uint32_t* ri = r + idx * k;
  uint32_t a[MAXW / 2 + 2];
  mw_add_shift(a, k, ri, nullptr, k, l, delta_a[idx] ? 0u : 1u);     // a = (r >> l) + (1 - delta_a)  (< N since r < N, l >= 1)
  // e = (N - a) mod N : a == 0 -> 0
  bool zero = true;
  for (int j = 0; j < k; j++) zero = zero && (a[j] == 0);
  uint32_t borrow = 0;
  for (int j = 0; j < k; j++) {
    const uint64_t d = static_cast<uint64_t>(keys->Nw[j]) - a[j] - borrow;
    e[idx * k + j] = zero ? 0u : static_cast<uint32_t>(d);
    borrow = static_cast<uint32_t>(d >> 63);
  }
  r_small[idx] = mw_less(ri, keys->halfw, k) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// Batched inversion with Montgomery's trick: chunks of C consecutive elements share one inversion.
// ------------------------------------------------------------------------------------------------
// pass 1: V[i] = in[i]~ ; P[first] = V[first]; P[i] = P[i-1]*V[i]; Tw[chunk] = canonical(P[last])
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_inv_prefix(const ModulusDev* __restrict__ m, const uint32_t* __restrict__ in, int k32, uint32_t* V, uint32_t* P, uint32_t* tmp_scratch,
             uint32_t* __restrict__ Tw, size_t M, int C) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  const size_t nchunks = (M + C - 1) / C;
  // two private slots per group: TMP (products of surplus / padded positions) and TMPV (their operands).
  // Every group executes the SAME sequence of mm() calls (mm shuffles across the whole warp), so padded
  // positions of a ragged chunk and surplus groups run the arithmetic on private dummies.
  uint32_t* TMP = tmp_scratch + slot * 2 * K;
  uint32_t* TMPV = TMP + K;
  const size_t rounds = (nchunks + total - 1) / total;
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t chunk = rd * total + slot;
    const bool active = chunk < nchunks;
    if (!active) chunk = nchunks - 1;
    const size_t first = chunk * C;
    for (int j = 0; j < C; j++) {
      size_t i = first + j;
      const bool valid = i < M;             // ragged last chunk: pad with the identity
      if (!valid) i = M - 1;
      const bool real = active && valid;
      uint32_t* Vi = real ? V + i * K : TMPV;
      uint32_t* Pi = real ? P + i * K : TMP;
      words_to_digits(g, Vi, in + i * k32, k32);
      mm<L, TPI>(m, g.sm, Vi, Vi, m->r2);
      const uint32_t* Vsrc = valid ? Vi : m->one;
      if (j == 0) {
        copy_digits(g, Pi, Vi);              // j == 0 is always a valid position
      } else {
        // previous prefix: the real one for active groups (also across padded positions, where it is
        // carried forward unchanged in TMP), the private slot for surplus groups
        const uint32_t* Pprev = !active ? TMP : (first + j - 1 < M ? P + (first + j - 1) * K : TMP);
        mm<L, TPI>(m, g.sm, Pi, Pprev, Vsrc);
      }
    }
    const size_t last = (first + C - 1 < M) ? first + C - 1 : M - 1;
    const uint32_t* Plast = (active && first + C - 1 < M) ? P + last * K : TMP;
    mm<L, TPI>(m, g.sm, TMP, Plast, m->unit);
    canonical_store(g, TMP, Tw + chunk * k32, k32, active);
  }
}

// pass 2: I = Iw[chunk]~ ; for i = last..first+1: OUT[i] = I*P[i-1]; I = I*V[i]; OUT[first] = I.
// OUT (Montgomery digits, [M][K]) may alias V.  ok_elem[i] = ok_chunk[chunk].
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_inv_backsub(const ModulusDev* __restrict__ m, const uint32_t* V, const uint32_t* P, const uint32_t* __restrict__ Iw, int k32,
              const uint8_t* __restrict__ ok_chunk, uint32_t* tmp_scratch, uint32_t* OUT, uint8_t* __restrict__ ok_elem, size_t M, int C) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  const size_t nchunks = (M + C - 1) / C;
  uint32_t* I = tmp_scratch + slot * 2 * K;
  uint32_t* DUMP = I + K;
  const size_t rounds = (nchunks + total - 1) / total;
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t chunk = rd * total + slot;
    const bool active = chunk < nchunks;
    if (!active) chunk = nchunks - 1;
    const size_t first = chunk * C;
    const size_t last = (first + C - 1 < M) ? first + C - 1 : M - 1;
    words_to_digits(g, I, Iw + chunk * k32, k32);
    mm<L, TPI>(m, g.sm, I, I, m->r2);
    for (int j = C - 1; j >= 1; j--) {
      const size_t i = first + j;
      const bool valid = i <= last;
      const size_t ii = valid ? i : last;
      // OUT[i] = I * P[i-1]   (must read V[i] before OUT[i] overwrites it when OUT aliases V)
      mm<L, TPI>(m, g.sm, DUMP, I, valid ? V + ii * K : m->one);            // I' = I * V[i]
      mm<L, TPI>(m, g.sm, (valid && active) ? OUT + ii * K : I, I, valid ? P + (ii - 1) * K : m->one);
      if (valid) copy_digits(g, I, DUMP); else copy_digits(g, I, I);
      if (valid && active && ok_elem && g.t == 0) ok_elem[ii] = ok_chunk[chunk];
    }
    if (active) {
      copy_digits(g, OUT + first * K, I);
      if (ok_elem && g.t == 0) ok_elem[first] = ok_chunk[chunk];
    }
  }
}

// words -> Montgomery digits (one group per element)
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_to_mont(const ModulusDev* __restrict__ m, const uint32_t* __restrict__ in, int k32, uint32_t* tmp_scratch, uint32_t* OUT, size_t M) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  const size_t rounds = (M + total - 1) / total;
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t i = rd * total + slot;
    const bool active = i < M;
    uint32_t* dst = active ? OUT + i * K : tmp_scratch + slot * K;
    if (!active) i = M - 1;
    words_to_digits(g, dst, in + i * k32, k32);
    mm<L, TPI>(m, g.sm, dst, dst, m->r2);
  }
}

// Montgomery digits -> canonical words (one group per element); used when an ABI call wants words.
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_from_mont(const ModulusDev* __restrict__ m, const uint32_t* D, uint32_t* tmp_scratch, uint32_t* __restrict__ out, int k32,
            const uint8_t* __restrict__ ok, size_t M) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  uint32_t* TMP = tmp_scratch + slot * K;
  const size_t rounds = (M + total - 1) / total;
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t i = rd * total + slot;
    const bool active = i < M;
    if (!active) i = M - 1;
    mm<L, TPI>(m, g.sm, TMP, D + i * K, m->unit);
    canonical_store(g, TMP, out + i * k32, k32, active);
    if (active && ok && !ok[i])
      for (int w = g.t; w < k32; w += TPI) out[i * k32 + w] = 0;
  }
}

// words -> Montgomery digits for a handful of constants (key setup).  premul (nullable, digits in
// Montgomery form): out = (in * premul)~.
template <int L, int TPI>
__global__ void __launch_bounds__(32) k_const_to_mont(const ModulusDev* __restrict__ m, const uint32_t* __restrict__ in, int k32, int count,
                                                      uint32_t* out, int use_r3) {
  extern __shared__ uint32_t smem[];
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, smem);
  const int grp = (threadIdx.x & 31) / TPI;
  // every group of the warp computes the same constants; only group 0 writes the real output, the
  // others write to a dump area behind it (keeps the warp converged for mm's shuffles).
  for (int c = 0; c < count; c++) {
    uint32_t* dst = out + (grp == 0 ? static_cast<size_t>(c) * K : static_cast<size_t>(count + grp - 1) * K);
    words_to_digits(g, dst, in + static_cast<size_t>(c) * k32, k32);
    mm<L, TPI>(m, g.sm, dst, dst, use_r3 ? m->r3 : m->r2);
  }
}

// dst[c] = a[c] * b[c] for `count` constant pairs (all Montgomery digits); same convergence scheme.
template <int L, int TPI>
__global__ void __launch_bounds__(32) k_const_mul(const ModulusDev* __restrict__ m, const uint32_t* a, const uint32_t* b, int count,
                                                  uint32_t* out, uint32_t* dump) {
  extern __shared__ uint32_t smem[];
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, smem);
  const int grp = (threadIdx.x & 31) / TPI;
  for (int c = 0; c < count; c++) {
    uint32_t* dst = grp == 0 ? out + static_cast<size_t>(c) * K : dump + static_cast<size_t>(grp - 1) * K;
    mm<L, TPI>(m, g.sm, dst, a + static_cast<size_t>(c) * K, b + static_cast<size_t>(c) * K);
  }
}

// ------------------------------------------------------------------------------------------------
// Paillier: encryption of a plaintext (1 + m N) and optional randomisation with rho^N
// ------------------------------------------------------------------------------------------------
// dst = (1 + m*N)~ for the plaintext words at `mw` (any value < 2^(29K); reduced mod N^2 implicitly)
template <int L, int TPI>
__device__ __forceinline__ void group_paillier_encode(const Grp<L, TPI>& g, const KeysDev* __restrict__ keys, uint32_t* dst,
                                                      const uint32_t* __restrict__ mw, int kw) {
  words_to_digits(g, dst, mw, kw);
  mm<L, TPI>(g.m, g.sm, dst, dst, keys->NR2);            // m*N*R
  lazy_add(g, dst, dst, g.m->one);                        // (1 + m N) R
}

// out[i][v] = (1 + plain[i][v] * N) * rho[i][v]^N mod N^2   for v < nv  (keyholder.py:274-287 + 126-128)
// plain: [B][nv][kN]; rho: [B][nv][kN] or null (no randomisation); out: [B][nv][2kN].
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_paillier_encrypt(const ModulusDev* __restrict__ m, const KeysDev* __restrict__ keys, const uint32_t* __restrict__ plain,
                   const uint32_t* __restrict__ rho, const uint32_t* __restrict__ rho_pow, int ebits, int wbits, uint32_t* scratch,
                   uint32_t* __restrict__ out, size_t M, unsigned long long* queue, const uint16_t* __restrict__ prog, int nprog, int pw) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot();
  const int kN = keys->kN;
  uint32_t* X = scratch + slot * modexp_scratch_words<L, TPI>(wbits);
  uint32_t* ACC = X + K;
  uint32_t* TMP = ACC + K;
  uint32_t* table = TMP + K;
  size_t i;
  bool active;
  while (GI::next(queue, M, i, active)) {
    if (rho != nullptr) {
      words_to_digits(g, TMP, rho + i * kN, kN);
      mm<L, TPI>(m, g.sm, X, TMP, m->r2);
      group_modexp_any(g, ACC, X, keys->Nw, kN, ebits, wbits, table, prog, nprog, pw);
      group_paillier_encode(g, keys, TMP, plain + i * kN, kN);
      mm<L, TPI>(m, g.sm, ACC, ACC, TMP);
    } else {
      group_paillier_encode(g, keys, ACC, plain + i * kN, kN);
      if (rho_pow != nullptr) {                       // rho^N already computed (FP64 engine): [M][2kN] words
        words_to_digits(g, TMP, rho_pow + i * 2 * kN, 2 * kN);
        mm<L, TPI>(m, g.sm, TMP, TMP, m->r2);
        mm<L, TPI>(m, g.sm, ACC, ACC, TMP);
      }
    }
    mm<L, TPI>(m, g.sm, TMP, ACC, m->unit);
    canonical_store(g, TMP, out + i * 2 * kN, 2 * kN, active);
  }
}

// Initiator.step_1 (+ randomize): z = y * x^-1 * (1 + (2^l + r) N) [* rho^N]  (initiator.py:228-258, 109)
// xinv: Montgomery digits [B][K] from the batched inversion.
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_step1(const ModulusDev* __restrict__ m, const KeysDev* __restrict__ keys, const uint32_t* __restrict__ y, const uint32_t* xinv,
        const uint32_t* __restrict__ r, const uint32_t* __restrict__ rho, const uint32_t* __restrict__ rho_pow, int l, int ebits, int wbits,
        uint32_t* scratch, uint32_t* __restrict__ out, size_t B, unsigned long long* queue, const uint16_t* __restrict__ prog, int nprog, int pw) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  const int kN = keys->kN, kc = 2 * kN;
  uint32_t* X = scratch + slot * modexp_scratch_words<L, TPI>(wbits);
  uint32_t* ACC = X + K;
  uint32_t* TMP = ACC + K;
  uint32_t* table = TMP + K;
  size_t i;
  bool active;
  while (GI::next(queue, B, i, active)) {
    // T = (1 + (2^l + r) N)~ : digits of r plus 2^l (lazy digit, < 2^30)
    words_to_digits(g, TMP, r + i * kN, kN);
    if (g.t == (l / W) / L) TMP[l / W] += 1u << (l % W);
    mm<L, TPI>(m, g.sm, TMP, TMP, keys->NR2);
    lazy_add(g, TMP, TMP, m->one);
    // Z = y~ * xinv~ * T~
    words_to_digits(g, X, y + i * kc, kc);
    mm<L, TPI>(m, g.sm, X, X, m->r2);
    mm<L, TPI>(m, g.sm, X, X, xinv + i * K);
    mm<L, TPI>(m, g.sm, TMP, X, TMP);
    if (rho != nullptr) {
      // park Z in the extra slot behind the window tables: X, ACC, TMP and the table are all rewritten by the exponentiation
      uint32_t* Z = scratch + (total * modexp_scratch_words<L, TPI>(wbits)) + slot * K;   // one extra slot per group behind the tables
      copy_digits(g, Z, TMP);
      words_to_digits(g, TMP, rho + i * kN, kN);
      mm<L, TPI>(m, g.sm, X, TMP, m->r2);
      group_modexp_any(g, ACC, X, keys->Nw, kN, ebits, wbits, table, prog, nprog, pw);
      mm<L, TPI>(m, g.sm, TMP, ACC, Z);
    } else if (rho_pow != nullptr) {                  // rho^N already computed (FP64 engine): [B][2kN] words
      words_to_digits(g, X, rho_pow + i * kc, kc);
      mm<L, TPI>(m, g.sm, X, X, m->r2);
      mm<L, TPI>(m, g.sm, TMP, TMP, X);
    }
    mm<L, TPI>(m, g.sm, TMP, TMP, m->unit);
    canonical_store(g, TMP, out + i * kc, kc, active);
  }
}

// Fused Initiator.step_6 + step_7 (initiator.py:518-564), see k_prep_step67 for the algebra.
// dbinv: Montgomery digits of delta_b_enc^-1 ([B][K]).
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_step67(const ModulusDev* __restrict__ m, const KeysDev* __restrict__ keys, const uint32_t* __restrict__ zeta1,
         const uint32_t* __restrict__ zeta2, size_t zstride, const uint32_t* __restrict__ delta_b_enc, const uint32_t* dbinv, const uint32_t* __restrict__ e,
         const uint8_t* __restrict__ r_small, const uint8_t* __restrict__ delta_a, uint32_t* scratch, uint32_t* __restrict__ out, size_t B) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  const int kN = keys->kN, kc = 2 * kN;
  uint32_t* A = scratch + slot * 2 * K;
  uint32_t* T = A + K;
  const size_t rounds = (B + total - 1) / total;
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t i = rd * total + slot;
    const bool active = i < B;
    if (!active) i = B - 1;
    const uint32_t* zeta = r_small[i] ? zeta1 + i * zstride : zeta2 + i * zstride;
    words_to_digits(g, A, zeta, kc);
    mm<L, TPI>(m, g.sm, A, A, m->r2);
    group_paillier_encode(g, keys, T, e + i * kN, kN);
    mm<L, TPI>(m, g.sm, A, A, T);
    // V: delta_a = 1 -> dbinv (Montgomery); delta_a = 0 -> delta_b_enc (ordinary words: a*R * v / R = a*v, then still *R)
    // (both variants are formed by every group: mm() shuffles across the whole warp, so it must never sit
    //  in a branch that depends on per-comparison data)
    words_to_digits(g, T, delta_b_enc + i * kc, kc);
    mm<L, TPI>(m, g.sm, T, T, m->r2);
    if (delta_a[i]) copy_digits(g, T, dbinv + i * K);
    mm<L, TPI>(m, g.sm, A, A, T);
    mm<L, TPI>(m, g.sm, A, A, m->unit);
    canonical_store(g, A, out + i * kc, kc, active);
  }
}

// ------------------------------------------------------------------------------------------------
// DGK: encryption of bits g^b * h^rho  (keyholder.py:212-216, 230-233, 106-108)
// ------------------------------------------------------------------------------------------------
// element (i, e): e = 0 is d, e = 1..l are beta_0..beta_{l-1}.  rho: [B][l+1][ew] or null.
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_dgk_encrypt_bits(const ModulusDev* __restrict__ m, const KeysDev* __restrict__ keys, const uint64_t* __restrict__ beta,
                   const uint8_t* __restrict__ dflag, int l, const uint32_t* fb_table, int nwin, int fbw, const uint32_t* __restrict__ rho, int ew,
                   uint32_t* scratch, uint32_t* __restrict__ out, size_t B, unsigned long long* queue) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot();
  const int kn = keys->kn;
  const size_t M = B * (l + 1);
  uint32_t* ACC = scratch + slot * K;
  size_t idx;
  bool active;
  while (GI::next(queue, M, idx, active)) {
    const size_t i = idx / (l + 1);
    const int e = static_cast<int>(idx % (l + 1));
    const int bit = (e == 0) ? dflag[i] : static_cast<int>((beta[i] >> (e - 1)) & 1ull);
    copy_digits(g, ACC, bit ? keys->g_m : m->one);
    if (rho != nullptr) group_fixed_base_mul(g, ACC, fb_table, nwin, rho + idx * ew, ew, fbw);
    mm<L, TPI>(m, g.sm, ACC, ACC, m->unit);
    canonical_store(g, ACC, out + idx * kn, kn, active);
  }
}

// ------------------------------------------------------------------------------------------------
// Initiator steps 4c..4i (+ randomize) for one comparison per group (initiator.py:272-516, 153-154)
// ------------------------------------------------------------------------------------------------
// scratch slots per group (K digits each):
//   D, DI, S, S2, T, U          (6)
//   Bm[l], BI[l] (prefix products first), Wm[l], Cm[l+1]
template <int L, int TPI>
__host__ __device__ constexpr size_t step4_scratch_words(int l, int wbits) {
  return static_cast<size_t>(Geo<L, TPI>::K) * (6 + 4 * static_cast<size_t>(l) + 1 + (static_cast<size_t>(1) << wbits));
}

// pass A: D = r_small ? 1~ : d~ (step 4c), Bm[i] = beta_i~, prefix products P_0 = D, P_{i+1} = P_i * Bm[i];
// total (canonical words) -> Tw[i].  The prefix products are parked in BI[0..l) (P_1..P_l) and DI (P_0 = D).
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_step4_prefix(const ModulusDev* __restrict__ m, const KeysDev* __restrict__ keys, const uint32_t* __restrict__ db_enc,
               const uint8_t* __restrict__ r_small, int l, int wbits, uint32_t* scratch,
               uint32_t* __restrict__ Tw, size_t B) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  const int kn = keys->kn;
  // NOTE: pass A and pass B must see the same (instance -> scratch) mapping, so the scratch is indexed
  // by INSTANCE here (one slot set per comparison), not by group slot.
  const size_t rounds = (B + total - 1) / total;
  const size_t stride = step4_scratch_words<L, TPI>(l, wbits);
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t i = rd * total + slot;
    const bool active = i < B;
    uint32_t* base = scratch + (active ? i : B + slot) * stride;      // surplus groups work in a private area behind
    if (!active) i = B - 1;
    uint32_t* D = base;
    uint32_t* DI = D + K;
    uint32_t* T = D + 4 * K;
    uint32_t* Bm = D + 6 * K;
    uint32_t* BI = Bm + static_cast<size_t>(l) * K;
    words_to_digits(g, D, db_enc + i * (l + 1) * kn, kn);
    mm<L, TPI>(m, g.sm, D, D, m->r2);                          // uniform across the warp (mm shuffles warp-wide)
    if (r_small[i]) copy_digits(g, D, m->one);                 // step 4c: [d] <- [0] = 1
    copy_digits(g, DI, D);
    const uint32_t* prev = DI;
    for (int j = 0; j < l; j++) {
      uint32_t* Bj = Bm + static_cast<size_t>(j) * K;
      words_to_digits(g, Bj, db_enc + (i * (l + 1) + 1 + j) * kn, kn);
      mm<L, TPI>(m, g.sm, Bj, Bj, m->r2);
      mm<L, TPI>(m, g.sm, BI + static_cast<size_t>(j) * K, prev, Bj);
      prev = BI + static_cast<size_t>(j) * K;
    }
    mm<L, TPI>(m, g.sm, T, prev, m->unit);
    canonical_store(g, T, Tw + i * kn, kn, active);
  }
}

// pass B: everything else.  TI: Montgomery digits of the inverse of the total product ([B][K]).
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_step4_main(const ModulusDev* __restrict__ m, const KeysDev* __restrict__ keys, const uint32_t* TI, const uint64_t* __restrict__ alpha,
             const uint64_t* __restrict__ alpha_tilde, const uint8_t* __restrict__ delta_a, const uint32_t* __restrict__ r_blind, int rb_words,
             const uint8_t* __restrict__ perm, const uint32_t* fb_table, int nwin, int fbw, const uint32_t* __restrict__ rho, int ew, int l,
             int ubits, int wbits, int do_blind, uint32_t* scratch, uint32_t* __restrict__ out, size_t B, unsigned long long* queue) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot();
  const int kn = keys->kn;
  const size_t stride = step4_scratch_words<L, TPI>(l, wbits);
  size_t i;
  bool active;
  while (GI::next(queue, B, i, active)) {
    uint32_t* base = scratch + (active ? i : B + slot) * stride;
    uint32_t* D = base;
    uint32_t* DI = D + K;
    uint32_t* S = D + 2 * K;
    uint32_t* S2 = D + 3 * K;
    uint32_t* T = D + 4 * K;
    uint32_t* U = D + 5 * K;
    uint32_t* Bm = D + 6 * K;
    uint32_t* BI = Bm + static_cast<size_t>(l) * K;
    uint32_t* Wm = BI + static_cast<size_t>(l) * K;
    uint32_t* Cm = Wm + static_cast<size_t>(l) * K;
    uint32_t* table = Cm + static_cast<size_t>(l + 1) * K;
    if (!active) {
      // surplus groups: make the private area well-defined (values are irrelevant, they are never stored)
      for (int s = 0; s < 6 + 4 * l + 1; s++) copy_digits(g, base + static_cast<size_t>(s) * K, m->one);
    }
    const uint64_t al = alpha[i], at = alpha_tilde[i];
    const int da = delta_a[i] ? 1 : 0;
    // ---- back-substitution: I = (D * prod B)^-1 ; BI[j] held prefix P_{j+1}; DI held P_0 = D
    copy_digits(g, T, active ? TI + i * K : m->one);
    for (int j = l - 1; j >= 0; j--) {
      const uint32_t* Pprev = (j == 0) ? DI : BI + static_cast<size_t>(j - 1) * K;
      mm<L, TPI>(m, g.sm, U, T, Pprev);                                           // beta_j^-1 = I * P_j
      mm<L, TPI>(m, g.sm, T, T, Bm + static_cast<size_t>(j) * K);               // I *= beta_j
      copy_digits(g, BI + static_cast<size_t>(j) * K, U);
    }
    copy_digits(g, DI, T);                                                         // d^-1
    // ---- steps 4d, 4e, 4f:  W_j = ((alpha_j ? g * beta_j^-1 : beta_j) * (alpha_j != alpha~_j ? d^-1 : 1))^(2^j)
    for (int j = 0; j < l; j++) {
      const int aj = static_cast<int>((al >> j) & 1), atj = static_cast<int>((at >> j) & 1);
      uint32_t* Wj = Wm + static_cast<size_t>(j) * K;
      mm<L, TPI>(m, g.sm, Wj, aj ? BI + static_cast<size_t>(j) * K : Bm + static_cast<size_t>(j) * K, aj ? keys->g_m : m->one);
      mm<L, TPI>(m, g.sm, Wj, Wj, (aj != atj) ? DI : m->one);
      for (int s = 0; s < j; s++) mm<L, TPI>(m, g.sm, Wj, Wj, Wj);
    }
    // ---- step 4h: S = prod_{j>i} W_j ; C[i+1] = c_sa[s][alpha_i] * d^(at_i - a_i) * beta_i^-1 * S^3 ; C[0] = g^delta_a * prod W
    copy_digits(g, S, m->one);
    for (int j = l - 1; j >= 0; j--) {
      const int aj = static_cast<int>((al >> j) & 1), atj = static_cast<int>((at >> j) & 1);
      const int ediff = atj - aj;
      uint32_t* Cj = Cm + static_cast<size_t>(j + 1) * K;
      mm<L, TPI>(m, g.sm, Cj, keys->c_sa[da][aj], ediff == 0 ? m->one : (ediff > 0 ? D : DI));
      mm<L, TPI>(m, g.sm, Cj, Cj, BI + static_cast<size_t>(j) * K);
      mm<L, TPI>(m, g.sm, S2, S, S);
      mm<L, TPI>(m, g.sm, S2, S2, S);
      mm<L, TPI>(m, g.sm, Cj, Cj, S2);
      mm<L, TPI>(m, g.sm, S, S, Wm + static_cast<size_t>(j) * K);
    }
    mm<L, TPI>(m, g.sm, Cm, S, da ? keys->g_m : m->one);
    // ---- step 4i: blind with r_j, shuffle, randomise, store.  Output position k takes masked[perm[k]].
    for (int k = 0; k <= l; k++) {
      const int src = perm ? perm[i * (l + 1) + k] : k;
      const uint32_t* Cs = Cm + static_cast<size_t>(src) * K;
      const uint32_t* res = Cs;
      if (do_blind) {
        group_modexp(g, T, Cs, r_blind + (i * (l + 1) + src) * rb_words, rb_words, ubits, wbits, table);
        res = T;
      }
      if (rho != nullptr) {
        copy_digits(g, U, res);
        group_fixed_base_mul(g, U, fb_table, nwin, rho + (i * (l + 1) + k) * ew, ew, fbw);
        res = U;
      }
      mm<L, TPI>(m, g.sm, S2, res, m->unit);
      canonical_store(g, S2, out + (i * (l + 1) + k) * kn, kn, active);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// KeyHolder-side Paillier with the factorisation (CRT over p^2, q^2): same results as the public formulas.
//   randomizer  rho^N mod N^2 :  b_P = (rho mod P)^(Q mod (P-1)) mod P,  s_P = b_P^P mod P^2   (P, Q) in {(p, q), (q, p)}
//                                rho^N = s_p * A + s_q * B mod N^2,  A = q^2 (q^-2 mod p^2), B = p^2 (p^-2 mod q^2)
//     ((rho mod p + k p)^p = (rho mod p)^p mod p^2, and rho^q mod p only needs the exponent mod p - 1.)
//   decryption  m = L_p(c^(p-1) mod p^2) h_p A' + L_q(c^(q-1) mod q^2) h_q B' mod N,  L_P(x) = (x - 1) / P,
//                                h_p = (-q)^-1 mod p,  A' = q (q^-1 mod p)   (and symmetrically)
// The exponentiations run in k_modexp on the p / p^2 / q / q^2 moduli; the kernels here recombine.
// ------------------------------------------------------------------------------------------------
// out[i] = (s_p[i] * A + s_q[i] * B) * (1 + plain[i] * N) mod N^2.   s_p, s_q: [M][kN] words; plain: [M][kN]; out: [M][2kN].
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_crt_encrypt_combine(const ModulusDev* __restrict__ m, const KeysDev* __restrict__ keys, const uint32_t* __restrict__ sp,
                      const uint32_t* __restrict__ sq, const uint32_t* __restrict__ plain, uint32_t* scratch, uint32_t* __restrict__ out, size_t M) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  const int kN = keys->kN;
  uint32_t* A = scratch + slot * 2 * K;
  uint32_t* T = A + K;
  const size_t rounds = (M + total - 1) / total;
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t i = rd * total + slot;
    const bool active = i < M;
    if (!active) i = M - 1;
    words_to_digits(g, A, sp + i * kN, kN);
    mm<L, TPI>(m, g.sm, A, A, keys->crt_enc[0]);            // (s_p A) R
    words_to_digits(g, T, sq + i * kN, kN);
    mm<L, TPI>(m, g.sm, T, T, keys->crt_enc[1]);            // (s_q B) R
    lazy_add(g, A, A, T);                                    // (rho^N) R
    group_paillier_encode(g, keys, T, plain + i * kN, kN);   // (1 + m N) R
    mm<L, TPI>(m, g.sm, A, A, T);
    mm<L, TPI>(m, g.sm, A, A, m->unit);
    canonical_store(g, A, out + i * 2 * kN, 2 * kN, active);
  }
}

// Lq[i][h] = (y_h[i] - 1) / P_h exactly (y = 1 mod P), Hensel division in radix 2^32.  y: [2][M][kN] words, Lq: [2][M][kP].
static __global__ void __launch_bounds__(64) k_crt_L(const KeysDev* __restrict__ keys, const uint32_t* __restrict__ y, uint32_t* __restrict__ Lq, size_t M) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= 2 * M) return;
  const int h = idx >= M ? 1 : 0;
  const int kP = keys->kP, kN = keys->kN;
  uint32_t t[MAXW / 4 + 2];
  const uint32_t* yi = y + idx * kN;
  const uint32_t* __restrict__ Pw = keys->Pw[h];
  uint32_t borrow = 1;
  for (int i = 0; i < kP; i++) {
    const uint32_t v = yi[i];
    t[i] = v - borrow;
    borrow = (v < borrow) ? 1u : 0u;
  }
  const uint32_t pinv = keys->Pinv32[h];
  for (int j = 0; j < kP; j++) {
    const uint32_t qj = t[j] * pinv;
    Lq[idx * kP + j] = qj;
    uint64_t carry = 0;
    for (int i = 0; i + j < kP; i++) {
      const uint64_t prod = static_cast<uint64_t>(qj) * Pw[i] + carry;
      const uint32_t lo = static_cast<uint32_t>(prod);
      carry = prod >> 32;
      const uint32_t old = t[i + j];
      t[i + j] = old - lo;
      carry += (old < lo) ? 1u : 0u;
    }
  }
}

// out[i] = Lp[i] * C_p + Lq[i] * C_q mod N   (one group per element, N configuration).  Lq: [2][M][kP]; out: [M][kN].
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_crt_decrypt_combine(const ModulusDev* __restrict__ m, const KeysDev* __restrict__ keys, const uint32_t* __restrict__ Lq, uint32_t* scratch,
                      uint32_t* __restrict__ out, size_t M) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  const int kN = keys->kN, kP = keys->kP;
  uint32_t* A = scratch + slot * 2 * K;
  uint32_t* T = A + K;
  const size_t rounds = (M + total - 1) / total;
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t i = rd * total + slot;
    const bool active = i < M;
    if (!active) i = M - 1;
    words_to_digits(g, A, Lq + i * kP, kP);
    mm<L, TPI>(m, g.sm, A, A, keys->crt_dec[0]);
    words_to_digits(g, T, Lq + (M + i) * kP, kP);
    mm<L, TPI>(m, g.sm, T, T, keys->crt_dec[1]);
    lazy_add(g, A, A, T);
    mm<L, TPI>(m, g.sm, A, A, m->unit);
    canonical_store(g, A, out + i * kN, kN, active);
  }
}

// flags[i] = (x[i] == 1) for k-word values (one thread per value): the comparison half of is_zero when the
// exponentiation c^{v_p} mod p ran on the FP64 engine
static __global__ void k_flag_is_one(const uint32_t* __restrict__ x, int k, uint8_t* __restrict__ flags, size_t M) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const uint32_t* xi = x + i * k;
  bool one = xi[0] == 1u;
  for (int j = 1; j < k && one; j++) one = xi[j] == 0u;
  flags[i] = one ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// DGK is_zero: flag = (c^{v_p} mod p == 1)   (keyholder.py:249 -> dgk.is_zero)
// ------------------------------------------------------------------------------------------------
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_dgk_is_zero(const ModulusDev* __restrict__ mp, const KeysDev* __restrict__ keys, const uint32_t* __restrict__ c, const uint32_t* __restrict__ vp,
              int vp_words, int vp_bits, int wbits, uint32_t* scratch, uint8_t* __restrict__ flags, size_t M, unsigned long long* queue,
              const uint16_t* __restrict__ prog, int nprog, int pw) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(mp, GI::warp_smem(smem));
  const size_t slot = GI::slot();
  const int kn = keys->kn;
  uint32_t* X = scratch + slot * modexp_scratch_words<L, TPI>(wbits);
  uint32_t* ACC = X + K;
  uint32_t* TMP = ACC + K;
  uint32_t* table = TMP + K;
  size_t i;
  bool active;
  while (GI::next(queue, M, i, active)) {
    load_mont(g, X, TMP, c + i * kn, kn);                        // (c mod p)~ from the double-width c
    group_modexp_any(g, ACC, X, vp, vp_words, vp_bits, wbits, table, prog, nprog, pw);
    mm<L, TPI>(mp, g.sm, TMP, ACC, mp->unit);
    stage_from(g, TMP);
    canonicalize_staged(g);
    const bool is_one = staged_equals_small(g, 1u);
    __syncwarp();
    if (active && g.t == 0) flags[i] = is_one ? 1 : 0;
  }
}

}  // namespace scb
