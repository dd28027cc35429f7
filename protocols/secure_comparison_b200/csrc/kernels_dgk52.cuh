// kernels_dgk52.cuh -- the DGK-side phases on the FP64 engine (bn52.cuh): fixed-base tables of h, bit encryption
// g^b * h^rho (keyholder.py:212-216, 230-233, 106-108) and the fused Initiator steps 4c..4i + randomize
// (initiator.py:272-516, 153-154).
//
// Same algorithms, same data flow and the same warp-uniformity rules as their IMAD-engine twins in kernels_generic.cuh /
// kernels_protocol.cuh (k_fb_*, k_dgk_encrypt_bits, k_step4_prefix, k_step4_main) -- digits are 52-bit integers held in
// doubles, every product is mmd() (memory to memory; one copy of the row loop per kernel).  The FP64 engine multiplies
// 1.36x faster than the IMAD one, and with the Paillier exponentiations in pair form these phases had grown to 14 % of a
// pass.  SCB_OPT_MODEXP_FP64 = 0 keeps everything on the IMAD kernels (bit-identical; bench.py compares the two every run).
#pragma once
#include "kernels_fb52.cuh"
#include "kernels_protocol.cuh"

namespace scb {

// Launch bounds.  Unlike the pair kernel these phases are mostly memory-to-memory products, and occupancy hides their loads:
// 2 blocks/SM (250 registers, a healthier row-loop schedule in tools/sass_deps.py) measured k52_step4_main 98 -> 116 ms and
// k52_dgk_encrypt_bits 26.8 -> 28.1 ms at B = 65,536 (profiles/r02k_bench_2blocks_dgk.json); 4 blocks/SM (128 registers, a
// few spills) measured k52_step4_main 97.7 -> 93.8 ms but k52_dgk_encrypt_bits 27.0 -> 29.0 ms (profiles/r02t_bounds.log).
// Hence: 4 for k52_step4_main, 3 for the others.  (With the single-copy form of step 4i, below: 3 / 4 / 5 blocks per SM =
// 90.2 / 84.6 / 126 ms -- at 5 the row loop spills; profiles/r02x_step4_blocks.log.)

// 52-bit Montgomery forms of the DGK constants the fused kernels use (device copy next to KeysDev)
struct Keys52Dev {
  double g_m[MAXK52];           // g~
  double c_sa[2][2][MAXK52];    // [s == -1][alpha_i] : (enc(s) * enc(alpha_i))~   (initiator.py:459-476)
};

// -------------------------------------------------------------------------------------------------
// DGK: encryption of bits g^b * h^rho.  element (i, e): e = 0 is d, e = 1..l are beta_0..beta_{l-1}.  rho: [B][l+1][ew] or null.
// -------------------------------------------------------------------------------------------------
template <int L, int TPI>
__global__ void __launch_bounds__(BLOCK_THREADS, (L <= 10 ? 3 : 2))
k52_dgk_encrypt_bits(const Modulus52Dev* __restrict__ m, const KeysDev* __restrict__ keys, const Keys52Dev* __restrict__ k52, const uint64_t* __restrict__ beta,
                     const uint8_t* __restrict__ dflag, int l, const double* fb_table, int nwin, int fbw, const uint32_t* __restrict__ rho, int ew,
                     double* scratch, uint32_t* __restrict__ out, size_t B, unsigned long long* queue) {
  extern __shared__ double smem52[];
  using GI = GroupIndex52<L, TPI>;
  constexpr int K = Geo52<L, TPI>::K;
  double* sm_group = GI::warp_smem(smem52) + (threadIdx.x & 31) / TPI;
  const size_t slot = GI::slot();
  const int kn = keys->kn;
  const size_t M = B * (l + 1);
  double* ACC = scratch + slot * 2 * K;
  double* TMP = ACC + K;
  size_t idx;
  bool active;
  while (GI::next(queue, M, idx, active)) {
    const size_t i = idx / (l + 1);
    const int e = static_cast<int>(idx % (l + 1));
    const int bit = (e == 0) ? dflag[i] : static_cast<int>((beta[i] >> (e - 1)) & 1ull);
    copy52<L, TPI>(ACC, bit ? k52->g_m : m->one);
    if (rho != nullptr) group_fixed_base_mul52<L, TPI>(m, sm_group, ACC, fb_table, nwin, rho + idx * ew, ew, fbw);
    mmd<L, TPI>(m, sm_group, TMP, ACC, m->unit);
    canonical_store52<L, TPI>(m, sm_group, TMP, out + idx * kn, kn, active);
  }
}

// acc = x~^e (fixed windows, per-group exponent) through memory-to-memory products only: the same operation sequence for
// every group of the warp (a zero window multiplies by 1~).  acc, xm: K-digit slots; table: (1 << wbits) slots.
// Used for the short blinding exponents of step 4i INSTEAD of the register-resident group_modexp52(): that one inlines a
// second 20 KB copy of the unrolled row loop next to mmd()'s, and this kernel alternates between the two all the time:
// k52_step4_main 93.8 -> 84.6 ms at B = 65,536 with this single-copy form although each product now goes through memory
// (profiles/r02w_step4_mem.log).  Most likely instruction-cache capacity (the opposite experiment: profiles/r02u_dual.log);
// not confirmed with a counter.
template <int L, int TPI>
__device__ __forceinline__ void group_modexp52_mem(const Modulus52Dev* __restrict__ m, double* sm_group, double* acc, const double* xm,
                                                   const uint32_t* __restrict__ e, int ewords, int ebits, int wbits, double* table) {
  constexpr int K = Geo52<L, TPI>::K;
  const int nwin = (ebits + wbits - 1) / wbits;
  if (nwin == 0) {
    copy52<L, TPI>(acc, m->one);
    return;
  }
  copy52<L, TPI>(table, m->one);
  copy52<L, TPI>(table + K, xm);
  for (int v = 2; v < (1 << wbits); v++) mmd<L, TPI>(m, sm_group, table + static_cast<size_t>(v) * K, table + static_cast<size_t>(v - 1) * K, xm);
  copy52<L, TPI>(acc, table + static_cast<size_t>(get_bits(e, ewords, (nwin - 1) * wbits, wbits)) * K);
  for (int wi = nwin - 2; wi >= 0; wi--) {
    for (int s = 0; s < wbits; s++) mmd<L, TPI>(m, sm_group, acc, acc, acc);
    mmd<L, TPI>(m, sm_group, acc, acc, table + static_cast<size_t>(get_bits(e, ewords, wi * wbits, wbits)) * K);
  }
}

// -------------------------------------------------------------------------------------------------
// Initiator steps 4c..4i (+ randomize) for one comparison per group.  Scratch slots per comparison (K digits each):
//   D, DI, S, S2, T, U (6), Bm[l], BI[l] (prefix products first), Wm[l], Cm[l+1], table[2^wbits] (blinding exponentiation)
// -------------------------------------------------------------------------------------------------
template <int L, int TPI>
__host__ __device__ constexpr size_t step4_scratch_doubles(int l, int wbits) {
  return static_cast<size_t>(Geo52<L, TPI>::K) * (6 + 4 * static_cast<size_t>(l) + 1 + (static_cast<size_t>(1) << wbits));
}

// pass A: D = r_small ? 1~ : d~ (step 4c), Bm[i] = beta_i~, prefix products P_0 = D, P_{i+1} = P_i * Bm[i]; total (canonical
// words) -> Tw[i].  The prefix products are parked in BI[0..l) (P_1..P_l) and DI (P_0 = D).  Scratch is indexed by INSTANCE.
template <int L, int TPI>
__global__ void __launch_bounds__(BLOCK_THREADS, (L <= 10 ? 3 : 2))
k52_step4_prefix(const Modulus52Dev* __restrict__ m, const KeysDev* __restrict__ keys, const uint32_t* __restrict__ db_enc, const uint8_t* __restrict__ r_small,
                 int l, int wbits, double* scratch, uint32_t* __restrict__ Tw, size_t B) {
  extern __shared__ double smem52[];
  using GI = GroupIndex52<L, TPI>;
  constexpr int K = Geo52<L, TPI>::K;
  double* sm_group = GI::warp_smem(smem52) + (threadIdx.x & 31) / TPI;
  const size_t slot = GI::slot(), total = static_cast<size_t>(gridDim.x) * GI::GROUPS_PER_BLOCK;
  const int kn = keys->kn;
  const size_t rounds = (B + total - 1) / total;
  const size_t stride = step4_scratch_doubles<L, TPI>(l, wbits);
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t i = rd * total + slot;
    const bool active = i < B;
    double* base = scratch + (active ? i : B + slot) * stride;      // surplus groups work in a private area behind
    if (!active) i = B - 1;
    double* D = base;
    double* DI = D + K;
    double* T = D + 4 * K;
    double* Bm = D + 6 * K;
    double* BI = Bm + static_cast<size_t>(l) * K;
    words_to_mont52<L, TPI>(m, sm_group, D, db_enc + i * (l + 1) * kn, kn);      // uniform across the warp (mmd shuffles warp-wide)
    if (r_small[i]) copy52<L, TPI>(D, m->one);                                    // step 4c: [d] <- [0] = 1
    copy52<L, TPI>(DI, D);
    const double* prev = DI;
    for (int j = 0; j < l; j++) {
      double* Bj = Bm + static_cast<size_t>(j) * K;
      words_to_mont52<L, TPI>(m, sm_group, Bj, db_enc + (i * (l + 1) + 1 + j) * kn, kn);
      mmd<L, TPI>(m, sm_group, BI + static_cast<size_t>(j) * K, prev, Bj);
      prev = BI + static_cast<size_t>(j) * K;
    }
    mmd<L, TPI>(m, sm_group, T, prev, m->unit);
    canonical_store52<L, TPI>(m, sm_group, T, Tw + i * kn, kn, active);
  }
}

// pass B: everything else.  TIw: the inverse of the total product as canonical words ([B][kn]).
template <int L, int TPI>
__global__ void __launch_bounds__(BLOCK_THREADS, (L <= 10 ? 4 : 2))
k52_step4_main(const Modulus52Dev* __restrict__ m, const KeysDev* __restrict__ keys, const Keys52Dev* __restrict__ k52, const uint32_t* __restrict__ TIw,
               const uint64_t* __restrict__ alpha, const uint64_t* __restrict__ alpha_tilde, const uint8_t* __restrict__ delta_a,
               const uint32_t* __restrict__ r_blind, int rb_words, const uint8_t* __restrict__ perm, const double* fb_table, int nwin, int fbw,
               const uint32_t* __restrict__ rho, int ew, int l, int ubits, int wbits, int do_blind, double* scratch, uint32_t* __restrict__ out, size_t B,
               unsigned long long* queue) {
  extern __shared__ double smem52[];
  using GI = GroupIndex52<L, TPI>;
  constexpr int K = Geo52<L, TPI>::K;
  double* sm_group = GI::warp_smem(smem52) + (threadIdx.x & 31) / TPI;
  const size_t slot = GI::slot();
  const int kn = keys->kn;
  const size_t stride = step4_scratch_doubles<L, TPI>(l, wbits);
  size_t i;
  bool active;
  while (GI::next(queue, B, i, active)) {
    double* base = scratch + (active ? i : B + slot) * stride;
    double* D = base;
    double* DI = D + K;
    double* S = D + 2 * K;
    double* S2 = D + 3 * K;
    double* T = D + 4 * K;
    double* U = D + 5 * K;
    double* Bm = D + 6 * K;
    double* BI = Bm + static_cast<size_t>(l) * K;
    double* Wm = BI + static_cast<size_t>(l) * K;
    double* Cm = Wm + static_cast<size_t>(l) * K;
    double* table = Cm + static_cast<size_t>(l + 1) * K;
    if (!active) {
      // surplus groups: make the private area well-defined (values are irrelevant, they are never stored)
      for (int s = 0; s < 6 + 4 * l + 1; s++) copy52<L, TPI>(base + static_cast<size_t>(s) * K, m->one);
    }
    const uint64_t al = alpha[i], at = alpha_tilde[i];
    const int da = delta_a[i] ? 1 : 0;
    // ---- back-substitution: I = (D * prod B)^-1 ; BI[j] held prefix P_{j+1}; DI held P_0 = D
    words_to_mont52<L, TPI>(m, sm_group, T, TIw + i * kn, kn);
    for (int j = l - 1; j >= 0; j--) {
      const double* Pprev = (j == 0) ? DI : BI + static_cast<size_t>(j - 1) * K;
      mmd<L, TPI>(m, sm_group, U, T, Pprev);                                       // beta_j^-1 = I * P_j
      mmd<L, TPI>(m, sm_group, T, T, Bm + static_cast<size_t>(j) * K);           // I *= beta_j
      copy52<L, TPI>(BI + static_cast<size_t>(j) * K, U);
    }
    copy52<L, TPI>(DI, T);                                                          // d^-1
    // ---- steps 4d, 4e, 4f:  W_j = ((alpha_j ? g * beta_j^-1 : beta_j) * (alpha_j != alpha~_j ? d^-1 : 1))^(2^j)
    for (int j = 0; j < l; j++) {
      const int aj = static_cast<int>((al >> j) & 1), atj = static_cast<int>((at >> j) & 1);
      double* Wj = Wm + static_cast<size_t>(j) * K;
      mmd<L, TPI>(m, sm_group, Wj, aj ? BI + static_cast<size_t>(j) * K : Bm + static_cast<size_t>(j) * K, aj ? k52->g_m : m->one);
      mmd<L, TPI>(m, sm_group, Wj, Wj, (aj != atj) ? DI : m->one);
      for (int s = 0; s < j; s++) mmd<L, TPI>(m, sm_group, Wj, Wj, Wj);
    }
    // ---- step 4h: S = prod_{j>i} W_j ; C[i+1] = c_sa[s][alpha_i] * d^(at_i - a_i) * beta_i^-1 * S^3 ; C[0] = g^delta_a * prod W
    copy52<L, TPI>(S, m->one);
    for (int j = l - 1; j >= 0; j--) {
      const int aj = static_cast<int>((al >> j) & 1), atj = static_cast<int>((at >> j) & 1);
      const int ediff = atj - aj;
      double* Cj = Cm + static_cast<size_t>(j + 1) * K;
      mmd<L, TPI>(m, sm_group, Cj, k52->c_sa[da][aj], ediff == 0 ? m->one : (ediff > 0 ? D : DI));
      mmd<L, TPI>(m, sm_group, Cj, Cj, BI + static_cast<size_t>(j) * K);
      mmd<L, TPI>(m, sm_group, S2, S, S);
      mmd<L, TPI>(m, sm_group, S2, S2, S);
      mmd<L, TPI>(m, sm_group, Cj, Cj, S2);
      mmd<L, TPI>(m, sm_group, S, S, Wm + static_cast<size_t>(j) * K);
    }
    mmd<L, TPI>(m, sm_group, Cm, S, da ? k52->g_m : m->one);
    // ---- step 4i: blind with r_j, shuffle, randomise, store.  Output position k takes masked[perm[k]].
    for (int k = 0; k <= l; k++) {
      const int src = perm ? perm[i * (l + 1) + k] : k;
      const double* Cs = Cm + static_cast<size_t>(src) * K;
      const double* res = Cs;
      if (do_blind) {
        group_modexp52_mem<L, TPI>(m, sm_group, T, Cs, r_blind + (i * (l + 1) + src) * rb_words, rb_words, ubits, wbits, table);
        res = T;
      }
      if (rho != nullptr) {
        copy52<L, TPI>(U, res);
        group_fixed_base_mul52<L, TPI>(m, sm_group, U, fb_table, nwin, rho + (i * (l + 1) + k) * ew, ew, fbw);
        res = U;
      }
      mmd<L, TPI>(m, sm_group, S2, res, m->unit);
      canonical_store52<L, TPI>(m, sm_group, S2, out + (i * (l + 1) + k) * kn, kn, active);
    }
  }
}

// Montgomery digits of the IMAD engine (what batch_modinv_mont produces) -> canonical words, for the hand-over of the
// batched inverse to the FP64 kernels.
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_mont29_to_words(const ModulusDev* __restrict__ m, const uint32_t* __restrict__ in, int kw, uint32_t* scratch, uint32_t* __restrict__ out, size_t B) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  uint32_t* A = scratch + slot * K;
  const size_t rounds = (B + total - 1) / total;
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t i = rd * total + slot;
    const bool active = i < B;
    if (!active) i = B - 1;
    mm<L, TPI>(m, g.sm, A, in + i * K, m->unit);
    canonical_store(g, A, out + i * kw, kw, active);
  }
}

}  // namespace scb
