// kernels_generic.cuh -- batched modexp / modmul / fixed-base / modinv kernels on top of bn29.cuh.
#pragma once
#include "bn29.cuh"

namespace scb {

constexpr int BLOCK_THREADS = 128;
constexpr int WARPS_PER_BLOCK = BLOCK_THREADS / 32;
// group kernels: cap registers so that 4 (L = 18) or 3 (L = 27) blocks of 128 threads fit on an SM;
// mm() itself needs ~110 / ~160 registers, the orchestration code around it may spill a little.
#define SCB_GROUP_BOUNDS(L) __launch_bounds__(BLOCK_THREADS, ((L) <= 18 ? 4 : 3))

template <int L, int TPI>
struct GroupIndex {
  static constexpr int GPW = Geo<L, TPI>::GPW;
  static constexpr int GROUPS_PER_BLOCK = GPW * WARPS_PER_BLOCK;
  __device__ static __forceinline__ size_t slot() {   // global group slot of this lane's group
    return (static_cast<size_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5)) * GPW + (threadIdx.x & 31) / TPI;
  }
  __device__ static __forceinline__ size_t total() { return static_cast<size_t>(gridDim.x) * GROUPS_PER_BLOCK; }
  // Dynamic work distribution for the long-running kernels: a warp takes the next GPW instances from a global
  // ticket counter (zeroed by the host before the launch), so no SM idles while another still has a queue
  // (static round-robin loses up to one whole modexp per group to the ragged last round).  Returns false when
  // the batch is exhausted; groups of the last warp past the end work on instance B-1 without storing.
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
  __device__ static __forceinline__ uint32_t* warp_smem(uint32_t* smem) {
    return smem + (threadIdx.x >> 5) * Geo<L, TPI>::WARP_SMEM_WORDS;
  }
};

// bits [pos, pos+w) of a little-endian word array (w <= 16), zero beyond `words`
__device__ __forceinline__ uint32_t get_bits(const uint32_t* __restrict__ e, int words, int pos, int w) {
  const int i = pos >> 5, sh = pos & 31;
  const uint32_t lo = (i < words) ? __ldg(e + i) : 0u;
  const uint32_t hi = (i + 1 < words) ? __ldg(e + i + 1) : 0u;
  return __funnelshift_r(lo, hi, sh) & ((1u << w) - 1u);
}

// -------------------------------------------------------------------------------------------------
// Fixed-window Montgomery exponentiation for ONE group: acc = x~^e (Montgomery form, lazy).
// xm, acc: K-digit slots; table: (1<<wbits) slots.  All in this group's global scratch.
//
// The running value and the modulus stay in REGISTERS across the whole exponentiation: a squaring re-stages the
// value into shared memory (it is its own b operand) and runs the row loop again -- no global-memory round trip,
// no operand reloads between the ~2,500 products of a 2048-bit exponent.  Only window products read their table
// entry from global scratch, and the table build stages x~ once for all 2^w - 2 products.  One software loop
// drives the three phases (table build, squarings, window products) so that the kernel holds a single inlined
// copy of the row loop; every branch depends on launch-uniform values only (exponent length, window size), the
// table INDEX is the only per-group quantity.
// -------------------------------------------------------------------------------------------------
template <int L, int TPI>
__device__ __forceinline__ void group_modexp(const Grp<L, TPI>& g, uint32_t* acc, const uint32_t* xm, const uint32_t* __restrict__ e,
                                             int ewords, int ebits, int wbits, uint32_t* table) {
  constexpr int K = Geo<L, TPI>::K;
  const int nwin = (ebits + wbits - 1) / wbits;
  if (nwin == 0) {
    copy_digits(g, acc, g.m->one);
    return;
  }
  const int t = g.t;
  uint32_t r[L], nv[L];
#pragma unroll
  for (int i = 0; i < L; i++) nv[i] = g.m->n[t * L + i];
  const uint32_t np = g.m->np;
  // table[0] = 1~, table[1] = x~; r = x~ and x~ staged as the b operand of the table build
  copy_digits(g, table, g.m->one);
#pragma unroll
  for (int i = 0; i < L; i++) r[i] = xm[t * L + i];
#pragma unroll
  for (int i = 0; i < L; i++) table[K + t * L + i] = r[i];
  stage_digits<L, TPI>(g.sm, r);
  const int ntab = (1 << wbits) - 2;                     // products of the table build: table[v] = table[v-1] * x~, v = 2..2^w-1
  const int nmain = (nwin - 1) * (wbits + 1);            // per remaining window: wbits squarings + 1 window product
  int pos = 0, wi = nwin - 2;                              // position inside the current window group / index of its window (no divisions in the loop)
#pragma unroll 1
  for (int step = 0; step < ntab + nmain; step++) {
    if (step >= ntab) {
      if (step == ntab) {
        // the table is complete (each lane wrote, and now reads, only its own digits): start from the top window
        const uint32_t win = get_bits(e, ewords, (nwin - 1) * wbits, wbits);
#pragma unroll
        for (int i = 0; i < L; i++) r[i] = table[static_cast<size_t>(win) * K + t * L + i];
      }
      if (pos < wbits) {
        stage_digits<L, TPI>(g.sm, r);                   // squaring: the value is its own b operand
      } else {
        const uint32_t win = get_bits(e, ewords, wi * wbits, wbits);
        uint32_t bv[L];
#pragma unroll
        for (int i = 0; i < L; i++) bv[i] = table[static_cast<size_t>(win) * K + t * L + i];
        stage_digits<L, TPI>(g.sm, bv);
      }
    }
    mm_rows<L, TPI>(r, nv, np, g.sm, r);
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
  if (nmain == 0) {                                      // a single window: the result is a table entry
    const uint32_t win = get_bits(e, ewords, 0, wbits);
#pragma unroll
    for (int i = 0; i < L; i++) r[i] = table[static_cast<size_t>(win) * K + t * L + i];
  }
#pragma unroll
  for (int i = 0; i < L; i++) acc[t * L + i] = r[i];
}

// acc = x~^e for a SHARED exponent compiled on the host into a sliding-window program (scb_ctx::register_exponent; format and
// rationale: group_modexp52_prog in kernels_f64.cuh).  table[i] = x~^(2i+1), i < 2^(pw-1).
template <int L, int TPI>
__device__ __forceinline__ void group_modexp_prog(const Grp<L, TPI>& g, uint32_t* acc, const uint32_t* xm, const uint16_t* __restrict__ prog,
                                                  int nprog, int pw, uint32_t* table) {
  constexpr int K = Geo<L, TPI>::K;
  const int t = g.t;
  const int T = 1 << (pw - 1);
  uint32_t r[L], nv[L];
#pragma unroll
  for (int i = 0; i < L; i++) nv[i] = g.m->n[t * L + i];
  const uint32_t np = g.m->np;
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
      stage_digits<L, TPI>(g.sm, r);                             // x~ * x~
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
        stage_digits<L, TPI>(g.sm, r);                           // squaring
        sq_left--;
      } else {
        uint32_t bv[L];
#pragma unroll
        for (int i = 0; i < L; i++) bv[i] = table[static_cast<size_t>(cur_idx) * K + t * L + i];
        stage_digits<L, TPI>(g.sm, bv);                          // window product
        need_mul = false;
      }
    }                                                            // phase 1: x~^2 is still staged
    mm_rows<L, TPI>(r, nv, np, g.sm, r);
    if (phase == 0) {
      stage_digits<L, TPI>(g.sm, r);                             // x~^2 stays staged for the whole table build
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

// one entry point for both loops: the program when the launch has one, fixed windows otherwise (launch-uniform choice)
template <int L, int TPI>
__device__ __forceinline__ void group_modexp_any(const Grp<L, TPI>& g, uint32_t* acc, const uint32_t* xm, const uint32_t* __restrict__ e,
                                                 int ewords, int ebits, int wbits, uint32_t* table, const uint16_t* __restrict__ prog,
                                                 int nprog, int pw) {
  if (prog != nullptr)
    group_modexp_prog(g, acc, xm, prog, nprog, pw, table);
  else
    group_modexp(g, acc, xm, e, ewords, ebits, wbits, table);
}

// dst = x~ (Montgomery form, lazy, < 4n) of the value x stored in `kw` words at `words`; x may be up to twice
// as wide as the modulus (x < 2^(58K)): x = lo + hi*R with K-digit halves, x*R = mm(lo, R^2) + mm(hi, R^3).
// Used for c mod p^2 / rho mod p in the KeyHolder's CRT paths and c mod p in is_zero.  The branch is on
// launch-uniform values only (mm shuffles across the whole warp).
template <int L, int TPI>
__device__ __forceinline__ void load_mont(const Grp<L, TPI>& g, uint32_t* dst, uint32_t* tmp, const uint32_t* __restrict__ words, int kw) {
  constexpr int K = Geo<L, TPI>::K;
  words_to_digits(g, dst, words, kw);
  mm<L, TPI>(g.m, g.sm, dst, dst, g.m->r2);
  if (32 * kw > W * K) {
#pragma unroll
    for (int i = 0; i < L; i++) tmp[g.t * L + i] = digit_from_words(words, kw, K + g.t * L + i);
    mm<L, TPI>(g.m, g.sm, tmp, tmp, g.m->r3);
    lazy_add(g, dst, dst, tmp);
  }
}

// scratch slots of one group in k_modexp: X, ACC, TMP, then the window table
template <int L, int TPI>
__host__ __device__ constexpr size_t modexp_scratch_words(int wbits) {
  return static_cast<size_t>(Geo<L, TPI>::K) * (3 + (static_cast<size_t>(1) << wbits));
}

// out[i] = (mul_in[i] *) base[i]^exp[i] mod n
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_modexp(const ModulusDev* __restrict__ m, const uint32_t* __restrict__ base, int kbase, const uint32_t* __restrict__ exp,
         int ewords, size_t exp_stride, int ebits, int wbits, uint32_t* scratch, const uint32_t* __restrict__ mul_in,
         uint32_t* __restrict__ out, int kout, size_t B, unsigned long long* queue, const uint16_t* __restrict__ prog, int nprog, int pw) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot();
  uint32_t* X = scratch + slot * modexp_scratch_words<L, TPI>(wbits);
  uint32_t* ACC = X + K;
  uint32_t* TMP = ACC + K;
  uint32_t* table = TMP + K;
  size_t inst;
  bool active;
  while (GI::next(queue, B, inst, active)) {
    load_mont(g, X, TMP, base + inst * kbase, kbase);            // x~ = x*R (reduces double-width bases)
    group_modexp_any(g, ACC, X, exp + inst * exp_stride, ewords, ebits, wbits, table, prog, nprog, pw);
    if (mul_in != nullptr) {
      words_to_digits(g, TMP, mul_in + inst * kout, kout);
      mm<L, TPI>(m, g.sm, TMP, TMP, m->r2);
      mm<L, TPI>(m, g.sm, ACC, ACC, TMP);
    }
    mm<L, TPI>(m, g.sm, TMP, ACC, m->unit);                     // leave Montgomery form
    canonical_store(g, TMP, out + inst * kout, kout, active);
  }
}

// out[i] = a[i] * b[i] mod n   (a_stride / b_stride = 0 broadcasts one operand)
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_modmul(const ModulusDev* __restrict__ m, const uint32_t* __restrict__ a, size_t a_stride, const uint32_t* __restrict__ b,
         size_t b_stride, int k32, uint32_t* scratch, uint32_t* __restrict__ out, size_t B) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot(), total = GI::total();
  uint32_t* A = scratch + slot * 2 * K;
  uint32_t* Bv = A + K;
  const size_t rounds = (B + total - 1) / total;
  for (size_t rd = 0; rd < rounds; rd++) {
    size_t inst = rd * total + slot;
    const bool active = inst < B;
    if (!active) inst = B - 1;
    words_to_digits(g, A, a + inst * a_stride, k32);
    words_to_digits(g, Bv, b + inst * b_stride, k32);
    mm<L, TPI>(m, g.sm, A, A, m->r2);                            // a*R
    mm<L, TPI>(m, g.sm, A, A, Bv);                               // a*R*b/R = a*b
    canonical_store(g, A, out + inst * k32, k32, active);
  }
}

// -------------------------------------------------------------------------------------------------
// Fixed-base tables: T[j][v] = base^(v * 256^j) in Montgomery form, j < nwin, v < 256.
// -------------------------------------------------------------------------------------------------
// step 1 (one group): T[j][1] for all j by repeated squaring.
template <int L, int TPI>
__global__ void __launch_bounds__(32) k_fb_bases(const ModulusDev* __restrict__ m, const uint32_t* __restrict__ base, int k32,
                                                  uint32_t* table, int nwin) {
  extern __shared__ uint32_t smem[];
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, smem);
  // every group of the warp computes the same thing into its own scratch copy; group 0's copy is the
  // table itself.  (Keeps all 32 lanes converged for mm's shuffles without special cases.)
  const int grp = (threadIdx.x & 31) / TPI;
  uint32_t* T = table + (grp == 0 ? 0 : static_cast<size_t>(nwin) * 256 * K + static_cast<size_t>(grp - 1) * K * 2);
  // group 0 writes T[j][1]; the other groups ping-pong in two private slots behind the table
  uint32_t* cur = (grp == 0) ? T + K : T;
  words_to_digits(g, cur, base, k32);
  mm<L, TPI>(m, g.sm, cur, cur, m->r2);
  for (int j = 1; j < nwin; j++) {
    uint32_t* nxt = (grp == 0) ? table + (static_cast<size_t>(j) * 256 + 1) * K : (cur == T ? T + K : T);
    mm<L, TPI>(m, g.sm, nxt, cur, cur);
    for (int s = 1; s < 8; s++) mm<L, TPI>(m, g.sm, nxt, nxt, nxt);
    cur = nxt;
  }
}

// step 2 (one group per window j): T[j][0] = 1~, T[j][v] = T[j][v-1] * T[j][1].
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L) k_fb_fill(const ModulusDev* __restrict__ m, uint32_t* table, int nwin, uint32_t* dump) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t j = GI::slot();
  const bool active = j < static_cast<size_t>(nwin);
  // inactive groups recompute window 0 into a private dump area (keeps the warp converged)
  uint32_t* T = active ? table + j * 256 * K : dump + (j - nwin) * 2 * K;
  const uint32_t* base1 = active ? T + K : table + K;
  if (active) copy_digits(g, T, m->one);
  const uint32_t* prev = base1;
  for (int v = 2; v < 256; v++) {
    uint32_t* dst = active ? T + static_cast<size_t>(v) * K : T + (v & 1) * K;
    mm<L, TPI>(m, g.sm, dst, prev, base1);
    prev = dst;
  }
}

// step 3 (optional): 16-bit windows from pairs of 8-bit windows, T16[j][v] = T8[2j][v & 255] * T8[2j+1][v >> 8]:
// nwin16 * 65536 entries (472 MB for DGK-2048's 400-bit randomizer exponents -- HBM capacity is what a B200 has
// plenty of), one independent product each, so h^rho costs 25 products instead of 50.
template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L) k_fb16_fill(const ModulusDev* __restrict__ m, const uint32_t* table8, uint32_t* table16, size_t M,
                                                uint32_t* dump, unsigned long long* queue) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot();
  size_t idx;
  bool active;
  while (GI::next(queue, M, idx, active)) {
    const size_t j = idx >> 16, v = idx & 0xffff;
    uint32_t* dst = active ? table16 + idx * K : dump + slot * K;
    mm<L, TPI>(m, g.sm, dst, table8 + ((2 * j) * 256 + (v & 255)) * K, table8 + ((2 * j + 1) * 256 + (v >> 8)) * K);
  }
}

// acc (Montgomery, lazy) *= base^exp using the window table (wbits = 8 or 16).  e: this group's exponent words.
template <int L, int TPI>
__device__ __forceinline__ void group_fixed_base_mul(const Grp<L, TPI>& g, uint32_t* acc, const uint32_t* table, int nwin,
                                                     const uint32_t* __restrict__ e, int ewords, int wbits = 8) {
  constexpr int K = Geo<L, TPI>::K;
  for (int j = 0; j < nwin; j++) {
    const uint32_t v = get_bits(e, ewords, j * wbits, wbits);
    mm<L, TPI>(g.m, g.sm, acc, acc, table + ((static_cast<size_t>(j) << wbits) + v) * K);
  }
}

template <int L, int TPI>
__global__ void SCB_GROUP_BOUNDS(L)
k_fixed_base(const ModulusDev* __restrict__ m, const uint32_t* table, int nwin, int fbw, const uint32_t* __restrict__ exp, int ewords,
             const uint32_t* __restrict__ mul_in, uint32_t* scratch, uint32_t* __restrict__ out, int k32, size_t B, unsigned long long* queue) {
  extern __shared__ uint32_t smem[];
  using GI = GroupIndex<L, TPI>;
  constexpr int K = Geo<L, TPI>::K;
  const Grp<L, TPI> g = grp_init<L, TPI>(m, GI::warp_smem(smem));
  const size_t slot = GI::slot();
  uint32_t* ACC = scratch + slot * K;
  size_t inst;
  bool active;
  while (GI::next(queue, B, inst, active)) {
    if (mul_in != nullptr) {
      words_to_digits(g, ACC, mul_in + inst * k32, k32);
      mm<L, TPI>(m, g.sm, ACC, ACC, m->r2);
    } else {
      copy_digits(g, ACC, m->one);
    }
    group_fixed_base_mul(g, ACC, table, nwin, exp + inst * ewords, ewords, fbw);
    mm<L, TPI>(m, g.sm, ACC, ACC, m->unit);
    canonical_store(g, ACC, out + inst * k32, k32, active);
  }
}

// -------------------------------------------------------------------------------------------------
// Modular inverse, one thread per element: Bernstein-Yang "safegcd" divsteps in batches of 30 (the
// transition-matrix formulation; data flow pinned by tests/safegcd_model.py).  Big numbers are signed
// 30-bit limbs in thread-local memory; one outer iteration = 30 branch-free divsteps on the low limbs,
// then two matrix-vector updates over the limbs ((f, g) exactly divided by 2^30, (d, e) mod n).
// ~2.05 divsteps per modulus bit, i.e. ~10 * len * (bits / 14.6) multiply-adds per inverse: microseconds
// of work where a bit-serial binary Euclid needs seconds.  Not on the throughput-critical path anyway:
// the protocol kernels batch their inversions with Montgomery's trick (one inverse per 32 values).
// -------------------------------------------------------------------------------------------------
constexpr int INV_MAXW = 260;     // 32-bit words of the widest operand
constexpr uint32_t M30 = (1u << 30) - 1u;

struct Trans30 {
  int32_t u, v, q, r;
};

// 30 divsteps on the low bits of f, g; zeta = -(delta + 1/2).  t * (f, g) = 2^30 * (f', g').
__device__ __forceinline__ int32_t divsteps_30(int32_t zeta, uint32_t f, uint32_t g, Trans30& t) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
#pragma unroll 6
  for (int i = 0; i < 30; i++) {
    uint32_t c1 = static_cast<uint32_t>(zeta >> 31);
    const uint32_t c2 = 0u - (g & 1u);
    const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
    g += x & c2;
    q += y & c2;
    r += z & c2;
    c1 &= c2;
    zeta = (zeta ^ static_cast<int32_t>(c1)) - 1;
    f += g & c1;
    u += q & c1;
    v += r & c1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  t.u = static_cast<int32_t>(u);
  t.v = static_cast<int32_t>(v);
  t.q = static_cast<int32_t>(q);
  t.r = static_cast<int32_t>(r);
  return zeta;
}

__device__ __forceinline__ void words_to_limbs30(int32_t* dst, int len, const uint32_t* __restrict__ w, int k) {
  uint64_t acc = 0;
  int have = 0, wi = 0;
  for (int i = 0; i < len; i++) {
    if (have < 30 && wi < k) {
      acc |= static_cast<uint64_t>(w[wi++]) << have;
      have += 32;
    }
    dst[i] = static_cast<int32_t>(static_cast<uint32_t>(acc) & M30);
    acc >>= 30;
    have = have >= 30 ? have - 30 : 0;
  }
}

static __global__ void __launch_bounds__(64) k_modinv(const ModulusDev* __restrict__ m, int k, const uint32_t* __restrict__ a,
                                                      uint32_t* __restrict__ out, uint8_t* __restrict__ ok, size_t B) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B) return;
  int32_t f[INV_MAXL], g[INV_MAXL], d[INV_MAXL], e[INV_MAXL];
  const int len = m->len30;
  const int32_t* __restrict__ mod = m->n30;
  const uint32_t ninv30 = m->ninv30;
  words_to_limbs30(g, len, a + idx * k, k);
  for (int i = 0; i < len; i++) {
    f[i] = mod[i];
    d[i] = 0;
    e[i] = 0;
  }
  e[0] = 1;
  int32_t zeta = -1;
  bool done = false;
  const int cap = m->inv_cap;
  for (int it = 0; it < cap && !done; it++) {
    Trans30 t;
    zeta = divsteps_30(zeta, static_cast<uint32_t>(f[0]) | (static_cast<uint32_t>(f[1]) << 30),
                       static_cast<uint32_t>(g[0]) | (static_cast<uint32_t>(g[1]) << 30), t);
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    // (d, e) <- t * (d, e) / 2^30 mod n, kept in (-2n, n)
    {
      const int32_t sd = d[len - 1] >> 31, se = e[len - 1] >> 31;
      int32_t md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);
      int64_t cd = u * d[0] + v * e[0], ce = q * d[0] + r * e[0];
      md -= static_cast<int32_t>((ninv30 * static_cast<uint32_t>(cd) + static_cast<uint32_t>(md)) & M30);
      me -= static_cast<int32_t>((ninv30 * static_cast<uint32_t>(ce) + static_cast<uint32_t>(me)) & M30);
      cd += static_cast<int64_t>(mod[0]) * md;
      ce += static_cast<int64_t>(mod[0]) * me;
      cd >>= 30;
      ce >>= 30;
      for (int i = 1; i < len; i++) {
        const int64_t di = d[i], ei = e[i], mi = mod[i];
        cd += u * di + v * ei + mi * md;
        ce += q * di + r * ei + mi * me;
        d[i - 1] = static_cast<int32_t>(static_cast<uint32_t>(cd) & M30);
        e[i - 1] = static_cast<int32_t>(static_cast<uint32_t>(ce) & M30);
        cd >>= 30;
        ce >>= 30;
      }
      d[len - 1] = static_cast<int32_t>(cd);
      e[len - 1] = static_cast<int32_t>(ce);
    }
    // (f, g) <- t * (f, g) / 2^30 (exact)
    {
      int64_t cf = u * f[0] + v * g[0], cg = q * f[0] + r * g[0];
      cf >>= 30;
      cg >>= 30;
      uint32_t any = 0;
      for (int i = 1; i < len; i++) {
        const int64_t fi = f[i], gi = g[i];
        cf += u * fi + v * gi;
        cg += q * fi + r * gi;
        f[i - 1] = static_cast<int32_t>(static_cast<uint32_t>(cf) & M30);
        const uint32_t gl = static_cast<uint32_t>(cg) & M30;
        g[i - 1] = static_cast<int32_t>(gl);
        any |= gl;
        cf >>= 30;
        cg >>= 30;
      }
      f[len - 1] = static_cast<int32_t>(cf);
      g[len - 1] = static_cast<int32_t>(cg);
      any |= static_cast<uint32_t>(cg);
      done = any == 0;
    }
  }
  // f = +-gcd(a, n): the inverse exists iff f = +-1
  const bool fneg = f[len - 1] < 0;
  bool unit = done;
  if (fneg) {
    for (int i = 0; i < len - 1 && unit; i++) unit = static_cast<uint32_t>(f[i]) == M30;
    unit = unit && f[len - 1] == -1;
  } else {
    unit = unit && f[0] == 1;
    for (int i = 1; i < len && unit; i++) unit = f[i] == 0;
  }
  // d * sign(f) mod n into [0, n)
  for (int pass = 0; pass < 2; pass++) {
    const int32_t add = d[len - 1] >> 31;
    const int32_t neg = (pass == 0 && fneg) ? -1 : 0;
    for (int i = 0; i < len; i++) {
      int32_t x = d[i] + (mod[i] & add);
      d[i] = (x ^ neg) - neg;
    }
    for (int i = 0; i < len - 1; i++) {
      d[i + 1] += d[i] >> 30;
      d[i] &= static_cast<int32_t>(M30);
    }
  }
  // limbs -> words
  uint64_t acc = 0;
  int have = 0, li = 0;
  for (int w = 0; w < k; w++) {
    while (have < 32 && li < len) {
      acc |= static_cast<uint64_t>(static_cast<uint32_t>(d[li++])) << have;
      have += 30;
    }
    out[idx * k + w] = unit ? static_cast<uint32_t>(acc) : 0u;
    acc >>= 32;
    have -= 32;
  }
  ok[idx] = unit ? 1 : 0;
}

}  // namespace scb
