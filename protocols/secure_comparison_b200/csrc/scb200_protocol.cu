// scb200_protocol.cu -- key loading and the fused protocol phases of the C ABI (include/scb200.h).
#include "../../../include/scb200.h"

#include <algorithm>
#include <cstddef>
#include <cstring>
#include <vector>

#include "host_ctx.hpp"
#include "kernels_protocol.cuh"
#include "kernels_dgk52.cuh"

using namespace scb;

namespace {

constexpr int INV_CHUNK = 32;

// scratch buffer roles (scb_ctx::scratch indices)
enum { SC_GROUP = 0, SC_SMALL = 1, SC_STAGE = 2, SC_INV_VP = 3, SC_INV_SMALL = 4, SC_PHASE = 5, SC_STEP4 = 6, SC_DIGITS = 7, SC_CRT = 8 };

int window_bits_for(int ebits) {
  if (ebits <= 6) return 1;
  if (ebits <= 24) return 2;
  if (ebits <= 80) return 3;
  if (ebits <= 240) return 4;
  if (ebits <= 1100) return 5;
  return 6;
}

template <int L, int TPI>
struct Cfg {
  using G = Geo<L, TPI>;
  static constexpr size_t SMEM = sizeof(uint32_t) * G::WARP_SMEM_WORDS * WARPS_PER_BLOCK;
  static constexpr int GPB = GroupIndex<L, TPI>::GROUPS_PER_BLOCK;
  static constexpr int K = G::K;
};

#define SCB_DISPATCH_K(K, CALL)                                                   \
  switch (K) {                                                                    \
    case 36: return CALL(18, 2);                                                  \
    case 54: return CALL(27, 2);                                                  \
    case 72: return CALL(18, 4);                                                  \
    case 108: return CALL(27, 4);                                                 \
    case 144: return CALL(18, 8);                                                 \
    case 216: return CALL(27, 8);                                                 \
    case 288: return CALL(18, 16);                                                \
    default: return fail(ctx, SCB_EINVAL, "unsupported digit count %d", K);       \
  }

inline size_t pad256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }

// bump allocator over one scratch buffer
struct Arena {
  char* base = nullptr;
  size_t off = 0;
  template <typename T>
  T* take(size_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += pad256(count * sizeof(T));
    return p;
  }
};

int get_arena(scb_ctx* ctx, int which, size_t bytes, Arena* a) {
  void* p = nullptr;
  int rc = ctx->scratch(which, bytes + 4096, &p);
  if (rc) return rc;
  a->base = static_cast<char*>(p);
  a->off = 0;
  return SCB_OK;
}

int need_keys(scb_ctx* ctx, bool paillier_pub, bool paillier_priv, bool dgk_pub, bool dgk_priv, const char* who) {
  if (!ctx) return SCB_EINVAL;
  if ((paillier_pub && !ctx->have_paillier_pub) || (paillier_priv && !ctx->have_paillier_priv) || (dgk_pub && !ctx->have_dgk_pub) ||
      (dgk_priv && !ctx->have_dgk_priv))
    return fail(ctx, SCB_ESTATE, "%s: required keys have not been loaded", who);
  cudaSetDevice(ctx->device);
  return SCB_OK;
}

int upload_keys(scb_ctx* ctx) {
  if (!ctx->keys_dev) {
    cudaError_t e = cudaMalloc(&ctx->keys_dev, sizeof(KeysDev));
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(keys)");
  }
  cudaError_t e = cudaMemcpyAsync(ctx->keys_dev, ctx->keys_host.data(), sizeof(KeysDev), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "upload keys");
  return SCB_OK;
}

// device address of the Paillier modulus N (kN words) inside the uploaded key block: the public exponent of rho^N
const uint32_t* keys_N_words(scb_ctx* ctx) {
  return reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(ctx->keys_dev) + offsetof(KeysDev, Nw));
}

KeysDev* host_keys(scb_ctx* ctx);

// (Re)compile the sliding-window programs of every shared exponent the library owns: the public exponent N of rho^N, lambda,
// the six CRT exponents, v_p.  Called whenever a key is loaded (device addresses and values may both have changed).
int rebuild_exponent_programs(scb_ctx* ctx) {
  ctx->clear_exponent_programs();
  int rc = SCB_OK;
  if (ctx->have_paillier_pub && ctx->keys_dev) {
    const KeysDev* hk = host_keys(ctx);
    if ((rc = ctx->register_exponent(keys_N_words(ctx), hk->Nw, hk->kN))) return rc;
  }
  if (ctx->secret_dev && !ctx->paillier_lambda.empty())
    if ((rc = ctx->register_exponent(ctx->secret_dev, ctx->paillier_lambda.data(), static_cast<int>(ctx->paillier_lambda.size())))) return rc;
  if (ctx->have_paillier_crt && ctx->crt_exp_dev) {
    const int kP = host_keys(ctx)->kP;
    for (int j = 0; j < 6; j++)
      if ((rc = ctx->register_exponent(ctx->crt_exp_dev + static_cast<size_t>(j) * kP, ctx->crt_exp_host.data() + static_cast<size_t>(j) * kP, kP))) return rc;
  }
  if (ctx->have_dgk_priv && ctx->vp_dev)
    if ((rc = ctx->register_exponent(ctx->vp_dev, ctx->dgk_vp.data(), static_cast<int>(ctx->dgk_vp.size())))) return rc;
  return SCB_OK;
}

KeysDev* host_keys(scb_ctx* ctx) {
  if (ctx->keys_host.size() != sizeof(KeysDev)) ctx->keys_host.assign(sizeof(KeysDev), 0);
  return reinterpret_cast<KeysDev*>(ctx->keys_host.data());
}

// Montgomery digits of `count` constants given as words (device copy made here); result copied back to host memory `out_host`.
template <int L, int TPI>
int const_to_mont(scb_ctx* ctx, const scb_ctx::Mod& md, const uint32_t* words_host, int k32, int count, int use_r3, uint32_t* out_host) {
  using C = Cfg<L, TPI>;
  Arena a;
  int rc = get_arena(ctx, SC_SMALL, pad256(sizeof(uint32_t) * k32 * count) + pad256(sizeof(uint32_t) * C::K * (count + 32)), &a);
  if (rc) return rc;
  uint32_t* dw = a.take<uint32_t>(static_cast<size_t>(k32) * count);
  uint32_t* dd = a.take<uint32_t>(static_cast<size_t>(C::K) * (count + 32));
  cudaError_t e = cudaMemcpyAsync(dw, words_host, sizeof(uint32_t) * k32 * count, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "upload constants");
  k_const_to_mont<L, TPI><<<1, 32, sizeof(uint32_t) * C::G::WARP_SMEM_WORDS, ctx->stream>>>(md.dev, dw, k32, count, dd, use_r3);
  rc = ctx->launched("k_const_to_mont");
  if (rc) return rc;
  e = cudaMemcpyAsync(out_host, dd, sizeof(uint32_t) * C::K * count, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "download constants");
  return SCB_OK;
}

template <int L, int TPI>
int const_mul(scb_ctx* ctx, const scb_ctx::Mod& md, const uint32_t* a_host, const uint32_t* b_host, int count, uint32_t* out_host) {
  using C = Cfg<L, TPI>;
  Arena ar;
  int rc = get_arena(ctx, SC_SMALL, 3 * pad256(sizeof(uint32_t) * C::K * count) + pad256(sizeof(uint32_t) * C::K * 32), &ar);
  if (rc) return rc;
  uint32_t* da = ar.take<uint32_t>(static_cast<size_t>(C::K) * count);
  uint32_t* db = ar.take<uint32_t>(static_cast<size_t>(C::K) * count);
  uint32_t* dout = ar.take<uint32_t>(static_cast<size_t>(C::K) * count);
  uint32_t* dump = ar.take<uint32_t>(static_cast<size_t>(C::K) * 32);
  cudaError_t e = cudaMemcpyAsync(da, a_host, sizeof(uint32_t) * C::K * count, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(db, b_host, sizeof(uint32_t) * C::K * count, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "upload constants");
  k_const_mul<L, TPI><<<1, 32, sizeof(uint32_t) * C::G::WARP_SMEM_WORDS, ctx->stream>>>(md.dev, da, db, count, dout, dump);
  rc = ctx->launched("k_const_mul");
  if (rc) return rc;
  e = cudaMemcpyAsync(out_host, dout, sizeof(uint32_t) * C::K * count, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "download constants");
  return SCB_OK;
}

// ------------------------------------------------------------------------------------------------
// batched inversion: in (device words [M][k32]) -> OUT (device Montgomery digits [M][K]).
// Falls back to one extended-Euclid per element when a chunk product is not invertible, so that
// ok_elem (device, [M]) is exact.  *any_fail tells the caller whether some element failed.
// ------------------------------------------------------------------------------------------------
template <int L, int TPI>
int batch_modinv_mont(scb_ctx* ctx, int slot, const uint32_t* in, size_t M, uint32_t* OUT, uint8_t* ok_elem, bool* any_fail) {
  using C = Cfg<L, TPI>;
  const scb_ctx::Mod& md = ctx->mod[slot];
  const int k32 = md.k32;
  const int chunk = static_cast<int>(std::min<size_t>(INV_CHUNK, M));
  const size_t G = (M + chunk - 1) / chunk;
  const int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_inv_prefix<L, TPI>), C::SMEM, G, C::GPB);
  const int grid2 = ctx->grid_for(reinterpret_cast<const void*>(&k_inv_backsub<L, TPI>), C::SMEM, G, C::GPB);
  const size_t groups = static_cast<size_t>(std::max(grid, grid2)) * C::GPB;
  Arena vp, sm;
  int rc = get_arena(ctx, SC_INV_VP, pad256(sizeof(uint32_t) * M * C::K), &vp);
  if (rc) return rc;
  rc = get_arena(ctx, SC_INV_SMALL, 2 * pad256(sizeof(uint32_t) * G * k32) + pad256(G) + pad256(sizeof(uint32_t) * groups * 2 * C::K), &sm);
  if (rc) return rc;
  uint32_t* P = vp.take<uint32_t>(M * C::K);
  uint32_t* Tw = sm.take<uint32_t>(G * k32);
  uint32_t* Iw = sm.take<uint32_t>(G * k32);
  uint8_t* okc = sm.take<uint8_t>(G);
  uint32_t* tmp = sm.take<uint32_t>(groups * 2 * C::K);
  uint32_t* V = OUT;   // the Montgomery forms of the inputs live in the output buffer until back-substitution replaces them
  k_inv_prefix<L, TPI><<<grid, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, in, k32, V, P, tmp, Tw, M, chunk);
  if ((rc = ctx->launched("k_inv_prefix"))) return rc;
  k_modinv<<<static_cast<unsigned>((G + 63) / 64), 64, 0, ctx->stream>>>(md.dev, k32, Tw, Iw, okc, G);
  if ((rc = ctx->launched("k_modinv"))) return rc;
  k_inv_backsub<L, TPI><<<grid2, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, V, P, Iw, k32, okc, tmp, OUT, ok_elem, M, chunk);
  if ((rc = ctx->launched("k_inv_backsub"))) return rc;
  // did any chunk fail?  (G bytes back to the host)
  std::vector<uint8_t> hok(G);
  cudaError_t e = cudaMemcpyAsync(hok.data(), okc, G, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "download inversion flags");
  bool failed = false;
  for (size_t i = 0; i < G; i++) failed = failed || !hok[i];
  if (any_fail) *any_fail = failed;
  if (!failed) return SCB_OK;
  // slow exact path: one inversion per element, then to Montgomery form
  Arena slow;
  rc = get_arena(ctx, SC_INV_VP, pad256(sizeof(uint32_t) * M * k32) + pad256(sizeof(uint32_t) * groups * C::K), &slow);
  if (rc) return rc;
  uint32_t* invw = slow.take<uint32_t>(M * k32);
  uint32_t* tmp2 = slow.take<uint32_t>(groups * C::K);
  k_modinv<<<static_cast<unsigned>((M + 63) / 64), 64, 0, ctx->stream>>>(md.dev, k32, in, invw, ok_elem, M);
  if ((rc = ctx->launched("k_modinv"))) return rc;
  const int grid3 = ctx->grid_for(reinterpret_cast<const void*>(&k_to_mont<L, TPI>), C::SMEM, M, C::GPB);
  k_to_mont<L, TPI><<<std::min(grid3, std::max(grid, grid2)), BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, invw, k32, tmp2, OUT, M);
  return ctx->launched("k_to_mont");
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// key loading
// ------------------------------------------------------------------------------------------------
extern "C" int scb_paillier_set_public(scb_ctx* ctx, const uint32_t* n, int kN) {
  if (!ctx || !n || kN <= 0 || 2 * kN > MAXW - 4) return fail(ctx, SCB_EINVAL, "scb_paillier_set_public: bad argument");
  cudaSetDevice(ctx->device);
  // N^2 by schoolbook multiplication
  std::vector<uint32_t> n2(2 * kN, 0);
  for (int i = 0; i < kN; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < kN; j++) {
      const uint64_t t = static_cast<uint64_t>(n[i]) * n[j] + n2[i + j] + carry;
      n2[i + j] = static_cast<uint32_t>(t);
      carry = t >> 32;
    }
    n2[i + kN] = static_cast<uint32_t>(carry);
  }
  int rc = scb_modulus_load(ctx, SCB_MOD_PAILLIER_N2, n2.data(), 2 * kN);
  if (rc) return rc;
  rc = scb_modulus_load(ctx, SCB_MOD_PAILLIER_N, n, kN);
  if (rc) return rc;
  if ((rc = scb::link_square_modulus(ctx, SCB_MOD_PAILLIER_N2, SCB_MOD_PAILLIER_N))) return rc;
  KeysDev* hk = host_keys(ctx);
  hk->kN = kN;
  memset(hk->Nw, 0, sizeof(hk->Nw));
  memset(hk->halfw, 0, sizeof(hk->halfw));
  memcpy(hk->Nw, n, sizeof(uint32_t) * kN);
  for (int i = 0; i < kN; i++) hk->halfw[i] = (n[i] >> 1) | (i + 1 < kN ? n[i + 1] << 31 : 0u);   // (N-1)/2 = N >> 1 for odd N
  uint32_t inv = n[0];
  for (int i = 0; i < 5; i++) inv *= 2u - n[0] * inv;
  hk->ninv32 = inv;
  // NR2 = N * R^2 mod N^2 = mm(N, R^3)
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_PAILLIER_N2];
  std::vector<uint32_t> nw(2 * kN, 0);
  memcpy(nw.data(), n, sizeof(uint32_t) * kN);
#define CALL(L, TPI) const_to_mont<L, TPI>(ctx, md, nw.data(), 2 * kN, 1, 1, hk->NR2)
  auto run = [&]() -> int { SCB_DISPATCH_K(md.K, CALL) };
#undef CALL
  rc = run();
  if (rc) return rc;
  ctx->have_paillier_pub = true;
  ctx->have_paillier_priv = false;
  ctx->have_paillier_crt = false;
  ctx->paillier_lambda.clear();
  if ((rc = upload_keys(ctx))) return rc;
  return rebuild_exponent_programs(ctx);
}

extern "C" int scb_paillier_set_private(scb_ctx* ctx, const uint32_t* lambda, int lambda_words, const uint32_t* mu) {
  int rc = need_keys(ctx, true, false, false, false, "scb_paillier_set_private");
  if (rc) return rc;
  if (!lambda || !mu || lambda_words <= 0) return fail(ctx, SCB_EINVAL, "scb_paillier_set_private: bad argument");
  const int kN = host_keys(ctx)->kN;
  ctx->paillier_lambda.assign(lambda, lambda + lambda_words);
  ctx->paillier_mu.assign(mu, mu + kN);
  int bits = 0;
  for (int i = lambda_words - 1; i >= 0 && !bits; i--)
    if (lambda[i]) bits = i * 32 + 32 - __builtin_clz(lambda[i]);
  ctx->paillier_lambda_bits = bits;
  // device copies
  if (ctx->secret_dev) cudaFree(ctx->secret_dev);
  cudaError_t e = cudaMalloc(&ctx->secret_dev, sizeof(uint32_t) * (lambda_words + kN));
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(secret)");
  e = cudaMemcpyAsync(ctx->secret_dev, lambda, sizeof(uint32_t) * lambda_words, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->secret_dev + lambda_words, mu, sizeof(uint32_t) * kN, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "upload secret");
  ctx->have_paillier_priv = true;
  return rebuild_exponent_programs(ctx);
}

extern "C" int scb_dgk_set_public(scb_ctx* ctx, const uint32_t* n, const uint32_t* g, const uint32_t* h, const uint32_t* g_neg, int kn,
                                  int u_bits, int l, int rand_bits) {
  if (!ctx || !n || !g || !h || kn <= 0 || kn > MAXW - 4 || l < 1 || l > MAX_L || u_bits < 2 || u_bits > 96 || rand_bits < 1)
    return fail(ctx, SCB_EINVAL, "scb_dgk_set_public: bad argument");
  cudaSetDevice(ctx->device);
  int rc = scb_modulus_load(ctx, SCB_MOD_DGK_N, n, kn);
  if (rc) return rc;
  rc = scb_fixed_base_load(ctx, 0, SCB_MOD_DGK_N, h, rand_bits);
  if (rc) return rc;
  KeysDev* hk = host_keys(ctx);
  hk->kn = kn;
  hk->l = l;
  hk->u_bits = u_bits;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_DGK_N];
  // g~, g_neg~ ; c_sa[s==-1][alpha] = enc(s) * enc(alpha): [0][0] = g, [0][1] = g*g, [1][0] = g_neg, [1][1] = g_neg*g
  std::vector<uint32_t> words(2 * kn), digs(2 * MAXK), prodA(2 * MAXK), prodB(2 * MAXK), prod(2 * MAXK);
  memcpy(words.data(), g, sizeof(uint32_t) * kn);
  if (g_neg != nullptr) {
    memcpy(words.data() + kn, g_neg, sizeof(uint32_t) * kn);
  } else {
    uint8_t ok = 0;                 // g^-1 mod n with the library's own inverse kernel
    if ((rc = scb_modinv(ctx, SCB_MOD_DGK_N, g, words.data() + kn, &ok, 1))) return rc;
  }
  const int K = md.K;
#define CALL(L, TPI) const_to_mont<L, TPI>(ctx, md, words.data(), kn, 2, 0, digs.data())
  auto run1 = [&]() -> int { SCB_DISPATCH_K(md.K, CALL) };
#undef CALL
  rc = run1();
  if (rc) return rc;
  memset(hk->g_m, 0, sizeof(hk->g_m));
  memset(hk->c_sa, 0, sizeof(hk->c_sa));
  memcpy(hk->g_m, digs.data(), sizeof(uint32_t) * K);
  memcpy(hk->c_sa[0][0], digs.data(), sizeof(uint32_t) * K);
  memcpy(hk->c_sa[1][0], digs.data() + K, sizeof(uint32_t) * K);
  memcpy(prodA.data(), digs.data(), sizeof(uint32_t) * K);            // g * g
  memcpy(prodB.data(), digs.data(), sizeof(uint32_t) * K);
  memcpy(prodA.data() + K, digs.data() + K, sizeof(uint32_t) * K);    // g_neg * g
  memcpy(prodB.data() + K, digs.data(), sizeof(uint32_t) * K);
#define CALL(L, TPI) const_mul<L, TPI>(ctx, md, prodA.data(), prodB.data(), 2, prod.data())
  auto run2 = [&]() -> int { SCB_DISPATCH_K(md.K, CALL) };
#undef CALL
  rc = run2();
  if (rc) return rc;
  memcpy(hk->c_sa[0][1], prod.data(), sizeof(uint32_t) * K);
  memcpy(hk->c_sa[1][1], prod.data() + K, sizeof(uint32_t) * K);
  ctx->dgk_rand_bits = rand_bits;
  ctx->have_dgk_pub = true;
  ctx->have_dgk_priv = false;
  if ((rc = upload_keys(ctx))) return rc;
  // the same constants in the FP64 engine's Montgomery form (x * R52 mod n), derived with the library's own batched product
  if (ctx->keys52_dev) {
    cudaFree(ctx->keys52_dev);
    ctx->keys52_dev = nullptr;
  }
  if (md.K52 > 0 && 2 * md.K52 <= 2 * MAXK52 && static_cast<int>(md.r52_words.size()) == kn) {
    const int K52 = md.K52;
    // canonical values: g, g_neg, g*g, g_neg*g
    std::vector<uint32_t> vals(4 * kn), a2(2 * kn), b2(2 * kn), r52x4(4 * kn), mont(4 * kn);
    memcpy(vals.data(), words.data(), sizeof(uint32_t) * 2 * kn);                  // g, g_neg
    memcpy(a2.data(), words.data(), sizeof(uint32_t) * 2 * kn);                    // (g, g_neg) * (g, g)
    memcpy(b2.data(), g, sizeof(uint32_t) * kn);
    memcpy(b2.data() + kn, g, sizeof(uint32_t) * kn);
    if ((rc = scb_modmul(ctx, SCB_MOD_DGK_N, a2.data(), b2.data(), vals.data() + 2 * kn, 2))) return rc;
    for (int i = 0; i < 4; i++) memcpy(r52x4.data() + static_cast<size_t>(i) * kn, md.r52_words.data(), sizeof(uint32_t) * kn);
    if ((rc = scb_modmul(ctx, SCB_MOD_DGK_N, vals.data(), r52x4.data(), mont.data(), 4))) return rc;
    std::vector<uint8_t> buf(sizeof(Keys52Dev), 0);
    Keys52Dev* h52 = reinterpret_cast<Keys52Dev*>(buf.data());
    auto to52 = [&](const uint32_t* x, double* digits) {
      for (int d = 0; d < K52; d++) {
        const int bit = d * 52, w = bit >> 5, sh = bit & 31;
        auto word = [&](int i) -> uint64_t { return i < kn ? x[i] : 0u; };
        uint64_t v = (word(w) | (word(w + 1) << 32)) >> sh;
        if (sh) v |= word(w + 2) << (64 - sh);
        digits[d] = static_cast<double>(v & ((1ull << 52) - 1ull));
      }
    };
    to52(mont.data(), h52->g_m);
    to52(mont.data(), h52->c_sa[0][0]);                                  // enc(+1) * enc(0) = g
    to52(mont.data() + static_cast<size_t>(2) * kn, h52->c_sa[0][1]);    // g * g
    to52(mont.data() + static_cast<size_t>(1) * kn, h52->c_sa[1][0]);    // g_neg
    to52(mont.data() + static_cast<size_t>(3) * kn, h52->c_sa[1][1]);    // g_neg * g
    cudaError_t ce = cudaMalloc(&ctx->keys52_dev, sizeof(Keys52Dev));
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "cudaMalloc(keys52)");
    ce = cudaMemcpyAsync(ctx->keys52_dev, h52, sizeof(Keys52Dev), cudaMemcpyHostToDevice, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "upload keys52");
  }
  return SCB_OK;
}

extern "C" int scb_dgk_set_private(scb_ctx* ctx, const uint32_t* p, int kp, const uint32_t* v_p, int vp_words, int vp_bits) {
  int rc = need_keys(ctx, false, false, true, false, "scb_dgk_set_private");
  if (rc) return rc;
  KeysDev* hk = host_keys(ctx);
  if (!p || !v_p || kp <= 0 || kp > hk->kn || vp_words <= 0 || vp_bits <= 0 || vp_bits > 32 * vp_words)
    return fail(ctx, SCB_EINVAL, "scb_dgk_set_private: bad argument");
  rc = scb_modulus_load(ctx, SCB_MOD_DGK_P, p, kp);
  if (rc) return rc;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_DGK_P];
  if (2 * W * md.K < 32 * hk->kn) return fail(ctx, SCB_EINVAL, "scb_dgk_set_private: p too short for the double-width reduction of c mod p");
  hk->kp = kp;
  ctx->dgk_vp.assign(v_p, v_p + vp_words);
  ctx->dgk_vp_bits = vp_bits;
  if (ctx->vp_dev) cudaFree(ctx->vp_dev);
  cudaError_t e = cudaMalloc(&ctx->vp_dev, sizeof(uint32_t) * vp_words);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(v_p)");
  e = cudaMemcpyAsync(ctx->vp_dev, v_p, sizeof(uint32_t) * vp_words, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "upload v_p");
  ctx->have_dgk_priv = true;
  if ((rc = upload_keys(ctx))) return rc;
  return rebuild_exponent_programs(ctx);
}

extern "C" int scb_set_option(scb_ctx* ctx, int option, int value) {
  if (!ctx) return SCB_EINVAL;
  switch (option) {
    case SCB_OPT_PAILLIER_CRT: ctx->use_crt = value != 0; return SCB_OK;
    case SCB_OPT_MODEXP_FP64: ctx->use_f64 = value != 0; return SCB_OK;
    case SCB_OPT_SLIDING_WINDOWS: ctx->use_sliding = value != 0; return SCB_OK;
    case SCB_OPT_SQUARE_MODULUS: ctx->use_pair = value != 0; return SCB_OK;
    case SCB_OPT_SMALL_BATCH: ctx->use_small_batch = value != 0; return SCB_OK;
    case SCB_OPT_OVERLAP_CRT: ctx->use_overlap = value != 0; return SCB_OK;
    case SCB_OPT_FIXED_BASE_WIDE: ctx->fb_wide = value != 0; return SCB_OK;   // takes effect at the next scb_fixed_base_load / scb_dgk_set_public
    default: return fail(ctx, SCB_EINVAL, "scb_set_option: unknown option %d", option);
  }
}

// ------------------------------------------------------------------------------------------------
// Paillier secret key in CRT form (KeyHolder)
// ------------------------------------------------------------------------------------------------
namespace {

typedef std::vector<uint32_t> HW;   // little-endian words, host side; only used once per key

HW hw_mul(const HW& a, const HW& b) {
  HW r(a.size() + b.size(), 0);
  for (size_t i = 0; i < a.size(); i++) {
    uint64_t carry = 0;
    for (size_t j = 0; j < b.size(); j++) {
      const uint64_t t = static_cast<uint64_t>(a[i]) * b[j] + r[i + j] + carry;
      r[i + j] = static_cast<uint32_t>(t);
      carry = t >> 32;
    }
    r[i + b.size()] = static_cast<uint32_t>(carry);
  }
  return r;
}
int hw_bits(const HW& x) {
  for (int i = static_cast<int>(x.size()) - 1; i >= 0; i--)
    if (x[i]) return i * 32 + 32 - __builtin_clz(x[i]);
  return 0;
}
bool hw_geq(const HW& a, const HW& b) {   // same length
  for (int i = static_cast<int>(a.size()) - 1; i >= 0; i--)
    if (a[i] != b[i]) return a[i] > b[i];
  return true;
}
void hw_sub(HW& a, const HW& b) {         // a -= b (same length, a >= b)
  uint64_t borrow = 0;
  for (size_t i = 0; i < a.size(); i++) {
    const uint64_t d = static_cast<uint64_t>(a[i]) - b[i] - borrow;
    a[i] = static_cast<uint32_t>(d);
    borrow = (d >> 63) & 1u;
  }
}
bool hw_is_zero(const HW& a) {
  for (uint32_t w : a)
    if (w) return false;
  return true;
}
// a mod m by shift-and-subtract (a one-off on 1024..4096-bit values)
HW hw_mod(const HW& a, const HW& m) {
  HW r(m.size() + 1, 0), mm_(m);
  mm_.push_back(0);
  for (int bit = hw_bits(a) - 1; bit >= 0; bit--) {
    uint32_t carry = (a[bit >> 5] >> (bit & 31)) & 1u;
    for (size_t i = 0; i < r.size(); i++) {
      const uint32_t nc = r[i] >> 31;
      r[i] = (r[i] << 1) | carry;
      carry = nc;
    }
    if (hw_geq(r, mm_)) hw_sub(r, mm_);
  }
  r.pop_back();
  return r;
}
HW hw_resized(const HW& a, size_t k) {
  HW r(a);
  r.resize(k, 0);
  return r;
}

}  // namespace

extern "C" int scb_paillier_set_private_crt(scb_ctx* ctx, const uint32_t* p, const uint32_t* q, int kP) {
  int rc = need_keys(ctx, true, false, false, false, "scb_paillier_set_private_crt");
  if (rc) return rc;
  KeysDev* hk = host_keys(ctx);
  const int kN = hk->kN;
  if (!p || !q || kP <= 0 || 2 * kP != kN || kP > MAXW / 4) return fail(ctx, SCB_EINVAL, "scb_paillier_set_private_crt: p and q must have kN/2 words each");
  const HW P[2] = {HW(p, p + kP), HW(q, q + kP)};
  const HW Nw(hk->Nw, hk->Nw + kN);
  if (hw_mul(P[0], P[1]) != Nw) return fail(ctx, SCB_EINVAL, "scb_paillier_set_private_crt: p * q != N");
  const HW P2[2] = {hw_mul(P[0], P[0]), hw_mul(P[1], P[1])};   // kN words each
  const int slotP[2] = {SCB_MOD_PAILLIER_P, SCB_MOD_PAILLIER_Q}, slotP2[2] = {SCB_MOD_PAILLIER_P2, SCB_MOD_PAILLIER_Q2};
  for (int h = 0; h < 2; h++) {
    if ((rc = scb_modulus_load(ctx, slotP[h], P[h].data(), kP))) return rc;
    if ((rc = scb_modulus_load(ctx, slotP2[h], P2[h].data(), kN))) return rc;
    if ((rc = scb::link_square_modulus(ctx, slotP2[h], slotP[h]))) return rc;
    if (2 * W * ctx->mod[slotP[h]].K < 32 * kN || 2 * W * ctx->mod[slotP2[h]].K < 64 * kN)
      return fail(ctx, SCB_EINVAL, "scb_paillier_set_private_crt: primes too short for the double-width reductions");
  }
  // exponents: e1[h] = Q mod (P - 1) (P - 1 when that is 0), P (for the lift b^P mod P^2), P - 1 (decryption)
  HW exps;                       // 6 arrays of kP words: e1p, p, p-1, e1q, q, q-1
  int e1_bits[2], P_bits[2];
  for (int h = 0; h < 2; h++) {
    HW pm1(P[h]);
    pm1[0] -= 1;                 // P is odd
    HW e1 = hw_mod(P[1 - h], pm1);
    if (hw_is_zero(e1)) e1 = pm1;
    e1_bits[h] = hw_bits(e1);
    P_bits[h] = hw_bits(P[h]);
    exps.insert(exps.end(), e1.begin(), e1.end());
    exps.insert(exps.end(), P[h].begin(), P[h].end());
    exps.insert(exps.end(), pm1.begin(), pm1.end());
  }
  if (ctx->crt_exp_dev) cudaFree(ctx->crt_exp_dev);
  cudaError_t e = cudaMalloc(&ctx->crt_exp_dev, sizeof(uint32_t) * exps.size());
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(crt exponents)");
  e = cudaMemcpyAsync(ctx->crt_exp_dev, exps.data(), sizeof(uint32_t) * exps.size(), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "upload crt exponents");
  ctx->crt_exp_host = exps;
  for (int h = 0; h < 2; h++) {
    ctx->crt_e1_bits[h] = e1_bits[h];
    ctx->crt_P_bits[h] = P_bits[h];
  }
  // recombination constants, derived with the library's own batched primitives (B = 1)
  hk->kP = kP;
  memset(hk->Pw, 0, sizeof(hk->Pw));
  memset(hk->crt_enc, 0, sizeof(hk->crt_enc));
  memset(hk->crt_dec, 0, sizeof(hk->crt_dec));
  HW enc_const, dec_const;       // 2 x 2kN words, 2 x kN words
  for (int h = 0; h < 2; h++) {
    memcpy(hk->Pw[h], P[h].data(), sizeof(uint32_t) * kP);
    uint32_t inv = P[h][0];
    for (int i = 0; i < 5; i++) inv *= 2u - P[h][0] * inv;
    hk->Pinv32[h] = inv;
    uint8_t ok = 0;
    // A_h = Q^2 * (Q^-2 mod P^2) < N^2
    HW inv2(kN, 0);
    if ((rc = scb_modinv(ctx, slotP2[h], P2[1 - h].data(), inv2.data(), &ok, 1))) return rc;
    HW A = hw_resized(hw_mul(P2[1 - h], inv2), 2 * kN);
    enc_const.insert(enc_const.end(), A.begin(), A.end());
    // C_h = h_P * Q * (Q^-1 mod P) mod N,  h_P = (-Q)^-1 mod P = P - (Q^-1 mod P)
    HW qinv(kP, 0);
    if ((rc = scb_modinv(ctx, slotP[h], P[1 - h].data(), qinv.data(), &ok, 1))) return rc;
    HW hP(P[h]);
    hw_sub(hP, qinv);
    HW Aprime = hw_resized(hw_mul(P[1 - h], qinv), kN);       // < N
    HW C(kN, 0);
    HW hP_N = hw_resized(hP, kN);
    if ((rc = scb_modmul(ctx, SCB_MOD_PAILLIER_N, Aprime.data(), hP_N.data(), C.data(), 1))) return rc;
    dec_const.insert(dec_const.end(), C.begin(), C.end());
  }
  {
    const scb_ctx::Mod& md = ctx->mod[SCB_MOD_PAILLIER_N2];
    std::vector<uint32_t> digs(2 * MAXK);
#define CALL(L, TPI) const_to_mont<L, TPI>(ctx, md, enc_const.data(), 2 * kN, 2, 1, digs.data())
    auto run = [&]() -> int { SCB_DISPATCH_K(md.K, CALL) };
#undef CALL
    if ((rc = run())) return rc;
    memcpy(hk->crt_enc[0], digs.data(), sizeof(uint32_t) * md.K);
    memcpy(hk->crt_enc[1], digs.data() + md.K, sizeof(uint32_t) * md.K);
  }
  {
    const scb_ctx::Mod& md = ctx->mod[SCB_MOD_PAILLIER_N];
    std::vector<uint32_t> digs(2 * MAXK);
#define CALL(L, TPI) const_to_mont<L, TPI>(ctx, md, dec_const.data(), kN, 2, 1, digs.data())
    auto run = [&]() -> int { SCB_DISPATCH_K(md.K, CALL) };
#undef CALL
    if ((rc = run())) return rc;
    memcpy(hk->crt_dec[0], digs.data(), sizeof(uint32_t) * md.K);
    memcpy(hk->crt_dec[1], digs.data() + md.K, sizeof(uint32_t) * md.K);
  }
  ctx->have_paillier_crt = true;
  ctx->have_paillier_priv = true;
  if ((rc = upload_keys(ctx))) return rc;
  return rebuild_exponent_programs(ctx);
}

// ------------------------------------------------------------------------------------------------
// phases, device pointers
// ------------------------------------------------------------------------------------------------
namespace {

// The launches issued while an AuxLane is alive go to the context's second stream (with their own per-group scratch), ordered
// after everything the main stream held when fork() was called; the destructor makes the main stream wait for them.  Used
// for the q half of the CRT forms below: the p and q chains are independent, and a launch of one chain fills the SMs that a
// draining launch of the other gives back (the last tickets of an exponentiation launch otherwise finish on a mostly idle
// GPU).  Off while per-kernel timing is on (one event stream) and with SCB_OPT_OVERLAP_CRT = 0.
struct AuxLane {
  scb_ctx* ctx;
  bool on = false;
  bool join_on_exit = true;
  static bool fork(scb_ctx* ctx) {
    if (!ctx->use_overlap || ctx->profiling) return false;
    if (!ctx->aux_stream) {
      cudaStream_t st = nullptr;
      cudaEvent_t a = nullptr, b = nullptr;
      if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess) {
        if (st) cudaStreamDestroy(st);
        if (a) cudaEventDestroy(a);
        cudaGetLastError();
        return false;
      }
      ctx->aux_stream = st;
      ctx->ev_fork = a;
      ctx->ev_join = b;
    }
    return cudaEventRecord(ctx->ev_fork, ctx->stream) == cudaSuccess && cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0) == cudaSuccess;
  }
  // join_on_exit = false: the caller issues more main-stream work that may overlap and calls join() itself
  AuxLane(scb_ctx* c, bool forked, bool join_at_exit = true) : ctx(c), on(forked), join_on_exit(join_at_exit) {
    if (!on) return;
    std::swap(ctx->stream, ctx->aux_stream);
    ctx->on_aux = true;
  }
  ~AuxLane() {
    if (!on) return;
    std::swap(ctx->stream, ctx->aux_stream);
    ctx->on_aux = false;
    if (join_on_exit) join(ctx);
  }
  static void join(scb_ctx* ctx) {
    cudaEventRecord(ctx->ev_join, ctx->aux_stream);
    cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
  }
};

template <int L, int TPI>
int phase_step1(scb_ctx* ctx, const uint32_t* x, const uint32_t* y, const uint32_t* r, const uint32_t* rho, int l, uint32_t* out, size_t B) {
  using C = Cfg<L, TPI>;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_PAILLIER_N2];
  const int kN = host_keys(ctx)->kN;
  Arena dg;
  int rc = get_arena(ctx, SC_DIGITS, pad256(sizeof(uint32_t) * B * C::K) + pad256(B), &dg);
  if (rc) return rc;
  uint32_t* xinv = dg.take<uint32_t>(B * C::K);
  uint8_t* ok = dg.take<uint8_t>(B);
  const int ebits = md.bits > 0 ? ctx->mod[SCB_MOD_PAILLIER_N].bits : 0;
  const int wbits = window_bits_for(ebits);
  // the randomizer rho^N mod N^2 on the FP64 engine when it is enabled: a separate launch, then k_step1 only multiplies.
  // It does not depend on x^-1, so it goes to the second stream FIRST and the (latency-bound, nearly empty) kernels of the
  // batched inversion move into the SMs its drain phase gives back; k_step1 waits for both.
  const uint32_t* rho_pow = nullptr;
  bool forked = false;
  if (rho != nullptr && ctx->use_f64 && md.K52 > 0) {
    Arena rp;
    rc = get_arena(ctx, SC_CRT, pad256(sizeof(uint32_t) * B * 2 * kN), &rp);
    if (rc) return rc;
    uint32_t* pw = rp.take<uint32_t>(B * 2 * kN);
    forked = AuxLane::fork(ctx);
    {
      AuxLane lane(ctx, forked, false);
      rc = scb::modexp_any(ctx, SCB_MOD_PAILLIER_N2, rho, kN, keys_N_words(ctx), kN, 0, ebits, nullptr, pw, 2 * kN, B);
    }
    if (rc) {
      if (forked) AuxLane::join(ctx);
      return rc;
    }
    rho_pow = pw;
    rho = nullptr;
  }
  bool any_fail = false;
  rc = batch_modinv_mont<L, TPI>(ctx, SCB_MOD_PAILLIER_N2, x, B, xinv, ok, &any_fail);
  if (forked) AuxLane::join(ctx);
  if (rc) return rc;
  int wb = wbits, nprog = 0, pw = 0;
  const uint16_t* prog = ctx->find_program(keys_N_words(ctx), &nprog, &pw, &wb);    // IMAD engine: sliding-window program of N
  const int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_step1<L, TPI>), C::SMEM, B, C::GPB);
  const size_t groups = static_cast<size_t>(grid) * C::GPB;
  void* scratch = nullptr;
  rc = ctx->scratch(SC_GROUP, sizeof(uint32_t) * groups * (modexp_scratch_words<L, TPI>(wb) + C::K), &scratch);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k_step1<L, TPI><<<grid, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, ctx->keys_dev, y, xinv, r, rho, rho_pow, l, ebits, wb,
                                                                 static_cast<uint32_t*>(scratch), out, B, queue, prog, nprog, pw);
  rc = ctx->launched("k_step1");
  if (rc) return rc;
  (void)kN;
  if (any_fail) return fail(ctx, SCB_ENOTINV, "scb_initiator_step1: an input ciphertext x has no inverse mod N^2");
  return SCB_OK;
}

template <int L, int TPI>
int phase_paillier_encrypt(scb_ctx* ctx, const uint32_t* plain, const uint32_t* rho, uint32_t* out, size_t M) {
  using C = Cfg<L, TPI>;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_PAILLIER_N2];
  const int ebits = ctx->mod[SCB_MOD_PAILLIER_N].bits;
  const int wbits = window_bits_for(ebits);
  const uint32_t* rho_pow = nullptr;
  if (rho != nullptr && ctx->use_f64 && md.K52 > 0) {
    const int kN = host_keys(ctx)->kN;
    Arena rp;
    int rc2 = get_arena(ctx, SC_CRT, pad256(sizeof(uint32_t) * M * 2 * kN), &rp);
    if (rc2) return rc2;
    uint32_t* pw = rp.take<uint32_t>(M * 2 * kN);
    rc2 = scb::modexp_any(ctx, SCB_MOD_PAILLIER_N2, rho, kN, keys_N_words(ctx), kN, 0, ebits, nullptr, pw, 2 * kN, M);
    if (rc2) return rc2;
    rho_pow = pw;
    rho = nullptr;
  }
  int wb = wbits, nprog = 0, pw = 0;
  const uint16_t* prog = ctx->find_program(keys_N_words(ctx), &nprog, &pw, &wb);
  const int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_paillier_encrypt<L, TPI>), C::SMEM, M, C::GPB);
  const size_t groups = static_cast<size_t>(grid) * C::GPB;
  void* scratch = nullptr;
  int rc = ctx->scratch(SC_GROUP, sizeof(uint32_t) * groups * modexp_scratch_words<L, TPI>(wb), &scratch);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k_paillier_encrypt<L, TPI><<<grid, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, ctx->keys_dev, plain, rho, rho_pow, ebits, wb,
                                                                            static_cast<uint32_t*>(scratch), out, M, queue, prog, nprog, pw);
  return ctx->launched("k_paillier_encrypt");
}

template <int L, int TPI>
int phase_dgk_encrypt_bits(scb_ctx* ctx, const uint64_t* beta, const uint8_t* dflag, int l, const uint32_t* rho, int ew, uint32_t* out, size_t B) {
  using C = Cfg<L, TPI>;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_DGK_N];
  const scb_ctx::FixedBase& fb = ctx->fb[0];
  const size_t M = B * (l + 1);
  const int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_dgk_encrypt_bits<L, TPI>), C::SMEM, M, C::GPB);
  void* scratch = nullptr;
  int rc = ctx->scratch(SC_GROUP, sizeof(uint32_t) * static_cast<size_t>(grid) * C::GPB * C::K, &scratch);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k_dgk_encrypt_bits<L, TPI><<<grid, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, ctx->keys_dev, beta, dflag, l, fb.use_table(), fb.use_nwin(), fb.use_wbits(), rho, ew,
                                                                            static_cast<uint32_t*>(scratch), out, B, queue);
  return ctx->launched("k_dgk_encrypt_bits");
}

template <int L, int TPI>
int phase_step4(scb_ctx* ctx, const uint32_t* db_enc, const uint32_t* r, const uint8_t* delta_a, const uint32_t* r_blind, int rb_words,
                const uint8_t* perm, const uint32_t* rho, int ew, int l, int do_blind, uint32_t* out, size_t B) {
  using C = Cfg<L, TPI>;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_DGK_N];
  const scb_ctx::FixedBase& fb = ctx->fb[0];
  const KeysDev* hk = host_keys(ctx);
  const int kn = hk->kn;
  const int ubits = hk->u_bits;
  const int wbits = window_bits_for(ubits);
  Arena ph;
  int rc = get_arena(ctx, SC_PHASE, 2 * pad256(sizeof(uint64_t) * B) + 2 * pad256(B) + pad256(sizeof(uint32_t) * B * kn), &ph);
  if (rc) return rc;
  uint64_t* alpha = ph.take<uint64_t>(B);
  uint64_t* alpha_t = ph.take<uint64_t>(B);
  uint8_t* r_small = ph.take<uint8_t>(B);
  uint8_t* r_bad = ph.take<uint8_t>(B);
  uint32_t* Tw = ph.take<uint32_t>(B * kn);
  k_prep_initiator<<<static_cast<unsigned>((B + 127) / 128), 128, 0, ctx->stream>>>(ctx->keys_dev, r, l, alpha, alpha_t, r_small, r_bad, B);
  if ((rc = ctx->launched("k_prep_initiator"))) return rc;
  const int gridA = ctx->grid_for(reinterpret_cast<const void*>(&k_step4_prefix<L, TPI>), C::SMEM, B, C::GPB);
  const int gridB = ctx->grid_for(reinterpret_cast<const void*>(&k_step4_main<L, TPI>), C::SMEM, B, C::GPB);
  const size_t groups = static_cast<size_t>(std::max(gridA, gridB)) * C::GPB;
  void* s4 = nullptr;
  rc = ctx->scratch(SC_STEP4, sizeof(uint32_t) * (B + groups) * step4_scratch_words<L, TPI>(l, wbits), &s4);
  if (rc) return rc;
  k_step4_prefix<L, TPI><<<gridA, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, ctx->keys_dev, db_enc, r_small, l, wbits,
                                                                         static_cast<uint32_t*>(s4), Tw, B);
  if ((rc = ctx->launched("k_step4_prefix"))) return rc;
  Arena dg;
  rc = get_arena(ctx, SC_DIGITS, pad256(sizeof(uint32_t) * B * C::K) + pad256(B), &dg);
  if (rc) return rc;
  uint32_t* TI = dg.take<uint32_t>(B * C::K);
  uint8_t* ok = dg.take<uint8_t>(B);
  bool any_fail = false;
  rc = batch_modinv_mont<L, TPI>(ctx, SCB_MOD_DGK_N, Tw, B, TI, ok, &any_fail);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k_step4_main<L, TPI><<<gridB, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, ctx->keys_dev, TI, alpha, alpha_t, delta_a, r_blind, rb_words, perm,
                                                                       fb.use_table(), fb.use_nwin(), fb.use_wbits(), rho, ew, l, ubits, wbits, do_blind,
                                                                       static_cast<uint32_t*>(s4), out, B, queue);
  if ((rc = ctx->launched("k_step4_main"))) return rc;
  if (any_fail) return fail(ctx, SCB_ENOTINV, "scb_initiator_step4: a DGK ciphertext has no inverse mod n");
  return SCB_OK;
}

// ---- the same two phases on the FP64 engine (kernels_dgk52.cuh); L29 / TPI29 = the IMAD geometry of the same modulus (batched inverse)
#define SCB_DISPATCH_K52(K52, CALL)                                               \
  switch (K52) {                                                                  \
    case 20: return CALL(10, 2);                                                  \
    case 30: return CALL(15, 2);                                                  \
    case 40: return CALL(10, 4);                                                  \
    case 60: return CALL(15, 4);                                                  \
    case 80: return CALL(10, 8);                                                  \
    case 120: return CALL(15, 8);                                                 \
    case 160: return CALL(10, 16);                                                \
    default: return fail(ctx, SCB_EINVAL, "unsupported FP64 digit count %d", K52); \
  }

template <int L, int TPI>
struct Cfg52 {
  using G = Geo52<L, TPI>;
  static constexpr size_t SMEM = sizeof(double) * G::WARP_SMEM_DOUBLES * WARPS_PER_BLOCK;
  static constexpr int GPB = GroupIndex52<L, TPI>::GROUPS_PER_BLOCK;
  static constexpr int K = G::K;
};

bool dgk52_ready(scb_ctx* ctx) {
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_DGK_N];
  return ctx->use_f64 && ctx->use_dgk52 && md.K52 > 0 && md.dev52 && ctx->keys52_dev && ctx->fb[0].loaded && ctx->fb[0].table52;
}

template <int L, int TPI>
int phase_dgk_encrypt_bits52(scb_ctx* ctx, const uint64_t* beta, const uint8_t* dflag, int l, const uint32_t* rho, int ew, uint32_t* out, size_t B) {
  using C = Cfg52<L, TPI>;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_DGK_N];
  const scb_ctx::FixedBase& fb = ctx->fb[0];
  const size_t M = B * (l + 1);
  const int grid = ctx->grid_for(reinterpret_cast<const void*>(&k52_dgk_encrypt_bits<L, TPI>), C::SMEM, M, C::GPB);
  void* scratch = nullptr;
  int rc = ctx->scratch(SC_GROUP, sizeof(double) * static_cast<size_t>(grid) * C::GPB * 2 * C::K, &scratch);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  const bool wide = fb.table16_52 != nullptr;
  k52_dgk_encrypt_bits<L, TPI><<<grid, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev52, ctx->keys_dev, ctx->keys52_dev, beta, dflag, l, fb.use_table52(),
                                                                              wide ? fb.nwin16 : fb.nwin, wide ? 16 : 8, rho, ew,
                                                                              static_cast<double*>(scratch), out, B, queue);
  return ctx->launched("k52_dgk_encrypt_bits");
}

template <int L29, int TPI29>
int mont29_to_words(scb_ctx* ctx, const scb_ctx::Mod& md, const uint32_t* digits, int kw, uint32_t* out, size_t B) {
  using C = Cfg<L29, TPI29>;
  const int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_mont29_to_words<L29, TPI29>), C::SMEM, B, C::GPB);
  void* scratch = nullptr;
  int rc = ctx->scratch(SC_GROUP, sizeof(uint32_t) * static_cast<size_t>(grid) * C::GPB * C::K, &scratch);
  if (rc) return rc;
  k_mont29_to_words<L29, TPI29><<<grid, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, digits, kw, static_cast<uint32_t*>(scratch), out, B);
  return ctx->launched("k_mont29_to_words");
}

template <int L, int TPI, int L29, int TPI29>
int phase_step4_52(scb_ctx* ctx, const uint32_t* db_enc, const uint32_t* r, const uint8_t* delta_a, const uint32_t* r_blind, int rb_words,
                   const uint8_t* perm, const uint32_t* rho, int ew, int l, int do_blind, uint32_t* out, size_t B) {
  using C = Cfg52<L, TPI>;
  using C29 = Cfg<L29, TPI29>;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_DGK_N];
  const scb_ctx::FixedBase& fb = ctx->fb[0];
  const KeysDev* hk = host_keys(ctx);
  const int kn = hk->kn;
  const int ubits = hk->u_bits;
  const int wbits = window_bits_for(ubits);
  Arena ph;
  int rc = get_arena(ctx, SC_PHASE, 2 * pad256(sizeof(uint64_t) * B) + 2 * pad256(B) + 2 * pad256(sizeof(uint32_t) * B * kn), &ph);
  if (rc) return rc;
  uint64_t* alpha = ph.take<uint64_t>(B);
  uint64_t* alpha_t = ph.take<uint64_t>(B);
  uint8_t* r_small = ph.take<uint8_t>(B);
  uint8_t* r_bad = ph.take<uint8_t>(B);
  uint32_t* Tw = ph.take<uint32_t>(B * kn);
  uint32_t* TIw = ph.take<uint32_t>(B * kn);
  k_prep_initiator<<<static_cast<unsigned>((B + 127) / 128), 128, 0, ctx->stream>>>(ctx->keys_dev, r, l, alpha, alpha_t, r_small, r_bad, B);
  if ((rc = ctx->launched("k_prep_initiator"))) return rc;
  const int gridA = ctx->grid_for(reinterpret_cast<const void*>(&k52_step4_prefix<L, TPI>), C::SMEM, B, C::GPB);
  const int gridB = ctx->grid_for(reinterpret_cast<const void*>(&k52_step4_main<L, TPI>), C::SMEM, B, C::GPB);
  const size_t groups = static_cast<size_t>(std::max(gridA, gridB)) * C::GPB;
  void* s4 = nullptr;
  rc = ctx->scratch(SC_STEP4, sizeof(double) * (B + groups) * step4_scratch_doubles<L, TPI>(l, wbits), &s4);
  if (rc) return rc;
  k52_step4_prefix<L, TPI><<<gridA, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev52, ctx->keys_dev, db_enc, r_small, l, wbits, static_cast<double*>(s4), Tw, B);
  if ((rc = ctx->launched("k52_step4_prefix"))) return rc;
  // one inverse per comparison through the batched-inversion trick of the IMAD engine, handed over as canonical words
  Arena dg;
  rc = get_arena(ctx, SC_DIGITS, pad256(sizeof(uint32_t) * B * C29::K) + pad256(B), &dg);
  if (rc) return rc;
  uint32_t* TI = dg.take<uint32_t>(B * C29::K);
  uint8_t* ok = dg.take<uint8_t>(B);
  bool any_fail = false;
  rc = batch_modinv_mont<L29, TPI29>(ctx, SCB_MOD_DGK_N, Tw, B, TI, ok, &any_fail);
  if (rc) return rc;
  if ((rc = mont29_to_words<L29, TPI29>(ctx, md, TI, kn, TIw, B))) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  const bool wide = fb.table16_52 != nullptr;
  k52_step4_main<L, TPI><<<gridB, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev52, ctx->keys_dev, ctx->keys52_dev, TIw, alpha, alpha_t, delta_a, r_blind, rb_words, perm,
                                                                         fb.use_table52(), wide ? fb.nwin16 : fb.nwin, wide ? 16 : 8, rho, ew, l, ubits, wbits, do_blind,
                                                                         static_cast<double*>(s4), out, B, queue);
  if ((rc = ctx->launched("k52_step4_main"))) return rc;
  if (any_fail) return fail(ctx, SCB_ENOTINV, "scb_initiator_step4: a DGK ciphertext has no inverse mod n");
  return SCB_OK;
}

// K52 -> FP64 geometry, K29 -> IMAD geometry of the same modulus (they always pair up: 20/36, 30/54, 40/72, 60/108, 80/144, 120/216, 160/288)
int dispatch_step4_52(scb_ctx* ctx, const uint32_t* db_enc, const uint32_t* r, const uint8_t* delta_a, const uint32_t* r_blind, int rb_words,
                      const uint8_t* perm, const uint32_t* rho, int ew, int l, int do_blind, uint32_t* out, size_t B) {
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_DGK_N];
#define S4(L, TPI, L29, TPI29) phase_step4_52<L, TPI, L29, TPI29>(ctx, db_enc, r, delta_a, r_blind, rb_words, perm, rho, ew, l, do_blind, out, B)
  if (md.K52 == 20 && md.K == 36) return S4(10, 2, 18, 2);
  if (md.K52 == 30 && md.K == 54) return S4(15, 2, 27, 2);
  if (md.K52 == 40 && md.K == 72) return S4(10, 4, 18, 4);
  if (md.K52 == 60 && md.K == 108) return S4(15, 4, 27, 4);
  if (md.K52 == 80 && md.K == 144) return S4(10, 8, 18, 8);
#undef S4
  return 1;      // no FP64 pairing for this size: the caller falls back to the IMAD kernels
}

template <int L, int TPI>
int phase_is_zero(scb_ctx* ctx, const uint32_t* c, uint8_t* flags, size_t M) {
  using C = Cfg<L, TPI>;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_DGK_P];
  int wbits = window_bits_for(ctx->dgk_vp_bits);
  int nprog = 0, pw = 0;
  const uint16_t* prog = ctx->find_program(ctx->vp_dev, &nprog, &pw, &wbits);
  const int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_dgk_is_zero<L, TPI>), C::SMEM, M, C::GPB);
  void* scratch = nullptr;
  int rc = ctx->scratch(SC_GROUP, sizeof(uint32_t) * static_cast<size_t>(grid) * C::GPB * modexp_scratch_words<L, TPI>(wbits), &scratch);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k_dgk_is_zero<L, TPI><<<grid, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, ctx->keys_dev, c, ctx->vp_dev, static_cast<int>(ctx->dgk_vp.size()),
                                                                       ctx->dgk_vp_bits, wbits, static_cast<uint32_t*>(scratch), flags, M, queue, prog, nprog, pw);
  return ctx->launched("k_dgk_is_zero");
}

template <int L, int TPI>
int phase_step67(scb_ctx* ctx, const uint32_t* enc3, const uint32_t* r, const uint8_t* delta_a, int l, uint32_t* out, size_t B) {
  using C = Cfg<L, TPI>;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_PAILLIER_N2];
  const int kN = host_keys(ctx)->kN, kc = 2 * kN;
  Arena ph, dg;
  int rc = get_arena(ctx, SC_PHASE, pad256(sizeof(uint32_t) * B * kN) + pad256(B) + pad256(sizeof(uint32_t) * B * kc), &ph);
  if (rc) return rc;
  uint32_t* e = ph.take<uint32_t>(B * kN);
  uint8_t* r_small = ph.take<uint8_t>(B);
  uint32_t* dbe = ph.take<uint32_t>(B * kc);       // delta_b_enc gathered contiguously for the batched inversion
  rc = get_arena(ctx, SC_DIGITS, pad256(sizeof(uint32_t) * B * C::K) + pad256(B), &dg);
  if (rc) return rc;
  uint32_t* dbinv = dg.take<uint32_t>(B * C::K);
  uint8_t* ok = dg.take<uint8_t>(B);
  cudaError_t ce = cudaMemcpy2DAsync(dbe, sizeof(uint32_t) * kc, enc3 + 2 * kc, sizeof(uint32_t) * 3 * kc, sizeof(uint32_t) * kc, B,
                                     cudaMemcpyDeviceToDevice, ctx->stream);
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "gather delta_b_enc");
  bool any_fail = false;
  rc = batch_modinv_mont<L, TPI>(ctx, SCB_MOD_PAILLIER_N2, dbe, B, dbinv, ok, &any_fail);
  if (rc) return rc;
  k_prep_step67<<<static_cast<unsigned>((B + 63) / 64), 64, 0, ctx->stream>>>(ctx->keys_dev, r, delta_a, l, e, r_small, B);
  if ((rc = ctx->launched("k_prep_step67"))) return rc;
  const int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_step67<L, TPI>), C::SMEM, B, C::GPB);
  void* scratch = nullptr;
  rc = ctx->scratch(SC_GROUP, sizeof(uint32_t) * static_cast<size_t>(grid) * C::GPB * 2 * C::K, &scratch);
  if (rc) return rc;
  // zeta1 / zeta2 / delta_b_enc live interleaved in enc3: [B][3][kc]; the kernel indexes i*kc, so hand it strided views
  k_step67<L, TPI><<<grid, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, ctx->keys_dev, enc3, enc3 + kc, static_cast<size_t>(3) * kc, dbe, dbinv, e, r_small, delta_a,
                                                                  static_cast<uint32_t*>(scratch), out, B);
  if ((rc = ctx->launched("k_step67"))) return rc;
  if (any_fail) return fail(ctx, SCB_ENOTINV, "scb_initiator_step6_7: [[delta_B]] has no inverse mod N^2");
  return SCB_OK;
}

}  // namespace

extern "C" int scb_initiator_step1_dev(scb_ctx* ctx, const uint32_t* x, const uint32_t* y, const uint32_t* r, const uint32_t* rho, int l,
                                       uint32_t* z_enc, size_t B) {
  int rc = need_keys(ctx, true, false, false, false, "scb_initiator_step1");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!x || !y || !r || !z_enc || l < 1 || l > MAX_L) return fail(ctx, SCB_EINVAL, "scb_initiator_step1: bad argument");
  // the reference asserts 2^(l+2) < N // 2 (initiator.py:249)
  if (l + 3 >= ctx->mod[SCB_MOD_PAILLIER_N].bits) return fail(ctx, SCB_EINVAL, "scb_initiator_step1: l too large for N (2^(l+2) < N//2 violated)");
  const int K = ctx->mod[SCB_MOD_PAILLIER_N2].K;
#define CALL(L, TPI) phase_step1<L, TPI>(ctx, x, y, r, rho, l, z_enc, B)
  SCB_DISPATCH_K(K, CALL)
#undef CALL
}

namespace {

template <int L, int TPI>
int phase_crt_encrypt_combine(scb_ctx* ctx, const uint32_t* sp, const uint32_t* sq, const uint32_t* plain, uint32_t* out, size_t M) {
  using C = Cfg<L, TPI>;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_PAILLIER_N2];
  const int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_crt_encrypt_combine<L, TPI>), C::SMEM, M, C::GPB);
  void* scratch = nullptr;
  int rc = ctx->scratch(SC_GROUP, sizeof(uint32_t) * static_cast<size_t>(grid) * C::GPB * 2 * C::K, &scratch);
  if (rc) return rc;
  k_crt_encrypt_combine<L, TPI><<<grid, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, ctx->keys_dev, sp, sq, plain, static_cast<uint32_t*>(scratch), out, M);
  return ctx->launched("k_crt_encrypt_combine");
}

template <int L, int TPI>
int phase_crt_decrypt_combine(scb_ctx* ctx, const uint32_t* Lq, uint32_t* out, size_t M) {
  using C = Cfg<L, TPI>;
  const scb_ctx::Mod& md = ctx->mod[SCB_MOD_PAILLIER_N];
  const int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_crt_decrypt_combine<L, TPI>), C::SMEM, M, C::GPB);
  void* scratch = nullptr;
  int rc = ctx->scratch(SC_GROUP, sizeof(uint32_t) * static_cast<size_t>(grid) * C::GPB * 2 * C::K, &scratch);
  if (rc) return rc;
  k_crt_decrypt_combine<L, TPI><<<grid, BLOCK_THREADS, C::SMEM, ctx->stream>>>(md.dev, ctx->keys_dev, Lq, static_cast<uint32_t*>(scratch), out, M);
  return ctx->launched("k_crt_decrypt_combine");
}

// KeyHolder with the factorisation: (1 + m N) * rho^N mod N^2 through p, p^2, q, q^2 (see kernels_protocol.cuh)
int paillier_encrypt_crt(scb_ctx* ctx, const uint32_t* plain, const uint32_t* rho, uint32_t* out, size_t M) {
  const KeysDev* hk = host_keys(ctx);
  const int kN = hk->kN, kP = hk->kP;
  Arena ph;
  int rc = get_arena(ctx, SC_CRT, 2 * pad256(sizeof(uint32_t) * M * kP) + 2 * pad256(sizeof(uint32_t) * M * kN), &ph);
  if (rc) return rc;
  uint32_t* b[2] = {ph.take<uint32_t>(M * kP), ph.take<uint32_t>(M * kP)};
  uint32_t* s[2] = {ph.take<uint32_t>(M * kN), ph.take<uint32_t>(M * kN)};
  const int slotP[2] = {SCB_MOD_PAILLIER_P, SCB_MOD_PAILLIER_Q}, slotP2[2] = {SCB_MOD_PAILLIER_P2, SCB_MOD_PAILLIER_Q2};
  const bool forked = AuxLane::fork(ctx);
  for (int h = 0; h < 2 && rc == SCB_OK; h++) {
    AuxLane lane(ctx, forked && h == 1);
    const uint32_t* ex = ctx->crt_exp_dev + static_cast<size_t>(3 * h) * kP;
    rc = scb::modexp_dev_ex(ctx, slotP[h], rho, kN, ex, kP, ctx->crt_e1_bits[h], b[h], kP, M);                         // (rho mod P)^(Q mod (P-1)) mod P
    if (rc == SCB_OK) rc = scb::modexp_dev_ex(ctx, slotP2[h], b[h], kP, ex + kP, kP, ctx->crt_P_bits[h], s[h], kN, M);  // b^P mod P^2
  }
  if (rc) return rc;
  const int K = ctx->mod[SCB_MOD_PAILLIER_N2].K;
#define CALL(L, TPI) phase_crt_encrypt_combine<L, TPI>(ctx, s[0], s[1], plain, out, M)
  SCB_DISPATCH_K(K, CALL)
#undef CALL
}

int paillier_decrypt_crt(scb_ctx* ctx, const uint32_t* c, uint32_t* out, size_t M) {
  const KeysDev* hk = host_keys(ctx);
  const int kN = hk->kN, kP = hk->kP;
  Arena ph;
  int rc = get_arena(ctx, SC_CRT, pad256(sizeof(uint32_t) * 2 * M * kN) + pad256(sizeof(uint32_t) * 2 * M * kP), &ph);
  if (rc) return rc;
  uint32_t* y = ph.take<uint32_t>(2 * M * kN);
  uint32_t* Lq = ph.take<uint32_t>(2 * M * kP);
  const int slotP2[2] = {SCB_MOD_PAILLIER_P2, SCB_MOD_PAILLIER_Q2};
  const bool forked = AuxLane::fork(ctx);
  for (int h = 0; h < 2 && rc == SCB_OK; h++) {
    AuxLane lane(ctx, forked && h == 1);
    const uint32_t* ex = ctx->crt_exp_dev + static_cast<size_t>(3 * h + 2) * kP;
    rc = scb::modexp_dev_ex(ctx, slotP2[h], c, 2 * kN, ex, kP, ctx->crt_P_bits[h], y + static_cast<size_t>(h) * M * kN, kN, M);   // c^(P-1) mod P^2
  }
  if (rc) return rc;
  k_crt_L<<<static_cast<unsigned>((2 * M + 63) / 64), 64, 0, ctx->stream>>>(ctx->keys_dev, y, Lq, M);
  if ((rc = ctx->launched("k_crt_L"))) return rc;
  const int K = ctx->mod[SCB_MOD_PAILLIER_N].K;
#define CALL(L, TPI) phase_crt_decrypt_combine<L, TPI>(ctx, Lq, out, M)
  SCB_DISPATCH_K(K, CALL)
#undef CALL
}

}  // namespace

extern "C" int scb_paillier_encrypt_dev(scb_ctx* ctx, const uint32_t* plain, const uint32_t* rho, uint32_t* out, size_t M) {
  int rc = need_keys(ctx, true, false, false, false, "scb_paillier_encrypt");
  if (rc) return rc;
  if (M == 0) return SCB_OK;
  if (!plain || !out) return fail(ctx, SCB_EINVAL, "scb_paillier_encrypt: null pointer");
  if (rho != nullptr && ctx->have_paillier_crt && ctx->use_crt) return paillier_encrypt_crt(ctx, plain, rho, out, M);
  const int K = ctx->mod[SCB_MOD_PAILLIER_N2].K;
#define CALL(L, TPI) phase_paillier_encrypt<L, TPI>(ctx, plain, rho, out, M)
  SCB_DISPATCH_K(K, CALL)
#undef CALL
}

extern "C" int scb_paillier_decrypt_dev(scb_ctx* ctx, const uint32_t* c, uint32_t* out, size_t M) {
  int rc = need_keys(ctx, true, true, false, false, "scb_paillier_decrypt");
  if (rc) return rc;
  if (M == 0) return SCB_OK;
  if (!c || !out) return fail(ctx, SCB_EINVAL, "scb_paillier_decrypt: null pointer");
  if (ctx->have_paillier_crt && ctx->use_crt) return paillier_decrypt_crt(ctx, c, out, M);
  if (ctx->paillier_lambda.empty()) return fail(ctx, SCB_ESTATE, "scb_paillier_decrypt: lambda / mu not loaded (and the CRT path is disabled)");
  const int kN = host_keys(ctx)->kN, kc = 2 * kN;
  Arena ph;
  rc = get_arena(ctx, SC_CRT, pad256(sizeof(uint32_t) * M * kc) + pad256(sizeof(uint32_t) * M * kN), &ph);
  if (rc) return rc;
  uint32_t* xw = ph.take<uint32_t>(M * kc);
  uint32_t* qw = ph.take<uint32_t>(M * kN);
  const int lw = static_cast<int>(ctx->paillier_lambda.size());
  rc = scb_modexp_dev(ctx, SCB_MOD_PAILLIER_N2, c, ctx->secret_dev, lw, 1, ctx->paillier_lambda_bits, nullptr, xw, M);
  if (rc) return rc;
  k_paillier_L<<<static_cast<unsigned>((M + 63) / 64), 64, 0, ctx->stream>>>(ctx->keys_dev, xw, qw, M);
  if ((rc = ctx->launched("k_paillier_L"))) return rc;
  return scb::modmul_dev_strided(ctx, SCB_MOD_PAILLIER_N, qw, kN, ctx->secret_dev + lw, 0, out, M);
}

extern "C" int scb_keyholder_step4ab_dev(scb_ctx* ctx, const uint32_t* z, int l, const uint32_t* rho, int rho_words, uint32_t* db_enc, size_t B) {
  int rc = need_keys(ctx, true, false, true, false, "scb_keyholder_step4ab");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!z || !db_enc || l < 1 || l > MAX_L || (rho && rho_words <= 0)) return fail(ctx, SCB_EINVAL, "scb_keyholder_step4ab: bad argument");
  if (host_keys(ctx)->u_bits < l + 3) return fail(ctx, SCB_EINVAL, "scb_keyholder_step4ab: u <= 2^(l+2) (keyholder.py:212)");
  Arena ph;
  rc = get_arena(ctx, SC_PHASE, pad256(sizeof(uint64_t) * B) + pad256(B), &ph);
  if (rc) return rc;
  uint64_t* beta = ph.take<uint64_t>(B);
  uint8_t* dflag = ph.take<uint8_t>(B);
  k_prep_keyholder_bits<<<static_cast<unsigned>((B + 127) / 128), 128, 0, ctx->stream>>>(ctx->keys_dev, z, l, beta, dflag, B);
  if ((rc = ctx->launched("k_prep_keyholder_bits"))) return rc;
  const int K = ctx->mod[SCB_MOD_DGK_N].K;
  if (dgk52_ready(ctx)) {
#define CALL(L, TPI) phase_dgk_encrypt_bits52<L, TPI>(ctx, beta, dflag, l, rho, rho_words, db_enc, B)
    SCB_DISPATCH_K52(ctx->mod[SCB_MOD_DGK_N].K52, CALL)
#undef CALL
  }
#define CALL(L, TPI) phase_dgk_encrypt_bits<L, TPI>(ctx, beta, dflag, l, rho, rho_words, db_enc, B)
  SCB_DISPATCH_K(K, CALL)
#undef CALL
}

extern "C" int scb_initiator_step4_dev(scb_ctx* ctx, const uint32_t* db_enc, const uint32_t* r, const uint8_t* delta_a, const uint32_t* r_blind,
                                       int rb_words, const uint8_t* perm, const uint32_t* rho, int rho_words, int l, uint32_t* c_out, size_t B) {
  int rc = need_keys(ctx, true, false, true, false, "scb_initiator_step4");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!db_enc || !r || !delta_a || !c_out || l < 1 || l > MAX_L || (r_blind && rb_words <= 0) || (rho && rho_words <= 0))
    return fail(ctx, SCB_EINVAL, "scb_initiator_step4: bad argument");
  const int K = ctx->mod[SCB_MOD_DGK_N].K;
  if (dgk52_ready(ctx)) {
    rc = dispatch_step4_52(ctx, db_enc, r, delta_a, r_blind, rb_words, perm, rho, rho_words, l, r_blind != nullptr, c_out, B);
    if (rc != 1) return rc;
  }
#define CALL(L, TPI) phase_step4<L, TPI>(ctx, db_enc, r, delta_a, r_blind, rb_words, perm, rho, rho_words, l, r_blind != nullptr, c_out, B)
  SCB_DISPATCH_K(K, CALL)
#undef CALL
}

extern "C" int scb_dgk_is_zero_dev(scb_ctx* ctx, const uint32_t* c, uint8_t* flags, size_t M) {
  int rc = need_keys(ctx, false, false, true, true, "scb_dgk_is_zero");
  if (rc) return rc;
  if (M == 0) return SCB_OK;
  if (!c || !flags) return fail(ctx, SCB_EINVAL, "scb_dgk_is_zero: null pointer");
  const scb_ctx::Mod& mdp = ctx->mod[SCB_MOD_DGK_P];
  if (ctx->use_f64 && mdp.K52 > 0 && ctx->dgk_vp_bits >= 64) {
    // c^{v_p} mod p on the FP64 engine (c is reduced mod p on load), then one compare per value
    const int kn = host_keys(ctx)->kn, kp = mdp.k32;
    Arena pw;
    rc = get_arena(ctx, SC_CRT, pad256(sizeof(uint32_t) * M * kp), &pw);
    if (rc) return rc;
    uint32_t* y = pw.take<uint32_t>(M * kp);
    rc = scb::modexp_any(ctx, SCB_MOD_DGK_P, c, kn, ctx->vp_dev, static_cast<int>(ctx->dgk_vp.size()), 0, ctx->dgk_vp_bits, nullptr, y, kp, M);
    if (rc) return rc;
    k_flag_is_one<<<static_cast<unsigned>((M + 127) / 128), 128, 0, ctx->stream>>>(y, kp, flags, M);
    return ctx->launched("k_flag_is_one");
  }
  const int K = mdp.K;
#define CALL(L, TPI) phase_is_zero<L, TPI>(ctx, c, flags, M)
  SCB_DISPATCH_K(K, CALL)
#undef CALL
}

extern "C" int scb_keyholder_step4j_5_dev(scb_ctx* ctx, const uint32_t* c, const uint32_t* z, const uint32_t* rho, int l, uint8_t* delta_b,
                                          uint32_t* enc3, size_t B) {
  int rc = need_keys(ctx, true, false, true, true, "scb_keyholder_step4j_5");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!c || !z || !delta_b || !enc3 || l < 1 || l > MAX_L) return fail(ctx, SCB_EINVAL, "scb_keyholder_step4j_5: bad argument");
  const int kN = host_keys(ctx)->kN;
  Arena ph;
  rc = get_arena(ctx, SC_PHASE, pad256(B * (l + 1)) + pad256(sizeof(uint32_t) * B * 3 * kN), &ph);
  if (rc) return rc;
  uint8_t* flags = ph.take<uint8_t>(B * (l + 1));
  uint32_t* plain = ph.take<uint32_t>(B * 3 * kN);
  rc = scb_dgk_is_zero_dev(ctx, c, flags, B * (l + 1));
  if (rc) return rc;
  k_prep_keyholder_step5<<<static_cast<unsigned>((B + 63) / 64), 64, 0, ctx->stream>>>(ctx->keys_dev, z, flags, l + 1, nullptr, l, delta_b, plain, B);
  if ((rc = ctx->launched("k_prep_keyholder_step5"))) return rc;
  return scb_paillier_encrypt_dev(ctx, plain, rho, enc3, 3 * B);
}

extern "C" int scb_keyholder_step5_dev(scb_ctx* ctx, const uint32_t* z, const uint8_t* delta_b, const uint32_t* rho, int l, uint32_t* enc3, size_t B) {
  int rc = need_keys(ctx, true, false, false, false, "scb_keyholder_step5");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!z || !delta_b || !enc3 || l < 1 || l > MAX_L) return fail(ctx, SCB_EINVAL, "scb_keyholder_step5: bad argument");
  const int kN = host_keys(ctx)->kN;
  Arena ph;
  rc = get_arena(ctx, SC_PHASE, pad256(sizeof(uint32_t) * B * 3 * kN), &ph);
  if (rc) return rc;
  uint32_t* plain = ph.take<uint32_t>(B * 3 * kN);
  k_prep_keyholder_step5<<<static_cast<unsigned>((B + 63) / 64), 64, 0, ctx->stream>>>(ctx->keys_dev, z, nullptr, 0, delta_b, l, nullptr, plain, B);
  if ((rc = ctx->launched("k_prep_keyholder_step5"))) return rc;
  return scb_paillier_encrypt_dev(ctx, plain, rho, enc3, 3 * B);
}

extern "C" int scb_initiator_step6_7_dev(scb_ctx* ctx, const uint32_t* enc3, const uint32_t* r, const uint8_t* delta_a, int l, uint32_t* result,
                                         size_t B) {
  int rc = need_keys(ctx, true, false, false, false, "scb_initiator_step6_7");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!enc3 || !r || !delta_a || !result || l < 1 || l > MAX_L) return fail(ctx, SCB_EINVAL, "scb_initiator_step6_7: bad argument");
  const int K = ctx->mod[SCB_MOD_PAILLIER_N2].K;
#define CALL(L, TPI) phase_step67<L, TPI>(ctx, enc3, r, delta_a, l, result, B)
  SCB_DISPATCH_K(K, CALL)
#undef CALL
}

// ------------------------------------------------------------------------------------------------
// host-pointer wrappers
// ------------------------------------------------------------------------------------------------
namespace {

struct HostStage {
  scb_ctx* ctx;
  int rc = SCB_OK;
  Arena a;
  HostStage(scb_ctx* c, size_t bytes) : ctx(c) { rc = get_arena(ctx, SC_STAGE, bytes, &a); }
  template <typename T>
  const T* up(const T* host, size_t count) {
    if (!host) return nullptr;
    T* p = a.take<T>(count);
    if (rc == SCB_OK) {
      cudaError_t e = cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) rc = cuda_fail(ctx, e, "H2D copy");
    }
    return p;
  }
  template <typename T>
  T* out(size_t count) { return a.take<T>(count); }
  template <typename T>
  void down(T* host, const T* dev, size_t count) {
    if (rc == SCB_OK && host) {
      cudaError_t e = cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
      if (e != cudaSuccess) rc = cuda_fail(ctx, e, "D2H copy");
    }
  }
  int finish(int call_rc) {
    // results are still copied back when the phase reported a non-invertible element (outputs of the
    // other elements are valid), but not after a hard error
    if (call_rc != SCB_OK && call_rc != SCB_ENOTINV) return call_rc;
    if (rc) return rc;
    int s = scb_sync(ctx);
    return s ? s : call_rc;
  }
};

}  // namespace

extern "C" int scb_initiator_step1(scb_ctx* ctx, const uint32_t* x, const uint32_t* y, const uint32_t* r, const uint32_t* rho, int l,
                                   uint32_t* z_enc, size_t B) {
  int rc = need_keys(ctx, true, false, false, false, "scb_initiator_step1");
  if (rc || B == 0) return rc;
  const size_t kN = host_keys(ctx)->kN, kc = 2 * kN;
  HostStage st(ctx, 3 * pad256(B * kc * 4) + 2 * pad256(B * kN * 4));
  if (st.rc) return st.rc;
  const uint32_t* dx = st.up(x, B * kc);
  const uint32_t* dy = st.up(y, B * kc);
  const uint32_t* dr = st.up(r, B * kN);
  const uint32_t* drho = st.up(rho, B * kN);
  uint32_t* dz = st.out<uint32_t>(B * kc);
  if (st.rc) return st.rc;
  rc = scb_initiator_step1_dev(ctx, dx, dy, dr, drho, l, dz, B);
  if (rc == SCB_OK || rc == SCB_ENOTINV) st.down(z_enc, dz, B * kc);
  return st.finish(rc);
}

extern "C" int scb_paillier_encrypt(scb_ctx* ctx, const uint32_t* plain, const uint32_t* rho, uint32_t* out, size_t M) {
  int rc = need_keys(ctx, true, false, false, false, "scb_paillier_encrypt");
  if (rc || M == 0) return rc;
  const size_t kN = host_keys(ctx)->kN, kc = 2 * kN;
  HostStage st(ctx, pad256(M * kc * 4) + 2 * pad256(M * kN * 4));
  if (st.rc) return st.rc;
  const uint32_t* dp = st.up(plain, M * kN);
  const uint32_t* drho = st.up(rho, M * kN);
  uint32_t* dout = st.out<uint32_t>(M * kc);
  if (st.rc) return st.rc;
  rc = scb_paillier_encrypt_dev(ctx, dp, drho, dout, M);
  if (rc == SCB_OK) st.down(out, dout, M * kc);
  return st.finish(rc);
}

extern "C" int scb_paillier_decrypt(scb_ctx* ctx, const uint32_t* c, uint32_t* out, size_t M) {
  int rc = need_keys(ctx, true, true, false, false, "scb_paillier_decrypt");
  if (rc || M == 0) return rc;
  const size_t kN = host_keys(ctx)->kN, kc = 2 * kN;
  HostStage st(ctx, pad256(M * kc * 4) + pad256(M * kN * 4));
  if (st.rc) return st.rc;
  const uint32_t* dc = st.up(c, M * kc);
  uint32_t* dout = st.out<uint32_t>(M * kN);
  if (st.rc) return st.rc;
  rc = scb_paillier_decrypt_dev(ctx, dc, dout, M);
  if (rc == SCB_OK) st.down(out, dout, M * kN);
  return st.finish(rc);
}

extern "C" int scb_keyholder_step4ab(scb_ctx* ctx, const uint32_t* z, int l, const uint32_t* rho, int rho_words, uint32_t* db_enc, size_t B) {
  int rc = need_keys(ctx, true, false, true, false, "scb_keyholder_step4ab");
  if (rc || B == 0) return rc;
  if (l < 1 || l > MAX_L) return fail(ctx, SCB_EINVAL, "scb_keyholder_step4ab: bad l");
  const size_t kN = host_keys(ctx)->kN, kn = host_keys(ctx)->kn, E = B * (l + 1);
  HostStage st(ctx, pad256(B * kN * 4) + pad256(E * kn * 4) + pad256(E * (rho ? rho_words : 0) * 4));
  if (st.rc) return st.rc;
  const uint32_t* dz = st.up(z, B * kN);
  const uint32_t* drho = st.up(rho, E * rho_words);
  uint32_t* dout = st.out<uint32_t>(E * kn);
  if (st.rc) return st.rc;
  rc = scb_keyholder_step4ab_dev(ctx, dz, l, drho, rho_words, dout, B);
  if (rc == SCB_OK) st.down(db_enc, dout, E * kn);
  return st.finish(rc);
}

extern "C" int scb_initiator_step4(scb_ctx* ctx, const uint32_t* db_enc, const uint32_t* r, const uint8_t* delta_a, const uint32_t* r_blind,
                                   int rb_words, const uint8_t* perm, const uint32_t* rho, int rho_words, int l, uint32_t* c_out, size_t B) {
  int rc = need_keys(ctx, true, false, true, false, "scb_initiator_step4");
  if (rc || B == 0) return rc;
  if (l < 1 || l > MAX_L) return fail(ctx, SCB_EINVAL, "scb_initiator_step4: bad l");
  const size_t kN = host_keys(ctx)->kN, kn = host_keys(ctx)->kn, E = B * (l + 1);
  HostStage st(ctx, 2 * pad256(E * kn * 4) + pad256(B * kN * 4) + pad256(B) + pad256(E * (r_blind ? rb_words : 0) * 4) + pad256(E) +
                        pad256(E * (rho ? rho_words : 0) * 4));
  if (st.rc) return st.rc;
  const uint32_t* ddb = st.up(db_enc, E * kn);
  const uint32_t* dr = st.up(r, B * kN);
  const uint8_t* dda = st.up(delta_a, B);
  const uint32_t* drb = st.up(r_blind, E * rb_words);
  const uint8_t* dperm = st.up(perm, E);
  const uint32_t* drho = st.up(rho, E * rho_words);
  uint32_t* dout = st.out<uint32_t>(E * kn);
  if (st.rc) return st.rc;
  rc = scb_initiator_step4_dev(ctx, ddb, dr, dda, drb, rb_words, dperm, drho, rho_words, l, dout, B);
  if (rc == SCB_OK || rc == SCB_ENOTINV) st.down(c_out, dout, E * kn);
  return st.finish(rc);
}

extern "C" int scb_dgk_is_zero(scb_ctx* ctx, const uint32_t* c, uint8_t* flags, size_t M) {
  int rc = need_keys(ctx, false, false, true, true, "scb_dgk_is_zero");
  if (rc || M == 0) return rc;
  const size_t kn = host_keys(ctx)->kn;
  HostStage st(ctx, pad256(M * kn * 4) + pad256(M));
  if (st.rc) return st.rc;
  const uint32_t* dc = st.up(c, M * kn);
  uint8_t* df = st.out<uint8_t>(M);
  if (st.rc) return st.rc;
  rc = scb_dgk_is_zero_dev(ctx, dc, df, M);
  if (rc == SCB_OK) st.down(flags, df, M);
  return st.finish(rc);
}

extern "C" int scb_keyholder_step4j_5(scb_ctx* ctx, const uint32_t* c, const uint32_t* z, const uint32_t* rho, int l, uint8_t* delta_b,
                                      uint32_t* enc3, size_t B) {
  int rc = need_keys(ctx, true, false, true, true, "scb_keyholder_step4j_5");
  if (rc || B == 0) return rc;
  if (l < 1 || l > MAX_L) return fail(ctx, SCB_EINVAL, "scb_keyholder_step4j_5: bad l");
  const size_t kN = host_keys(ctx)->kN, kc = 2 * kN, kn = host_keys(ctx)->kn, E = B * (l + 1);
  HostStage st(ctx, pad256(E * kn * 4) + pad256(B * kN * 4) + pad256(3 * B * kN * 4) + pad256(B) + pad256(3 * B * kc * 4));
  if (st.rc) return st.rc;
  const uint32_t* dc = st.up(c, E * kn);
  const uint32_t* dz = st.up(z, B * kN);
  const uint32_t* drho = st.up(rho, 3 * B * kN);
  uint8_t* ddb = st.out<uint8_t>(B);
  uint32_t* dout = st.out<uint32_t>(3 * B * kc);
  if (st.rc) return st.rc;
  rc = scb_keyholder_step4j_5_dev(ctx, dc, dz, drho, l, ddb, dout, B);
  if (rc == SCB_OK) {
    st.down(delta_b, ddb, B);
    st.down(enc3, dout, 3 * B * kc);
  }
  return st.finish(rc);
}

extern "C" int scb_keyholder_step5(scb_ctx* ctx, const uint32_t* z, const uint8_t* delta_b, const uint32_t* rho, int l, uint32_t* enc3, size_t B) {
  int rc = need_keys(ctx, true, false, false, false, "scb_keyholder_step5");
  if (rc || B == 0) return rc;
  if (l < 1 || l > MAX_L) return fail(ctx, SCB_EINVAL, "scb_keyholder_step5: bad l");
  const size_t kN = host_keys(ctx)->kN, kc = 2 * kN;
  HostStage st(ctx, pad256(B * kN * 4) + pad256(3 * B * kN * 4) + pad256(B) + pad256(3 * B * kc * 4));
  if (st.rc) return st.rc;
  const uint32_t* dz = st.up(z, B * kN);
  const uint8_t* ddb = st.up(delta_b, B);
  const uint32_t* drho = st.up(rho, 3 * B * kN);
  uint32_t* dout = st.out<uint32_t>(3 * B * kc);
  if (st.rc) return st.rc;
  rc = scb_keyholder_step5_dev(ctx, dz, ddb, drho, l, dout, B);
  if (rc == SCB_OK) st.down(enc3, dout, 3 * B * kc);
  return st.finish(rc);
}

extern "C" int scb_initiator_step6_7(scb_ctx* ctx, const uint32_t* enc3, const uint32_t* r, const uint8_t* delta_a, int l, uint32_t* result,
                                     size_t B) {
  int rc = need_keys(ctx, true, false, false, false, "scb_initiator_step6_7");
  if (rc || B == 0) return rc;
  const size_t kN = host_keys(ctx)->kN, kc = 2 * kN;
  HostStage st(ctx, pad256(3 * B * kc * 4) + pad256(B * kN * 4) + pad256(B) + pad256(B * kc * 4));
  if (st.rc) return st.rc;
  const uint32_t* de = st.up(enc3, 3 * B * kc);
  const uint32_t* dr = st.up(r, B * kN);
  const uint8_t* dda = st.up(delta_a, B);
  uint32_t* dout = st.out<uint32_t>(B * kc);
  if (st.rc) return st.rc;
  rc = scb_initiator_step6_7_dev(ctx, de, dr, dda, l, dout, B);
  if (rc == SCB_OK || rc == SCB_ENOTINV) st.down(result, dout, B * kc);
  return st.finish(rc);
}
