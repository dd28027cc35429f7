// scb200_rng.cu -- device-side CSPRNG for the protocol's nonces.
//
// The reference draws r = randbelow(N), delta_A = randbelow(2), r_i = randbelow(u-1)+1, the shuffle
// (initiator.py:223,250,420,512) from `secrets`, and the randomizer bases / exponents from the schemes' randomness
// pools (initiator.py:205-210, keyholder.py:174-179).  In the batched engine those are arrays of up to 10^6..10^7 big
// integers per pass; generating them on the host and copying them over PCIe costs more than the GPU arithmetic they
// feed.  This file generates them on the device with ChaCha20 in counter mode (RFC 8439 block function; 256-bit key
// from the host's `secrets`), one independent keystream position per (call, element, attempt, block):
//     state[12] = attempt << 16 | block     state[13] = element (low 32)
//     state[14] = element (high 32) ^ nonce0     state[15] = call id ^ nonce1
// so every output is a deterministic function of (key, nonce, call order) -- which is what lets the tests pin the
// samplers bit for bit against a Python model (tests/chacha_model.py) -- and no two blocks ever share a counter.
// Uniform values below a bound use rejection sampling on bit-length-masked candidates (exactly uniform).
// Injected host nonces remain the parity path; this is the production path.
#include "../../../include/scb200.h"

#include <cstring>

#include "host_ctx.hpp"

using namespace scb;

namespace {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

#define SCB_QR(a, b, c, d) \
  a += b; d ^= a; d = rotl32(d, 16); \
  c += d; b ^= c; b = rotl32(b, 12); \
  a += b; d ^= a; d = rotl32(d, 8);  \
  c += d; b ^= c; b = rotl32(b, 7);

// one ChaCha20 block: key[8], words 12..15 of the state -> 16 output words
__device__ __forceinline__ void chacha20_block(const uint32_t* __restrict__ key, uint32_t s12, uint32_t s13, uint32_t s14, uint32_t s15,
                                               uint32_t (&out)[16]) {
  uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                     key[4], key[5], key[6], key[7], s12, s13, s14, s15};
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = in[i];
#pragma unroll
  for (int r = 0; r < 10; r++) {
    SCB_QR(x[0], x[4], x[8], x[12]) SCB_QR(x[1], x[5], x[9], x[13]) SCB_QR(x[2], x[6], x[10], x[14]) SCB_QR(x[3], x[7], x[11], x[15])
    SCB_QR(x[0], x[5], x[10], x[15]) SCB_QR(x[1], x[6], x[11], x[12]) SCB_QR(x[2], x[7], x[8], x[13]) SCB_QR(x[3], x[4], x[9], x[14])
  }
#pragma unroll
  for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

struct RngParams {
  uint32_t key[8];
  uint32_t nonce0, nonce1, call;
};

// raw block (known-answer tests): out[16] for explicit state words
__global__ void k_chacha_block(RngParams p, uint32_t s12, uint32_t s13, uint32_t s14, uint32_t s15, uint32_t* out) {
  uint32_t o[16];
  chacha20_block(p.key, s12, s13, s14, s15, o);
  for (int i = 0; i < 16; i++) out[i] = o[i];
}

// out[e] = uniform integer in [min_one, bound) of k words; `bits` = bit length of bound.  One thread per element.
__global__ void __launch_bounds__(128) k_rng_below(RngParams p, const uint32_t* __restrict__ bound, int k, int bits, int min_one,
                                                   uint32_t* __restrict__ out, size_t B) {
  const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= B) return;
  uint32_t* dst = out + e * k;
  const int top = (bits - 1) >> 5;                               // index of the top significant word
  const uint32_t top_mask = (bits & 31) ? ((1u << (bits & 31)) - 1u) : 0xffffffffu;
  const uint32_t e_lo = static_cast<uint32_t>(e), e_hi = static_cast<uint32_t>(e >> 32);
  for (uint32_t attempt = 0;; attempt++) {
    for (int blk = 0; blk * 16 <= top; blk++) {
      uint32_t o[16];
      chacha20_block(p.key, (attempt << 16) | static_cast<uint32_t>(blk), e_lo, e_hi ^ p.nonce0, p.call ^ p.nonce1, o);
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int w = blk * 16 + i;
        if (w <= top) dst[w] = (w == top) ? (o[i] & top_mask) : o[i];
      }
    }
    for (int w = top + 1; w < k; w++) dst[w] = 0u;
    // accept iff min_one <= candidate < bound (compare from the most significant word)
    int cmp = 0;
    bool nonzero = false;
    for (int w = k - 1; w >= 0; w--) {
      const uint32_t c = dst[w], b = bound[w];
      nonzero = nonzero || c != 0u;
      if (cmp == 0 && c != b) cmp = c < b ? -1 : 1;
    }
    if (cmp < 0 && (nonzero || !min_one)) break;
    if (attempt == 0xffffu) break;                               // unreachable in practice (probability < 2^-65535)
  }
}

// out[e] = `bits` random bits in `words` words
__global__ void __launch_bounds__(128) k_rng_bits(RngParams p, int bits, int words, uint32_t* __restrict__ out, size_t B) {
  const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= B) return;
  uint32_t* dst = out + e * words;
  const uint32_t e_lo = static_cast<uint32_t>(e), e_hi = static_cast<uint32_t>(e >> 32);
  for (int blk = 0; blk * 16 < words; blk++) {
    uint32_t o[16];
    chacha20_block(p.key, static_cast<uint32_t>(blk), e_lo, e_hi ^ p.nonce0, p.call ^ p.nonce1, o);
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int w = blk * 16 + i;
      if (w < words) {
        const int lo_bit = 32 * w;
        uint32_t v = o[i];
        if (lo_bit >= bits) v = 0u;
        else if (bits - lo_bit < 32) v &= (1u << (bits - lo_bit)) - 1u;
        dst[w] = v;
      }
    }
  }
}

// out[e] = one uniform bit (0/1) per element; 512 elements share a block
__global__ void __launch_bounds__(128) k_rng_coin(RngParams p, uint8_t* __restrict__ out, size_t B) {
  const size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;     // group of 512 coins
  if (g * 512 >= B) return;
  uint32_t o[16];
  chacha20_block(p.key, 0u, static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32) ^ p.nonce0, p.call ^ p.nonce1, o);
  for (int i = 0; i < 512; i++) {
    const size_t e = g * 512 + i;
    if (e < B) out[e] = (o[i >> 5] >> (i & 31)) & 1u;
  }
}

// out[e][0..n) = uniform permutation of 0..n-1 (Fisher-Yates, bounded integers by mask-and-reject on keystream words)
__global__ void __launch_bounds__(128) k_rng_perm(RngParams p, int n, uint8_t* __restrict__ out, size_t B) {
  const size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= B) return;
  uint8_t perm[256];
  for (int i = 0; i < n; i++) perm[i] = static_cast<uint8_t>(i);
  const uint32_t e_lo = static_cast<uint32_t>(e), e_hi = static_cast<uint32_t>(e >> 32);
  uint32_t o[16];
  uint32_t blk = 0;
  int have = 0;
  for (int j = n - 1; j >= 1; j--) {
    uint32_t mask = static_cast<uint32_t>(j);
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8;
    uint32_t r;
    do {
      if (have == 0) {
        chacha20_block(p.key, blk++, e_lo, e_hi ^ p.nonce0, p.call ^ p.nonce1, o);
        have = 16;
      }
      r = o[16 - have] & mask;
      have--;
    } while (r > static_cast<uint32_t>(j));
    const uint8_t t = perm[j];
    perm[j] = perm[r];
    perm[r] = t;
  }
  for (int i = 0; i < n; i++) out[e * n + i] = perm[i];
}

RngParams params_for_call(scb_ctx* ctx) {
  RngParams p;
  memcpy(p.key, ctx->rng_key, sizeof(p.key));
  p.nonce0 = ctx->rng_nonce[0];
  p.nonce1 = ctx->rng_nonce[1];
  p.call = ctx->rng_call++;
  return p;
}

int need_rng(scb_ctx* ctx, const char* who) {
  if (!ctx) return SCB_EINVAL;
  if (!ctx->rng_seeded) return fail(ctx, SCB_ESTATE, "%s: scb_rng_seed has not been called", who);
  cudaSetDevice(ctx->device);
  return SCB_OK;
}

}  // namespace

extern "C" int scb_rng_seed(scb_ctx* ctx, const uint32_t key[8], const uint32_t nonce[2]) {
  if (!ctx || !key || !nonce) return SCB_EINVAL;
  memcpy(ctx->rng_key, key, sizeof(ctx->rng_key));
  ctx->rng_nonce[0] = nonce[0];
  ctx->rng_nonce[1] = nonce[1];
  ctx->rng_call = 0;
  ctx->rng_seeded = true;
  return SCB_OK;
}

extern "C" int scb_rng_chacha_block_dev(scb_ctx* ctx, const uint32_t counter_nonce[4], uint32_t* out16) {
  int rc = need_rng(ctx, "scb_rng_chacha_block");
  if (rc) return rc;
  if (!counter_nonce || !out16) return fail(ctx, SCB_EINVAL, "scb_rng_chacha_block: null pointer");
  RngParams p;
  memcpy(p.key, ctx->rng_key, sizeof(p.key));
  p.nonce0 = p.nonce1 = p.call = 0;
  k_chacha_block<<<1, 1, 0, ctx->stream>>>(p, counter_nonce[0], counter_nonce[1], counter_nonce[2], counter_nonce[3], out16);
  return ctx->launched("k_chacha_block");
}

extern "C" int scb_rng_below_dev(scb_ctx* ctx, const uint32_t* bound, int k, int min_one, uint32_t* out, size_t B) {
  int rc = need_rng(ctx, "scb_rng_below");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!bound || !out || k <= 0 || k > 512) return fail(ctx, SCB_EINVAL, "scb_rng_below: bad argument");
  int bits = 0;
  for (int i = k - 1; i >= 0 && !bits; i--)
    if (bound[i]) bits = i * 32 + 32 - __builtin_clz(bound[i]);
  if (bits == 0 || (min_one && bits == 1 && bound[0] == 1)) return fail(ctx, SCB_EINVAL, "scb_rng_below: empty range");
  void* dbound = nullptr;
  rc = ctx->scratch(9, sizeof(uint32_t) * k, &dbound);
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(dbound, bound, sizeof(uint32_t) * k, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "upload bound");
  k_rng_below<<<static_cast<unsigned>((B + 127) / 128), 128, 0, ctx->stream>>>(params_for_call(ctx), static_cast<const uint32_t*>(dbound), k, bits,
                                                                                min_one ? 1 : 0, out, B);
  return ctx->launched("k_rng_below");
}

extern "C" int scb_rng_bits_dev(scb_ctx* ctx, int bits, int words, uint32_t* out, size_t B) {
  int rc = need_rng(ctx, "scb_rng_bits");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!out || bits <= 0 || words <= 0 || bits > 32 * words || words > 65535 * 16) return fail(ctx, SCB_EINVAL, "scb_rng_bits: bad argument");
  k_rng_bits<<<static_cast<unsigned>((B + 127) / 128), 128, 0, ctx->stream>>>(params_for_call(ctx), bits, words, out, B);
  return ctx->launched("k_rng_bits");
}

extern "C" int scb_rng_coins_dev(scb_ctx* ctx, uint8_t* out, size_t B) {
  int rc = need_rng(ctx, "scb_rng_coins");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!out) return fail(ctx, SCB_EINVAL, "scb_rng_coins: null pointer");
  const size_t groups = (B + 511) / 512;
  k_rng_coin<<<static_cast<unsigned>((groups + 127) / 128), 128, 0, ctx->stream>>>(params_for_call(ctx), out, B);
  return ctx->launched("k_rng_coin");
}

extern "C" int scb_rng_permutations_dev(scb_ctx* ctx, int n, uint8_t* out, size_t B) {
  int rc = need_rng(ctx, "scb_rng_permutations");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!out || n < 1 || n > 256) return fail(ctx, SCB_EINVAL, "scb_rng_permutations: bad argument");
  k_rng_perm<<<static_cast<unsigned>((B + 127) / 128), 128, 0, ctx->stream>>>(params_for_call(ctx), n, out, B);
  return ctx->launched("k_rng_perm");
}

// host-buffer convenience: generate on the device, copy back (tests; callers that want GPU-made nonces on the host)
extern "C" int scb_rng_fetch(scb_ctx* ctx, int what, const uint32_t* bound, int k, int arg, void* out_host, size_t B) {
  int rc = need_rng(ctx, "scb_rng_fetch");
  if (rc) return rc;
  if (!out_host) return fail(ctx, SCB_EINVAL, "scb_rng_fetch: null pointer");
  size_t bytes = 0;
  switch (what) {
    case SCB_RNG_BELOW: bytes = B * k * sizeof(uint32_t); break;
    case SCB_RNG_BITS: bytes = B * k * sizeof(uint32_t); break;
    case SCB_RNG_COINS: bytes = B; break;
    case SCB_RNG_PERMUTATIONS: bytes = B * static_cast<size_t>(arg); break;
    case SCB_RNG_BLOCK: bytes = 16 * sizeof(uint32_t); break;
    default: return fail(ctx, SCB_EINVAL, "scb_rng_fetch: unknown kind %d", what);
  }
  if (bytes == 0) return SCB_OK;
  void* dev = nullptr;
  rc = ctx->scratch(2, bytes + 256, &dev);
  if (rc) return rc;
  switch (what) {
    case SCB_RNG_BELOW: rc = scb_rng_below_dev(ctx, bound, k, arg, static_cast<uint32_t*>(dev), B); break;
    case SCB_RNG_BITS: rc = scb_rng_bits_dev(ctx, arg, k, static_cast<uint32_t*>(dev), B); break;
    case SCB_RNG_COINS: rc = scb_rng_coins_dev(ctx, static_cast<uint8_t*>(dev), B); break;
    case SCB_RNG_PERMUTATIONS: rc = scb_rng_permutations_dev(ctx, arg, static_cast<uint8_t*>(dev), B); break;
    case SCB_RNG_BLOCK: rc = scb_rng_chacha_block_dev(ctx, bound, static_cast<uint32_t*>(dev)); break;
  }
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(out_host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "D2H copy");
  return scb_sync(ctx);
}
