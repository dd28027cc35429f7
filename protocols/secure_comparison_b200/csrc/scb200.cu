// scb200.cu -- host side of the C ABI declared in include/scb200.h (context, modulus constants,
// launch geometry) for the sm_100a kernels in kernels_generic.cuh / kernels_protocol.cuh.
#include "../../../include/scb200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "host_ctx.hpp"
#include "kernels_generic.cuh"
#include "kernels_f64.cuh"
#include "kernels_f64_pair.cuh"
#include "kernels_fb52.cuh"

using namespace scb;

// -------------------------------------------------------------------------------------------------
// small host-side big-integer helpers (only used once per modulus, for the Montgomery constants)
// -------------------------------------------------------------------------------------------------
namespace {

typedef std::vector<uint32_t> Words;

int words_bits(const Words& x) {
  for (int i = static_cast<int>(x.size()) - 1; i >= 0; i--)
    if (x[i]) return i * 32 + 32 - __builtin_clz(x[i]);
  return 0;
}
bool words_geq(const Words& a, const Words& b) {   // same length
  for (int i = static_cast<int>(a.size()) - 1; i >= 0; i--)
    if (a[i] != b[i]) return a[i] > b[i];
  return true;
}
void words_sub(Words& a, const Words& b) {
  uint64_t borrow = 0;
  for (size_t i = 0; i < a.size(); i++) {
    const uint64_t d = static_cast<uint64_t>(a[i]) - b[i] - borrow;
    a[i] = static_cast<uint32_t>(d);
    borrow = (d >> 63) & 1u;
  }
}
// x = 2x mod n   (x < n; arrays have one spare top word)
void words_dbl_mod(Words& x, const Words& n) {
  uint32_t carry = 0;
  for (size_t i = 0; i < x.size(); i++) {
    const uint32_t nc = x[i] >> 31;
    x[i] = (x[i] << 1) | carry;
    carry = nc;
  }
  if (words_geq(x, n)) words_sub(x, n);
}
void words_to_digits(const Words& x, uint32_t* digits, int K) {
  for (int d = 0; d < K; d++) {
    const int bit = d * W, w = bit >> 5, sh = bit & 31;
    uint64_t v = (static_cast<size_t>(w) < x.size()) ? x[w] : 0u;
    if (static_cast<size_t>(w + 1) < x.size()) v |= static_cast<uint64_t>(x[w + 1]) << 32;
    digits[d] = static_cast<uint32_t>(v >> sh) & DMASK;
  }
}

}  // namespace

int scb::config_K_for_bits(int bits) {
  // R = 2^(29K) must exceed 4n; supported (L, TPI) products
  static const int Ks[] = {36, 54, 72, 108, 144, 216, 288};
  for (int K : Ks)
    if (29 * K >= bits + 2) return K;
  return -1;
}

static int config_K52_for_bits(int bits) {
  // R = 2^(52K) must exceed 16n; geometries L x TPI: 10x2, 15x2, 10x4, 15x4, 10x8, 15x8, 10x16
  static const int Ks[] = {20, 30, 40, 60, 80, 120, 160};
  for (int K : Ks)
    if (52 * K >= bits + 4) return K;
  return 0;
}

int scb::fail(scb_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    ctx->error = buf;
  }
  return code;
}

int scb::cuda_fail(scb_ctx* ctx, cudaError_t e, const char* what) {
  return fail(ctx, e == cudaErrorMemoryAllocation ? SCB_ENOMEM : SCB_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

int scb_ctx::scratch(int which, size_t bytes, void** out) {
  if (which < 0 || which >= NSCRATCH) return fail(this, SCB_EINVAL, "bad scratch index");
  if (which == 0 && on_aux) which = SCRATCH_AUX_GROUP;      // per-group tables of a launch on the second stream
  if (scratch_bytes[which] < bytes) {
    if (scratch_ptr[which]) {
      cudaStreamSynchronize(stream);
      cudaFree(scratch_ptr[which]);
      scratch_ptr[which] = nullptr;
      scratch_bytes[which] = 0;
    }
    const size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&scratch_ptr[which], want);
    if (e != cudaSuccess) return cuda_fail(this, e, "cudaMalloc(scratch)");
    scratch_bytes[which] = want;
  }
  *out = scratch_ptr[which];
  return SCB_OK;
}

// Compile an exponent into a left-to-right sliding-window program (format: kernels_f64.cuh, group_modexp52_prog).
// Pure host code.  Returns an empty program for exponents below 2^63 (they stay on the fixed-window loop).
static std::vector<uint16_t> compile_exponent_program(const uint32_t* host_words, int nwords, int* w_out) {
  std::vector<uint16_t> prog;
  int top = -1;
  for (int i = nwords - 1; i >= 0 && top < 0; i--)
    if (host_words[i]) top = i * 32 + 31 - __builtin_clz(host_words[i]);
  *w_out = 0;
  if (top < 63) return prog;
  const int bits = top + 1;
  int w = 2;
  double best = 1e30;
  for (int cand = 2; cand <= 7; cand++) {
    const double cost = (1 << (cand - 1)) + static_cast<double>(bits) / (cand + 1);
    if (cost < best) { best = cost; w = cand; }
  }
  auto bit = [&](int i) -> int { return (host_words[i >> 5] >> (i & 31)) & 1u; };
  int i = top, pending = 0;
  bool first = true;
  auto emit = [&](int nsq, int idx) {
    while (nsq > 255) {
      prog.push_back(static_cast<uint16_t>((255 << 8) | 0xFF));
      nsq -= 255;
    }
    prog.push_back(static_cast<uint16_t>((nsq << 8) | idx));
  };
  while (i >= 0) {
    if (!bit(i)) {
      pending++;
      i--;
      continue;
    }
    int s = i - w + 1 < 0 ? 0 : i - w + 1;
    while (!bit(s)) s++;
    int val = 0;
    for (int j = i; j >= s; j--) val = (val << 1) | bit(j);
    const int idx = (val - 1) / 2;
    if (first) {
      prog.push_back(static_cast<uint16_t>(idx));
      first = false;
    } else {
      emit(pending + (i - s + 1), idx);
    }
    pending = 0;
    i = s - 1;
  }
  if (pending > 0) emit(pending, 0xFF);
  *w_out = w;
  return prog;
}

// host-only test hook: the program the library would run for a shared exponent (no GPU involved)
extern "C" int scb_debug_compile_exponent(const uint32_t* exp, int exp_words, uint16_t* prog_out, int prog_capacity, int* window_bits) {
  if (!exp || exp_words <= 0 || !prog_out || !window_bits) return SCB_EINVAL;
  int w = 0;
  const std::vector<uint16_t> prog = compile_exponent_program(exp, exp_words, &w);
  *window_bits = w;
  if (static_cast<int>(prog.size()) > prog_capacity) return SCB_EINVAL;
  for (size_t i = 0; i < prog.size(); i++) prog_out[i] = prog[i];
  return static_cast<int>(prog.size());
}

// Keep the program of a shared exponent on the device, keyed by the device address of the exponent words.
int scb_ctx::register_exponent(const uint32_t* dev_words, const uint32_t* host_words, int nwords) {
  auto it = exp_programs.find(dev_words);
  if (it != exp_programs.end()) {
    cudaFree(it->second.dev);
    exp_programs.erase(it);
  }
  int w = 0;
  const std::vector<uint16_t> prog = compile_exponent_program(host_words, nwords, &w);
  if (prog.empty()) return SCB_OK;                 // short (or zero) exponents stay on the fixed-window loop
  ExpProgram ep;
  ep.n = static_cast<int>(prog.size());
  ep.w = w;
  cudaError_t e = cudaMalloc(&ep.dev, sizeof(uint16_t) * prog.size());
  if (e != cudaSuccess) return cuda_fail(this, e, "cudaMalloc(exponent program)");
  e = cudaMemcpyAsync(ep.dev, prog.data(), sizeof(uint16_t) * prog.size(), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) {
    cudaFree(ep.dev);
    return cuda_fail(this, e, "upload exponent program");
  }
  exp_programs[dev_words] = ep;
  return SCB_OK;
}

const uint16_t* scb_ctx::find_program(const uint32_t* dev_words, int* nprog, int* pw, int* wbits) const {
  *nprog = 0;
  *pw = 0;
  if (!use_sliding) return nullptr;
  auto it = exp_programs.find(dev_words);
  if (it == exp_programs.end()) return nullptr;
  *nprog = it->second.n;
  *pw = it->second.w;
  if (wbits && it->second.w - 1 > *wbits) *wbits = it->second.w - 1;
  return it->second.dev;
}

void scb_ctx::unregister_exponent(const uint32_t* dev_words) {
  auto it = exp_programs.find(dev_words);
  if (it == exp_programs.end()) return;
  cudaStreamSynchronize(stream);
  cudaFree(it->second.dev);
  exp_programs.erase(it);
}

void scb_ctx::clear_exponent_programs() {
  for (auto& kv : exp_programs) cudaFree(kv.second.dev);
  exp_programs.clear();
}

unsigned long long* scb_ctx::fresh_queue() {
  constexpr int RING = 64;
  if (!queue_ring && cudaMalloc(&queue_ring, sizeof(unsigned long long) * RING) != cudaSuccess) return nullptr;
  unsigned long long* q = queue_ring + queue_next;
  queue_next = (queue_next + 1) % RING;
  if (cudaMemsetAsync(q, 0, sizeof(unsigned long long), stream) != cudaSuccess) return nullptr;
  return q;
}

// -------------------------------------------------------------------------------------------------
// context
// -------------------------------------------------------------------------------------------------
extern "C" int scb_version(void) { return 200; }
#ifndef SCB_BUILD_ID
#define SCB_BUILD_ID "unversioned"
#endif
extern "C" const char* scb_build_id(void) { return SCB_BUILD_ID; }

extern "C" int scb_ctx_create(int device, scb_ctx** out) {
  if (!out) return SCB_EINVAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || device < 0 || device >= count) return SCB_ECUDA;
  scb_ctx* ctx = new scb_ctx();
  ctx->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) {
    delete ctx;
    return SCB_ECUDA;
  }
  *out = ctx;
  return SCB_OK;
}

extern "C" void scb_ctx_destroy(scb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->aux_stream) {
    cudaStreamSynchronize(ctx->aux_stream);
    cudaStreamDestroy(ctx->aux_stream);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
  }
  for (auto& m : ctx->mod) {
    if (m.dev) cudaFree(m.dev);
    if (m.dev52) cudaFree(m.dev52);
    if (m.pair52) cudaFree(m.pair52);
  }
  for (auto& t : ctx->fb) {
    if (t.table) cudaFree(t.table);
    if (t.table16) cudaFree(t.table16);
    if (t.table52) cudaFree(t.table52);
    if (t.table16_52) cudaFree(t.table16_52);
  }
  for (int i = 0; i < scb_ctx::NSCRATCH; i++)
    if (ctx->scratch_ptr[i]) cudaFree(ctx->scratch_ptr[i]);
  if (ctx->keys_dev) cudaFree(ctx->keys_dev);
  if (ctx->keys52_dev) cudaFree(ctx->keys52_dev);
  if (ctx->secret_dev) cudaFree(ctx->secret_dev);
  if (ctx->vp_dev) cudaFree(ctx->vp_dev);
  if (ctx->crt_exp_dev) cudaFree(ctx->crt_exp_dev);
  if (ctx->queue_ring) cudaFree(ctx->queue_ring);
  ctx->clear_exponent_programs();
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* scb_last_error(const scb_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }
extern "C" void* scb_stream(scb_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }
extern "C" unsigned long long scb_launch_count(const scb_ctx* ctx) { return ctx ? ctx->launches : 0ull; }

extern "C" int scb_sync(scb_ctx* ctx) {
  if (!ctx) return SCB_EINVAL;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaStreamSynchronize");
  return SCB_OK;
}

extern "C" int scb_host_alloc(size_t bytes, void** out) {
  if (!out) return SCB_EINVAL;
  return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? SCB_OK : SCB_ENOMEM;
}
extern "C" int scb_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? SCB_OK : SCB_ECUDA; }

extern "C" int scb_dev_alloc(scb_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return SCB_EINVAL;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(scb_dev_alloc)");
  return SCB_OK;
}
extern "C" int scb_dev_free(scb_ctx* ctx, void* ptr) {
  if (!ctx) return SCB_EINVAL;
  if (!ptr) return SCB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaFree(scb_dev_free)");
  return SCB_OK;
}
extern "C" int scb_copy_h2d(scb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  if (!ctx || (bytes && (!dst_dev || !src_host))) return fail(ctx, SCB_EINVAL, "scb_copy_h2d: null pointer");
  if (!bytes) return SCB_OK;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "H2D copy");
  return SCB_OK;
}
extern "C" int scb_copy_d2h(scb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  if (!ctx || (bytes && (!dst_host || !src_dev))) return fail(ctx, SCB_EINVAL, "scb_copy_d2h: null pointer");
  if (!bytes) return SCB_OK;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "D2H copy");
  return SCB_OK;
}

// -------------------------------------------------------------------------------------------------
// modulus constants
// -------------------------------------------------------------------------------------------------
extern "C" int scb_modulus_load(scb_ctx* ctx, int slot, const uint32_t* n, int k32) {
  if (!ctx || !n || slot < 0 || slot >= SCB_MOD_SLOTS || k32 <= 0 || k32 > INV_MAXW - 2 || (32 * k32 + 2 + 29) / 30 > INV_MAXL) return fail(ctx, SCB_EINVAL, "scb_modulus_load: bad argument");
  if ((n[0] & 1u) == 0) return fail(ctx, SCB_EINVAL, "scb_modulus_load: modulus must be odd");
  Words nn(n, n + k32);
  const int bits = words_bits(nn);
  if (bits < 2) return fail(ctx, SCB_EINVAL, "scb_modulus_load: modulus too small");
  const int K = config_K_for_bits(bits);
  if (K < 0) return fail(ctx, SCB_EINVAL, "scb_modulus_load: %d-bit modulus exceeds the supported 8350 bits", bits);
  cudaSetDevice(ctx->device);
  ModulusDev h;
  memset(&h, 0, sizeof(h));
  h.K = K;
  h.k32 = k32;
  h.bits = bits;
  // n' = -n^-1 mod 2^29 (Newton on 32 bits)
  uint32_t inv = n[0];
  for (int i = 0; i < 5; i++) inv *= 2u - n[0] * inv;
  h.np = (0u - inv) & DMASK;
  Words nw(nn);
  nw.push_back(0);
  words_to_digits(nw, h.n, K);
  h.unit[0] = 1;
  // x = 2^(bits-1) mod n = 2^(bits-1) (< n), then double up to 2^(29K), 2^(58K), 2^(87K)
  Words x(k32 + 1, 0);
  x[(bits - 1) >> 5] = 1u << ((bits - 1) & 31);
  int e = bits - 1;
  const int targets[3] = {W * K, 2 * W * K, 3 * W * K};
  uint32_t* dst[3] = {h.one, h.r2, h.r3};
  for (int t = 0; t < 3; t++) {
    for (; e < targets[t]; e++) words_dbl_mod(x, nw);
    words_to_digits(x, dst[t], K);
  }
  // safegcd constants of k_modinv: n in 30-bit limbs, n^-1 mod 2^30; operands are k32-word values
  h.len30 = (32 * k32 + 2 + 29) / 30;
  h.inv_cap = (49 * 32 * k32 + 57) / 17 / 30 + 2;
  h.ninv30 = inv & ((1u << 30) - 1u);
  for (int i = 0; i < h.len30; i++) {
    const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
    uint64_t v = (w < k32) ? n[w] : 0u;
    if (w + 1 < k32) v |= static_cast<uint64_t>(n[w + 1]) << 32;
    h.n30[i] = static_cast<int32_t>(static_cast<uint32_t>(v >> sh) & ((1u << 30) - 1u));
  }
  scb_ctx::Mod& md = ctx->mod[slot];
  if (!md.dev) {
    cudaError_t ce = cudaMalloc(&md.dev, sizeof(ModulusDev));
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "cudaMalloc(modulus)");
  }
  cudaError_t ce = cudaMemcpyAsync(md.dev, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "upload modulus");
  md.K = K;
  md.k32 = k32;
  md.bits = bits;
  md.words.assign(n, n + k32);
  md.loaded = true;
  md.root_slot = -1;                                      // a reloaded modulus is no longer a linked square ...
  for (auto& other : ctx->mod)
    if (other.root_slot == slot) other.root_slot = -1;    // ... nor the root of one
  // ---- FP64 engine constants (52-bit digits as doubles): n, n' = -n^-1 mod 2^52, R, R^2, R^3 mod n with R = 2^(52 K52)
  md.K52 = config_K52_for_bits(bits);
  if (md.K52 > 0 && 32 * k32 <= 52 * md.K52) {
    const int K52 = md.K52;
    std::vector<uint8_t> buf(sizeof(Modulus52Dev), 0);
    Modulus52Dev* h52 = reinterpret_cast<Modulus52Dev*>(buf.data());
    h52->K = K52;
    h52->k32 = k32;
    uint64_t n64 = n[0] | (k32 > 1 ? static_cast<uint64_t>(n[1]) << 32 : 0ull), inv64 = n64;
    for (int i = 0; i < 6; i++) inv64 *= 2ull - n64 * inv64;
    h52->np = static_cast<double>((0ull - inv64) & ((1ull << 52) - 1ull));
    auto to52 = [&](const Words& x, double* digits) {
      for (int d = 0; d < K52; d++) {
        const int bit = d * 52, w = bit >> 5, sh = bit & 31;
        auto word = [&](int i) -> uint64_t { return static_cast<size_t>(i) < x.size() ? x[i] : 0u; };
        uint64_t v = (word(w) | (word(w + 1) << 32)) >> sh;
        if (sh) v |= word(w + 2) << (64 - sh);
        digits[d] = static_cast<double>(v & ((1ull << 52) - 1ull));
      }
    };
    to52(nw, h52->n);
    h52->unit[0] = 1.0;
    Words y(k32 + 1, 0);
    y[(bits - 1) >> 5] = 1u << ((bits - 1) & 31);
    int ee = bits - 1;
    const int targets52[3] = {52 * K52, 104 * K52, 156 * K52};
    double* dst52[3] = {h52->one, h52->r2, h52->r3};
    for (int t = 0; t < 3; t++) {
      for (; ee < targets52[t]; ee++) words_dbl_mod(y, nw);
      to52(y, dst52[t]);
      if (t == 0) md.r52_words.assign(y.begin(), y.begin() + k32);
    }
    if (!md.dev52) {
      cudaError_t ce2 = cudaMalloc(&md.dev52, sizeof(Modulus52Dev));
      if (ce2 != cudaSuccess) return cuda_fail(ctx, ce2, "cudaMalloc(modulus52)");
    }
    cudaError_t ce2 = cudaMemcpyAsync(md.dev52, h52, sizeof(Modulus52Dev), cudaMemcpyHostToDevice, ctx->stream);
    if (ce2 == cudaSuccess) ce2 = cudaStreamSynchronize(ctx->stream);
    if (ce2 != cudaSuccess) return cuda_fail(ctx, ce2, "upload modulus52");
  } else {
    md.K52 = 0;
    md.r52_words.clear();
  }
  return SCB_OK;
}

// -------------------------------------------------------------------------------------------------
// square moduli: constants of the base-n pair arithmetic (kernels_f64_pair.cuh), derived once per key on the host
// -------------------------------------------------------------------------------------------------
namespace {

Words words_mul(const Words& a, const Words& b) {
  Words r(a.size() + b.size(), 0);
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

// a = q * n + r by shift-and-subtract (one-off, 4096..8192-bit values); n has no leading zero WORD requirement
void words_divmod(const Words& a, const Words& n, Words& q, Words& r) {
  Words nn(n);
  nn.push_back(0);
  r.assign(nn.size(), 0);
  q.assign(a.size(), 0);
  for (int bit = words_bits(a) - 1; bit >= 0; bit--) {
    uint32_t carry = (a[bit >> 5] >> (bit & 31)) & 1u;
    for (size_t i = 0; i < r.size(); i++) {
      const uint32_t nc = r[i] >> 31;
      r[i] = (r[i] << 1) | carry;
      carry = nc;
    }
    if (words_geq(r, nn)) {
      words_sub(r, nn);
      q[bit >> 5] |= 1u << (bit & 31);
    }
  }
  r.pop_back();
}

}  // namespace

int scb::link_square_modulus(scb_ctx* ctx, int square_slot, int root_slot) {
  if (!ctx) return SCB_EINVAL;
  if (square_slot < 0 || square_slot >= SCB_MOD_SLOTS || root_slot < 0 || root_slot >= SCB_MOD_SLOTS || square_slot == root_slot)
    return fail(ctx, SCB_EINVAL, "scb_modulus_link_square: bad slot");
  scb_ctx::Mod& sq = ctx->mod[square_slot];
  const scb_ctx::Mod& rt = ctx->mod[root_slot];
  if (!sq.loaded || !rt.loaded) return fail(ctx, SCB_ESTATE, "scb_modulus_link_square: both moduli must be loaded first");
  Words n2 = words_mul(rt.words, rt.words);
  n2.resize(std::max(n2.size(), sq.words.size()), 0);
  Words have(sq.words);
  have.resize(n2.size(), 0);
  if (n2 != have) return fail(ctx, SCB_EINVAL, "scb_modulus_link_square: slot %d does not hold the square of slot %d", square_slot, root_slot);
  sq.root_slot = -1;
  if (rt.K52 <= 0 || 2 * rt.K52 > MAXK52) return SCB_OK;      // no FP64 geometry for the root: n^2 stays an ordinary modulus
  cudaSetDevice(ctx->device);
  const int K = rt.K52;
  std::vector<uint8_t> buf(sizeof(PairConsts52Dev), 0);
  PairConsts52Dev* h = reinterpret_cast<PairConsts52Dev*>(buf.data());
  h->K = K;
  h->w = rt.bits - 1;
  Words n2w(n2);
  while (n2w.size() > 1 && n2w.back() == 0) n2w.pop_back();
  n2w.push_back(0);                                            // spare top word for the doubling
  auto to52 = [&](const Words& x, double* digits) {
    for (int d = 0; d < K; d++) {
      const int bit = d * 52, w = bit >> 5, sh = bit & 31;
      auto word = [&](int i) -> uint64_t { return static_cast<size_t>(i) < x.size() ? x[i] : 0u; };
      uint64_t v = (word(w) | (word(w + 1) << 32)) >> sh;
      if (sh) v |= word(w + 2) << (64 - sh);
      digits[d] = static_cast<double>(v & ((1ull << 52) - 1ull));
    }
  };
  // 2^e mod n^2 by e doublings of 1, split into base-n digits (a, b) = (v mod n, v div n)
  Words x(n2w.size(), 0);
  x[0] = 1;
  int e = 0;
  const int targets[3] = {52 * K, 52 * K + h->w, 104 * K};      // R, 2^w R, R^2   (w < 52 K)
  double* dstA[3] = {h->oneA, h->sA, h->r2A};
  double* dstB[3] = {h->oneB, h->sB, h->r2B};
  for (int tgt = 0; tgt < 3; tgt++) {
    for (; e < targets[tgt]; e++) words_dbl_mod(x, n2w);
    Words q, r;
    words_divmod(x, rt.words, q, r);
    to52(r, dstA[tgt]);
    to52(q, dstB[tgt]);
  }
  if (!sq.pair52) {
    cudaError_t ce = cudaMalloc(&sq.pair52, sizeof(PairConsts52Dev));
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "cudaMalloc(pair constants)");
  }
  cudaError_t ce = cudaMemcpyAsync(sq.pair52, h, sizeof(PairConsts52Dev), cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "upload pair constants");
  sq.root_slot = root_slot;
  return SCB_OK;
}

extern "C" int scb_modulus_link_square(scb_ctx* ctx, int square_slot, int root_slot) { return scb::link_square_modulus(ctx, square_slot, root_slot); }

// -------------------------------------------------------------------------------------------------
// launch helpers
// -------------------------------------------------------------------------------------------------
namespace {

int window_bits_for(int ebits) {
  if (ebits <= 6) return 1;
  if (ebits <= 24) return 2;
  if (ebits <= 80) return 3;
  if (ebits <= 240) return 4;
  if (ebits <= 1100) return 5;
  return 6;
}

template <int L, int TPI>
int launch_modexp(scb_ctx* ctx, const scb_ctx::Mod& md, const uint32_t* base, int kbase, const uint32_t* exp, int ewords, size_t estride, int ebits,
                  const uint32_t* mul_in, uint32_t* out, int kout, size_t B) {
  using G = Geo<L, TPI>;
  int wbits = window_bits_for(ebits);
  int nprog = 0, pw = 0;
  const uint16_t* prog = estride == 0 ? ctx->find_program(exp, &nprog, &pw, &wbits) : nullptr;
  const size_t smem = sizeof(uint32_t) * G::WARP_SMEM_WORDS * WARPS_PER_BLOCK;
  const int gpb = GroupIndex<L, TPI>::GROUPS_PER_BLOCK;
  int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_modexp<L, TPI>), smem, B, gpb);
  void* scratch = nullptr;
  int rc = ctx->scratch(0, static_cast<size_t>(grid) * gpb * modexp_scratch_words<L, TPI>(wbits) * sizeof(uint32_t), &scratch);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k_modexp<L, TPI><<<grid, BLOCK_THREADS, smem, ctx->stream>>>(md.dev, base, kbase, exp, ewords, estride, ebits, wbits,
                                                               static_cast<uint32_t*>(scratch), mul_in, out, kout, B, queue, prog, nprog, pw);
  char label[64];
  snprintf(label, sizeof(label), "k_modexp<%d,%d>[e=%d]", L, TPI, ebits);   // profile label: geometry and exponent length
  return ctx->launched(label);
}

template <int L, int TPI>
int launch_modexp52(scb_ctx* ctx, const scb_ctx::Mod& md, const uint32_t* base, int kbase, const uint32_t* exp, int ewords, size_t estride, int ebits,
                    const uint32_t* mul_in, uint32_t* out, int kout, size_t B) {
  using G = Geo52<L, TPI>;
  int wbits = window_bits_for(ebits);
  // shared exponent with a registered sliding-window program?
  int nprog = 0, pw = 0;
  const uint16_t* prog = estride == 0 ? ctx->find_program(exp, &nprog, &pw, &wbits) : nullptr;
  size_t smem = sizeof(double) * G::WARP_SMEM_DOUBLES * WARPS_PER_BLOCK;
  // tuning aid: SCB_F64_EXTRA_SMEM=<bytes> pads the dynamic shared memory to lower the number of resident blocks
  static const size_t extra_smem = getenv("SCB_F64_EXTRA_SMEM") ? static_cast<size_t>(atol(getenv("SCB_F64_EXTRA_SMEM"))) : 0;
  if (extra_smem) {
    smem += extra_smem;
    cudaFuncSetAttribute(reinterpret_cast<const void*>(&k_modexp52<L, TPI>), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  }
  const int gpb = GroupIndex52<L, TPI>::GROUPS_PER_BLOCK;
  int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_modexp52<L, TPI>), smem, B, gpb);
  void* scratch = nullptr;
  int rc = ctx->scratch(0, static_cast<size_t>(grid) * gpb * modexp52_scratch_doubles<L, TPI>(wbits) * sizeof(double), &scratch);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k_modexp52<L, TPI><<<grid, BLOCK_THREADS, smem, ctx->stream>>>(md.dev52, base, kbase, exp, ewords, estride, ebits, wbits,
                                                                 static_cast<double*>(scratch), mul_in, out, kout, B, queue, prog, nprog, pw);
  char label[64];
  snprintf(label, sizeof(label), "k_modexp52<%d,%d>[e=%d]", L, TPI, ebits);
  return ctx->launched(label);
}

// fixed-base tables in the FP64 engine's digits (kernels_fb52.cuh)
template <int L, int TPI>
int launch_fb52_build(scb_ctx* ctx, const scb_ctx::Mod& md, const uint32_t* base_dev, double* table, int nwin) {
  using G = Geo52<L, TPI>;
  k52_fb_bases<L, TPI><<<1, 32, sizeof(double) * G::WARP_SMEM_DOUBLES, ctx->stream>>>(md.dev52, base_dev, md.k32, table, nwin);
  int rc = ctx->launched("k52_fb_bases");
  if (rc) return rc;
  const size_t smem = sizeof(double) * G::WARP_SMEM_DOUBLES * WARPS_PER_BLOCK;
  const int gpb = GroupIndex52<L, TPI>::GROUPS_PER_BLOCK;
  k52_fb_fill<L, TPI><<<(nwin + gpb - 1) / gpb, BLOCK_THREADS, smem, ctx->stream>>>(md.dev52, table, nwin, table + static_cast<size_t>(nwin) * 256 * G::K);
  return ctx->launched("k52_fb_fill");
}

template <int L, int TPI>
int launch_fb52_16_build(scb_ctx* ctx, const scb_ctx::Mod& md, const double* table8, double* table16, int nwin16) {
  using G = Geo52<L, TPI>;
  const size_t smem = sizeof(double) * G::WARP_SMEM_DOUBLES * WARPS_PER_BLOCK;
  const int gpb = GroupIndex52<L, TPI>::GROUPS_PER_BLOCK;
  const size_t M = static_cast<size_t>(nwin16) << 16;
  int grid = ctx->grid_for(reinterpret_cast<const void*>(&k52_fb16_fill<L, TPI>), smem, M, gpb);
  void* dump = nullptr;
  int rc = ctx->scratch(0, static_cast<size_t>(grid) * gpb * G::K * sizeof(double), &dump);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k52_fb16_fill<L, TPI><<<grid, BLOCK_THREADS, smem, ctx->stream>>>(md.dev52, table8, table16, M, static_cast<double*>(dump), queue);
  return ctx->launched("k52_fb16_fill");
}

// exponentiation modulo the square of `rt` (pair form, kernels_f64_pair.cuh)
template <int L, int TPI>
int launch_modexp52p(scb_ctx* ctx, const scb_ctx::Mod& sq, const scb_ctx::Mod& rt, const uint32_t* base, int kbase, const uint32_t* exp, int ewords,
                     size_t estride, int ebits, uint32_t* out, int kout, size_t B) {
  using GP = GeoPair<L, TPI>;
  int wbits = window_bits_for(ebits);
  int nprog = 0, pw = 0;
  const uint16_t* prog = estride == 0 ? ctx->find_program(exp, &nprog, &pw, &wbits) : nullptr;
  const size_t smem = sizeof(double) * GP::WARP_SMEM_WORDS64 * WARPS_PER_BLOCK;
  const void* kfn = reinterpret_cast<const void*>(&k_modexp52p<L, TPI>);
  if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const int gpb = GroupIndex52<L, TPI>::GROUPS_PER_BLOCK;
  int grid = ctx->grid_for(kfn, smem, B, gpb);
  void* scratch = nullptr;
  int rc = ctx->scratch(0, static_cast<size_t>(grid) * gpb * modexp52p_scratch_doubles<L, TPI>(wbits) * sizeof(double), &scratch);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k_modexp52p<L, TPI><<<grid, BLOCK_THREADS, smem, ctx->stream>>>(rt.dev52, sq.pair52, base, kbase, exp, ewords, estride, ebits, wbits,
                                                                  static_cast<double*>(scratch), out, kout, B, queue, prog, nprog, pw);
  char label[64];
  snprintf(label, sizeof(label), "k_modexp52p<%d,%d>[e=%d]", L, TPI, ebits);
  return ctx->launched(label);
}

template <int L, int TPI>
int launch_modmul(scb_ctx* ctx, const scb_ctx::Mod& md, const uint32_t* a, size_t as, const uint32_t* b, size_t bs, uint32_t* out, size_t B) {
  using G = Geo<L, TPI>;
  const size_t smem = sizeof(uint32_t) * G::WARP_SMEM_WORDS * WARPS_PER_BLOCK;
  const int gpb = GroupIndex<L, TPI>::GROUPS_PER_BLOCK;
  int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_modmul<L, TPI>), smem, B, gpb);
  void* scratch = nullptr;
  int rc = ctx->scratch(0, static_cast<size_t>(grid) * gpb * 2 * G::K * sizeof(uint32_t), &scratch);
  if (rc) return rc;
  k_modmul<L, TPI><<<grid, BLOCK_THREADS, smem, ctx->stream>>>(md.dev, a, as, b, bs, md.k32, static_cast<uint32_t*>(scratch), out, B);
  return ctx->launched("k_modmul");
}

template <int L, int TPI>
int launch_fb_build(scb_ctx* ctx, const scb_ctx::Mod& md, const uint32_t* base_dev, uint32_t* table, int nwin) {
  using G = Geo<L, TPI>;
  k_fb_bases<L, TPI><<<1, 32, sizeof(uint32_t) * G::WARP_SMEM_WORDS, ctx->stream>>>(md.dev, base_dev, md.k32, table, nwin);
  int rc = ctx->launched("k_fb_bases");
  if (rc) return rc;
  const size_t smem = sizeof(uint32_t) * G::WARP_SMEM_WORDS * WARPS_PER_BLOCK;
  const int gpb = GroupIndex<L, TPI>::GROUPS_PER_BLOCK;
  k_fb_fill<L, TPI><<<(nwin + gpb - 1) / gpb, BLOCK_THREADS, smem, ctx->stream>>>(md.dev, table, nwin,
                                                                                table + static_cast<size_t>(nwin) * 256 * G::K);
  return ctx->launched("k_fb_fill");
}

template <int L, int TPI>
int launch_fb16_build(scb_ctx* ctx, const scb_ctx::Mod& md, const uint32_t* table8, uint32_t* table16, int nwin16) {
  using G = Geo<L, TPI>;
  const size_t smem = sizeof(uint32_t) * G::WARP_SMEM_WORDS * WARPS_PER_BLOCK;
  const int gpb = GroupIndex<L, TPI>::GROUPS_PER_BLOCK;
  const size_t M = static_cast<size_t>(nwin16) << 16;
  int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_fb16_fill<L, TPI>), smem, M, gpb);
  void* dump = nullptr;
  int rc = ctx->scratch(0, static_cast<size_t>(grid) * gpb * G::K * sizeof(uint32_t), &dump);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k_fb16_fill<L, TPI><<<grid, BLOCK_THREADS, smem, ctx->stream>>>(md.dev, table8, table16, M, static_cast<uint32_t*>(dump), queue);
  return ctx->launched("k_fb16_fill");
}

template <int L, int TPI>
int launch_fixed_base(scb_ctx* ctx, const scb_ctx::Mod& md, const scb_ctx::FixedBase& fb, const uint32_t* exp, int ewords,
                      const uint32_t* mul_in, uint32_t* out, size_t B) {
  using G = Geo<L, TPI>;
  const size_t smem = sizeof(uint32_t) * G::WARP_SMEM_WORDS * WARPS_PER_BLOCK;
  const int gpb = GroupIndex<L, TPI>::GROUPS_PER_BLOCK;
  int grid = ctx->grid_for(reinterpret_cast<const void*>(&k_fixed_base<L, TPI>), smem, B, gpb);
  void* scratch = nullptr;
  int rc = ctx->scratch(0, static_cast<size_t>(grid) * gpb * G::K * sizeof(uint32_t), &scratch);
  if (rc) return rc;
  unsigned long long* queue = ctx->fresh_queue();
  if (!queue) return fail(ctx, SCB_ECUDA, "cannot set up the work queue");
  k_fixed_base<L, TPI><<<grid, BLOCK_THREADS, smem, ctx->stream>>>(md.dev, fb.use_table(), fb.use_nwin(), fb.use_wbits(), exp, ewords, mul_in,
                                                                   static_cast<uint32_t*>(scratch), out, md.k32, B, queue);
  return ctx->launched("k_fixed_base");
}

}  // namespace

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

#define SCB_DISPATCH_K52(K52, CALL)                                              \
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

int scb_ctx::grid_for(const void* kernel, size_t smem, size_t B, int groups_per_block) {
  auto it = occupancy_cache.find(kernel);
  int per_sm;
  if (it == occupancy_cache.end()) {
    per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BLOCK_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    occupancy_cache[kernel] = per_sm;
  } else {
    per_sm = it->second;
  }
  const size_t need = (B + groups_per_block - 1) / groups_per_block;
  const size_t cap = static_cast<size_t>(per_sm) * num_sms;
  return static_cast<int>(std::max<size_t>(1, std::min(need, cap)));
}

int scb_ctx::launched(const char* what) {
  launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(this, e, what);
  if (profiling) {
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) == cudaSuccess) {
      cudaEventRecord(ev, stream);
      prof_events.push_back({what, ev});
    }
  }
  return SCB_OK;
}

extern "C" int scb_profile_begin(scb_ctx* ctx) {
  if (!ctx) return SCB_EINVAL;
  cudaSetDevice(ctx->device);
  for (auto& pe : ctx->prof_events) cudaEventDestroy(pe.second);
  ctx->prof_events.clear();
  ctx->profiling = true;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return SCB_ECUDA;
  cudaEventRecord(ev, ctx->stream);
  ctx->prof_events.push_back({"<begin>", ev});
  return SCB_OK;
}

extern "C" int scb_profile_end(scb_ctx* ctx, char* out, size_t out_bytes) {
  if (!ctx || !out || out_bytes < 2) return SCB_EINVAL;
  ctx->profiling = false;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaStreamSynchronize");
  std::map<std::string, std::pair<double, int>> acc;   // name -> (ms, launches)
  for (size_t i = 1; i < ctx->prof_events.size(); i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->prof_events[i - 1].second, ctx->prof_events[i].second);
    auto& a = acc[ctx->prof_events[i].first];
    a.first += ms;
    a.second += 1;
  }
  for (auto& pe : ctx->prof_events) cudaEventDestroy(pe.second);
  ctx->prof_events.clear();
  std::string js = "{";
  bool first = true;
  for (auto& kv : acc) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"ms\": %.4f, \"launches\": %d}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
    js += buf;
    first = false;
  }
  js += "}";
  if (js.size() + 1 > out_bytes) return fail(ctx, SCB_EINVAL, "scb_profile_end: buffer too small");
  memcpy(out, js.c_str(), js.size() + 1);
  return SCB_OK;
}

// -------------------------------------------------------------------------------------------------
// generic primitives, device-pointer versions
// -------------------------------------------------------------------------------------------------
static int check_mod(scb_ctx* ctx, int slot, const char* who) {
  if (!ctx) return SCB_EINVAL;
  if (slot < 0 || slot >= SCB_MOD_SLOTS) return fail(ctx, SCB_EINVAL, "%s: bad modulus slot %d", who, slot);
  if (!ctx->mod[slot].loaded) return fail(ctx, SCB_ESTATE, "%s: modulus slot %d not loaded", who, slot);
  cudaSetDevice(ctx->device);
  return SCB_OK;
}

int scb::modexp_any(scb_ctx* ctx, int slot, const uint32_t* base, int kbase, const uint32_t* exp, int exp_words, size_t estride, int exp_bits,
                    const uint32_t* mul_in, uint32_t* out, int kout, size_t B) {
  const scb_ctx::Mod& md = ctx->mod[slot];
  if (kout < md.k32) return fail(ctx, SCB_EINVAL, "modexp: output narrower than the modulus");
  // Small batches: when the batch fills less than a quarter of the resident groups of the usual geometry (L = 10 digits per
  // lane) the launch is bound by the latency of ONE exponentiation, and the same number is spread over twice as many lanes
  // (L = 5): measured 1.42x at B = 1,024 and 2,048 (2048-bit), break-even near B = 4,096 (profiles/r02c_small_batch_*.jsonl).
  auto small_batch = [&](int tpi_big) -> bool {
    const size_t resident = static_cast<size_t>(ctx->num_sms) * 3 * WARPS_PER_BLOCK * (32 / tpi_big);
    return ctx->use_small_batch && B * 4 <= resident;
  };
  // square modulus linked to its root: base-root pair form on the FP64 engine (half the digit products)
  if (ctx->use_f64 && ctx->use_pair && md.root_slot >= 0 && md.pair52 && mul_in == nullptr && exp_bits >= 64) {
    const scb_ctx::Mod& rt = ctx->mod[md.root_slot];
#define PAIRCALL(L, TPI) launch_modexp52p<L, TPI>(ctx, md, rt, base, kbase, exp, exp_words, estride, exp_bits, out, kout, B)
    switch (rt.K52) {
      case 20: return small_batch(2) ? PAIRCALL(5, 4) : PAIRCALL(10, 2);
      case 30: return PAIRCALL(15, 2);
      case 40: return small_batch(4) ? PAIRCALL(5, 8) : PAIRCALL(10, 4);
      case 60: return PAIRCALL(15, 4);
      case 80: return small_batch(8) ? PAIRCALL(5, 16) : PAIRCALL(10, 8);
      default: break;
    }
#undef PAIRCALL
  }
  // FP64 engine for real exponentiations (it has no advantage for the few-bit exponents of the DGK blinding)
  if (ctx->use_f64 && md.K52 > 0 && exp_bits >= 64 && 32 * kbase <= 104 * md.K52) {
#define F64CALL(L, TPI) launch_modexp52<L, TPI>(ctx, md, base, kbase, exp, exp_words, estride, exp_bits, mul_in, out, kout, B)
    switch (md.K52) {
      case 20: return small_batch(2) ? F64CALL(5, 4) : F64CALL(10, 2);
      case 30: return F64CALL(15, 2);
      case 40: return small_batch(4) ? F64CALL(5, 8) : F64CALL(10, 4);
      case 60: return F64CALL(15, 4);
      case 120: return F64CALL(15, 8);
      case 80: return small_batch(8) ? F64CALL(5, 16) : F64CALL(10, 8);
      case 160: return F64CALL(10, 16);
      default: break;
    }
#undef F64CALL
  }
  if (32 * kbase > 2 * W * md.K) return fail(ctx, SCB_EINVAL, "modexp: base wider than twice the modulus");
#define CALL(L, TPI) launch_modexp<L, TPI>(ctx, md, base, kbase, exp, exp_words, estride, exp_bits, mul_in, out, kout, B)
  SCB_DISPATCH_K(md.K, CALL)
#undef CALL
}

extern "C" int scb_modexp_dev(scb_ctx* ctx, int slot, const uint32_t* base, const uint32_t* exp, int exp_words, int exp_shared,
                              int exp_bits, const uint32_t* mul_in, uint32_t* out, size_t B) {
  int rc = check_mod(ctx, slot, "scb_modexp");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!base || !exp || !out || exp_words <= 0 || exp_bits < 0 || exp_bits > exp_words * 32) return fail(ctx, SCB_EINVAL, "scb_modexp: bad argument");
  const scb_ctx::Mod& md = ctx->mod[slot];
  const size_t estride = exp_shared ? 0 : static_cast<size_t>(exp_words);
  return modexp_any(ctx, slot, base, md.k32, exp, exp_words, estride, exp_bits, mul_in, out, md.k32, B);
}

// base: [B][kbase] words (kbase up to twice the modulus width: reduced on load); out: [B][kout] words
int scb::modexp_dev_ex(scb_ctx* ctx, int slot, const uint32_t* base, int kbase, const uint32_t* exp, int exp_words, int exp_bits,
                       uint32_t* out, int kout, size_t B) {
  return modexp_any(ctx, slot, base, kbase, exp, exp_words, 0, exp_bits, nullptr, out, kout, B);
}

extern "C" int scb_modmul_dev(scb_ctx* ctx, int slot, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t B) {
  int rc = check_mod(ctx, slot, "scb_modmul");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!a || !b || !out) return fail(ctx, SCB_EINVAL, "scb_modmul: null pointer");
  const scb_ctx::Mod& md = ctx->mod[slot];
#define CALL(L, TPI) launch_modmul<L, TPI>(ctx, md, a, md.k32, b, md.k32, out, B)
  SCB_DISPATCH_K(md.K, CALL)
#undef CALL
}

int scb::modmul_dev_strided(scb_ctx* ctx, int slot, const uint32_t* a, size_t a_stride, const uint32_t* b, size_t b_stride, uint32_t* out, size_t B) {
  const scb_ctx::Mod& md = ctx->mod[slot];
#define CALL(L, TPI) launch_modmul<L, TPI>(ctx, md, a, a_stride, b, b_stride, out, B)
  SCB_DISPATCH_K(md.K, CALL)
#undef CALL
}

int scb::modinv_dev_raw(scb_ctx* ctx, int slot, const uint32_t* a, uint32_t* out, uint8_t* ok, size_t B) {
  const scb_ctx::Mod& md = ctx->mod[slot];
  k_modinv<<<static_cast<unsigned>((B + 63) / 64), 64, 0, ctx->stream>>>(md.dev, md.k32, a, out, ok, B);
  return ctx->launched("k_modinv");
}

extern "C" int scb_modinv_dev(scb_ctx* ctx, int slot, const uint32_t* a, uint32_t* out, uint8_t* ok, size_t B) {
  int rc = check_mod(ctx, slot, "scb_modinv");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!a || !out || !ok) return fail(ctx, SCB_EINVAL, "scb_modinv: null pointer");
  return modinv_dev_raw(ctx, slot, a, out, ok, B);
}

// test hook: carry resolution of the FP64 engine on crafted column values, in any geometry the engine instantiates
// (digits per lane x lanes per number); cols [G][K], hi [G][lanes], out [G][K], K = digits_per_lane * lanes
extern "C" int scb_debug_resolve52_geo(scb_ctx* ctx, int digits_per_lane, int lanes, const uint64_t* cols, const uint64_t* hi, uint64_t* out, size_t G) {
  if (!ctx || !cols || !hi || !out) return SCB_EINVAL;
  if (G == 0) return SCB_OK;
  cudaSetDevice(ctx->device);
  const int K = digits_per_lane * lanes, TPI = lanes;
  void* dev = nullptr;
  const size_t nb_cols = sizeof(uint64_t) * G * K, nb_hi = sizeof(uint64_t) * G * TPI;
  int rc = ctx->scratch(2, 2 * nb_cols + nb_hi + 1024, &dev);
  if (rc) return rc;
  uint64_t* dcols = static_cast<uint64_t*>(dev);
  uint64_t* dout = dcols + G * K;
  uint64_t* dhi = dout + G * K;
  cudaError_t e = cudaMemcpyAsync(dcols, cols, nb_cols, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dhi, hi, nb_hi, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "H2D copy");
#define GEO(L_, T_)                                                                         \
  if (digits_per_lane == L_ && lanes == T_) {                                               \
    k_debug_resolve52<L_, T_><<<1, 32, 0, ctx->stream>>>(dcols, dhi, dout, G);              \
  } else
  GEO(10, 2) GEO(10, 4) GEO(10, 8) GEO(10, 16) GEO(5, 4) GEO(5, 8) GEO(5, 16) GEO(15, 2) GEO(15, 4) GEO(15, 8)
  return fail(ctx, SCB_EINVAL, "scb_debug_resolve52_geo: geometry %d x %d is not one the engine uses", digits_per_lane, lanes);
#undef GEO
  if ((rc = ctx->launched("k_debug_resolve52"))) return rc;
  e = cudaMemcpyAsync(out, dout, nb_cols, cudaMemcpyDeviceToHost, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "D2H copy");
  return scb_sync(ctx);
}

// the same for K = 40: 10 digits x 4 lanes
extern "C" int scb_debug_resolve52(scb_ctx* ctx, const uint64_t* cols, const uint64_t* hi, uint64_t* out, size_t G) {
  return scb_debug_resolve52_geo(ctx, 10, 4, cols, hi, out, G);
}

extern "C" int scb_fixed_base_load(scb_ctx* ctx, int tslot, int slot, const uint32_t* base, int exp_bits) {
  int rc = check_mod(ctx, slot, "scb_fixed_base_load");
  if (rc) return rc;
  if (tslot < 0 || tslot >= scb_ctx::NFB || !base || exp_bits <= 0 || exp_bits > 8192) return fail(ctx, SCB_EINVAL, "scb_fixed_base_load: bad argument");
  const scb_ctx::Mod& md = ctx->mod[slot];
  scb_ctx::FixedBase& fb = ctx->fb[tslot];
  const int nwin16 = (exp_bits + 15) / 16;
  const int nwin = 2 * nwin16;            // an even number of 8-bit windows, so that they pair up into 16-bit ones
  // + a tail of 128 two-slot private areas used by the surplus groups of the build kernels
  const size_t bytes = (static_cast<size_t>(nwin) * 256 + 256) * md.K * sizeof(uint32_t);
  if (fb.table || fb.table16) cudaStreamSynchronize(ctx->stream);
  if (fb.table) {
    cudaFree(fb.table);
    fb.table = nullptr;
  }
  if (fb.table16) {
    cudaFree(fb.table16);
    fb.table16 = nullptr;
  }
  if (fb.table52) {
    cudaFree(fb.table52);
    fb.table52 = nullptr;
  }
  if (fb.table16_52) {
    cudaFree(fb.table16_52);
    fb.table16_52 = nullptr;
  }
  cudaError_t e = cudaMalloc(&fb.table, bytes);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(fixed-base table)");
  void* bdev = nullptr;
  rc = ctx->scratch(1, sizeof(uint32_t) * md.k32, &bdev);
  if (rc) return rc;
  e = cudaMemcpyAsync(bdev, base, sizeof(uint32_t) * md.k32, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "upload fixed base");
  fb.nwin = nwin;
  fb.slot = slot;
  fb.exp_bits = exp_bits;
  uint32_t* table = fb.table;
  const uint32_t* bd = static_cast<const uint32_t*>(bdev);
#define CALL(L, TPI) launch_fb_build<L, TPI>(ctx, md, bd, table, nwin)
  auto run = [&]() -> int { SCB_DISPATCH_K(md.K, CALL) };
#undef CALL
  rc = run();
  if (rc) return rc;
  // 16-bit windows when the table fits comfortably (it is a capacity trade: HBM is 180 GB); otherwise stay with 8-bit
  fb.nwin16 = nwin16;
  const size_t bytes16 = (static_cast<size_t>(nwin16) << 16) * md.K * sizeof(uint32_t);
  if (ctx->fb_wide && exp_bits >= 32 && bytes16 <= (static_cast<size_t>(8) << 30)) {
    e = cudaMalloc(&fb.table16, bytes16);
    if (e != cudaSuccess) {
      cudaGetLastError();
      fb.table16 = nullptr;
    } else {
      uint32_t* t16 = fb.table16;
#define CALL(L, TPI) launch_fb16_build<L, TPI>(ctx, md, table, t16, nwin16)
      auto run16 = [&]() -> int { SCB_DISPATCH_K(md.K, CALL) };
#undef CALL
      rc = run16();
      if (rc) return rc;
    }
  }
  // the same tables in the FP64 engine's digits, for the DGK phases on that engine (kernels_dgk52.cuh)
  if (ctx->use_f64 && md.K52 > 0 && md.dev52) {
    const size_t bytes52 = (static_cast<size_t>(nwin) * 256 + 256) * md.K52 * sizeof(double);
    e = cudaMalloc(&fb.table52, bytes52);
    if (e != cudaSuccess) {
      cudaGetLastError();
      fb.table52 = nullptr;
    } else {
      double* t52 = fb.table52;
#define CALL(L, TPI) launch_fb52_build<L, TPI>(ctx, md, bd, t52, nwin)
      auto run52 = [&]() -> int { SCB_DISPATCH_K52(md.K52, CALL) };
#undef CALL
      if ((rc = run52())) return rc;
      const size_t bytes16_52 = (static_cast<size_t>(nwin16) << 16) * md.K52 * sizeof(double);
      if (fb.table16 && bytes16_52 <= (static_cast<size_t>(8) << 30)) {
        e = cudaMalloc(&fb.table16_52, bytes16_52);
        if (e != cudaSuccess) {
          cudaGetLastError();
          fb.table16_52 = nullptr;
        } else {
          double* t16 = fb.table16_52;
#define CALL(L, TPI) launch_fb52_16_build<L, TPI>(ctx, md, t52, t16, nwin16)
          auto run52b = [&]() -> int { SCB_DISPATCH_K52(md.K52, CALL) };
#undef CALL
          if ((rc = run52b())) return rc;
        }
      }
    }
  }
  rc = scb_sync(ctx);
  if (rc) return rc;
  fb.loaded = true;
  return SCB_OK;
}

extern "C" int scb_fixed_base_exp_dev(scb_ctx* ctx, int tslot, const uint32_t* exp, int exp_words, const uint32_t* mul_in,
                                      uint32_t* out, size_t B) {
  if (!ctx) return SCB_EINVAL;
  if (tslot < 0 || tslot >= scb_ctx::NFB || !ctx->fb[tslot].loaded) return fail(ctx, SCB_ESTATE, "scb_fixed_base_exp: table %d not loaded", tslot);
  const scb_ctx::FixedBase& fb = ctx->fb[tslot];
  if (B == 0) return SCB_OK;
  if (!exp || !out || exp_words <= 0 || exp_words * 32 > fb.nwin * 8 + 31) return fail(ctx, SCB_EINVAL, "scb_fixed_base_exp: exponent of %d words exceeds the table (%d bits)", exp_words, fb.exp_bits);
  cudaSetDevice(ctx->device);
  const scb_ctx::Mod& md = ctx->mod[fb.slot];
#define CALL(L, TPI) launch_fixed_base<L, TPI>(ctx, md, fb, exp, exp_words, mul_in, out, B)
  SCB_DISPATCH_K(md.K, CALL)
#undef CALL
}

// -------------------------------------------------------------------------------------------------
// host-pointer wrappers: stage through device scratch, run, copy back, synchronise
// -------------------------------------------------------------------------------------------------
namespace {

struct Staging {
  scb_ctx* ctx;
  int rc = SCB_OK;
  size_t offset = 0;
  char* base = nullptr;
  explicit Staging(scb_ctx* c, size_t total_bytes) : ctx(c) {
    void* p = nullptr;
    rc = ctx->scratch(2, total_bytes + 4096, &p);
    base = static_cast<char*>(p);
  }
  template <typename T>
  T* reserve(size_t count) {
    T* p = reinterpret_cast<T*>(base + offset);
    offset += (count * sizeof(T) + 255) & ~static_cast<size_t>(255);
    return p;
  }
  template <typename T>
  T* upload(const T* host, size_t count) {
    if (!host) return nullptr;
    T* p = reserve<T>(count);
    if (rc == SCB_OK) {
      cudaError_t e = cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) rc = cuda_fail(ctx, e, "H2D copy");
    }
    return p;
  }
  template <typename T>
  void download(T* host, const T* dev, size_t count) {
    if (rc == SCB_OK) {
      cudaError_t e = cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
      if (e != cudaSuccess) rc = cuda_fail(ctx, e, "D2H copy");
    }
  }
};
inline size_t pad256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }

}  // namespace

extern "C" int scb_modexp(scb_ctx* ctx, int slot, const uint32_t* base, const uint32_t* exp, int exp_words, int exp_shared,
                          int exp_bits, const uint32_t* mul_in, uint32_t* out, size_t B) {
  int rc = check_mod(ctx, slot, "scb_modexp");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!base || !exp || !out || exp_words <= 0) return fail(ctx, SCB_EINVAL, "scb_modexp: bad argument");
  const size_t k = ctx->mod[slot].k32, ne = exp_shared ? exp_words : B * exp_words;
  Staging st(ctx, 3 * pad256(B * k * 4) + pad256(ne * 4));
  if (st.rc) return st.rc;
  const uint32_t* dbase = st.upload(base, B * k);
  const uint32_t* dexp = st.upload(exp, ne);
  const uint32_t* dmul = st.upload(mul_in, B * k);
  uint32_t* dout = st.reserve<uint32_t>(B * k);
  if (st.rc) return st.rc;
  // a shared exponent is known on the host here: compile it into a sliding-window program for this call
  if (exp_shared && (rc = ctx->register_exponent(dexp, exp, exp_words))) return rc;
  rc = scb_modexp_dev(ctx, slot, dbase, dexp, exp_words, exp_shared, exp_bits, dmul, dout, B);
  if (rc == SCB_OK) {
    st.download(out, dout, B * k);
    rc = st.rc ? st.rc : scb_sync(ctx);
  }
  if (exp_shared) ctx->unregister_exponent(dexp);
  return rc;
}

extern "C" int scb_modmul(scb_ctx* ctx, int slot, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t B) {
  int rc = check_mod(ctx, slot, "scb_modmul");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!a || !b || !out) return fail(ctx, SCB_EINVAL, "scb_modmul: null pointer");
  const size_t k = ctx->mod[slot].k32;
  Staging st(ctx, 3 * pad256(B * k * 4));
  if (st.rc) return st.rc;
  const uint32_t* da = st.upload(a, B * k);
  const uint32_t* db = st.upload(b, B * k);
  uint32_t* dout = st.reserve<uint32_t>(B * k);
  if (st.rc) return st.rc;
  rc = scb_modmul_dev(ctx, slot, da, db, dout, B);
  if (rc) return rc;
  st.download(out, dout, B * k);
  if (st.rc) return st.rc;
  return scb_sync(ctx);
}

extern "C" int scb_modinv(scb_ctx* ctx, int slot, const uint32_t* a, uint32_t* out, uint8_t* ok, size_t B) {
  int rc = check_mod(ctx, slot, "scb_modinv");
  if (rc) return rc;
  if (B == 0) return SCB_OK;
  if (!a || !out || !ok) return fail(ctx, SCB_EINVAL, "scb_modinv: null pointer");
  const size_t k = ctx->mod[slot].k32;
  Staging st(ctx, 2 * pad256(B * k * 4) + pad256(B));
  if (st.rc) return st.rc;
  const uint32_t* da = st.upload(a, B * k);
  uint32_t* dout = st.reserve<uint32_t>(B * k);
  uint8_t* dok = st.reserve<uint8_t>(B);
  if (st.rc) return st.rc;
  rc = scb_modinv_dev(ctx, slot, da, dout, dok, B);
  if (rc) return rc;
  st.download(out, dout, B * k);
  st.download(ok, dok, B);
  if (st.rc) return st.rc;
  rc = scb_sync(ctx);
  if (rc) return rc;
  for (size_t i = 0; i < B; i++)
    if (!ok[i]) return fail(ctx, SCB_ENOTINV, "scb_modinv: element %zu has no inverse", i);
  return SCB_OK;
}

extern "C" int scb_fixed_base_exp(scb_ctx* ctx, int tslot, const uint32_t* exp, int exp_words, const uint32_t* mul_in,
                                  uint32_t* out, size_t B) {
  if (!ctx) return SCB_EINVAL;
  if (tslot < 0 || tslot >= scb_ctx::NFB || !ctx->fb[tslot].loaded) return fail(ctx, SCB_ESTATE, "scb_fixed_base_exp: table %d not loaded", tslot);
  if (B == 0) return SCB_OK;
  if (!exp || !out || exp_words <= 0) return fail(ctx, SCB_EINVAL, "scb_fixed_base_exp: bad argument");
  cudaSetDevice(ctx->device);
  const size_t k = ctx->mod[ctx->fb[tslot].slot].k32;
  Staging st(ctx, 2 * pad256(B * k * 4) + pad256(B * exp_words * 4));
  if (st.rc) return st.rc;
  const uint32_t* dexp = st.upload(exp, B * exp_words);
  const uint32_t* dmul = st.upload(mul_in, B * k);
  uint32_t* dout = st.reserve<uint32_t>(B * k);
  if (st.rc) return st.rc;
  int rc = scb_fixed_base_exp_dev(ctx, tslot, dexp, exp_words, dmul, dout, B);
  if (rc) return rc;
  st.download(out, dout, B * k);
  if (st.rc) return st.rc;
  return scb_sync(ctx);
}
