// host_ctx.hpp -- the opaque scb_ctx behind the C ABI (one per GPU; owns all device memory).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace scb {
struct ModulusDev;
struct Modulus52Dev;
struct PairConsts52Dev;
struct KeysDev;
struct Keys52Dev;
}  // namespace scb

struct scb_ctx {
  static constexpr int NSCRATCH = 11;
  static constexpr int SCRATCH_AUX_GROUP = 10;   // what scratch(0, ..) hands out while the launches go to the second stream
  static constexpr int NFB = 4;

  struct Mod {
    scb::ModulusDev* dev = nullptr;
    scb::Modulus52Dev* dev52 = nullptr;   // constants of the FP64 engine (bn52.cuh); K52 = 0 when the size has no FP64 geometry
    int K = 0, k32 = 0, bits = 0, K52 = 0;
    bool loaded = false;
    std::vector<uint32_t> words;
    std::vector<uint32_t> r52_words;      // R mod n of the FP64 engine (R = 2^(52 K52)) as k32 words: x * r52 mod n = Montgomery form of x
    // this modulus is the SQUARE of the modulus in slot `root_slot` (scb_modulus_link_square): exponentiations run in
    // base-root pair form on the FP64 engine (kernels_f64_pair.cuh); pair52 = their constants
    int root_slot = -1;
    scb::PairConsts52Dev* pair52 = nullptr;
  };
  struct FixedBase {
    uint32_t* table = nullptr;   // [nwin][256][K] Montgomery digits (8-bit windows)
    uint32_t* table16 = nullptr; // [nwin16][65536][K] (16-bit windows; optional, built from `table`)
    double* table52 = nullptr;   // the same two tables in the FP64 engine's digits ([..][K52] doubles), for the DGK phases on that engine
    double* table16_52 = nullptr;
    const double* use_table52() const { return table16_52 ? table16_52 : table52; }
    int nwin = 0, nwin16 = 0, slot = -1, exp_bits = 0;
    // what the exponentiation kernels use
    const uint32_t* use_table() const { return table16 ? table16 : table; }
    int use_nwin() const { return table16 ? nwin16 : nwin; }
    int use_wbits() const { return table16 ? 16 : 8; }
    bool loaded = false;
  };

  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  // Second stream for the q half of the KeyHolder's CRT forms (scb200_protocol.cu, AuxLane).  An exponentiation launch ends with
  // a drain phase -- the last tickets finish on a mostly idle GPU, about the time of one solo exponentiation -- and the p and q
  // chains are independent, so with one chain per stream the other chain's blocks fill the SMs a draining launch gives back.
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool use_overlap = true;            // SCB_OPT_OVERLAP_CRT
  bool on_aux = false;                // launches currently go to aux_stream (`stream` and `aux_stream` are swapped)
  std::string error;
  unsigned long long launches = 0;
  Mod mod[16];
  FixedBase fb[NFB];
  void* scratch_ptr[NSCRATCH] = {nullptr};
  size_t scratch_bytes[NSCRATCH] = {0};
  std::map<const void*, int> occupancy_cache;
  scb::KeysDev* keys_dev = nullptr;   // protocol-level key material (kernels_protocol.cuh)
  scb::Keys52Dev* keys52_dev = nullptr;   // DGK constants in the FP64 engine's Montgomery form (kernels_dgk52.cuh); null = not available
  bool use_dgk52 = true;              // DGK-side phases on the FP64 engine when its tables exist (follows use_f64)
  std::vector<uint8_t> keys_host;     // host mirror of KeysDev
  bool have_paillier_pub = false, have_paillier_priv = false, have_dgk_pub = false, have_dgk_priv = false;
  std::vector<uint32_t> paillier_lambda, paillier_mu, dgk_vp;
  int paillier_lambda_bits = 0, dgk_vp_bits = 0, dgk_rand_bits = 0;
  uint32_t* secret_dev = nullptr;     // lambda words followed by mu words
  uint32_t* vp_dev = nullptr;         // v_p words
  // Paillier CRT form: device exponents [e1p, p, p-1, e1q, q, q-1] (kP words each) and their bit lengths
  bool have_paillier_crt = false, use_crt = true;
  bool use_f64 = true;                // large exponentiations on the FP64 engine (bn52.cuh) instead of the IMAD engine
  bool use_small_batch = true;        // spread one number over twice the lanes when the batch cannot fill the GPU (L = 5 geometries)
  bool use_pair = true;               // exponentiations modulo a linked square n^2 in base-n pair form (half the digit products)
  bool fb_wide = true;                // build 16-bit-window fixed-base tables when memory allows
  uint32_t* crt_exp_dev = nullptr;
  int crt_e1_bits[2] = {0, 0}, crt_P_bits[2] = {0, 0};

  // optional per-kernel timing: one event after every launch; a kernel's time is the gap to the previous event
  bool profiling = false;
  std::vector<std::pair<std::string, cudaEvent_t>> prof_events;

  // device CSPRNG (scb200_rng.cu): ChaCha20 key / nonce from the host, one stream per draw call
  uint32_t rng_key[8] = {0}, rng_nonce[2] = {0, 0}, rng_call = 0;
  bool rng_seeded = false;

  // sliding-window programs of the SHARED exponents the library itself owns (N, the CRT exponents, v_p, lambda), keyed by
  // the device address of the exponent words; consulted by modexp_any for shared-exponent launches on the FP64 engine
  struct ExpProgram {
    uint16_t* dev = nullptr;
    int n = 0, w = 0;
  };
  std::map<const uint32_t*, ExpProgram> exp_programs;
  bool use_sliding = true;
  int register_exponent(const uint32_t* dev_words, const uint32_t* host_words, int nwords);
  // program of a shared exponent (nullptr / 0 when none is registered or sliding windows are switched off); *wbits is
  // raised when the program's table (2^(w-1) entries) needs more scratch than the fixed-window table
  const uint16_t* find_program(const uint32_t* dev_words, int* nprog, int* pw, int* wbits) const;
  void unregister_exponent(const uint32_t* dev_words);
  void clear_exponent_programs();
  std::vector<uint32_t> crt_exp_host;

  // ticket counters of the dynamically scheduled kernels (GroupIndex::next): a ring of device words, one zeroed per launch
  unsigned long long* queue_ring = nullptr;
  int queue_next = 0;
  unsigned long long* fresh_queue();

  int scratch(int which, size_t bytes, void** out);
  int grid_for(const void* kernel, size_t smem, size_t B, int groups_per_block);
  int launched(const char* what);
};

namespace scb {
int fail(scb_ctx* ctx, int code, const char* fmt, ...);
int cuda_fail(scb_ctx* ctx, cudaError_t e, const char* what);
int config_K_for_bits(int bits);
// declare slot `square_slot` (already loaded with n^2) to be the square of slot `root_slot` (already loaded with n)
int link_square_modulus(scb_ctx* ctx, int square_slot, int root_slot);
int modinv_dev_raw(scb_ctx* ctx, int slot, const uint32_t* a, uint32_t* out, uint8_t* ok, size_t B);
// out = (mul_in *) base^exp with either engine (FP64 when enabled and the modulus has an FP64 geometry)
int modexp_any(scb_ctx* ctx, int slot, const uint32_t* base, int kbase, const uint32_t* exp, int exp_words, size_t exp_stride, int exp_bits,
               const uint32_t* mul_in, uint32_t* out, int kout, size_t B);
// shared-exponent modexp with a base of kbase words (up to twice the modulus width) and kout-word results
int modexp_dev_ex(scb_ctx* ctx, int slot, const uint32_t* base, int kbase, const uint32_t* exp, int exp_words, int exp_bits, uint32_t* out, int kout, size_t B);
// out[i] = a[i*a_stride..] * b[i*b_stride..] mod n (stride 0 broadcasts an operand); device pointers
int modmul_dev_strided(scb_ctx* ctx, int slot, const uint32_t* a, size_t a_stride, const uint32_t* b, size_t b_stride, uint32_t* out, size_t B);
}  // namespace scb
