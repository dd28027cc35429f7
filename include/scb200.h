/* scb200.h -- C ABI of the B200-native batched big-integer engine for the DGK/Veugen secure comparison.
 *
 * The reference (tno.mpc.protocols.secure_comparison 4.4.0) has no FFI: its hot path is reached through
 * Python operator overloading on Paillier/DGK ciphertext objects which end in gmpy2.powmod / gmpy2.invert
 * (GMP mpz_powm / mpz_invert / mpz_mul+mpz_mod).  Every entry point below names the reference interface
 * it replaces (file:line under /root/reference/src/tno/mpc/protocols/secure_comparison/).
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative SCB_E* code and never throws;
 *     scb_last_error(ctx) gives the message of the last failure on that context.
 *   - big integers are little-endian arrays of uint32_t words ("limbs"), least significant word first;
 *     a batch is a C-contiguous [B][k] array.  The caller owns every buffer; outputs are written into
 *     caller-provided buffers; no pointer returned by the library outlives its context.
 *   - "_dev" variants take DEVICE pointers (same layouts) and enqueue on the context's stream without
 *     synchronising (scb_sync waits); the others take HOST pointers and are synchronous.
 *   - one context per GPU; a context is not re-entrant.  Work is independent per batch element, so
 *     multi-GPU use is one context (and one host thread or process) per device on a slice of the batch.
 *   - results are canonical residues in [0, modulus), bit-identical to what GMP / CPython return.
 */
#ifndef SCB200_H
#define SCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scb_ctx scb_ctx;

enum {
  SCB_OK = 0,
  SCB_EINVAL = -1,   /* bad argument (null pointer, even modulus, size out of range, unknown slot) */
  SCB_ECUDA = -2,    /* CUDA runtime error (message in scb_last_error) */
  SCB_ENOMEM = -3,   /* device or host allocation failed */
  SCB_ENOTINV = -4,  /* at least one element had no modular inverse (per-element flags tell which) */
  SCB_ESTATE = -5    /* keys / tables required by the call have not been loaded */
};

/* Modulus slots.  SCB_MOD_USER0.. are free for the generic primitives; the others are filled by the
 * key-loading calls below. */
enum {
  SCB_MOD_PAILLIER_N2 = 0, /* N^2  : Paillier ciphertext space (paillier.public_key.n_squared) */
  SCB_MOD_PAILLIER_N = 1,  /* N    : Paillier plaintext space */
  SCB_MOD_DGK_N = 2,       /* n    : DGK ciphertext space (dgk.public_key.n) */
  SCB_MOD_DGK_P = 3,       /* p    : DGK secret prime (dgk.secret_key.p), is_zero test */
  SCB_MOD_PAILLIER_P2 = 4, /* p^2  : CRT half of KeyHolder-side Paillier work */
  SCB_MOD_PAILLIER_Q2 = 5, /* q^2 */
  SCB_MOD_PAILLIER_P = 6,  /* p */
  SCB_MOD_PAILLIER_Q = 7,  /* q */
  SCB_MOD_USER0 = 8,
  SCB_MOD_SLOTS = 16
};

/* ---- context ------------------------------------------------------------------------------------ */
int scb_ctx_create(int device_ordinal, scb_ctx** out);
void scb_ctx_destroy(scb_ctx* ctx);
const char* scb_last_error(const scb_ctx* ctx);
int scb_sync(scb_ctx* ctx);                       /* wait for everything enqueued by *_dev calls */
void* scb_stream(scb_ctx* ctx);                   /* the context's cudaStream_t (for event timing) */
int scb_version(void);
/* SHA-256 prefix over the library's sources, the header and the compiler flags it was built from (build.py rebuilds when it
 * differs from the tree's). */
const char* scb_build_id(void);
/* number of kernels launched by this context since creation (bench.py reports it as gpu_launches) */
unsigned long long scb_launch_count(const scb_ctx* ctx);
/* Per-kernel device timing.  Between scb_profile_begin and scb_profile_end one CUDA event is recorded on the
 * context's stream after every kernel launch; _end synchronises and writes a JSON object
 * {"kernel": {"ms": total, "launches": n}, ...} (time of a kernel = gap to the previous event). */
int scb_profile_begin(scb_ctx* ctx);
int scb_profile_end(scb_ctx* ctx, char* json_out, size_t json_bytes);
/* Measured issue-rate ceiling of one multiplier pipe on THIS GPU, in thread-level operations per second: a register-resident
 * stream of accumulating IMAD.WIDE.U32 (one 32x32+64 multiply-add each: the integer-multiply roofline of bench.py) or of DFMA
 * (the FP64 engine's multiplier).  Takes a few milliseconds. */
enum { SCB_PEAK_IMAD_WIDE = 0, SCB_PEAK_DFMA = 1 };
int scb_measure_peak(scb_ctx* ctx, int which, double* ops_per_second);
/* pinned host memory for callers that want DMA-able staging buffers */
int scb_host_alloc(size_t bytes, void** out);
int scb_host_free(void* ptr);
/* Device buffers for callers that keep data resident between *_dev calls (the batched Initiator / KeyHolder drivers hold
 * r, delta_A and z on the device across the phases of a session).  Copies are enqueued on the context's stream:
 * scb_sync before reading the destination of a device-to-host copy or reusing the source of a host-to-device copy. */
int scb_dev_alloc(scb_ctx* ctx, size_t bytes, void** out);
int scb_dev_free(scb_ctx* ctx, void* ptr);
int scb_copy_h2d(scb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int scb_copy_d2h(scb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- generic batched primitives -------------------------------------------------------------------
 * Replace tno.mpc.encryption_schemes.utils.pow_mod / mod_inv -> gmpy2.powmod / gmpy2.invert and the
 * mulmod inside every scheme.add.                                      */

/* Load an odd modulus of k32 words into a slot (Montgomery constants are derived on the host). */
int scb_modulus_load(scb_ctx* ctx, int slot, const uint32_t* n, int k32);

/* Declare that the modulus in `square_slot` is the SQUARE of the modulus in `root_slot` (both loaded; checked).  Exponentiations
 * in square_slot then run in base-root pair form (x = a + b n, three half-size Montgomery passes per product, two per
 * squaring) with bit-identical results.  The key-loading calls below link N^2 -> N, p^2 -> p and q^2 -> q themselves:
 * rho^N mod N^2 (initiator.py:109), c^(p-1) mod p^2 and b^p mod p^2 (keyholder.py:126-128, 195).  Loading either slot
 * again drops the link. */
int scb_modulus_link_square(scb_ctx* ctx, int square_slot, int root_slot);

/* out[i] = base[i] ^ exp[i] mod n.   base: [B][k32]; exp: [B][exp_words] or, when exp_shared != 0, a
 * single [exp_words] exponent used for every element; exp_bits = number of significant exponent bits
 * to scan (the same for all elements; higher bits must be zero).  mul_in (nullable, [B][k32]):
 * out[i] = mul_in[i] * base[i]^exp[i] mod n -- the "randomize" fusion c * rho^N mod N^2
 * (initiator.py:109, keyholder.py:126-128).                                                          */
int scb_modexp(scb_ctx* ctx, int slot, const uint32_t* base, const uint32_t* exp, int exp_words, int exp_shared,
               int exp_bits, const uint32_t* mul_in, uint32_t* out, size_t B);
int scb_modexp_dev(scb_ctx* ctx, int slot, const uint32_t* base, const uint32_t* exp, int exp_words, int exp_shared,
                   int exp_bits, const uint32_t* mul_in, uint32_t* out, size_t B);

/* out[i] = a[i] * b[i] mod n  (scheme.add on raw ciphertext values). */
int scb_modmul(scb_ctx* ctx, int slot, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t B);
int scb_modmul_dev(scb_ctx* ctx, int slot, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t B);

/* out[i] = a[i]^-1 mod n  (scheme.neg).  ok[i] = 1 when the inverse exists, else 0 and out[i] = 0
 * (GMP mpz_invert returns 0; gmpy2 / CPython raise ZeroDivisionError / ValueError).  Returns
 * SCB_ENOTINV when any element failed (all other elements are still valid).                          */
int scb_modinv(scb_ctx* ctx, int slot, const uint32_t* a, uint32_t* out, uint8_t* ok, size_t B);
int scb_modinv_dev(scb_ctx* ctx, int slot, const uint32_t* a, uint32_t* out, uint8_t* ok, size_t B);

/* Fixed-base exponentiation with a precomputed window table (8-bit windows, widened to 16-bit ones when the
 * table fits: (exp_bits/16) * 65536 residues): out[i] = mul_in[i] * base^exp[i].
 * scb_fixed_base_load builds the table for `base` (k32 words) and exponents of at most exp_bits bits
 * in table slot `tslot` (0..3) over modulus slot `slot`.  This is DGK's randomizer h^rho
 * (keyholder.py:106-108, initiator.py:153-154).                                                      */
int scb_fixed_base_load(scb_ctx* ctx, int tslot, int slot, const uint32_t* base, int exp_bits);
int scb_fixed_base_exp(scb_ctx* ctx, int tslot, const uint32_t* exp, int exp_words, const uint32_t* mul_in,
                       uint32_t* out, size_t B);
int scb_fixed_base_exp_dev(scb_ctx* ctx, int tslot, const uint32_t* exp, int exp_words, const uint32_t* mul_in,
                           uint32_t* out, size_t B);

/* Test hook (not part of the product surface): the exact carry resolution that ends every product of the FP64 engine,
 * on crafted column values.  cols: [G][40] column sums (< 2^63), hi: [G][4] extra high accumulator per lane, out: [G][40]
 * normalised 52-bit digits.  Its second-order ripple path is unreachable with random operands, so tests drive it here. */
int scb_debug_resolve52(scb_ctx* ctx, const uint64_t* cols, const uint64_t* hi, uint64_t* out, size_t G);
/* the same hook for any geometry the FP64 engine instantiates: digits_per_lane x lanes in {10x2, 10x4, 10x8, 10x16, 5x4, 5x8,
 * 5x16, 15x2, 15x4, 15x8}; cols [G][K], hi [G][lanes], out [G][K] with K = digits_per_lane * lanes */
int scb_debug_resolve52_geo(scb_ctx* ctx, int digits_per_lane, int lanes, const uint64_t* cols, const uint64_t* hi, uint64_t* out, size_t G);

/* Test hook, host only (no GPU): the sliding-window program the library compiles for a shared exponent.  Returns the number
 * of entries written (0 = the exponent is below 2^63 and stays on the fixed-window loop), negative on error.
 * prog[0] = table index of the leading window; prog[i>=1] = (squarings << 8) | table index (0xFF = no product);
 * table[j] = base^(2j+1), j < 2^(window_bits-1). */
int scb_debug_compile_exponent(const uint32_t* exp, int exp_words, uint16_t* prog_out, int prog_capacity, int* window_bits);

/* ---- scheme keys ------------------------------------------------------------------------------------
 * Replace the scheme objects the reference threads through every step (tno.mpc.encryption_schemes
 * Paillier / DGK: public_key.{n, n_squared, g} / secret_key.{lambda_, mu}; public_key.{n, g, h, u, t} /
 * secret_key.{v_p, p}).  Key GENERATION stays on the host (keyholder.py:135-172, one-off).            */

/* Paillier public key: N of kN words (g = N + 1 is implied, reference U1).  Loads slots N^2 and N. */
int scb_paillier_set_public(scb_ctx* ctx, const uint32_t* n, int kN);
/* Paillier secret key: lambda (any word count) and mu = lambda^-1 mod N (kN words). */
int scb_paillier_set_private(scb_ctx* ctx, const uint32_t* lambda, int lambda_words, const uint32_t* mu);
/* Paillier secret key in CRT form: the primes p, q (kN/2 words each, p*q = N).  The KeyHolder's Paillier work
 * then runs over p, p^2, q, q^2 (about 3-4x fewer multiply-adds, bit-identical results): decrypt() through
 * L_p(c^(p-1) mod p^2), randomize() through rho^N = CRT((rho mod p)^(q mod p-1) lifted by ^p mod p^2, ...).
 * Implies scb_paillier_set_private (lambda, mu are not needed).  scb_set_option(SCB_OPT_PAILLIER_CRT, 0)
 * switches back to the public formulas (c^lambda mod N^2, rho^N mod N^2) when those were loaded too.     */
int scb_paillier_set_private_crt(scb_ctx* ctx, const uint32_t* p, const uint32_t* q, int kP);
enum {
  SCB_OPT_PAILLIER_CRT = 1,     /* KeyHolder Paillier work over p, p^2, q, q^2 (default 1 when the primes are loaded) */
  SCB_OPT_FIXED_BASE_WIDE = 2,  /* build 16-bit-window fixed-base tables (default 1; 0 = 8-bit windows only, 128x less memory) */
  SCB_OPT_SLIDING_WINDOWS = 4,  /* sliding-window programs for the shared exponents the library owns (N, CRT exponents, v_p, lambda) on the
                                   FP64 engine (default 1; 0 = fixed windows everywhere) */
  SCB_OPT_MODEXP_FP64 = 3,      /* exponentiations with >= 64-bit exponents AND the DGK-side fused phases (bit encryption, steps 4c-4i) on the
                                   FP64-pipe multiplier (52-bit digits, exact) instead of the IMAD one (default 1; results are bit-identical
                                   either way; the 52-bit fixed-base tables are built when the DGK key is loaded with the option on) */
  SCB_OPT_SMALL_BATCH = 6,      /* FP64 engine: batches that cannot fill half of the GPU's resident groups run with 5 digits per lane on
                                   twice as many lanes per number (default 1; results are bit-identical either way) */
  SCB_OPT_OVERLAP_CRT = 7,      /* independent launch chains inside one call on two streams of the context -- the p and q halves of
                                   the KeyHolder's CRT forms; the Initiator's rho^N next to the batched inversion of step 1 -- so
                                   that one chain fills the SMs the other's draining launch gives back (default 1; off while
                                   scb_profile_begin is active; results are bit-identical either way) */
  SCB_OPT_SQUARE_MODULUS = 5    /* exponentiations modulo a linked square n^2 (N^2, p^2, q^2: scb_modulus_link_square) in base-n pair
                                   form on the FP64 engine: half the digit products of a plain product mod n^2 (default 1; 0 = treat
                                   n^2 as an ordinary modulus; results are bit-identical either way) */
};
int scb_set_option(scb_ctx* ctx, int option, int value);
/* DGK public key: n, g, h of kn words; g_neg = the raw ciphertext the scheme produces for plaintext -1
 * (unsafe_encrypt(-1, apply_encoding=False), initiator.py:460; pow_mod gives g^-1 mod n -- pass NULL to have the
 * library compute exactly that on the device); u_bits = bit
 * length of u (blinding exponents are < u); l = comparison bit length; rand_bits = bit length of the
 * randomizer exponents rho (h^rho table size).                                                        */
int scb_dgk_set_public(scb_ctx* ctx, const uint32_t* n, const uint32_t* g, const uint32_t* h, const uint32_t* g_neg, int kn,
                       int u_bits, int l, int rand_bits);
/* DGK secret key: prime factor p (kp words) and v_p (vp_words words, vp_bits significant bits). */
int scb_dgk_set_private(scb_ctx* ctx, const uint32_t* p, int kp, const uint32_t* v_p, int vp_words, int vp_bits);

/* ---- Paillier scheme operations ---------------------------------------------------------------------*/
/* out[i] = (1 + plain[i]*N) * rho[i]^N mod N^2.  plain, rho: [M][kN]; rho nullable = unsafe_encrypt without
 * randomize().  Replaces Paillier.unsafe_encrypt(m, apply_encoding=False) (+ .randomize()).             */
int scb_paillier_encrypt(scb_ctx* ctx, const uint32_t* plain, const uint32_t* rho, uint32_t* out, size_t M);
int scb_paillier_encrypt_dev(scb_ctx* ctx, const uint32_t* plain, const uint32_t* rho, uint32_t* out, size_t M);
/* out[i] = L(c[i]^lambda mod N^2) * mu mod N  (raw plaintext in [0, N)).  Replaces
 * Paillier.decrypt(c, apply_encoding=False): KeyHolder.step_2 (keyholder.py:195) and the final decrypt. */
int scb_paillier_decrypt(scb_ctx* ctx, const uint32_t* c, uint32_t* out, size_t M);
int scb_paillier_decrypt_dev(scb_ctx* ctx, const uint32_t* c, uint32_t* out, size_t M);
/* flags[i] = (c[i]^{v_p} mod p == 1).  Replaces DGK.is_zero (keyholder.py:249). */
int scb_dgk_is_zero(scb_ctx* ctx, const uint32_t* c, uint8_t* flags, size_t M);
int scb_dgk_is_zero_dev(scb_ctx* ctx, const uint32_t* c, uint8_t* flags, size_t M);

/* ---- device-side CSPRNG for the nonces ---------------------------------------------------------------------
 * Replaces, at batch scale, secrets.randbelow / secrets.choice (initiator.py:223, 250, 420, 512) and the schemes'
 * randomness pools (initiator.py:205-210, keyholder.py:174-179): ChaCha20 in counter mode on the device, keyed from
 * the host, one independent keystream position per (call, element, attempt, block); uniform values below a bound by
 * rejection sampling.  Outputs are deterministic functions of (key, nonce, call order).  The arrays feed the
 * *_dev phases directly, so production runs need no host->device nonce traffic; injected host nonces stay
 * available for parity testing.                                                                              */
int scb_rng_seed(scb_ctx* ctx, const uint32_t key[8], const uint32_t nonce[2]);
/* out[e] = uniform integer in [min_one ? 1 : 0, bound), k words each (bound: HOST array of k words; out: device) */
int scb_rng_below_dev(scb_ctx* ctx, const uint32_t* bound, int k, int min_one, uint32_t* out, size_t B);
/* out[e] = `bits` uniform bits in `words` words (DGK randomizer exponents rho) */
int scb_rng_bits_dev(scb_ctx* ctx, int bits, int words, uint32_t* out, size_t B);
/* out[e] = uniform bit (delta_A = randbelow(2)) */
int scb_rng_coins_dev(scb_ctx* ctx, uint8_t* out, size_t B);
/* out[e][0..n) = uniform permutation of 0..n-1 (the shuffle of step_4i: output position k takes element out[e][k]) */
int scb_rng_permutations_dev(scb_ctx* ctx, int n, uint8_t* out, size_t B);
/* one raw ChaCha20 block for explicit state words 12..15 (known-answer tests against RFC 8439) */
int scb_rng_chacha_block_dev(scb_ctx* ctx, const uint32_t counter_nonce[4], uint32_t* out16);
/* host-buffer convenience: run one of the generators above and copy the result back.
 *   SCB_RNG_BELOW: (bound, k, arg = min_one)   SCB_RNG_BITS: (k = words, arg = bits)   SCB_RNG_COINS
 *   SCB_RNG_PERMUTATIONS: (arg = n)            SCB_RNG_BLOCK: (bound = counter_nonce[4])                     */
enum { SCB_RNG_BELOW = 0, SCB_RNG_BITS = 1, SCB_RNG_COINS = 2, SCB_RNG_PERMUTATIONS = 3, SCB_RNG_BLOCK = 4 };
int scb_rng_fetch(scb_ctx* ctx, int what, const uint32_t* bound, int k, int arg, void* out_host, size_t B);

/* ---- fused protocol phases (one call per party-phase of perform_secure_comparison) -------------------
 * Layouts: Paillier ciphertexts [..][2kN]; Paillier plaintexts / r / Paillier nonces rho [..][kN]; DGK
 * ciphertexts [..][kn]; per-comparison DGK vectors are [B][l+1][kn] with element 0 = [d] and elements
 * 1..l = [beta_0]..[beta_{l-1}] (phase B1 -> A2) or [c_0]..[c_l] in the order they are sent (A2 -> B2).
 * Every nonce array is nullable: NULL skips the corresponding randomize() (the step-by-step flow of
 * README.md:117-141).                                                                                  */

/* A1: Initiator.step_1 (initiator.py:228-258) + z_enc.randomize() (:109).
 *     z_enc = y * x^-1 * (1 + (2^l + r) N) * rho^N mod N^2.  r: [B][kN] (the caller draws it: randbelow(N)). */
int scb_initiator_step1(scb_ctx* ctx, const uint32_t* x, const uint32_t* y, const uint32_t* r, const uint32_t* rho, int l,
                        uint32_t* z_enc, size_t B);
int scb_initiator_step1_dev(scb_ctx* ctx, const uint32_t* x, const uint32_t* y, const uint32_t* r, const uint32_t* rho, int l,
                            uint32_t* z_enc, size_t B);
/* B1: KeyHolder.step_4a + step_4b (keyholder.py:198-233) + randomize (:106-108) from the plaintext z of
 *     step_2 (= scb_paillier_decrypt).  rho: [B][l+1][rho_words] (d first).  db_enc: [B][l+1][kn].       */
int scb_keyholder_step4ab(scb_ctx* ctx, const uint32_t* z, int l, const uint32_t* rho, int rho_words, uint32_t* db_enc, size_t B);
int scb_keyholder_step4ab_dev(scb_ctx* ctx, const uint32_t* z, int l, const uint32_t* rho, int rho_words, uint32_t* db_enc, size_t B);
/* A2: Initiator.step_3, 4c, 4d, 4e, 4f, 4h, 4i (initiator.py:260-516) + randomize (:153-154).
 *     delta_a: [B] (0/1; s = 1 - 2 delta_a); r_blind: [B][l+1][rb_words] exponents in [1, u-1] in list order
 *     (NULL = skip the blinding exponentiation, i.e. stop after step_4h); perm: [B][l+1], output position k
 *     takes the masked value perm[k] (NULL = identity = do_shuffle False); rho: [B][l+1][rho_words] applied in
 *     OUTPUT order.  c_out: [B][l+1][kn].                                                              */
int scb_initiator_step4(scb_ctx* ctx, const uint32_t* db_enc, const uint32_t* r, const uint8_t* delta_a, const uint32_t* r_blind,
                        int rb_words, const uint8_t* perm, const uint32_t* rho, int rho_words, int l, uint32_t* c_out, size_t B);
int scb_initiator_step4_dev(scb_ctx* ctx, const uint32_t* db_enc, const uint32_t* r, const uint8_t* delta_a, const uint32_t* r_blind,
                            int rb_words, const uint8_t* perm, const uint32_t* rho, int rho_words, int l, uint32_t* c_out, size_t B);
/* B2: KeyHolder.step_4j + step_5 (keyholder.py:235-287) + randomize (:126-128).  rho: [B][3][kN] for
 *     (zeta_1, zeta_2, delta_B).  delta_b: [B]; enc3: [B][3][2kN] = ([[zeta_1]], [[zeta_2]], [[delta_B]]).     */
int scb_keyholder_step4j_5(scb_ctx* ctx, const uint32_t* c, const uint32_t* z, const uint32_t* rho, int l, uint8_t* delta_b,
                           uint32_t* enc3, size_t B);
int scb_keyholder_step4j_5_dev(scb_ctx* ctx, const uint32_t* c, const uint32_t* z, const uint32_t* rho, int l, uint8_t* delta_b,
                               uint32_t* enc3, size_t B);
/*     KeyHolder.step_5 alone (delta_b given). */
int scb_keyholder_step5(scb_ctx* ctx, const uint32_t* z, const uint8_t* delta_b, const uint32_t* rho, int l, uint32_t* enc3, size_t B);
int scb_keyholder_step5_dev(scb_ctx* ctx, const uint32_t* z, const uint8_t* delta_b, const uint32_t* rho, int l, uint32_t* enc3, size_t B);
/* A3: Initiator.step_6 + step_7 (initiator.py:518-564).  result: [B][2kN] = [[x <= y]] (not randomised). */
int scb_initiator_step6_7(scb_ctx* ctx, const uint32_t* enc3, const uint32_t* r, const uint8_t* delta_a, int l, uint32_t* result,
                          size_t B);
int scb_initiator_step6_7_dev(scb_ctx* ctx, const uint32_t* enc3, const uint32_t* r, const uint8_t* delta_a, int l, uint32_t* result,
                              size_t B);

#ifdef __cplusplus
}
#endif
#endif /* SCB200_H */
