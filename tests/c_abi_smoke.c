/* Plain-C use of the C ABI (include/scb200.h): no Python, no C++ -- what a cgo / JNI / FFI binding would do.
 * Computes 3^e mod n for a few 256-bit cases with known answers and a modular inverse; exits 0 on success.
 *   gcc -std=c99 -I include tests/c_abi_smoke.c -L protocols/secure_comparison_b200 -lscb200 -o c_abi_smoke   */
#include <stdio.h>
#include <string.h>
#include "scb200.h"

int main(void) {
  scb_ctx* ctx = NULL;
  if (scb_ctx_create(0, &ctx) != SCB_OK) { fprintf(stderr, "no CUDA device\n"); return 2; }
  /* n = 2^255 - 19 (odd), k = 8 words */
  uint32_t n[8] = {0xffffffedu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x7fffffffu};
  if (scb_modulus_load(ctx, SCB_MOD_USER0, n, 8) != SCB_OK) { fprintf(stderr, "%s\n", scb_last_error(ctx)); return 3; }
  /* batch of 3: base^exp with (2, 255) -> 2^255 mod n = 19 ; (5, 0) -> 1 ; (n-1, 2) -> 1 */
  uint32_t base[3][8] = {{2}, {5}, {0}}, exp[3][8] = {{255}, {0}, {2}}, out[3][8];
  memcpy(base[2], n, sizeof(n));
  base[2][0] -= 1;
  if (scb_modexp(ctx, SCB_MOD_USER0, &base[0][0], &exp[0][0], 8, 0, 256, NULL, &out[0][0], 3) != SCB_OK) {
    fprintf(stderr, "%s\n", scb_last_error(ctx));
    return 4;
  }
  const uint32_t want0[3] = {19, 1, 1};
  for (int i = 0; i < 3; i++) {
    if (out[i][0] != want0[i]) { fprintf(stderr, "modexp case %d: got %u\n", i, out[i][0]); return 5; }
    for (int j = 1; j < 8; j++)
      if (out[i][j] != 0) { fprintf(stderr, "modexp case %d: high words not zero\n", i); return 5; }
  }
  /* inverse of 2 mod n = (n + 1) / 2 = 2^254 - 9; 0 has no inverse (SCB_ENOTINV, ok flag 0) */
  uint32_t a[2][8] = {{2}, {0}}, inv[2][8];
  uint8_t ok[2] = {9, 9};
  int rc = scb_modinv(ctx, SCB_MOD_USER0, &a[0][0], &inv[0][0], ok, 2);
  if (rc != SCB_ENOTINV || ok[0] != 1 || ok[1] != 0) { fprintf(stderr, "modinv rc=%d ok=%d,%d\n", rc, ok[0], ok[1]); return 6; }
  if (inv[0][0] != 0xfffffff7u || inv[0][7] != 0x3fffffffu) { fprintf(stderr, "modinv value\n"); return 7; }
  printf("c_abi_smoke ok (version %d, %llu kernel launches)\n", scb_version(), scb_launch_count(ctx));
  scb_ctx_destroy(ctx);
  return 0;
}
