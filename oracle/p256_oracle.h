/* oracle/p256_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of aguinet/ecsimd's P-256 hot path (see p256_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Flat "lane" layout everywhere: a 256-bit value = 8 x u32 (== 4 x u64 on a
 * little-endian host), least-significant word first; Jacobian point = X|Y|Z
 * (24 x u32); affine point = x|y (16 x u32).
 */
#ifndef P256_ORACLE_H
#define P256_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int  orc_abi_version(void);

/* single-element primitives (exposed for unit tests) */
void orc1_mod_add(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]);
void orc1_mod_sub(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]);
void orc1_mod_shl1(uint32_t r[8], const uint32_t a[8]);
void orc1_mul512(uint32_t r[16], const uint32_t a[8], const uint32_t b[8]);
void orc1_square512(uint32_t r[16], const uint32_t a[8]);      /* WITH the lost-carry quirk */
int  orc1_square_quirk_hits(const uint32_t a[8]);               /* number of silent wraps */
void orc1_mgry_reduce(uint32_t r[8], const uint32_t t[16]);

/* batched field ops; nt = host threads */
void orc_mgry_add(uint32_t* o, const uint32_t* a, const uint32_t* b, size_t n, int nt);
void orc_mgry_sub(uint32_t* o, const uint32_t* a, const uint32_t* b, size_t n, int nt);
void orc_mgry_shl1(uint32_t* o, const uint32_t* a, size_t n, int nt);
void orc_mgry_mul(uint32_t* o, const uint32_t* a, const uint32_t* b, size_t n, int nt);
void orc_mgry_sqr(uint32_t* o, const uint32_t* a, size_t n, int nt);
void orc_opposite(uint32_t* o, const uint32_t* a, size_t n, int nt);
void orc_from_classical(uint32_t* o, const uint32_t* a, size_t n, int nt);
void orc_to_classical(uint32_t* o, const uint32_t* a, size_t n, int nt);
void orc_inverse(uint32_t* o, const uint32_t* a, size_t n, int nt);
void orc_mul512(uint32_t* o, const uint32_t* a, const uint32_t* b, size_t n);
void orc_square512(uint32_t* o, const uint32_t* a, size_t n);
void orc_mgry_reduce(uint32_t* o, const uint32_t* t, size_t n);

/* batched point ops */
void orc_dblu(uint32_t* outP, uint32_t* out2, const uint32_t* P, size_t n, int nt);
void orc_zaddu(uint32_t* outP, uint32_t* outR, const uint32_t* P, const uint32_t* O, size_t n, int nt);
void orc_zdau(uint32_t* outQ, uint32_t* outR, const uint32_t* P, const uint32_t* Q, size_t n, int nt);
void orc_add_z2_1(uint32_t* outR, const uint32_t* A, const uint32_t* B, size_t n, int nt);
void orc_trplu(uint32_t* outP, uint32_t* out3, const uint32_t* P, size_t n, int nt);
void orc_scalar_mult(uint32_t* out, const uint32_t* k, const uint32_t* P, size_t n, int nt);
void orc_scalar_mult_wraps(uint32_t* out, uint32_t* wraps, const uint32_t* k, const uint32_t* P, size_t n, int nt);
void orc_from_affine(uint32_t* outJ, const uint32_t* xy, size_t n, int nt);
void orc_to_affine(uint32_t* xy, const uint32_t* J, size_t n, int nt);
/* y = sqrt(x^3-3x+b) per lane; ok[i] = 1 iff lane i is a square (the reference
 * answers per 4-lane pack: a pack is valid iff all 4 of its lanes are) */
void orc_from_x(uint32_t* y, uint8_t* ok, const uint32_t* x, size_t n, int nt);
/* run-time modulus p (8 words, odd, bit 255 set); op: 0 mod_add 1 mod_sub 2 mod_shift_left_one
 * 3 mgry_mul 4 mgry_sqr 5 from_classical 6 to_classical 7 mgry_pow(e) 8 opposite */
void orc_gen_op(int op, uint32_t* o, const uint32_t* a, const uint32_t* b, const uint32_t* e, const uint32_t* p, size_t n);
void orc_constants(uint32_t* out /* 64 x u32: P, R, R^2, (p-1)R, Am, Bm, Gx_m, Gy_m */);

/* instrumentation: counts of field ops executed by the calling thread since
 * the last reset: [0]=mul [1]=sqr [2]=add [3]=sub [4]=shl1 [5]=quirk wraps */
void orc_counters_reset(void);
void orc_counters_get(uint64_t out[6]);

#ifdef __cplusplus
}
#endif
#endif
