/* ecb200.h -- C ABI of the B200-native batched NIST P-256 engine.
 *
 * This is the drop-in boundary for ONE hot path of aguinet/ecsimd: Montgomery
 * field ops mod p256 -> co-Z Jacobian point ops -> scalar_mult_p256.  Every
 * entry point names the reference interface it replaces (paths relative to the
 * reference checkout).  The reference is a header-only C++20 template library
 * with a single compiled symbol (lib/scalar_mult_p256.cpp:12-14) operating on
 * one 4-lane AVX2 pack per call; this ABI is the batched form of the same
 * functions: n independent lanes per call, plain pointers and sizes.
 *
 * All functions return ECB200_OK (0) or a negative error code and never throw;
 * ecb200_last_error() gives the message for the calling thread.  Like the
 * reference, no input validation is performed on values: out-of-contract
 * inputs (non-canonical field elements, Z != R, ...) give the same
 * deterministic garbage as the reference does.
 *
 * DATA LAYOUTS (flags & ECB200_LAYOUT_MASK); a "value" is a 256-bit integer:
 *   ECB200_LAYOUT_LANE   value i = 8 x u32 (= 4 x u64, little-endian host),
 *                        least-significant word first, at byte offset 32*i.
 *                        Points: X|Y|Z (96 B) or x|y (64 B) per lane.
 *   ECB200_LAYOUT_PACK4  the reference's in-memory pack layout
 *                        (include/ecsimd/bignum.h:101-102, eve::wide<struct>):
 *                        lanes are grouped by 4; inside a 128-byte pack the
 *                        u64 word index is limb*4 + lane.  A Jacobian pack
 *                        (include/ecsimd/jacobian_curve_point.h:64-67) is
 *                        X-pack | Y-pack | Z-pack = 384 B; an affine pack
 *                        (curve_point.h:40-42) is x-pack | y-pack = 256 B.
 *                        n must be a multiple of 4.  An array of
 *                        wide_bignum<bignum_256> / wide_jacobian_curve_point
 *                        objects can be passed as is.
 *   ECB200_LAYOUT_SOA    device-native planar layout: a buffer of n values is
 *                        two planes of n x uint4 (16 B): plane 0 = words 0..3,
 *                        plane 1 = words 4..7.  A buffer of n points with C
 *                        coordinates is 2*C planes (X lo, X hi, Y lo, ...).
 *                        Every warp-wide access is a fully coalesced 128-bit
 *                        load/store.
 * MEMORY SPACE (flags & ECB200_MEM_MASK):
 *   ECB200_MEM_HOST      pointers are host memory; the call stages the data
 *                        through device buffers owned by the library and
 *                        returns when the outputs are written (synchronous).
 *   ECB200_MEM_DEVICE    pointers are device memory on the current CUDA device;
 *                        kernels are enqueued on `stream` (a cudaStream_t, may
 *                        be NULL for the default stream) and the call returns
 *                        without synchronising.
 * SIDE CHANNELS: this library reproduces the reference's VALUES, not its constant-time property.  The
 * variable-base ladder runs the same field operations for every scalar and addresses memory independently
 * of it, but the squaring-defect filter and the 2^-32 corner cases of the conditional subtractions are
 * data-dependent branches, and the fixed-base entry point indexes a table with scalar bits 1..16 unless
 * ECB200_NO_BASE_TABLE is given.
 * ECB200_NO_QUIRK        compute mathematically exact squares instead of
 *                        reproducing the reference's lost-carry squaring defect
 *                        (include/ecsimd/mul.h:192-206).  Default (flag clear)
 *                        is bit-exact parity with the reference.
 */
#ifndef ECB200_H
#define ECB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define ECB200_ABI_VERSION 1

#define ECB200_OK 0
#define ECB200_ERR_ARG (-1)
#define ECB200_ERR_CUDA (-2)
#define ECB200_ERR_NOMEM (-3)

#define ECB200_LAYOUT_LANE 0u
#define ECB200_LAYOUT_PACK4 1u
#define ECB200_LAYOUT_SOA 2u
#define ECB200_LAYOUT_MASK 0xfu
#define ECB200_MEM_HOST 0u
#define ECB200_MEM_DEVICE 0x10u
#define ECB200_MEM_MASK 0x10u
#define ECB200_NO_QUIRK 0x100u
/* ecb200_scalar_mult_p256_base only: run the plain ladder on G instead of starting from the
 * fixed-base table of ladder states (same results; for A/B measurements and tests). */
#define ECB200_NO_BASE_TABLE 0x200u

/* ---- library ------------------------------------------------------------- */
int ecb200_abi_version(void);
/* Select the CUDA device used by the calling thread for subsequent calls, check that it is an
 * sm_100 part, create the library's memory pool on it and build the fixed-base tables of
 * ecb200_scalar_mult_p256_base (20 MiB, ~1 ms).  A device used without this call is set up by the first call
 * that needs it (the first fixed-base call then synchronises its stream once). */
int ecb200_init(int device);
/* Initialise several devices for single-process use -- the reference's caller is one process
 * (benchs/curve_group.cpp:23-60).  Afterwards a HOST-memory batch (ECB200_LAYOUT_LANE or _PACK4) given to
 * ecb200_scalar_mult_p256, ecb200_scalar_mult_p256_base or ecb200_scalar_mult_p256_affine is cut into one
 * contiguous index range per device (lanes are independent: include/ecsimd/bignum.h:101-102), each range
 * running on its own host thread through that device's three-stream pipeline; results are bit-identical to
 * the single-device call.  ECB200_MEM_DEVICE calls keep using the calling thread's current device, which
 * becomes devices[0].  count <= 1 restores single-device behaviour. */
int ecb200_init_devices(const int* devices, int count);
/* Number of devices host-memory batches are currently cut over (>= 1). */
int ecb200_device_count(void);
/* Release what the library owns on the current device: its memory pool, the fixed-base tables, the pipeline
 * streams and pinned bounce buffers (all re-created by the next call that needs them). */
int ecb200_shutdown(void);
const char* ecb200_last_error(void);
/* Number of kernels launched by this library in this process so far. */
uint64_t ecb200_launch_count(void);

/* ---- Montgomery field ops (include/ecsimd/mgry_ops.h) ----------------------- */
/* out = mgry_add(a,b)            mgry_ops.h:10-13 (modular.h:10-15) */
int ecb200_mgry_add(void* out, const void* a, const void* b, size_t n, uint32_t flags, void* stream);
/* out = mgry_sub(a,b)            mgry_ops.h:26-29 (modular.h:24-41) */
int ecb200_mgry_sub(void* out, const void* a, const void* b, size_t n, uint32_t flags, void* stream);
/* out = mgry_mul(a,b)            mgry_ops.h:31-35 (mul.h:150-158, mgry_mul.h:84-121) */
int ecb200_mgry_mul(void* out, const void* a, const void* b, size_t n, uint32_t flags, void* stream);
/* out = mgry_sqr(a)              mgry_ops.h:37-42 (mul.h:160-221) */
int ecb200_mgry_sqr(void* out, const void* a, size_t n, uint32_t flags, void* stream);
/* out = mgry_shift_left<count>(a), count in 1..8   mgry_ops.h:15-24 */
int ecb200_mgry_shift_left(void* out, const void* a, int count, size_t n, uint32_t flags, void* stream);
/* out = GFp::opposite(a)         gfp.h:60-64 */
int ecb200_gfp_opposite(void* out, const void* a, size_t n, uint32_t flags, void* stream);
/* out = wide_mgry_bignum::from_classical(a) / to_classical()   mgry.h:47-55 */
int ecb200_from_classical(void* out, const void* a, size_t n, uint32_t flags, void* stream);
int ecb200_to_classical(void* out, const void* a, size_t n, uint32_t flags, void* stream);
/* out = GFp::inverse(a) = a^(p-2)   gfp.h:42-44 (mgry_ops.h:44-86) */
int ecb200_gfp_inverse(void* out, const void* a, size_t n, uint32_t flags, void* stream);
/* Repeated multiply, register resident: out = a * b^iters (iters Montgomery
 * multiplications per lane with one load and one store) -- the kernel the IMAD
 * roofline of the multiplier is measured with. */
int ecb200_mgry_mul_chain(void* out, const void* a, const void* b, int iters, size_t n, uint32_t flags, void* stream);

/* ---- co-Z Jacobian point ops (include/ecsimd/curve_group.h) ---------------- */
/* Points are Jacobian, Montgomery form, 3 coordinates (X,Y,Z). */
/* out2 = DBLU(P) (= 2P), outP = P rewritten              curve_group.h:64-87 */
int ecb200_dblu(void* outP, void* out2, const void* P, size_t n, uint32_t flags, void* stream);
/* outR = ZADDU(P,O) (= P+O), outP = P rewritten          curve_group.h:91-116 */
int ecb200_zaddu(void* outP, void* outR, const void* P, const void* O, size_t n, uint32_t flags, void* stream);
/* outR = ZDAU(P,Q) (= 2P+Q), outQ = Q rewritten          curve_group.h:120-153 */
int ecb200_zdau(void* outQ, void* outR, const void* P, const void* Q, size_t n, uint32_t flags, void* stream);
/* outR = ADD_Z2_1(A,B), Z(B) == R                         curve_group.h:155-179 */
int ecb200_add_z2_1(void* outR, const void* A, const void* B, size_t n, uint32_t flags, void* stream);
/* out3 = TRPLU(P) (= 3P), outP = P rewritten             curve_group.h:183-186 */
int ecb200_trplu(void* outP, void* out3, const void* P, size_t n, uint32_t flags, void* stream);

/* ---- scalar multiplication ---------------------------------------------------- */
/* out[i] = scalar_mult_p256(k[i], P[i])   lib/scalar_mult_p256.cpp:12-14
 *        = curve_group<curve_nist_p256>::scalar_mult   curve_group.h:189-218
 * k: n raw 256-bit scalars (not reduced mod the group order);
 * P: n Jacobian points with Z == R (as produced by from_affine); out: n Jacobian
 * points, Montgomery form, the same (X:Y:Z) representative as the reference. */
int ecb200_scalar_mult_p256(void* out, const void* k, const void* P, size_t n, uint32_t flags, void* stream);
/* Same with P = the generator for every lane (curve_group::WJG(), curve_group.h:39-41).
 * The ladder is right-to-left, so its state after the steps for scalar bits 1..16 depends only on
 * those bits: the library keeps the 2^16 states of G in a 10 MiB device table (built on first use,
 * per device) and resumes the ladder at bit 17 -- the reference's Jacobian (X, Y, Z) bit for bit,
 * 16.5 of 256 steps cheaper.  ECB200_NO_BASE_TABLE (or ECB200_BASE_TABLE=0 in the environment)
 * runs the plain ladder. */
int ecb200_scalar_mult_p256_base(void* out, const void* k, size_t n, uint32_t flags, void* stream);
/* scalar_mult_1s: one scalar (8 x u32, host memory) for all lanes   curve_group.h:221-251 */
int ecb200_scalar_mult_p256_1s(void* out, const uint32_t* k1, const void* P, size_t n, uint32_t flags, void* stream);

/* ---- affine <-> Jacobian ------------------------------------------------------- */
/* outJ = wide_jacobian_curve_point::from_affine(xy)   jacobian_curve_point.h:25-31
 * xy: classical affine (x,y), 2 coordinates. */
int ecb200_from_affine(void* outJ, const void* xy, size_t n, uint32_t flags, void* stream);
/* xy = J.to_affine()                                   jacobian_curve_point.h:33-42 */
int ecb200_to_affine(void* xy, const void* J, size_t n, uint32_t flags, void* stream);
/* xy = scalar_mult(k, P).to_affine() in one call -- the pair of operations the reference's own bench
 * times (benchs/curve_group.cpp:28-46, tests/curve_group.cpp:127-131); same values as
 * ecb200_scalar_mult_p256 followed by ecb200_to_affine, but the Jacobian result never leaves the
 * device (host callers move 128 B in and 64 B out per lane).  P == NULL: the generator for every
 * lane (ecb200_scalar_mult_p256_base, with its table unless ECB200_NO_BASE_TABLE). */
int ecb200_scalar_mult_p256_affine(void* xy, const void* k, const void* P, size_t n, uint32_t flags, void* stream);

/* y = wide_curve_point::from_x(x).y(): point decompression, y = sqrt(x^3 - 3x + b) (classical x in,
 * classical y out)   curve_point_ops.h:12-22, curve_group.h:43-58, gfp.h:46-54.
 * ok[i] (n bytes, same memory space as the other pointers) = 1 iff lane i has a square root; the
 * reference's std::optional answers per 4-lane pack: a pack is valid iff all four ok bytes are 1. */
int ecb200_from_x(void* y, uint8_t* ok, const void* x, size_t n, uint32_t flags, void* stream);

/* ---- the field layer for a modulus given at run time -------------------------------------------
 * The reference's field templates take the prime as a parameter, and its own field tests use the
 * secp256k1 prime (tests/mgry.cpp:25-27, tests/ops.cpp:221-252).  p8: the modulus as 8 x u32 in HOST
 * memory, least-significant word first; it must be odd with bit 255 set (R = 2^256).  Same value
 * semantics as the templates for any 256-bit input, including the squaring defect. */
int ecb200_gen_mod_add(void* out, const void* a, const void* b, const uint32_t* p8, size_t n, uint32_t flags, void* stream);          /* modular.h:10-15 */
int ecb200_gen_mod_sub(void* out, const void* a, const void* b, const uint32_t* p8, size_t n, uint32_t flags, void* stream);          /* modular.h:24-41 */
int ecb200_gen_mod_shift_left_one(void* out, const void* a, const uint32_t* p8, size_t n, uint32_t flags, void* stream);              /* modular.h:17-22 */
int ecb200_gen_mgry_mul(void* out, const void* a, const void* b, const uint32_t* p8, size_t n, uint32_t flags, void* stream);         /* mgry_ops.h:31-35 */
int ecb200_gen_mgry_sqr(void* out, const void* a, const uint32_t* p8, size_t n, uint32_t flags, void* stream);                        /* mgry_ops.h:37-42 */
int ecb200_gen_from_classical(void* out, const void* a, const uint32_t* p8, size_t n, uint32_t flags, void* stream);                  /* mgry.h:47-50 */
int ecb200_gen_to_classical(void* out, const void* a, const uint32_t* p8, size_t n, uint32_t flags, void* stream);                    /* mgry.h:52-55 */
/* out = mgry_pow(a, e): e8 = exponent, 8 x u32 in HOST memory (inverse: p-2, sqrt: (p+1)/4)   mgry_ops.h:44-86 */
int ecb200_gen_mgry_pow(void* out, const void* a, const uint32_t* e8, const uint32_t* p8, size_t n, uint32_t flags, void* stream);
int ecb200_gen_opposite(void* out, const void* a, const uint32_t* p8, size_t n, uint32_t flags, void* stream);                        /* gfp.h:60-64 */
/* out16 = mul(a,b): the exact 256x256 -> 512-bit product (mul.h:150-158), and square(a) with the
 * reference's lost carry (mul.h:214-221); ECB200_LAYOUT_LANE only, 16 x u32 per lane out. */
int ecb200_mul512(void* out16, const void* a, const void* b, size_t n, uint32_t flags, void* stream);
int ecb200_square512(void* out16, const void* a, size_t n, uint32_t flags, void* stream);

/* ---- layout and byte-string adapters ----------------------------------------------------- */
/* Re-lay a buffer of n lanes x ncoord coordinates (1 value, 2 affine, 3 Jacobian) from src_layout
 * to dst_layout (ECB200_LAYOUT_*); `flags` carries only the memory space. */
int ecb200_convert_layout(void* dst, uint32_t dst_layout, const void* src, uint32_t src_layout, int ncoord, size_t n, uint32_t flags, void* stream);
/* bn_from_bytes_BE / bn_to_bytes_BE (include/ecsimd/serialization.h:12-48): `bytes` holds n*ncoord
 * big-endian 32-byte strings back to back (lane-major; x|y for ncoord = 2, i.e. SEC1 uncompressed
 * coordinates without the 0x04 prefix); `vals` is a value buffer in the layout given by `flags`. */
int ecb200_bn_from_bytes_be(void* vals, const void* bytes, int ncoord, size_t n, uint32_t flags, void* stream);
int ecb200_bn_to_bytes_be(void* bytes, const void* vals, int ncoord, size_t n, uint32_t flags, void* stream);

/* ---- synthetic inputs, generated on the device (bench / large parity runs) ------ */
/* value i = 4 x splitmix64 words of counter (seed * 0x100000001B3 + 4*(start+i) + limb);
 * kind 0: raw 256 bits (scalars); kind 1: canonical field element (minus p once if >= p). */
int ecb200_synth_values(void* out, uint64_t seed, uint64_t start, int kind, size_t n, uint32_t flags, void* stream);
/* XOR-fold of all 32-bit words of a device buffer into 8 words (order independent
 * checksum used by the multi-GPU parity tests); out8: host memory, 8 x u32. */
int ecb200_checksum(uint32_t* out8, const void* buf, size_t nwords, void* stream);

/* ---- integer-pipe micro-benchmarks (roofline denominators) ------------------------ */
/* which: 0 IMAD.WIDE.U32 (independent chains)  1 IMAD.LO+IMAD.HI pairs  2 IADD3 chains
 *        3 IMAD.WIDE + IADD3 interleaved 1:1     4 IMAD.WIDE.U32.X carry chains
 *        5 VIMNMX3                               6 IMAD.WIDE + IADD3 interleaved 1:2
 * Runs `blocks` x `threads` threads for `iters` loop trips on `stream`, returns the
 * number of counted instructions per thread per trip in *ops_per_iter and the
 * elapsed milliseconds (CUDA events) in *ms. */
int ecb200_microbench(int which, int blocks, int threads, int iters, double* ops_per_iter, float* ms, void* stream);

/* Pipe-mix probe `combo` (0 .. ecb200_microbench_mix_count()-1): per step NW x IMAD.WIDE.U32,
 * NL x IMAD, NH x IMAD.HI.U32, NA x (IADD3 + IADD3.X); 8 steps per loop trip.  counts4 receives
 * {NW, NL, NH, NA}. */
int ecb200_microbench_mix(int combo, int blocks, int threads, int iters, int* counts4, float* ms, void* stream);
int ecb200_microbench_mix_count(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ECB200_H */
