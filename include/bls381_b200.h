/* bls381_b200.h -- C ABI of the B200-native batched BLS12-381 pairing engine.
 *
 * Drop-in boundary for the pairing / verify / verifyBatch / aggregate* / sign path of
 * paulmillr/noble-bls12-381 v1.4.0.  The reference has no FFI of its own (pure TypeScript,
 * SURVEY.md section 8b): each entry point below states which reference function it replaces
 * (file:line relative to the reference tree) and INTEGRATION.md shows the N-API / TypeScript and
 * ctypes bindings a maintainer adds on top.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Host entry points take HOST pointers and copy to/from the
 *     device internally; `_dev` entry points take DEVICE pointers (inputs already resident in HBM)
 *     and run on the caller's CUDA stream.
 *   - Wire format = the reference's own: every field element is 48 bytes big-endian
 *     (Fp.toBytes, math.ts:288-290).
 *       G1 affine   : 96 B  = x || y                                     (index.ts:377-378)
 *       G2 affine   : 192 B = x.c0 || x.c1 || y.c0 || y.c1   (Fp2.toBytes order, math.ts:547-549)
 *       Fp12        : 576 B = 12 coefficients in the flat order of Fp12.fromBigTwelve /
 *                     Fp12.toBytes (math.ts:709-714, 882-884)
 *   - All functions return 0 on success, a negative BLS381_E* code on call-level failure
 *     (bad argument, CUDA error, library not initialised); bls381_last_error() has the text.
 *     Per-item conditions are reported through `status[i]` (BLS381_ST_*), so that the wrapper can
 *     re-create the reference's throw-vs-false behaviour (index.ts:716, :385-386, :635-636, :802-820).
 *   - The library never keeps caller pointers after a call returns.  Thread-safe per call (calls are serialised per
 *     device context).  `_dev` entry points are asynchronous on the caller's stream; their internal scratch (far slots,
 *     partial products, product tree) is kept per stream, so calls on different streams do not share buffers.
 *   - msg_off[0] must be 0 and msg_off non-decreasing (BLS381_EINVAL otherwise).
 *   - There is NO CPU fallback: every entry point fails with BLS381_ENODEV without a CUDA device.
 */
#ifndef BLS381_B200_H
#define BLS381_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLS381_OK 0
#define BLS381_EINVAL (-1)   /* bad argument */
#define BLS381_ENODEV (-2)   /* no CUDA device / driver */
#define BLS381_ECUDA (-3)    /* CUDA runtime error */
#define BLS381_ENOINIT (-4)  /* bls381_init() not called or failed */
#define BLS381_EPROGRAM (-5) /* tower-VM program file missing or malformed */

/* per-item status codes */
#define BLS381_ST_OK 0
#define BLS381_ST_INFINITY 1        /* 'No pairings at point of Infinity'           index.ts:716 */
#define BLS381_ST_NOT_ON_CURVE 2    /* 'Invalid G1/G2 point: not on curve'          index.ts:385,635 */
#define BLS381_ST_NOT_IN_SUBGROUP 3 /* '... must be of prime-order subgroup'        index.ts:386,636 */
#define BLS381_ST_BAD_ENCODING 4    /* 'Invalid compressed G1 point', bad flags     index.ts:312 */
#define BLS381_ST_NO_SQRT 5         /* 'Failed to find a square root'               index.ts:518 */

/* Library lifetime.  `device` = CUDA ordinal (a second call with a different ordinal fails with BLS381_EINVAL).  `program_dir` = directory holding the tower-VM program
 * files (*.b2vm) produced at build time; NULL = "<dir of this shared object>/programs". */
int bls381_init(int device, const char* program_dir);
int bls381_shutdown(void);
const char* bls381_last_error(void);
/* number of streaming multiprocessors of the initialised device (0 if not initialised) */
int bls381_sm_count(void);

/* pairing(P, Q, withFinalExponent)                               replaces index.ts:715-722
 * (the Miller loop math.ts:1331-1388 with fused line evaluation + finalExponentiate math.ts:856-874)
 * n independent pairings e(P_i, Q_i) of affine points.
 *   g1  : n x 96 B,  g2 : n x 192 B,  out : n x 576 B
 *   status : n x int32, filled in the reference's order of checks (index.ts:716-718): INFINITY if P_i or Q_i is the affine
 *            image (0, 0) of the point at infinity (math.ts:955), else the result of P_i.assertValidity(), else that of
 *            Q_i.assertValidity() (NOT_ON_CURVE / NOT_IN_SUBGROUP); the 576 output bytes of a failing item are zero.
 *            NULL = the caller vouches for the points (e.g. they come from bls381_g1_decompress_batch): no checks run. */
int bls381_pairing_batch(const uint8_t* g1, const uint8_t* g2, size_t n, int with_final_exp,
                         uint8_t* out_fp12, int32_t* status);
/* same, device pointers, asynchronous on `cuda_stream` (a cudaStream_t cast to void*) */
int bls381_pairing_batch_dev(const uint8_t* d_g1, const uint8_t* d_g2, size_t n, int with_final_exp,
                             uint8_t* d_out_fp12, void* cuda_stream);

/* Fp12#finalExponentiate()                                      replaces math.ts:856-874
 *   in / out : n x 576 B                                                                         */
int bls381_final_exp_batch(const uint8_t* in_fp12, size_t n, uint8_t* out_fp12);
int bls381_final_exp_batch_dev(const uint8_t* d_in_fp12, size_t n, uint8_t* d_out_fp12, void* cuda_stream);

/* prod_i millerLoop(P_i, Q_i), optionally followed by one shared final exponentiation: the core of
 * verify (index.ts:763-766) and verifyBatch (index.ts:812-817).
 *   out : 576 B.  The product over the batch is an exact product in Fp12, so any sharding/ordering of
 *   the items (warps, CTAs, GPUs) gives the identical 576 bytes.                                   */
int bls381_miller_product(const uint8_t* g1, const uint8_t* g2, size_t n, int with_final_exp,
                          uint8_t* out_fp12);
int bls381_miller_product_dev(const uint8_t* d_g1, const uint8_t* d_g2, size_t n, int with_final_exp,
                              uint8_t* d_out_fp12, void* cuda_stream);

/* PointG1.fromHex for 48-byte compressed public keys incl. assertValidity     replaces index.ts:298-327, 383-388
 *   out96: affine x || y; status[i]: OK / INFINITY (flag bit set, the reference returns ZERO) /
 *   BAD_ENCODING ('Invalid compressed G1 point') / NOT_IN_SUBGROUP.                                  */
int bls381_g1_decompress_batch(const uint8_t* in48, size_t n, uint8_t* out96, int32_t* status);
/* PointG2.fromSignature for 96-byte compressed signatures incl. assertValidity  replaces index.ts:500-530, 633-638
 *   out192: affine x.c0 || x.c1 || y.c0 || y.c1; status: OK / INFINITY / NO_SQRT / NOT_IN_SUBGROUP       */
int bls381_g2_decompress_batch(const uint8_t* in96, size_t n, uint8_t* out192, int32_t* status);
/* PointG1#toHex / toRawBytes (index.ts:355-381) and PointG2#toHex / toSignature (index.ts:586-631) for affine points in
 * the C-ABI layout; the affine image (0, 0) of ZERO encodes the point at infinity (0xc0 / 0x40 flag byte).
 *   G1: compressed 48 B (flag bits: 0x80 compressed, 0x20 = (y * 2) / P), uncompressed 96 B = x || y.
 *   G2: compressed 96 B = x.c1 (+ flags; sign from y.c1, or y.c0 when y.c1 = 0) || x.c0; uncompressed 192 B = x.c1 x.c0 y.c1 y.c0 */
int bls381_g1_encode_batch(const uint8_t* g1_affine, size_t n, int compressed, uint8_t* out);
int bls381_g2_encode_batch(const uint8_t* g2_affine, size_t n, int compressed, uint8_t* out);
/* PointG1.fromHex for 96-byte and PointG2.fromHex for 192-byte UNCOMPRESSED encodings incl. assertValidity
 *                                                                            replaces index.ts:315-325, 533-538, 565-579
 *   out: canonical affine coordinates in the C-ABI layout; status: OK / INFINITY (flag 0x40) / NOT_ON_CURVE /
 *   NOT_IN_SUBGROUP / BAD_ENCODING (G2: 'Invalid encoding flag', or the compression bit set on a 192-byte input)      */
int bls381_g1_from_uncompressed_batch(const uint8_t* in96, size_t n, uint8_t* out96, int32_t* status);
int bls381_g2_from_uncompressed_batch(const uint8_t* in192, size_t n, uint8_t* out192, int32_t* status);
/* PointG2.hashToCurve(msg, {DST})                                              replaces index.ts:481-490
 * (expand_message_xmd index.ts:207-231, hash_to_field :240-267, SWU math.ts:1220-1267, 3-isogeny
 *  math.ts:1315-1325, clearCofactor index.ts:659-672).  msgs: packed bytes, msg_off[n+1] byte offsets.
 *   out192: affine H(m_i).                                                                             */
int bls381_hash_to_g2_batch(const uint8_t* msgs, const uint64_t* msg_off, size_t n, const uint8_t* dst,
                            size_t dst_len, uint8_t* out192);
/* PointG1.hashToCurve(msg, {DST}) (min-signature deployments)                   replaces index.ts:331-339
 * (hash_to_field with m = 1, map_to_curve_simple_swu_3mod4 math.ts:1270-1313, the 11-isogeny math.ts:1327 + 1612-1790,
 *  clearCofactor index.ts:401-405).   out96: affine H(m_i).                                                        */
int bls381_hash_to_g1_batch(const uint8_t* msgs, const uint64_t* msg_off, size_t n, const uint8_t* dst, size_t dst_len,
                            uint8_t* out96);
/* verifyBatch(signature, messages, publicKeys) for byte inputs                  replaces index.ts:792-821
 *   sig96: aggregated signature; pks48: n compressed public keys; status: n + 1 codes (public keys, then signature).
 *   *verdict: 1 true, 0 false, -1 = the reference would throw (a decoding / validity error, see status).   */
int bls381_verify_batch(const uint8_t* sig96, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* pks48,
                        size_t n, const uint8_t* dst, size_t dst_len, int* verdict, int32_t* status);

/* One shard of verifyBatch for multi-GPU runs: the un-exponentiated product of the shard's Miller
 * loops e(pk_i, H(m_i)), times e(-G1, sig) when sig96_or_null != NULL (rank 0).  Partials of all ranks are
 * all-gathered and combined with bls381_fp12_product(..., with_final_exp = 1).  status: n (+1 with a signature). */
int bls381_verify_batch_partial(const uint8_t* sig96_or_null, const uint8_t* msgs, const uint64_t* msg_off,
                                const uint8_t* pks48, size_t n, const uint8_t* dst, size_t dst_len,
                                uint8_t* out_fp12, int32_t* status);
/* the same, the product stays in device memory (d_out_fp12 = device pointer): a multi-process run all-gathers the partials
 * device to device and finishes with bls381_fp12_product_dev */
int bls381_verify_batch_partial_dev(const uint8_t* sig96_or_null, const uint8_t* msgs, const uint64_t* msg_off,
                                    const uint8_t* pks48, size_t n, const uint8_t* dst, size_t dst_len,
                                    uint8_t* d_out_fp12, int32_t* status);

/* In-process multi-GPU (north_star: "large verifyBatch workloads shard by signature index across the 8 GPUs of
 * one box with a single all-reduce of the partial Gt products over NVLink").
 *   bls381_init_devices(mask): one context per CUDA device whose bit is set (bit d = ordinal d); the first one is also
 *   the context of every single-device entry point.  NCCL is loaded at run time (dlopen libnccl.so.2, ncclCommInitAll);
 *   bls381_multi_transport() says "nccl" or, if that failed, "peer-copy" (cudaMemcpyPeer of the 576-byte partials).
 *   bls381_verify_batch_multi: verifyBatch (index.ts:792-821) with the items sharded by index over the contexts (contiguous
 *   blocks, one host thread per device), ONE all-gather of W x 576 bytes device to device -- NCCL has no modular-product
 *   reduction op; gather + product is the all-reduce of this monoid -- then the product and ONE final exponentiation on the
 *   first device.  Same arguments, verdict and status layout as bls381_verify_batch; same result bit for bit.        */
int bls381_init_devices(uint32_t device_mask, const char* program_dir);
int bls381_device_count(void);
const char* bls381_multi_transport(void);
int bls381_verify_batch_multi(const uint8_t* sig96, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* pks48,
                              size_t n, const uint8_t* dst, size_t dst_len, int* verdict, int32_t* status);

/* sign(message, privateKey) for byte messages                                   replaces index.ts:746-752
 * (hashToCurve + constant-time scalar multiplication math.ts:1061-1078 + toSignature index.ts:586-598).
 *   sks32: n x 32 B big-endian scalars already normalised to 0 < sk < r (normalizePrivKey index.ts:269-279 is
 *   host-side argument checking; larger 32-byte values are reduced mod r on the device, sk = 0 yields the encoding
 *   of the point at infinity); out_sig96: n x 96 B compressed signatures.                                   */
int bls381_sign_batch(const uint8_t* sks32, const uint8_t* msgs, const uint64_t* msg_off, size_t n,
                      const uint8_t* dst, size_t dst_len, uint8_t* out_sig96);
/* getPublicKey(privateKey)                                                      replaces index.ts:738-740 (351-353)
 *   sks32: n x 32 B big-endian scalars (reduced mod r on the device without secret-dependent branches, like
 *   normalizePrivKey index.ts:269-279; the reference's fixed-base wNAF table math.ts:1086-1167 yields the same point);
 *   out48: n compressed public keys (sk = 0 mod r gives the encoding of the point at infinity; the wrapper rejects it). */
int bls381_get_public_key_batch(const uint8_t* sks32, size_t n, uint8_t* out48);
/* aggregatePublicKeys(Hex[]) / aggregateSignatures(Hex[])                       replaces index.ts:773-788
 *   n compressed points in, one compressed point out; status[i] per input (any status other than OK / INFINITY
 *   means the reference would have thrown while decoding that element).                                    */
int bls381_aggregate_g1(const uint8_t* pks48, size_t n, uint8_t* out48, int32_t* status);
int bls381_aggregate_g2(const uint8_t* sigs96, size_t n, uint8_t* out96, int32_t* status);
/* PointG1#assertValidity / PointG2#assertValidity on affine points           replaces index.ts:383-388, 633-638
 * (isOnCurve :408-414 / :675-681, isTorsionFree :444-448 / :688-690).  status: OK / NOT_ON_CURVE / NOT_IN_SUBGROUP */
int bls381_g1_validate_batch(const uint8_t* g1_affine, size_t n, int32_t* status);
int bls381_g2_validate_batch(const uint8_t* g2_affine, size_t n, int32_t* status);
/* PointG2#multiply(scalar) (math.ts:1061-1078), used by sign(PointG2, key) index.ts:749-750.
 *   scalars32: 0 < k <= r big-endian; out192 affine; flags[i] bit1 = result is the point at infinity          */
int bls381_g2_scalar_mul_batch(const uint8_t* g2_affine, const uint8_t* scalars32, size_t n, uint8_t* out192,
                               int32_t* flags);
/* PointG1#multiply / multiplyUnsafe (math.ts:1048-1078): same for G1 (96-byte affine points).                  */
int bls381_g1_scalar_mul_batch(const uint8_t* g1_affine, const uint8_t* scalars32, size_t n, uint8_t* out96,
                               int32_t* flags);
/* prod_i f_i in Fp12 (+ optional finalExponentiate): combines the per-GPU partial products of a sharded
 * verifyBatch after the all-gather (index.ts:815-816).  in: n x 576 B, out: 576 B.                          */
int bls381_fp12_product(const uint8_t* in_fp12, size_t n, int with_final_exp, uint8_t* out_fp12);
int bls381_fp12_product_dev(const uint8_t* d_in_fp12, size_t n, int with_final_exp, uint8_t* d_out_fp12, void* cuda_stream);

/* Generic tower-VM launch (used by the tests and by the entry points above).
 *   program  : name of a loaded program ("pairing", "miller", "final_exp", ...)
 *   bufs     : up to 8 DEVICE buffers; strides[i] = bytes per item in buffer i                      */
int bls381_vm_run_dev(const char* program, uint8_t* const* d_bufs, const uint32_t* strides, int nbuf,
                      size_t n_items, void* cuda_stream);
/* Loads a program from memory (file image) under `name`, replacing any previous one. */
int bls381_vm_load(const char* name, const uint8_t* image, size_t len);

/* Engine tuning knobs (same as the BLS381_B200_* environment variables read by bls381_init): "dynamic_batches" (1 = CTAs
 * claim 32-item batches from a global counter, 0 = round-robin), "ctas_per_sm" (0 = automatic), "poll_sleep_ns",
 * "no_tma" (1 = read wire-format inputs directly from global memory), "pairs_per_lane" (1..4, default 3: the
 * Miller-product entry points give every lane that many consecutive items, which share the Fp12 squarings),
 * "swu_kernel" / "tail_kernels" / "g1_kernel" / "g2_kernel" / "validate_kernels" (default 1: hash_to_field + SWU, the tail of hash-to-curve and
 * the sign ladder, G1 key decompression, batches of compressed signatures and the validity checks of affine points run as hand-written
 * per-item kernels; 0 = the tower-VM programs of the same functions,
 * kept as the A/B path), "fixed_base" (default 1: getPublicKey adds 64 points of a fixed-base table of G1, one per nibble of
 * the key, every table entry read for every key; 0 = the constant-time ladder program g1_scalar_mul), "pipeline_copies" (default 1: bls381_pairing_batch without a status array cuts batches of at
 * least five rounds of resident CTAs into three chunks on two streams, so that only the first chunk's copy in and the last
 * chunk's copy out are exposed; 0 = one copy in, one launch, one copy out).  Results never depend on them.           */
int bls381_set_option(const char* name, int value);

/* Measurement aids (bench.py): number of kernel launches (tower-VM and per-item kernels) since init, and a dependent-free
 * IMAD.WIDE.U32 issue-rate microbenchmark (returns multiply-adds per second on the whole device). */
uint64_t bls381_launch_count(void);
int bls381_imad_peak(double* imad_per_second);
/* device time in milliseconds of the last tower-VM kernel launched through a host entry point */
double bls381_last_kernel_ms(void);
/* effective SM clock (MHz) during the last tower-VM launch: cycles / nanoseconds seen by CTA 0 (clock64 vs globaltimer) */
double bls381_last_kernel_sm_mhz(void);
/* the same microbenchmark launched back to back for `seconds`; reports the rate of the second half (the sustained
 * figure, with power and clocks settled) and the effective SM clock of its last launch */
int bls381_imad_peak_sustained(double seconds, double* imad_per_second, double* sm_mhz);

#ifdef __cplusplus
}
#endif
#endif /* BLS381_B200_H */
