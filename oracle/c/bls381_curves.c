/* bls381_curves.c -- plain-C CPU restatement of the reference's sign / hash-to-curve / key path.  TEST INFRASTRUCTURE ONLY.
 *
 * Extends bls381_oracle.c (included below: the Fp / Fp2 / Fp12 arithmetic is shared) with restatements of
 *   ProjectivePoint double / add / multiply            math.ts:974-1025, 1048-1078
 *   sgn0_fp2, sqrt_div_fp2, map_to_curve_simple_swu_9mod16   math.ts:1179-1267
 *   isogenyMapG2 + its coefficient tables              math.ts:1315-1325, 1547-1610
 *   psi / psi2                                         math.ts:1398-1408
 *   expand_message_xmd, hash_to_field                  index.ts:207-267   (SHA-256: FIPS 180-4, the reference takes it
 *                                                      from node:crypto / WebCrypto, index.ts:39-48)
 *   PointG2.hashToCurve, clearCofactor, toSignature    index.ts:481-490, 659-672, 586-598
 *   PointG1.toHex(compressed), getPublicKey            index.ts:359-371, 738-740
 *   sign                                               index.ts:746-752
 * so that the full-size configurations (1 048 576 signatures, BASELINE config 4) can be byte-compared against a CPU
 * checker that finishes in minutes (the Python oracle needs 73 ms per signature).
 * Parity: pinned -- tests/test_oracle_c.py checks it against the reference's 559 sign KATs
 * (test/bls12-381-g2-test-vectors.txt), the RFC / kilic hash-to-curve vectors (test/hashToCurve.test.ts) and the
 * zkcrypto G1 encodings (test/zkcrypto), all committed under tests/golden/.
 * Only tests/, __graft_entry__.smoke() and bench.py's parity / cpu_baseline legs may use it.
 */
#include "bls381_oracle.c"

/* ---------------------------------------------------------------- SHA-256 (FIPS 180-4) */
typedef struct { uint32_t h[8]; uint8_t buf[64]; uint64_t len; uint32_t fill; } sha256_t;
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha_block(sha256_t* s, const uint8_t* p) {
    uint32_t w[64], a, b, c, d, e, f, g, h;
    for (int i = 0; i < 16; ++i) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    a = s->h[0]; b = s->h[1]; c = s->h[2]; d = s->h[3]; e = s->h[4]; f = s->h[5]; g = s->h[6]; h = s->h[7];
    for (int i = 0; i < 64; ++i) {
        uint32_t t1 = h + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i];
        uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    s->h[0] += a; s->h[1] += b; s->h[2] += c; s->h[3] += d; s->h[4] += e; s->h[5] += f; s->h[6] += g; s->h[7] += h;
}
static void sha_init(sha256_t* s) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(s->h, iv, sizeof(iv)); s->len = 0; s->fill = 0;
}
static void sha_update(sha256_t* s, const uint8_t* p, size_t n) {
    s->len += n;
    while (n) {
        size_t k = 64 - s->fill; if (k > n) k = n;
        memcpy(s->buf + s->fill, p, k); s->fill += (uint32_t)k; p += k; n -= k;
        if (s->fill == 64) { sha_block(s, s->buf); s->fill = 0; }
    }
}
static void sha_final(sha256_t* s, uint8_t out[32]) {
    uint64_t bits = s->len * 8;
    uint8_t pad = 0x80; sha_update(s, &pad, 1);
    uint8_t z = 0; while (s->fill != 56) sha_update(s, &z, 1);
    uint8_t lb[8]; for (int i = 0; i < 8; ++i) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    sha_update(s, lb, 8);
    for (int i = 0; i < 8; ++i) { out[4 * i] = (uint8_t)(s->h[i] >> 24); out[4 * i + 1] = (uint8_t)(s->h[i] >> 16); out[4 * i + 2] = (uint8_t)(s->h[i] >> 8); out[4 * i + 3] = (uint8_t)s->h[i]; }
}

/* expand_message_xmd (index.ts:207-231), SHA-256, len_in_bytes <= 255*32 */
static int expand_message_xmd(const uint8_t* msg, size_t mlen, const uint8_t* dst, size_t dlen, size_t len_in_bytes, uint8_t* out) {
    uint8_t dprime[256 + 1]; size_t dpl;
    if (dlen > 255) { /* index.ts:214 */
        sha256_t s; sha_init(&s); sha_update(&s, (const uint8_t*)"H2C-OVERSIZE-DST-", 17); sha_update(&s, dst, dlen); sha_final(&s, dprime); dpl = 32;
    } else { memcpy(dprime, dst, dlen); dpl = dlen; }
    dprime[dpl] = (uint8_t)dpl; ++dpl;
    const size_t ell = (len_in_bytes + 31) / 32;
    if (ell > 255) return -1;
    uint8_t zpad[64] = {0}, b0[32], bi[32], tmp[32];
    uint8_t lib[3] = {(uint8_t)(len_in_bytes >> 8), (uint8_t)len_in_bytes, 0};
    sha256_t s; sha_init(&s); sha_update(&s, zpad, 64); sha_update(&s, msg, mlen); sha_update(&s, lib, 3); sha_update(&s, dprime, dpl); sha_final(&s, b0);
    uint8_t ctr = 1;
    sha_init(&s); sha_update(&s, b0, 32); sha_update(&s, &ctr, 1); sha_update(&s, dprime, dpl); sha_final(&s, bi);
    size_t done = 0;
    for (size_t i = 1;; ++i) {
        size_t k = len_in_bytes - done; if (k > 32) k = 32;
        memcpy(out + done, bi, k); done += k;
        if (done == len_in_bytes) break;
        for (int j = 0; j < 32; ++j) tmp[j] = b0[j] ^ bi[j];
        ctr = (uint8_t)(i + 1);
        sha_init(&s); sha_update(&s, tmp, 32); sha_update(&s, &ctr, 1); sha_update(&s, dprime, dpl); sha_final(&s, bi);
    }
    return 0;
}

/* ---------------------------------------------------------------- small helpers on top of bls381_oracle.c */
static fp2 FP2_ONE_, FP2_ZERO_;
static void fp_plain(fp r, const fp a) { fp one = {1, 0, 0, 0, 0, 0}; fp_mul(r, a, one); } /* out of Montgomery form */
static void fp_set_u64(fp r, uint64_t v) { fp t = {v, 0, 0, 0, 0, 0}; fp_mul(r, t, R2); }
static void fp_from_hex(fp r, const char* hex) { /* big-endian hex (up to 96 digits) */
    uint8_t b[48] = {0}; size_t n = strlen(hex);
    for (size_t i = 0; i < n; ++i) {
        char ch = hex[n - 1 - i]; int v = ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10;
        b[47 - i / 2] |= (uint8_t)(v << (4 * (i & 1)));
    }
    fp_from_bytes(r, b);
}
static int fp2_is_zero(const fp2* a) { return fp_is_zero(a->c0) && fp_is_zero(a->c1); }
static int fp2_eq(const fp2* a, const fp2* b) { return fp_eq(a->c0, b->c0) && fp_eq(a->c1, b->c1); }
static void fp2_pow_u64(fp2* r, const fp2* a, uint64_t e) { uint8_t be[8]; for (int i = 0; i < 8; ++i) be[i] = (uint8_t)(e >> (56 - 8 * i)); fp2_pow_big(r, a, be, 8); }
/* (y*2)/P for a canonical value: 1 iff y > (p-1)/2 (index.ts:313, 594) */
static int fp_gt_half(const fp a) {
    fp t; fp_plain(t, a);
    uint64_t d[6]; uint64_t carry = 0;
    for (int i = 0; i < 6; ++i) { uint64_t hi = t[i] >> 63; d[i] = (t[i] << 1) | carry; carry = hi; }
    if (carry) return 1;
    return geq_p(d);
}

/* ---------------------------------------------------------------- projective points over Fp2 (math.ts:893-1078) */
typedef struct { fp2 x, y, z; } g2p;
static int g2_is_zero(const g2p* p) { return fp2_is_zero(&p->z); }                       /* :903-905 */
static void g2_zero(g2p* p) { p->x = FP2_ONE_; p->y = FP2_ONE_; p->z = FP2_ZERO_; }        /* :910-912 (1, 1, 0) */
static void g2_neg(g2p* r, const g2p* p) { r->x = p->x; fp2_neg(&r->y, &p->y); r->z = p->z; } /* :929-931 */
static void g2_double(g2p* r, const g2p* p) { /* :974-989 dbl-1998-cmo-2 */
    fp2 W, S, SS, SSS, B, H, t, u;
    fp2_sqr(&t, &p->x); fp2_mul_small(&W, &t, 3);
    fp2_mul(&S, &p->y, &p->z);
    fp2_sqr(&SS, &S);
    fp2_mul(&SSS, &SS, &S);
    fp2_mul(&t, &p->x, &p->y); fp2_mul(&B, &t, &S);
    fp2_sqr(&t, &W); fp2_mul_small(&u, &B, 8); fp2_sub(&H, &t, &u);
    g2p o;
    fp2_mul(&t, &H, &S); fp2_mul_small(&o.x, &t, 2);
    fp2_mul_small(&t, &B, 4); fp2_sub(&t, &t, &H); fp2_mul(&t, &W, &t);
    fp2_sqr(&u, &p->y); fp2_mul_small(&u, &u, 8); fp2_mul(&u, &u, &SS);
    fp2_sub(&o.y, &t, &u);
    fp2_mul_small(&o.z, &SSS, 8);
    *r = o;
}
static void g2_add(g2p* r, const g2p* p1, const g2p* p2) { /* :993-1025 add-1998-cmo-2 */
    if (g2_is_zero(p1)) { *r = *p2; return; }
    if (g2_is_zero(p2)) { *r = *p1; return; }
    fp2 U1, U2, V1, V2, U, V, VV, VVV, V2VV, W, A, t, u;
    fp2_mul(&U1, &p2->y, &p1->z); fp2_mul(&U2, &p1->y, &p2->z);
    fp2_mul(&V1, &p2->x, &p1->z); fp2_mul(&V2, &p1->x, &p2->z);
    if (fp2_eq(&V1, &V2) && fp2_eq(&U1, &U2)) { g2_double(r, p1); return; }
    if (fp2_eq(&V1, &V2)) { g2_zero(r); return; }
    fp2_sub(&U, &U1, &U2); fp2_sub(&V, &V1, &V2);
    fp2_sqr(&VV, &V); fp2_mul(&VVV, &VV, &V); fp2_mul(&V2VV, &V2, &VV);
    fp2_mul(&W, &p1->z, &p2->z);
    fp2_sqr(&t, &U); fp2_mul(&t, &t, &W); fp2_sub(&t, &t, &VVV); fp2_mul_small(&u, &V2VV, 2); fp2_sub(&A, &t, &u);
    g2p o;
    fp2_mul(&o.x, &V, &A);
    fp2_sub(&t, &V2VV, &A); fp2_mul(&t, &U, &t); fp2_mul(&u, &VVV, &U2); fp2_sub(&o.y, &t, &u);
    fp2_mul(&o.z, &VVV, &W);
    *r = o;
}
static void g2_sub(g2p* r, const g2p* a, const g2p* b) { g2p n; g2_neg(&n, b); g2_add(r, a, &n); } /* :1027-1033 */
/* multiplyUnsafe (:1048-1058) for a scalar given as big-endian bytes; the constant-time ladder of :1061-1078 computes the
   same group element (it adds into a discarded `fake` point for zero bits) */
static void g2_mul_be(g2p* r, const g2p* p, const uint8_t* k, int nbytes) {
    g2p acc, d = *p; g2_zero(&acc);
    for (int i = nbytes - 1; i >= 0; --i)
        for (int b = 0; b < 8; ++b) {
            if ((k[i] >> b) & 1) g2_add(&acc, &acc, &d);
            g2_double(&d, &d);
        }
    *r = acc;
}
static void g2_to_affine(fp2* ax, fp2* ay, const g2p* p) { /* :949-958 (ZERO -> (0, 0)) */
    if (g2_is_zero(p)) { *ax = FP2_ZERO_; *ay = FP2_ZERO_; return; }
    fp2 zi; fp2_inv(&zi, &p->z); fp2_mul(ax, &p->x, &zi); fp2_mul(ay, &p->y, &zi);
}
static const uint8_t X_BE[8] = {0xd2, 0x01, 0x00, 0x00, 0x00, 0x01, 0x00, 0x00}; /* |x| math.ts:48 */
static void g2_mul_curve_x(g2p* r, const g2p* p) { g2p t; g2_mul_be(&t, p, X_BE, 8); g2_neg(r, &t); } /* index.ts:651-653 */

/* psi / psi2 (math.ts:1398-1408): the untwist-Frobenius-twist map reduces to conj(x)*cx, conj(y)*cy with
   cx = 1/xi^((p-1)/3), cy = 1/xi^((p-1)/2) (oracle/noble_oracle.py asserts the equality with the Fp12 route) */
static fp2 PSI_CX, PSI_CY; static fp PSI2_C1;
static void g2_psi(g2p* r, const g2p* p) { /* index.ts:641-643 */
    fp2 x, y, t; g2_to_affine(&x, &y, p);
    fp2_conj(&t, &x); fp2_mul(&r->x, &t, &PSI_CX);
    fp2_conj(&t, &y); fp2_mul(&r->y, &t, &PSI_CY);
    r->z = FP2_ONE_;
}
static void g2_psi2(g2p* r, const g2p* p) { /* index.ts:646-648, math.ts:1406-1408 */
    fp2 x, y; g2_to_affine(&x, &y, p);
    fp2_mul_fp(&r->x, &x, PSI2_C1); fp2_neg(&r->y, &y); r->z = FP2_ONE_;
}
static void g2_clear_cofactor(g2p* r, const g2p* p) { /* index.ts:659-672 */
    g2p t1, t2, t3;
    g2_mul_curve_x(&t1, p);
    g2_psi(&t2, p);
    g2_double(&t3, p);
    g2_psi2(&t3, &t3);
    g2_sub(&t3, &t3, &t2);
    g2_add(&t2, &t1, &t2);
    g2_mul_curve_x(&t2, &t2);
    g2_add(&t3, &t3, &t2);
    g2_sub(&t3, &t3, &t1);
    g2_sub(r, &t3, p);
}

/* ---------------------------------------------------------------- hash to curve G2 (math.ts:1179-1325) */
static fp2 ROOTS4[4], ETAS[4], ISO_A, ISO_B, ISO_Z, ISO3[4][4];
static uint8_t EXP_P2M9_16[96]; /* (p^2 - 9) / 16, big-endian */
static int sgn0_fp2(const fp2* x) { /* :1179-1185 */
    fp a, b; fp_plain(a, x->c0); fp_plain(b, x->c1);
    int sign_0 = (int)(a[0] & 1), zero_0 = fp_is_zero(a), sign_1 = (int)(b[0] & 1);
    return sign_0 || (zero_0 && sign_1);
}
static int sqrt_div_fp2(fp2* result, const fp2* u, const fp2* v) { /* :1195-1214 */
    fp2 v7, uv7, uv15, gamma, t, cand;
    fp2_pow_u64(&v7, v, 7);
    fp2_mul(&uv7, u, &v7);
    fp2_mul(&t, &v7, v); fp2_mul(&uv15, &uv7, &t);
    fp2_pow_big(&t, &uv15, EXP_P2M9_16, 96); fp2_mul(&gamma, &t, &uv7);
    int success = 0; *result = gamma;
    for (int k = 0; k < 4; ++k) {
        fp2_mul(&cand, &ROOTS4[k], &gamma);
        fp2_sqr(&t, &cand); fp2_mul(&t, &t, v); fp2_sub(&t, &t, u);
        if (fp2_is_zero(&t) && !success) { success = 1; *result = cand; }
    }
    return success;
}
static int swu_g2(fp2* xo, fp2* yo, const fp2* tt) { /* map_to_curve_simple_swu_9mod16 :1220-1267 */
    fp2 t2, zt2, ztzt, den, num, v, u, t, w, y, cand_x1, cx;
    fp2_sqr(&t2, tt);
    fp2_mul(&zt2, &ISO_Z, &t2);
    fp2_sqr(&t, &zt2); fp2_add(&ztzt, &zt2, &t);
    fp2_mul(&t, &ISO_A, &ztzt); fp2_neg(&den, &t);
    fp2_add(&t, &ztzt, &FP2_ONE_); fp2_mul(&num, &ISO_B, &t);
    if (fp2_is_zero(&den)) fp2_mul(&den, &ISO_Z, &ISO_A);
    fp2_pow_u64(&v, &den, 3);
    fp2_pow_u64(&u, &num, 3);
    fp2_mul(&t, &ISO_A, &num); fp2_sqr(&w, &den); fp2_mul(&t, &t, &w); fp2_add(&u, &u, &t);
    fp2_mul(&t, &ISO_B, &v); fp2_add(&u, &u, &t);
    fp2 cand;
    int success = sqrt_div_fp2(&cand, &u, &v), have_y = 0;
    if (success) { y = cand; have_y = 1; }
    fp2_pow_u64(&t, tt, 3); fp2_mul(&cand_x1, &cand, &t);
    fp2_pow_u64(&t, &zt2, 3); fp2_mul(&u, &t, &u);
    int success2 = 0;
    for (int k = 0; k < 4; ++k) {
        fp2_mul(&cx, &ETAS[k], &cand_x1);
        fp2_sqr(&t, &cx); fp2_mul(&t, &t, &v); fp2_sub(&t, &t, &u);
        if (fp2_is_zero(&t) && !success && !success2) { y = cx; success2 = 1; have_y = 1; }
    }
    if (!have_y) return -1; /* 'Hash to Curve - Optimized SWU failure' */
    if (success2) fp2_mul(&num, &num, &zt2);
    if (sgn0_fp2(tt) != sgn0_fp2(&y)) fp2_neg(&y, &y);
    fp2_inv(&t, &den); fp2_mul(xo, &num, &t); /* numerator.div(denominator) */
    *yo = y;
    return 0;
}
static void isogeny_map_g2(fp2* xo, fp2* yo, const fp2* x, const fp2* y) { /* :1315-1325, Horner over the four tables */
    fp2 acc[4], t;
    for (int k = 0; k < 4; ++k) {
        acc[k] = ISO3[k][0];
        for (int j = 1; j < 4; ++j) { fp2_mul(&t, &acc[k], x); fp2_add(&acc[k], &t, &ISO3[k][j]); }
    }
    fp2_inv(&t, &acc[1]); fp2_mul(xo, &acc[0], &t);
    fp2_inv(&t, &acc[3]); fp2_mul(&t, &acc[2], &t); fp2_mul(yo, y, &t);
}
/* os2ip(64 bytes) mod p (index.ts:257-263): hi * 2^256 + lo with both halves < 2^256 < p */
static fp TWO256;
static void fp_from_64(fp r, const uint8_t* b) {
    uint8_t t[48]; fp hi, lo;
    memset(t, 0, 16); memcpy(t + 16, b, 32); fp_from_bytes(hi, t);
    memcpy(t + 16, b + 32, 32); fp_from_bytes(lo, t);
    fp_mul(hi, hi, TWO256); fp_add(r, hi, lo);
}
static int hash_to_g2(g2p* out, const uint8_t* msg, size_t mlen, const uint8_t* dst, size_t dlen) { /* index.ts:481-490 */
    uint8_t prb[256];
    if (expand_message_xmd(msg, mlen, dst, dlen, 256, prb)) return -1;   /* hash_to_field: count 2, m 2, L 64 */
    fp2 u0, u1; g2p q0, q1, s, e;
    fp_from_64(u0.c0, prb); fp_from_64(u0.c1, prb + 64); fp_from_64(u1.c0, prb + 128); fp_from_64(u1.c1, prb + 192);
    if (swu_g2(&q0.x, &q0.y, &u0) || swu_g2(&q1.x, &q1.y, &u1)) return -1;
    q0.z = FP2_ONE_; q1.z = FP2_ONE_;
    g2_add(&s, &q0, &q1);
    fp2 ax, ay; g2_to_affine(&ax, &ay, &s);
    isogeny_map_g2(&e.x, &e.y, &ax, &ay); e.z = FP2_ONE_;
    g2_clear_cofactor(out, &e);
    return 0;
}
static void g2_to_signature(uint8_t* out96, const g2p* p) { /* index.ts:586-598 */
    if (g2_is_zero(p)) { memset(out96, 0, 96); out96[0] = 0xc0; return; }
    fp2 x, y; g2_to_affine(&x, &y, p);
    int aflag = fp_is_zero(y.c1) ? fp_gt_half(y.c0) : fp_gt_half(y.c1);
    fp_to_bytes(out96, x.c1); fp_to_bytes(out96 + 48, x.c0);
    out96[0] |= (uint8_t)(0x80 | (aflag << 5));
}

/* ---------------------------------------------------------------- G1 (math.ts:974-1058 over Fp; index.ts:359-371) */
typedef struct { fp x, y, z; } g1p;
static int g1_is_zero(const g1p* p) { return fp_is_zero(p->z); }
static void g1_zero(g1p* p) { fp_copy(p->x, R1); fp_copy(p->y, R1); fp_zero(p->z); }
static void g1_double(g1p* r, const g1p* p) { /* :974-989 */
    fp W, S, SS, SSS, B, H, t, u; g1p o;
    fp_sqr(t, p->x); fp_mul_small(W, t, 3);
    fp_mul(S, p->y, p->z); fp_sqr(SS, S); fp_mul(SSS, SS, S);
    fp_mul(t, p->x, p->y); fp_mul(B, t, S);
    fp_sqr(t, W); fp_mul_small(u, B, 8); fp_sub(H, t, u);
    fp_mul(t, H, S); fp_mul_small(o.x, t, 2);
    fp_mul_small(t, B, 4); fp_sub(t, t, H); fp_mul(t, W, t);
    fp_sqr(u, p->y); fp_mul_small(u, u, 8); fp_mul(u, u, SS);
    fp_sub(o.y, t, u);
    fp_mul_small(o.z, SSS, 8);
    *r = o;
}
static void g1_add(g1p* r, const g1p* p1, const g1p* p2) { /* :993-1025 */
    if (g1_is_zero(p1)) { *r = *p2; return; }
    if (g1_is_zero(p2)) { *r = *p1; return; }
    fp U1, U2, V1, V2, U, V, VV, VVV, V2VV, W, A, t, u; g1p o;
    fp_mul(U1, p2->y, p1->z); fp_mul(U2, p1->y, p2->z); fp_mul(V1, p2->x, p1->z); fp_mul(V2, p1->x, p2->z);
    if (fp_eq(V1, V2) && fp_eq(U1, U2)) { g1_double(r, p1); return; }
    if (fp_eq(V1, V2)) { g1_zero(r); return; }
    fp_sub(U, U1, U2); fp_sub(V, V1, V2);
    fp_sqr(VV, V); fp_mul(VVV, VV, V); fp_mul(V2VV, V2, VV); fp_mul(W, p1->z, p2->z);
    fp_sqr(t, U); fp_mul(t, t, W); fp_sub(t, t, VVV); fp_mul_small(u, V2VV, 2); fp_sub(A, t, u);
    fp_mul(o.x, V, A);
    fp_sub(t, V2VV, A); fp_mul(t, U, t); fp_mul(u, VVV, U2); fp_sub(o.y, t, u);
    fp_mul(o.z, VVV, W);
    *r = o;
}
static void g1_mul_be(g1p* r, const g1p* p, const uint8_t* k, int nbytes) { /* :1048-1058 */
    g1p acc, d = *p; g1_zero(&acc);
    for (int i = nbytes - 1; i >= 0; --i)
        for (int b = 0; b < 8; ++b) {
            if ((k[i] >> b) & 1) g1_add(&acc, &acc, &d);
            g1_double(&d, &d);
        }
    *r = acc;
}
static g1p G1_BASE_;
static void g1_to_hex_compressed(uint8_t* out48, const g1p* p) { /* index.ts:359-371 */
    if (g1_is_zero(p)) { memset(out48, 0, 48); out48[0] = 0xc0; return; }
    fp zi, x, y; fp_inv(zi, p->z); fp_mul(x, p->x, zi); fp_mul(y, p->y, zi);
    fp_to_bytes(out48, x);
    out48[0] |= (uint8_t)(0x80 | (fp_gt_half(y) << 5));
}

/* ---------------------------------------------------------------- scalars mod r (index.ts:269-279) */
static const uint64_t R_ORD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static int normalize_priv_key(uint8_t out[32], const uint8_t sk[32]) { /* sk mod r, big-endian; returns -1 if the result is 0 */
    uint64_t w[4];
    for (int i = 0; i < 4; ++i) { uint64_t v = 0; for (int k = 0; k < 8; ++k) v = (v << 8) | sk[8 * (3 - i) + k]; w[i] = v; }
    for (;;) {
        int ge = 1;
        for (int i = 3; i >= 0; --i) if (w[i] != R_ORD[i]) { ge = w[i] > R_ORD[i]; break; }
        if (!ge) break;
        uint64_t borrow = 0;
        for (int i = 0; i < 4; ++i) { uint64_t a = w[i], s1 = a - R_ORD[i], s2 = s1 - borrow; borrow = (a < R_ORD[i]) | (s1 < borrow); w[i] = s2; }
    }
    for (int i = 0; i < 4; ++i) for (int k = 0; k < 8; ++k) out[8 * (3 - i) + k] = (uint8_t)(w[i] >> (56 - 8 * k));
    return (w[0] | w[1] | w[2] | w[3]) ? 0 : -1;
}

/* ---------------------------------------------------------------- init + exported checkers */
static int g_init2 = 0;
static void set_fp2_hex(fp2* r, const char* c0, const char* c1) { fp_from_hex(r->c0, c0); fp_from_hex(r->c1, c1); }
static void neg_fp2_parts(fp2* r, int neg0, int neg1) { if (neg0) fp_neg(r->c0, r->c0); if (neg1) fp_neg(r->c1, r->c1); }
void oracle_curves_init(void) {
    if (g_init2) return;
    oracle_init();
    memset(&FP2_ZERO_, 0, sizeof(fp2)); FP2_ONE_ = FP2_ZERO_; fp_copy(FP2_ONE_.c0, R1);
    /* roots of unity / etas (math.ts:1415-1452) */
    const char* rv1 = "6af0e0437ff400b6831e36d6bd17ffe48395dabc2d3435e77f76e17009241c5ee67992f72ec05f4c81084fbede3cc09";
    const char* ev1 = "699be3b8c6870965e5bf892ad5d2cc7b0e85a117402dfd83b7f4a947e02d978498255a2aaec0ac627b5afbdf1bf1c90";
    const char* ev2 = "8157cd83046453f5dd0972b6e3949e4288020b5b8a9cc99ca07e27089a2ce2436d965026adad3ef7baba37f2183e9b5";
    const char* ev3 = "ab1c2ffdd6c253ca155231eb3e71ba044fd562f6f72bc5bad5ec46a0b7a3b0247cf08ce6c6317f40edbc653a72dee17";
    const char* ev4 = "aa404866706722864480885d68ad0ccac1967c7544b447873cc37e0181271e006df72162a3d3e0287bf597fbf7f8fc1";
    ROOTS4[0] = FP2_ONE_;
    set_fp2_hex(&ROOTS4[1], rv1, rv1); neg_fp2_parts(&ROOTS4[1], 0, 1);   /* (rv1, -rv1) */
    ROOTS4[2] = FP2_ZERO_; fp_copy(ROOTS4[2].c1, R1);                      /* (0, 1) */
    set_fp2_hex(&ROOTS4[3], rv1, rv1);                                     /* (rv1, rv1) */
    set_fp2_hex(&ETAS[0], ev1, ev2);
    set_fp2_hex(&ETAS[1], ev2, ev1); neg_fp2_parts(&ETAS[1], 1, 0);        /* (-ev2, ev1) */
    set_fp2_hex(&ETAS[2], ev3, ev4);
    set_fp2_hex(&ETAS[3], ev4, ev3); neg_fp2_parts(&ETAS[3], 1, 0);        /* (-ev4, ev3) */
    /* SWU parameters of the 3-isogenous curve (math.ts:1221-1223): A' = 240u, B' = 1012(1+u), Z = -(2+u) */
    ISO_A = FP2_ZERO_; fp_set_u64(ISO_A.c1, 240);
    fp_set_u64(ISO_B.c0, 1012); fp_set_u64(ISO_B.c1, 1012);
    fp_set_u64(ISO_Z.c0, 2); fp_set_u64(ISO_Z.c1, 1); fp2_neg(&ISO_Z, &ISO_Z);
    /* (p^2 - 9) / 16 (math.ts:1191) */
    { big e; big_set(&e, 1); big_mul_p(&e); big_mul_p(&e);
      for (int k = 0; k < 9; ++k) big_sub1(&e);
      big_div_small(&e, 16);
      memset(EXP_P2M9_16, 0, 96);
      for (int i = 0; i < e.n * 4 && i < 96; ++i) EXP_P2M9_16[95 - i] = (uint8_t)(e.w[i / 4] >> (8 * (i % 4))); }
    /* 3-isogeny tables, leading coefficient first (math.ts:1547-1610) */
    const char* k1 = "11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71e";
    const char* k2 = "8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38d";
    const char* k3 = "11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71a";
    const char* k4 = "5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97d6";
    set_fp2_hex(&ISO3[0][0], "171d6541fa38ccfaed6dea691f5fb614cb14b4e7f4e810aa22d6108f142b85757098e38d0f671c7188e2aaaaaaaa5ed1", "0");
    set_fp2_hex(&ISO3[0][1], k1, k2);
    set_fp2_hex(&ISO3[0][2], "0", k3);
    set_fp2_hex(&ISO3[0][3], k4, k4);
    ISO3[1][0] = FP2_ZERO_;
    ISO3[1][1] = FP2_ONE_;
    fp_set_u64(ISO3[1][2].c0, 12); fp_set_u64(ISO3[1][2].c1, 12); fp_neg(ISO3[1][2].c1, ISO3[1][2].c1);   /* (12, -12) */
    ISO3[1][3] = FP2_ZERO_; fp_set_u64(ISO3[1][3].c1, 72); fp_neg(ISO3[1][3].c1, ISO3[1][3].c1);            /* (0, -72) */
    set_fp2_hex(&ISO3[2][0], "124c9ad43b6cf79bfbf7043de3811ad0761b0f37a1e26286b0e977c69aa274524e79097a56dc4bd9e1b371c71c718b10", "0");
    set_fp2_hex(&ISO3[2][1], "11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71c",
                "8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38f");
    set_fp2_hex(&ISO3[2][2], "0", "5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97be");
    set_fp2_hex(&ISO3[2][3], "1530477c7ab4113b59a4c18b076d11930f7da5d4a07f649bf54439d87d27e500fc8c25ebf8c92f6812cfc71c71c6d706",
                "1530477c7ab4113b59a4c18b076d11930f7da5d4a07f649bf54439d87d27e500fc8c25ebf8c92f6812cfc71c71c6d706");
    ISO3[3][0] = FP2_ONE_;
    fp_set_u64(ISO3[3][1].c0, 18); fp_set_u64(ISO3[3][1].c1, 18); fp_neg(ISO3[3][1].c1, ISO3[3][1].c1);   /* (18, -18) */
    ISO3[3][2] = FP2_ZERO_; fp_set_u64(ISO3[3][2].c1, 216); fp_neg(ISO3[3][2].c1, ISO3[3][2].c1);          /* (0, -216) */
    fp_set_u64(ISO3[3][3].c0, 432); fp_neg(ISO3[3][3].c0, ISO3[3][3].c0); fp_copy(ISO3[3][3].c1, ISO3[3][3].c0); /* (-432, -432) */
    /* psi constants: cx = 1 / xi^((p-1)/3), cy = 1 / xi^((p-1)/2) ; psi2 constant math.ts:1411 */
    { fp2 t; xi_pow(&t, 1, 1, 3); fp2_inv(&PSI_CX, &t); xi_pow(&t, 1, 1, 2); fp2_inv(&PSI_CY, &t); }
    fp_from_hex(PSI2_C1, "1a0111ea397fe699ec02408663d4de85aa0d857d89759ad4897d29650fb85f9b409427eb4f49fffd8bfd00000000aaac");
    { fp t; fp_set_u64(t, 1); for (int i = 0; i < 256; ++i) fp_dbl(t, t); fp_copy(TWO256, t); }
    /* G1 generator (math.ts:18-21) */
    fp_from_hex(G1_BASE_.x, "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb");
    fp_from_hex(G1_BASE_.y, "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1");
    fp_copy(G1_BASE_.z, R1);
    g_init2 = 1;
}

/* PointG2.hashToCurve(msg).toAffine() as x.c0 || x.c1 || y.c0 || y.c1 (192 B); returns 0 on success */
int oracle_hash_to_g2(const uint8_t* msg, size_t mlen, const uint8_t* dst, size_t dlen, uint8_t* out192) {
    oracle_curves_init();
    g2p h; if (hash_to_g2(&h, msg, mlen, dst, dlen)) return -1;
    fp2 x, y; g2_to_affine(&x, &y, &h);
    fp_to_bytes(out192, x.c0); fp_to_bytes(out192 + 48, x.c1); fp_to_bytes(out192 + 96, y.c0); fp_to_bytes(out192 + 144, y.c1);
    return 0;
}
/* sign(msg, sk) (index.ts:746-752): 96-byte compressed signature; returns 0, or -1 for an invalid key / SWU failure */
int oracle_sign(const uint8_t* sk32, const uint8_t* msg, size_t mlen, const uint8_t* dst, size_t dlen, uint8_t* out96) {
    oracle_curves_init();
    uint8_t k[32];
    if (normalize_priv_key(k, sk32)) return -1;
    g2p h, s; if (hash_to_g2(&h, msg, mlen, dst, dlen)) return -1;
    g2_mul_be(&s, &h, k, 32);
    g2_to_signature(out96, &s);
    return 0;
}
/* getPublicKey(sk) (index.ts:738-740): 48-byte compressed sk * G1 */
int oracle_get_public_key(const uint8_t* sk32, uint8_t* out48) {
    oracle_curves_init();
    uint8_t k[32];
    if (normalize_priv_key(k, sk32)) return -1;
    g1p p; g1_mul_be(&p, &G1_BASE_, k, 32);
    g1_to_hex_compressed(out48, &p);
    return 0;
}
/* (a * G1, b * G2) as affine wire bytes (96 B, 192 B): input generator + checker for the random-pair configuration */
int oracle_scalar_mul_bases(const uint8_t* a32, const uint8_t* b32, uint8_t* g1_96, uint8_t* g2_192) {
    oracle_curves_init();
    static const char* g2c[4] = {
        "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8",
        "13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e",
        "0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801",
        "0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be"};
    uint8_t ka[32], kb[32];
    if (normalize_priv_key(ka, a32) || normalize_priv_key(kb, b32)) return -1;
    g1p p; g1_mul_be(&p, &G1_BASE_, ka, 32);
    fp zi, x, y; fp_inv(zi, p.z); fp_mul(x, p.x, zi); fp_mul(y, p.y, zi);
    fp_to_bytes(g1_96, x); fp_to_bytes(g1_96 + 48, y);
    g2p q, s; set_fp2_hex(&q.x, g2c[0], g2c[1]); set_fp2_hex(&q.y, g2c[2], g2c[3]); q.z = FP2_ONE_;
    g2_mul_be(&s, &q, kb, 32);
    fp2 ax, ay; g2_to_affine(&ax, &ay, &s);
    fp_to_bytes(g2_192, ax.c0); fp_to_bytes(g2_192 + 48, ax.c1); fp_to_bytes(g2_192 + 96, ay.c0); fp_to_bytes(g2_192 + 144, ay.c1);
    return 0;
}

typedef struct { int kind; const uint8_t *a, *b, *msgs; const uint64_t* off; const uint8_t* dst; size_t dlen; uint8_t *o1, *o2; size_t lo, hi; int bad; } cjob_t;
static void* cworker(void* p) {
    cjob_t* j = (cjob_t*)p;
    for (size_t i = j->lo; i < j->hi; ++i) {
        int rc = 0;
        if (j->kind == 0) rc = oracle_sign(j->a + 32 * i, j->msgs + j->off[i], (size_t)(j->off[i + 1] - j->off[i]), j->dst, j->dlen, j->o1 + 96 * i);
        else if (j->kind == 1) rc = oracle_get_public_key(j->a + 32 * i, j->o1 + 48 * i);
        else if (j->kind == 2) rc = oracle_scalar_mul_bases(j->a + 32 * i, j->b + 32 * i, j->o1 + 96 * i, j->o2 + 192 * i);
        else rc = oracle_hash_to_g2(j->msgs + j->off[i], (size_t)(j->off[i + 1] - j->off[i]), j->dst, j->dlen, j->o1 + 192 * i);
        if (rc) j->bad++;
    }
    return NULL;
}
static int crun(cjob_t proto, size_t n, int threads) {
    oracle_curves_init();
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; cjob_t jobs[256];
    for (int t = 0; t < threads; ++t) { jobs[t] = proto; jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads; jobs[t].bad = 0; pthread_create(&th[t], NULL, cworker, &jobs[t]); }
    int bad = 0;
    for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); bad += jobs[t].bad; }
    return bad;
}
/* batch forms on `threads` host threads; return the number of failed items */
int oracle_sign_batch(const uint8_t* sks32, const uint8_t* msgs, const uint64_t* off, size_t n, const uint8_t* dst, size_t dlen, uint8_t* out96, int threads) {
    cjob_t j = {0, sks32, NULL, msgs, off, dst, dlen, out96, NULL, 0, 0, 0}; return crun(j, n, threads);
}
int oracle_get_public_key_batch(const uint8_t* sks32, size_t n, uint8_t* out48, int threads) {
    cjob_t j = {1, sks32, NULL, NULL, NULL, NULL, 0, out48, NULL, 0, 0, 0}; return crun(j, n, threads);
}
int oracle_scalar_mul_bases_batch(const uint8_t* a32, const uint8_t* b32, size_t n, uint8_t* g1_96, uint8_t* g2_192, int threads) {
    cjob_t j = {2, a32, b32, NULL, NULL, NULL, 0, g1_96, g2_192, 0, 0, 0}; return crun(j, n, threads);
}
int oracle_hash_to_g2_batch(const uint8_t* msgs, const uint64_t* off, size_t n, const uint8_t* dst, size_t dlen, uint8_t* out192, int threads) {
    cjob_t j = {3, NULL, NULL, msgs, off, dst, dlen, out192, NULL, 0, 0, 0}; return crun(j, n, threads);
}
