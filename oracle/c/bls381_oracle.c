/* bls381_oracle.c -- plain-C CPU restatement of the reference's pairing path.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates paulmillr/noble-bls12-381 v1.4.0 (file:line relative to /root/reference):
 *   Fp    math.ts:215-291   (here: 6 x 64-bit limbs, Montgomery form, unsigned __int128 products)
 *   Fp2   math.ts:403-550   Fp6 math.ts:554-700   Fp12 math.ts:705-885 (formulas as coded there:
 *         Karatsuba Fp2/Fp6/Fp12 multiply, complex squaring, Fp4Square cyclotomic squaring)
 *   calcPairingPrecomputes + millerLoop  math.ts:1331-1388   finalExponentiate math.ts:856-874
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 * Parity: pinned -- tests/test_oracle_c.py checks it against the reference's golden vectors
 * (tests/golden/pairing_kilic_1000.bin, pairing_kats.json) and against oracle/noble_oracle.py.
 *
 * This is an independent implementation from the CUDA engine on purpose: 64-bit limbs and the
 * reference's own (Karatsuba, eagerly reduced) formulas, versus 32-bit limbs + lazy schoolbook MACs.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t fp[6];
typedef struct { fp c0, c1; } fp2;
typedef struct { fp2 c0, c1, c2; } fp6;
typedef struct { fp6 c0, c1; } fp12;

static const uint64_t P[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                              0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t N0 = 0x89f3fffcfffcfffdULL; /* -p^-1 mod 2^64 */
static fp R1, R2; /* R mod p, R^2 mod p (computed at init) */
static fp INV2;   /* 1/2 in Montgomery form */
static int g_init = 0;

/* ---------------------------------------------------------------- Fp */
static int fp_is_zero(const fp a) { return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5]) == 0; }
static int fp_eq(const fp a, const fp b) { return memcmp(a, b, sizeof(fp)) == 0; }
static void fp_copy(fp r, const fp a) { memcpy(r, a, sizeof(fp)); }
static void fp_zero(fp r) { memset(r, 0, sizeof(fp)); }

static int geq_p(const uint64_t* a) {
    for (int i = 5; i >= 0; --i) {
        if (a[i] > P[i]) return 1;
        if (a[i] < P[i]) return 0;
    }
    return 1;
}
static void sub_p(uint64_t* a) {
    u128 b = 0;
    for (int i = 0; i < 6; ++i) {
        u128 t = (u128)a[i] - P[i] - (uint64_t)b;
        a[i] = (uint64_t)t;
        b = (t >> 64) & 1;
    }
}
static void fp_add(fp r, const fp a, const fp b) { /* math.ts:243 */
    u128 c = 0;
    uint64_t t[6];
    for (int i = 0; i < 6; ++i) {
        c += (u128)a[i] + b[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    if (c || geq_p(t)) sub_p(t);
    memcpy(r, t, sizeof(fp));
}
static void fp_sub(fp r, const fp a, const fp b) { /* math.ts:266 */
    u128 br = 0;
    uint64_t t[6];
    for (int i = 0; i < 6; ++i) {
        u128 d = (u128)a[i] - b[i] - (uint64_t)br;
        t[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 6; ++i) {
            c += (u128)t[i] + P[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r, t, sizeof(fp));
}
static void fp_neg(fp r, const fp a) { /* math.ts:235 */
    fp z;
    fp_zero(z);
    fp_sub(r, z, a);
}
static void fp_mul(fp r, const fp a, const fp b) { /* math.ts:270 (Montgomery CIOS) */
    uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; ++i) {
        u128 c = 0;
        for (int j = 0; j < 6; ++j) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[6];
        t[6] = (uint64_t)c;
        t[7] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * N0;
        c = (u128)m * P[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 6; ++j) {
            c += (u128)m * P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[6];
        t[5] = (uint64_t)c;
        t[6] = t[7] + (uint64_t)(c >> 64);
    }
    if (t[6] || geq_p(t)) sub_p(t);
    memcpy(r, t, sizeof(fp));
}
static void fp_sqr(fp r, const fp a) { fp_mul(r, a, a); } /* math.ts:247 */
static void fp_dbl(fp r, const fp a) { fp_add(r, a, a); }
static void fp_pow(fp r, const fp a, const uint64_t* e, int nlimbs) { /* math.ts:90-100 */
    fp acc, base;
    fp_copy(acc, R1);
    fp_copy(base, a);
    for (int i = 0; i < nlimbs * 64; ++i) {
        if ((e[i / 64] >> (i % 64)) & 1) fp_mul(acc, acc, base);
        fp_sqr(base, base);
    }
    fp_copy(r, acc);
}
static void fp_inv(fp r, const fp a) { /* math.ts:239 (canonical residue: a^(p-2)) */
    uint64_t e[6];
    memcpy(e, P, sizeof(e));
    e[0] -= 2;
    fp_pow(r, a, e, 6);
}
static void fp_from_bytes(fp r, const uint8_t* b) { /* 48 B big-endian -> Montgomery */
    fp t;
    for (int i = 0; i < 6; ++i) {
        uint64_t w = 0;
        for (int k = 0; k < 8; ++k) w = (w << 8) | b[(5 - i) * 8 + k];
        t[i] = w;
    }
    fp_mul(r, t, R2);
}
static void fp_to_bytes(uint8_t* b, const fp a) { /* math.ts:288-290 */
    fp one = {1, 0, 0, 0, 0, 0}, t;
    fp_mul(t, a, one);
    for (int i = 0; i < 6; ++i)
        for (int k = 0; k < 8; ++k) b[(5 - i) * 8 + k] = (uint8_t)(t[i] >> (56 - 8 * k));
}
static void fp_mul_small(fp r, const fp a, int k) { /* multiply(bigint) for k in 2,3,4,8 */
    fp t;
    fp_copy(t, a);
    fp acc;
    fp_zero(acc);
    while (k) {
        if (k & 1) fp_add(acc, acc, t);
        fp_dbl(t, t);
        k >>= 1;
    }
    fp_copy(r, acc);
}

/* ---------------------------------------------------------------- Fp2 (math.ts:403-550) */
static void fp2_add(fp2* r, const fp2* a, const fp2* b) { fp_add(r->c0, a->c0, b->c0); fp_add(r->c1, a->c1, b->c1); }
static void fp2_sub(fp2* r, const fp2* a, const fp2* b) { fp_sub(r->c0, a->c0, b->c0); fp_sub(r->c1, a->c1, b->c1); }
static void fp2_neg(fp2* r, const fp2* a) { fp_neg(r->c0, a->c0); fp_neg(r->c1, a->c1); }
static void fp2_mul(fp2* r, const fp2* a, const fp2* b) { /* :451-462 Karatsuba */
    fp t1, t2, s1, s2, m;
    fp_mul(t1, a->c0, b->c0);
    fp_mul(t2, a->c1, b->c1);
    fp_add(s1, a->c0, a->c1);
    fp_add(s2, b->c0, b->c1);
    fp_mul(m, s1, s2);
    fp_sub(r->c0, t1, t2);
    fp_add(t1, t1, t2);
    fp_sub(r->c1, m, t1);
}
static void fp2_sqr(fp2* r, const fp2* a) { /* :477-484 */
    fp s, d, c;
    fp_add(s, a->c0, a->c1);
    fp_sub(d, a->c0, a->c1);
    fp_add(c, a->c0, a->c0);
    fp_mul(r->c1, c, a->c1);
    fp_mul(r->c0, s, d);
}
static void fp2_mul_fp(fp2* r, const fp2* a, const fp k) { fp_mul(r->c0, a->c0, k); fp_mul(r->c1, a->c1, k); }
static void fp2_mul_small(fp2* r, const fp2* a, int k) { fp_mul_small(r->c0, a->c0, k); fp_mul_small(r->c1, a->c1, k); }
static void fp2_mul_xi(fp2* r, const fp2* a) { /* :471-475 */
    fp t;
    fp_sub(t, a->c0, a->c1);
    fp_add(r->c1, a->c0, a->c1);
    fp_copy(r->c0, t);
}
static void fp2_inv(fp2* r, const fp2* a) { /* :522-526 */
    fp n, t;
    fp_sqr(n, a->c0);
    fp_sqr(t, a->c1);
    fp_add(n, n, t);
    fp_inv(n, n);
    fp_mul(r->c0, a->c0, n);
    fp_neg(t, a->c1);
    fp_mul(r->c1, t, n);
}
static void fp2_mul_by_b(fp2* r, const fp2* a) { /* :532-539 */
    fp t0, t1;
    fp_mul_small(t0, a->c0, 4);
    fp_mul_small(t1, a->c1, 4);
    fp_sub(r->c0, t0, t1);
    fp_add(r->c1, t0, t1);
}
static void fp2_half(fp2* r, const fp2* a) { fp2_mul_fp(r, a, INV2); } /* div(2n) :466-469 */
static void fp2_conj(fp2* r, const fp2* a) { fp_copy(r->c0, a->c0); fp_neg(r->c1, a->c1); }

/* ---------------------------------------------------------------- Fp6 (math.ts:554-700) */
static void fp6_add(fp6* r, const fp6* a, const fp6* b) { fp2_add(&r->c0, &a->c0, &b->c0); fp2_add(&r->c1, &a->c1, &b->c1); fp2_add(&r->c2, &a->c2, &b->c2); }
static void fp6_sub(fp6* r, const fp6* a, const fp6* b) { fp2_sub(&r->c0, &a->c0, &b->c0); fp2_sub(&r->c1, &a->c1, &b->c1); fp2_sub(&r->c2, &a->c2, &b->c2); }
static void fp6_neg(fp6* r, const fp6* a) { fp2_neg(&r->c0, &a->c0); fp2_neg(&r->c1, &a->c1); fp2_neg(&r->c2, &a->c2); }
static void fp6_mul_v(fp6* r, const fp6* a) { /* :627-629 */
    fp2 t;
    fp2_mul_xi(&t, &a->c2);
    r->c2 = a->c1;
    r->c1 = a->c0;
    r->c0 = t;
}
static void fp6_mul(fp6* r, const fp6* a, const fp6* b) { /* :601-618 */
    fp2 t0, t1, t2, s, u, m, x;
    fp6 o;
    fp2_mul(&t0, &a->c0, &b->c0);
    fp2_mul(&t1, &a->c1, &b->c1);
    fp2_mul(&t2, &a->c2, &b->c2);
    fp2_add(&s, &a->c1, &a->c2); fp2_add(&u, &b->c1, &b->c2); fp2_mul(&m, &s, &u);
    fp2_add(&x, &t1, &t2); fp2_sub(&m, &m, &x); fp2_mul_xi(&m, &m); fp2_add(&o.c0, &t0, &m);
    fp2_add(&s, &a->c0, &a->c1); fp2_add(&u, &b->c0, &b->c1); fp2_mul(&m, &s, &u);
    fp2_add(&x, &t0, &t1); fp2_sub(&m, &m, &x); fp2_mul_xi(&x, &t2); fp2_add(&o.c1, &m, &x);
    fp2_add(&s, &a->c0, &a->c2); fp2_add(&u, &b->c0, &b->c2); fp2_mul(&m, &s, &u);
    fp2_add(&x, &t0, &t2); fp2_sub(&m, &m, &x); fp2_add(&o.c2, &t1, &m);
    *r = o;
}
static void fp6_mul_by_1(fp6* r, const fp6* a, const fp2* b1) { /* :631-637 */
    fp6 o;
    fp2_mul(&o.c0, &a->c2, b1); fp2_mul_xi(&o.c0, &o.c0);
    fp2_mul(&o.c1, &a->c0, b1);
    fp2_mul(&o.c2, &a->c1, b1);
    *r = o;
}
static void fp6_mul_by_01(fp6* r, const fp6* a, const fp2* b0, const fp2* b1) { /* :639-651 */
    fp2 t0, t1, s, m, u;
    fp6 o;
    fp2_mul(&t0, &a->c0, b0);
    fp2_mul(&t1, &a->c1, b1);
    fp2_add(&s, &a->c1, &a->c2); fp2_mul(&m, &s, b1); fp2_sub(&m, &m, &t1); fp2_mul_xi(&m, &m); fp2_add(&o.c0, &m, &t0);
    fp2_add(&s, b0, b1); fp2_add(&u, &a->c0, &a->c1); fp2_mul(&m, &s, &u); fp2_sub(&m, &m, &t0); fp2_sub(&o.c1, &m, &t1);
    fp2_add(&s, &a->c0, &a->c2); fp2_mul(&m, &s, b0); fp2_sub(&m, &m, &t0); fp2_add(&o.c2, &m, &t1);
    *r = o;
}
static void fp6_sqr(fp6* r, const fp6* a) { /* :658-670 */
    fp2 t0, t1, t3, t4, s, x;
    fp6 o;
    fp2_sqr(&t0, &a->c0);
    fp2_mul(&t1, &a->c0, &a->c1); fp2_mul_small(&t1, &t1, 2);
    fp2_mul(&t3, &a->c1, &a->c2); fp2_mul_small(&t3, &t3, 2);
    fp2_sqr(&t4, &a->c2);
    fp2_mul_xi(&x, &t3); fp2_add(&o.c0, &x, &t0);
    fp2_mul_xi(&x, &t4); fp2_add(&o.c1, &x, &t1);
    fp2_sub(&s, &a->c0, &a->c1); fp2_add(&s, &s, &a->c2); fp2_sqr(&s, &s);
    fp2_add(&x, &t1, &s); fp2_add(&x, &x, &t3); fp2_sub(&x, &x, &t0); fp2_sub(&o.c2, &x, &t4);
    *r = o;
}
static void fp6_inv(fp6* r, const fp6* a) { /* :672-680 */
    fp2 t0, t1, t2, t4, x, y;
    fp2_sqr(&t0, &a->c0); fp2_mul(&x, &a->c2, &a->c1); fp2_mul_xi(&x, &x); fp2_sub(&t0, &t0, &x);
    fp2_sqr(&t1, &a->c2); fp2_mul_xi(&t1, &t1); fp2_mul(&x, &a->c0, &a->c1); fp2_sub(&t1, &t1, &x);
    fp2_sqr(&t2, &a->c1); fp2_mul(&x, &a->c0, &a->c2); fp2_sub(&t2, &t2, &x);
    fp2_mul(&x, &a->c2, &t1); fp2_mul(&y, &a->c1, &t2); fp2_add(&x, &x, &y); fp2_mul_xi(&x, &x);
    fp2_mul(&y, &a->c0, &t0); fp2_add(&x, &x, &y);
    fp2_inv(&t4, &x);
    fp2_mul(&r->c0, &t4, &t0); fp2_mul(&r->c1, &t4, &t1); fp2_mul(&r->c2, &t4, &t2);
}

/* ---------------------------------------------------------------- Frobenius tables (math.ts:1428-1543) */
static fp2 FROB6_C1[6], FROB6_C2[6], FROB12[12];
static void fp2_pow_big(fp2* r, const fp2* a, const uint8_t* e_be, int nbytes) {
    fp2 acc, base = *a;
    memset(&acc, 0, sizeof(acc));
    fp_copy(acc.c0, R1);
    for (int i = nbytes - 1; i >= 0; --i)
        for (int b = 0; b < 8; ++b) {
            if ((e_be[i] >> b) & 1) fp2_mul(&acc, &acc, &base);
            fp2_sqr(&base, &base);
        }
    *r = acc;
}

/* ---------------------------------------------------------------- Fp12 (math.ts:705-885) */
static void fp12_mul(fp12* r, const fp12* a, const fp12* b) { /* :748-759 */
    fp6 t1, t2, s, u, m, x;
    fp12 o;
    fp6_mul(&t1, &a->c0, &b->c0);
    fp6_mul(&t2, &a->c1, &b->c1);
    fp6_mul_v(&x, &t2); fp6_add(&o.c0, &t1, &x);
    fp6_add(&s, &a->c0, &a->c1); fp6_add(&u, &b->c0, &b->c1); fp6_mul(&m, &s, &u);
    fp6_add(&x, &t1, &t2); fp6_sub(&o.c1, &m, &x);
    *r = o;
}
static void fp12_mul_by_014(fp12* r, const fp12* a, const fp2* o0, const fp2* o1, const fp2* o4) { /* :768-777 */
    fp6 t0, t1, s, x;
    fp2 o14;
    fp12 o;
    fp6_mul_by_01(&t0, &a->c0, o0, o1);
    fp6_mul_by_1(&t1, &a->c1, o4);
    fp6_mul_v(&x, &t1); fp6_add(&o.c0, &x, &t0);
    fp6_add(&s, &a->c1, &a->c0); fp2_add(&o14, o1, o4);
    fp6_mul_by_01(&x, &s, o0, &o14); fp6_sub(&x, &x, &t0); fp6_sub(&o.c1, &x, &t1);
    *r = o;
}
static void fp12_sqr(fp12* r, const fp12* a) { /* :783-791 */
    fp6 ab, s, u, x;
    fp12 o;
    fp6_mul(&ab, &a->c0, &a->c1);
    fp6_mul_v(&s, &a->c1); fp6_add(&s, &s, &a->c0);
    fp6_add(&u, &a->c0, &a->c1);
    fp6_mul(&x, &s, &u); fp6_sub(&x, &x, &ab); fp6_mul_v(&s, &ab); fp6_sub(&o.c0, &x, &s);
    fp6_add(&o.c1, &ab, &ab);
    *r = o;
}
static void fp12_conj(fp12* r, const fp12* a) { r->c0 = a->c0; fp6_neg(&r->c1, &a->c1); } /* :799-801 */
static void fp12_inv(fp12* r, const fp12* a) { /* :793-797 */
    fp6 t, x;
    fp6_sqr(&t, &a->c0); fp6_sqr(&x, &a->c1); fp6_mul_v(&x, &x); fp6_sub(&t, &t, &x);
    fp6_inv(&t, &t);
    fp6_mul(&r->c0, &a->c0, &t);
    fp6_mul(&x, &a->c1, &t); fp6_neg(&r->c1, &x);
}
static void fp2_frob(fp2* r, const fp2* a, int power) { if (power & 1) fp2_conj(r, a); else *r = *a; } /* :529-531 */
static void fp6_frob(fp6* r, const fp6* a, int power) { /* :682-688 */
    fp2 t;
    fp2_frob(&r->c0, &a->c0, power);
    fp2_frob(&t, &a->c1, power); fp2_mul(&r->c1, &t, &FROB6_C1[power % 6]);
    fp2_frob(&t, &a->c2, power); fp2_mul(&r->c2, &t, &FROB6_C2[power % 6]);
}
static void fp12_frob(fp12* r, const fp12* a, int power) { /* :804-809 */
    fp6 t;
    fp6_frob(&r->c0, &a->c0, power);
    fp6_frob(&t, &a->c1, power);
    fp2_mul(&r->c1.c0, &t.c0, &FROB12[power % 12]);
    fp2_mul(&r->c1.c1, &t.c1, &FROB12[power % 12]);
    fp2_mul(&r->c1.c2, &t.c2, &FROB12[power % 12]);
}
static void fp4_square(fp2* first, fp2* second, const fp2* a, const fp2* b) { /* :811-818 */
    fp2 a2, b2, s;
    fp2_sqr(&a2, a);
    fp2_sqr(&b2, b);
    fp2_add(&s, a, b); fp2_sqr(&s, &s); fp2_sub(&s, &s, &a2); fp2_sub(second, &s, &b2);
    fp2_mul_xi(&s, &b2); fp2_add(first, &s, &a2);
}
static void cyc_out(fp2* r, const fp2* t, const fp2* c, int plus) { /* 2*(t -/+ c) + t */
    fp2 x;
    if (plus) fp2_add(&x, t, c); else fp2_sub(&x, t, c);
    fp2_mul_small(&x, &x, 2);
    fp2_add(r, &x, t);
}
static void fp12_cyc_sqr(fp12* r, const fp12* f) { /* :824-843 */
    fp2 t3, t4, t5, t6, t7, t8, t9;
    fp12 o;
    fp4_square(&t3, &t4, &f->c0.c0, &f->c1.c1);
    fp4_square(&t5, &t6, &f->c1.c0, &f->c0.c2);
    fp4_square(&t7, &t8, &f->c0.c1, &f->c1.c2);
    fp2_mul_xi(&t9, &t8);
    cyc_out(&o.c0.c0, &t3, &f->c0.c0, 0);
    cyc_out(&o.c0.c1, &t5, &f->c0.c1, 0);
    cyc_out(&o.c0.c2, &t7, &f->c0.c2, 0);
    cyc_out(&o.c1.c0, &t9, &f->c1.c0, 1);
    cyc_out(&o.c1.c1, &t4, &f->c1.c1, 1);
    cyc_out(&o.c1.c2, &t6, &f->c1.c2, 1);
    *r = o;
}
#define X_PARAM 0xd201000000010000ULL
static void fp12_one(fp12* r) { memset(r, 0, sizeof(*r)); fp_copy(r->c0.c0.c0, R1); }
static void fp12_cyc_exp(fp12* r, const fp12* f) { /* :845-852 */
    fp12 z;
    fp12_one(&z);
    for (int i = 63; i >= 0; --i) {
        fp12_cyc_sqr(&z, &z);
        if ((X_PARAM >> i) & 1) fp12_mul(&z, &z, f);
    }
    *r = z;
}
static void fp12_final_exp(fp12* r, const fp12* f) { /* :856-874 */
    fp12 t0, t1, t2, t3, t4, t5, t6, t7, a, b, c, d, x;
    fp12_frob(&a, f, 6); fp12_inv(&x, f); fp12_mul(&t0, &a, &x);
    fp12_frob(&a, &t0, 2); fp12_mul(&t1, &a, &t0);
    fp12_cyc_exp(&a, &t1); fp12_conj(&t2, &a);
    fp12_cyc_sqr(&a, &t1); fp12_conj(&a, &a); fp12_mul(&t3, &a, &t2);
    fp12_cyc_exp(&a, &t3); fp12_conj(&t4, &a);
    fp12_cyc_exp(&a, &t4); fp12_conj(&t5, &a);
    fp12_cyc_exp(&a, &t5); fp12_conj(&a, &a); fp12_cyc_sqr(&x, &t2); fp12_mul(&t6, &a, &x);
    fp12_cyc_exp(&a, &t6); fp12_conj(&t7, &a);
    fp12_mul(&x, &t2, &t5); fp12_frob(&a, &x, 2);
    fp12_mul(&x, &t4, &t1); fp12_frob(&b, &x, 3);
    fp12_conj(&x, &t1); fp12_mul(&x, &t6, &x); fp12_frob(&c, &x, 1);
    fp12_conj(&x, &t3); fp12_mul(&x, &t7, &x); fp12_mul(&d, &x, &t1);
    fp12_mul(&x, &a, &b); fp12_mul(&x, &x, &c); fp12_mul(r, &x, &d);
}

/* ---------------------------------------------------------------- Miller loop (math.ts:1331-1388) */
static void miller(fp12* out, const fp Px, const fp Py, const fp2* Qx, const fp2* Qy) {
    fp2 Rx = *Qx, Ry = *Qy, Rz, t0, t1, t2, t3, t4, e0, e1, e2, x, y;
    fp12 f;
    memset(&Rz, 0, sizeof(Rz));
    fp_copy(Rz.c0, R1);
    fp12_one(&f);
    for (int i = 62; i >= 0; --i) {
        fp2_sqr(&t0, &Ry);
        fp2_sqr(&t1, &Rz);
        fp2_mul_small(&x, &t1, 3); fp2_mul_by_b(&t2, &x);
        fp2_mul_small(&t3, &t2, 3);
        fp2_add(&x, &Ry, &Rz); fp2_sqr(&x, &x); fp2_sub(&x, &x, &t1); fp2_sub(&t4, &x, &t0);
        fp2_sub(&e0, &t2, &t0);
        fp2_sqr(&x, &Rx); fp2_mul_small(&e1, &x, 3);
        fp2_neg(&e2, &t4);
        fp2_sub(&x, &t0, &t3); fp2_mul(&x, &x, &Rx); fp2_mul(&x, &x, &Ry); fp2_half(&x, &x);
        fp2_add(&y, &t0, &t3); fp2_half(&y, &y); fp2_sqr(&y, &y);
        fp2 z; fp2_sqr(&z, &t2); fp2_mul_small(&z, &z, 3); fp2_sub(&Ry, &y, &z);
        Rx = x;
        fp2_mul(&Rz, &t0, &t4);
        fp2_mul_fp(&e1, &e1, Px); fp2_mul_fp(&e2, &e2, Py);
        fp12_mul_by_014(&f, &f, &e0, &e1, &e2);
        if ((X_PARAM >> i) & 1) {
            fp2_mul(&x, Qy, &Rz); fp2_sub(&t0, &Ry, &x);
            fp2_mul(&x, Qx, &Rz); fp2_sub(&t1, &Rx, &x);
            fp2_mul(&x, &t0, Qx); fp2_mul(&y, &t1, Qy); fp2_sub(&e0, &x, &y);
            fp2_neg(&e1, &t0);
            e2 = t1;
            fp2_sqr(&t2, &t1);
            fp2_mul(&t3, &t2, &t1);
            fp2_mul(&t4, &t2, &Rx);
            fp2 t5; fp2_mul_small(&x, &t4, 2); fp2_sub(&t5, &t3, &x); fp2_sqr(&x, &t0); fp2_mul(&x, &x, &Rz); fp2_add(&t5, &t5, &x);
            fp2_mul(&Rx, &t1, &t5);
            fp2_sub(&x, &t4, &t5); fp2_mul(&x, &x, &t0); fp2_mul(&y, &t3, &Ry); fp2_sub(&Ry, &x, &y);
            fp2_mul(&Rz, &Rz, &t3);
            fp2_mul_fp(&e1, &e1, Px); fp2_mul_fp(&e2, &e2, Py);
            fp12_mul_by_014(&f, &f, &e0, &e1, &e2);
        }
        if (i != 0) fp12_sqr(&f, &f);
    }
    fp12_conj(out, &f);
}

/* ---------------------------------------------------------------- init + wire format */
static void fp12_from_bytes(fp12* f, const uint8_t* b) {
    fp2* c[6] = {&f->c0.c0, &f->c0.c1, &f->c0.c2, &f->c1.c0, &f->c1.c1, &f->c1.c2};
    for (int i = 0; i < 6; ++i) { fp_from_bytes(c[i]->c0, b + 96 * i); fp_from_bytes(c[i]->c1, b + 96 * i + 48); }
}
static void fp12_to_bytes(uint8_t* b, const fp12* f) {
    const fp2* c[6] = {&f->c0.c0, &f->c0.c1, &f->c0.c2, &f->c1.c0, &f->c1.c1, &f->c1.c2};
    for (int i = 0; i < 6; ++i) { fp_to_bytes(b + 96 * i, c[i]->c0); fp_to_bytes(b + 96 * i + 48, c[i]->c1); }
}

/* big-endian byte-string exponent (p^k - 1)/d computed with schoolbook big arithmetic */
typedef struct { uint32_t w[160]; int n; } big;
static void big_set(big* a, uint32_t v) { memset(a, 0, sizeof(*a)); a->w[0] = v; a->n = 1; }
static void big_mul_p(big* a) {
    big r; memset(&r, 0, sizeof(r));
    for (int i = 0; i < a->n; ++i) {
        uint64_t c = 0;
        for (int j = 0; j < 12; ++j) {
            uint32_t pj = (uint32_t)(P[j / 2] >> (32 * (j & 1)));
            c += (uint64_t)a->w[i] * pj + r.w[i + j];
            r.w[i + j] = (uint32_t)c; c >>= 32;
        }
        int k = i + 12;
        while (c) { c += r.w[k]; r.w[k] = (uint32_t)c; c >>= 32; ++k; }
    }
    r.n = a->n + 12;
    while (r.n > 1 && r.w[r.n - 1] == 0) --r.n;
    *a = r;
}
static void big_sub1(big* a) { int i = 0; while (a->w[i] == 0) { a->w[i] = 0xffffffffu; ++i; } a->w[i]--; }
static void big_div_small(big* a, uint32_t d) { uint64_t r = 0; for (int i = a->n - 1; i >= 0; --i) { r = (r << 32) | a->w[i]; a->w[i] = (uint32_t)(r / d); r %= d; } }
static void big_mul_small(big* a, uint32_t m) { uint64_t c = 0; for (int i = 0; i < a->n; ++i) { c += (uint64_t)a->w[i] * m; a->w[i] = (uint32_t)c; c >>= 32; } if (c) a->w[a->n++] = (uint32_t)c; }
static void xi_pow(fp2* r, int k, uint32_t num, uint32_t den) { /* xi^(num*(p^k-1)/den) */
    big e; big_set(&e, 1);
    for (int i = 0; i < k; ++i) big_mul_p(&e);
    big_sub1(&e); big_div_small(&e, den); big_mul_small(&e, num);
    uint8_t bytes[640];
    int nb = e.n * 4;
    for (int i = 0; i < nb; ++i) bytes[nb - 1 - i] = (uint8_t)(e.w[i / 4] >> (8 * (i % 4)));
    fp2 xi; fp_copy(xi.c0, R1); fp_copy(xi.c1, R1);
    fp2_pow_big(r, &xi, bytes, nb);
}

void oracle_init(void) {
    if (g_init) return;
    /* R mod p by doubling 1 384 times; R^2 by doubling 384 more */
    fp t = {1, 0, 0, 0, 0, 0};
    for (int i = 0; i < 384; ++i) fp_dbl(t, t);
    fp_copy(R1, t);
    for (int i = 0; i < 384; ++i) fp_dbl(t, t);
    fp_copy(R2, t);
    fp two; fp_dbl(two, R1); fp_inv(INV2, two);
    for (int k = 0; k < 6; ++k) { xi_pow(&FROB6_C1[k], k, 1, 3); xi_pow(&FROB6_C2[k], k, 2, 3); }
    for (int k = 0; k < 12; ++k) xi_pow(&FROB12[k], k, 1, 6);
    g_init = 1;
}

/* e(P, Q) for affine wire-format inputs (96 B, 192 B) -> 576 B; with_fe: apply finalExponentiate */
void oracle_pairing(const uint8_t* g1, const uint8_t* g2, int with_fe, uint8_t* out) {
    fp Px, Py; fp2 Qx, Qy; fp12 f;
    oracle_init();
    fp_from_bytes(Px, g1); fp_from_bytes(Py, g1 + 48);
    fp_from_bytes(Qx.c0, g2); fp_from_bytes(Qx.c1, g2 + 48); fp_from_bytes(Qy.c0, g2 + 96); fp_from_bytes(Qy.c1, g2 + 144);
    miller(&f, Px, Py, &Qx, &Qy);
    if (with_fe) fp12_final_exp(&f, &f);
    fp12_to_bytes(out, &f);
}
void oracle_final_exp(const uint8_t* in, uint8_t* out) {
    fp12 f; oracle_init(); fp12_from_bytes(&f, in); fp12_final_exp(&f, &f); fp12_to_bytes(out, &f);
}
void oracle_fp12_mul(const uint8_t* a, const uint8_t* b, uint8_t* out) {
    fp12 x, y; oracle_init(); fp12_from_bytes(&x, a); fp12_from_bytes(&y, b); fp12_mul(&x, &x, &y); fp12_to_bytes(out, &x);
}

typedef struct { const uint8_t *g1, *g2; uint8_t* out; size_t lo, hi; int fe; } job_t;
static void* worker(void* p) {
    job_t* j = (job_t*)p;
    for (size_t i = j->lo; i < j->hi; ++i) oracle_pairing(j->g1 + 96 * i, j->g2 + 192 * i, j->fe, j->out + 576 * i);
    return NULL;
}
/* n pairings on `threads` host threads */
void oracle_pairing_batch(const uint8_t* g1, const uint8_t* g2, size_t n, int with_fe, uint8_t* out, int threads) {
    oracle_init();
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; job_t jobs[256];
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (job_t){g1, g2, out, n * t / threads, n * (t + 1) / threads, with_fe};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}
