// hash_to_field + map_to_curve_simple_swu_9mod16 for G2 as ONE hand-written kernel (a thread per field element u).
//
// Replaces, for the hash-to-curve front of verifyBatch / sign / PointG2.hashToCurve, the part of the reference that is a
// long SERIAL chain per item: hash_to_field's reduction of the 64-byte chunks (index.ts:240-267) and the simplified SWU
// map with its square root (math.ts:1195-1267).  In the tower-VM these chains are one multiply per record on one warp
// (profiles/r2_notes.md: the chains bound the schedule of hash_to_g2); here they stay in registers and the kernel is bound
// by the multiplier.  The two points of E' per message then go to the tower-VM program `h2g2_tail` (point addition,
// 3-isogeny, clearCofactor, affine) or `sign_tail`.
//
// Square root: the reference computes sqrt(u/v) (or sqrt(Z^3 t^6 u/v) when u/v is not a square) with one 758-bit Fp2
// exponentiation and eight candidate checks.  Here: the "complex method" on W = u v (u/v is a square iff W is) -- two Fp
// exponentiations a^((p-3)/4) serve both cases:
//     n = norm(W), e = n^((p-3)/4), delta = n e;   delta^2 = n  <=>  W is a square in Fp2.
//     otherwise delta^2 = -n and W' = Z^3 W IS a square with norm(W') = norm(Z)^3 n = (c delta)^2, c = sqrt(-norm(Z)^3).
//     gamma = (a0 + delta)/2 (gamma = 0 only for a = a0 a non-residue of Fp: then gamma := a0), t = gamma^((p-3)/4),
//     s = gamma t:  s^2 = gamma: root r = s + (a1 t/2) i;   s^2 = -gamma: r = -(a1 t/2) + s i.
//     1/r = conj(r)/norm(r), norm(r) = +-delta, 1/delta = +-e:  sqrt(u/v) = u/r = +-u conj(r) e.
// The sign of y is then fixed by sgn0 exactly as in math.ts:1262-1263, so the point is the reference's point, bit for bit
// (tests: RFC 9380 / kilic vectors, the 559 sign KATs, random messages against the oracle).
//
// The same source compiles for the host (tests/emu) so that the logic is checked without a GPU.
#pragma once
#include "fp_core.cuh"
#include "swu_g2_gen.cuh"

namespace swu {

#if defined(__CUDACC__)
#define SWU_FN __device__ __noinline__
#define SWU_INL __device__ __forceinline__
#else
#define SWU_FN static
#define SWU_INL static inline
#endif

#if !defined(__CUDACC__)
// host build only (tests/emu): how many 384-bit products / Montgomery reductions one call executes -- bench.py's issued
// multiply count of these kernels is pinned to these counters by tests/test_vm_ingest_emu.py
static long g_products = 0, g_reductions = 0;
#define SWU_COUNT(p, r) do { g_products += (p); g_reductions += (r); } while (0)
#else
#define SWU_COUNT(p, r) do { } while (0)
#endif

struct Fe { uint32_t v[12]; };
struct Fe2 { Fe c0, c1; };

FPC_DEV void fe_set(Fe& r, const uint32_t* k) { fpc::copy12(r.v, k); }
FPC_DEV void fe_zero(Fe& r) { fpc::zero12(r.v); }
FPC_DEV bool fe_is_zero(const Fe& a) {
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) any |= a.v[k];
    return any == 0;
}
FPC_DEV bool fe_eq(const Fe& a, const Fe& b) {
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) any |= a.v[k] ^ b.v[k];
    return any == 0;
}
FPC_DEV void fe_sel(Fe& r, bool c, const Fe& a, const Fe& b) {
#pragma unroll
    for (int k = 0; k < 12; ++k) r.v[k] = c ? a.v[k] : b.v[k];
}
// Small functions are calls on purpose: the kernels of this file and of g2_kernels.cuh are long straight-line sequences of
// field operations, and what decides their speed is whether the hot code stays in the instruction cache (measured: with
// four inlined multiply-accumulate bodies the tail kernel ran at 43 % of the multiplier, with one at 71 %).
SWU_FN void fe_add(Fe& r, const Fe& a, const Fe& b) {
    Fe t;
    fpc::add_mod(t.v, a.v, b.v);
    r = t;
}
SWU_FN void fe_sub(Fe& r, const Fe& a, const Fe& b) {
    Fe t;
    fpc::sub_mod(t.v, a.v, b.v);
    r = t;
}
SWU_INL void fe_neg(Fe& r, const Fe& a) {
    Fe z;
    fe_zero(z);
    fe_sub(r, z, a);
}

// THE multiplier: r = sum_{k < n} x_k * y_k / R mod p, canonical, ONE reduction (n <= 4; every operand <= 2p and the sum of
// the products below 4p^2, so the reduced value is below 1.41p).  One multiply-accumulate body and one reduction body for
// all callers.
SWU_FN void fe_dotn(Fe& r, int n, const uint32_t* const* xs, const uint32_t* const* ys) {
    fpc::Acc A;
    fpc::acc_zero(A);
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        uint32_t x[12], y[12];
        fpc::copy12(x, xs[k]);
        fpc::copy12(y, ys[k]);
        fpc::acc_mac(A, x, y);
    }
    Fe t;
    fpc::acc_redc(A, t.v);
    fpc::correct(t.v, 1);
    r = t;
    SWU_COUNT(n, 1);
}
SWU_INL void fe_mul(Fe& r, const Fe& a, const Fe& b) {
    const uint32_t* xs[1] = {a.v};
    const uint32_t* ys[1] = {b.v};
    fe_dotn(r, 1, xs, ys);
}
// r = a0*b0 + a1*b1
SWU_INL void fe_dot2(Fe& r, const uint32_t* a0, const uint32_t* b0, const uint32_t* a1, const uint32_t* b1) {
    const uint32_t* xs[2] = {a0, a1};
    const uint32_t* ys[2] = {b0, b1};
    fe_dotn(r, 2, xs, ys);
}

// r = x0*y0 + x1*y1 + x2*y2 + x3*y3 (one reduction); operands <= p
SWU_INL void fe_dot4(Fe& r, const uint32_t* x0, const uint32_t* y0, const uint32_t* x1, const uint32_t* y1, const uint32_t* x2,
                     const uint32_t* y2, const uint32_t* x3, const uint32_t* y3) {
    const uint32_t* xs[4] = {x0, x1, x2, x3};
    const uint32_t* ys[4] = {y0, y1, y2, y3};
    fe_dotn(r, 4, xs, ys);
}

SWU_INL void fe2_add(Fe2& r, const Fe2& a, const Fe2& b) { fe_add(r.c0, a.c0, b.c0); fe_add(r.c1, a.c1, b.c1); }
SWU_INL void fe2_neg(Fe2& r, const Fe2& a) { fe_neg(r.c0, a.c0); fe_neg(r.c1, a.c1); }
FPC_DEV void fe2_sel(Fe2& r, bool c, const Fe2& a, const Fe2& b) { fe_sel(r.c0, c, a.c0, b.c0); fe_sel(r.c1, c, a.c1, b.c1); }
FPC_DEV bool fe2_is_zero(const Fe2& a) { return fe_is_zero(a.c0) && fe_is_zero(a.c1); }
FPC_DEV void fe2_const(Fe2& r, const uint32_t (*k)[12]) { fe_set(r.c0, k[0]); fe_set(r.c1, k[1]); }

// (a0 + a1 i)(b0 + b1 i) = (a0 b0 - a1 b1) + (a0 b1 + a1 b0) i      (math.ts:451-462 without the Karatsuba step)
SWU_FN void fe2_mul(Fe2& r, const Fe2& a, const Fe2& b) {
    uint32_t n1[12];
    fpc::neg_raw(n1, a.c1.v);  // p - a1 in [1, p]
    Fe t0, t1;
    fe_dot2(t0, a.c0.v, b.c0.v, n1, b.c1.v);
    fe_dot2(t1, a.c0.v, b.c1.v, a.c1.v, b.c0.v);
    r.c0 = t0;
    r.c1 = t1;
}
// r = a b + c d  (sub = false)  or  a b - c d  (sub = true): eight products, ONE reduction per coefficient
SWU_FN void fe2_mul2(Fe2& r, const Fe2& a, const Fe2& b, const Fe2& c, const Fe2& d, bool sub) {
    uint32_t na1[12], nc0[12], nc1[12];
    fpc::neg_raw(na1, a.c1.v);
    fpc::neg_raw(nc0, c.c0.v);
    fpc::neg_raw(nc1, c.c1.v);
    Fe t0, t1;
    // real: a0 b0 - a1 b1 +- (c0 d0 - c1 d1);   imaginary: a0 b1 + a1 b0 +- (c0 d1 + c1 d0)
    fe_dot4(t0, a.c0.v, b.c0.v, na1, b.c1.v, sub ? nc0 : c.c0.v, d.c0.v, sub ? c.c1.v : nc1, d.c1.v);
    fe_dot4(t1, a.c0.v, b.c1.v, a.c1.v, b.c0.v, sub ? nc0 : c.c0.v, d.c1.v, sub ? nc1 : c.c1.v, d.c0.v);
    r.c0 = t0;
    r.c1 = t1;
}
// (a0 + a1 i)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 i: two products instead of four   (math.ts:477-484)
SWU_FN void fe2_sqr(Fe2& r, const Fe2& a) {
    uint32_t s[12], d[12], n1[12], a2[12];
    (void)fpc::add12(s, a.c0.v, a.c1.v);        // < 2p
    fpc::neg_raw(n1, a.c1.v);
    (void)fpc::add12(d, a.c0.v, n1);            // a0 + (p - a1) <= 2p
    (void)fpc::add12(a2, a.c0.v, a.c0.v);       // 2 a0 < 2p
    Fe t0, t1;
    {
        const uint32_t* xs[1] = {s};
        const uint32_t* ys[1] = {d};
        fe_dotn(t0, 1, xs, ys);
    }
    {
        const uint32_t* xs[1] = {a2};
        const uint32_t* ys[1] = {a.c1.v};
        fe_dotn(t1, 1, xs, ys);
    }
    r.c0 = t0;
    r.c1 = t1;
}
SWU_INL void fe2_mul_fe(Fe2& r, const Fe2& a, const Fe& k) { fe_mul(r.c0, a.c0, k); fe_mul(r.c1, a.c1, k); }

// a^((p-3)/4): sliding windows over the odd powers; everything but the table look-ups stays in registers
SWU_FN void fe_pow_p34(Fe& r, const Fe& a) {
    Fe tab[8];
    Fe a2;
    fe_mul(a2, a, a);
    tab[0] = a;
#pragma unroll 1
    for (int k = 1; k < 8; ++k) fe_mul(tab[k], tab[k - 1], a2);
    Fe acc = tab[kPowIdx[0]];
#pragma unroll 1
    for (int w = 1; w <= kPowWindows; ++w) {   // the last round only runs the kPowTail squarings
        const int nsq = w < kPowWindows ? kPowSq[w] : kPowTail;
#pragma unroll 1
        for (int s = 0; s < nsq; ++s) {        // the hot loop: the ONE inlined product of this function, in registers
            Fe t;
            fpc::mont_mul(t.v, acc.v, acc.v);
            acc = t;
            SWU_COUNT(1, 1);
        }
        if (w < kPowWindows) fe_mul(acc, acc, tab[kPowIdx[w]]);
    }
    r = acc;
}

// 64 big-endian bytes mod p, Montgomery form   (hash_to_field, index.ts:253-263: os2ip(tv) mod p)
SWU_FN void fe_from_be64(Fe& r, const uint8_t* p) {
    Fe hi, lo, c;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint8_t* q = p + 4 * (7 - k);
        hi.v[k] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
        q += 32;
        lo.v[k] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
#pragma unroll
    for (int k = 8; k < 12; ++k) hi.v[k] = lo.v[k] = 0;
    fe_set(c, k2_256R2);
    fe_mul(hi, hi, c);
    fe_set(c, kR2);
    fe_mul(lo, lo, c);
    fe_add(r, hi, lo);
}

// Montgomery residue as 48 big-endian bytes (read back by the tower-VM with inp_bytes(.., montgomery=False))
FPC_DEV void fe_store_be(uint8_t* p, const Fe& a) {
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const uint32_t w = a.v[11 - k];
        p[4 * k] = (uint8_t)(w >> 24); p[4 * k + 1] = (uint8_t)(w >> 16); p[4 * k + 2] = (uint8_t)(w >> 8); p[4 * k + 3] = (uint8_t)w;
    }
}

// sgn0 of an Fp2 element (math.ts:1179-1185) -- parity of the PLAIN integers, so leave Montgomery form first
SWU_FN bool fe2_sgn0(const Fe2& a) {
    Fe one_plain, x0, x1;
    fe_zero(one_plain);
    one_plain.v[0] = 1;
    fe_mul(x0, a.c0, one_plain);
    fe_mul(x1, a.c1, one_plain);
    const bool s0 = x0.v[0] & 1u, z0 = fe_is_zero(x0), s1 = x1.v[0] & 1u;
    return s0 || (z0 && s1);
}

// map_to_curve_simple_swu_9mod16 (math.ts:1220-1267): t -> point (X : Y : Z) of E' with x = X/Z, y = Y/Z
SWU_FN void map_to_curve_g2(Fe2& X, Fe2& Y, Fe2& Zc, const Fe2& tt) {
    Fe2 cA, cB, cZ, t2, z_t2, ztzt, den, num, tmp, d2, v, u, W;
    fe2_const(cA, kA);
    fe2_const(cB, kB);
    fe2_const(cZ, kZ);
    fe2_sqr(t2, tt);
    fe2_mul(z_t2, t2, cZ);
    fe2_sqr(tmp, z_t2);
    fe2_add(ztzt, z_t2, tmp);
    fe2_mul(den, ztzt, cA);
    fe2_neg(den, den);                         // denominator = -A (Z t^2 + Z^2 t^4)
    Fe one;
    fe_set(one, kOne);
    tmp = ztzt;
    fe_add(tmp.c0, tmp.c0, one);
    fe2_mul(num, tmp, cB);                     // numerator = B (Z t^2 + Z^2 t^4 + 1)
    {
        Fe2 za;
        fe2_const(za, kZA);
        fe2_sel(den, fe2_is_zero(den), za, den);   // math.ts:1231: exceptional case -> Z A
    }
    fe2_sqr(d2, den);
    fe2_mul(v, d2, den);                       // v = D^3
    {
        Fe2 n2, n3, nd, and_, bv;
        fe2_sqr(n2, num);
        fe2_mul(n3, n2, num);
        fe2_mul(nd, num, d2);
        fe2_mul(and_, nd, cA);
        fe2_mul(bv, v, cB);
        fe2_add(u, n3, and_);
        fe2_add(u, u, bv);                     // u = N^3 + A N D^2 + B D^3
    }
    fe2_mul(W, u, v);
    Fe n, e, delta, dd;
    fe_dot2(n, W.c0.v, W.c0.v, W.c1.v, W.c1.v);
    fe_pow_p34(e, n);
    fe_mul(delta, n, e);
    fe_mul(dd, delta, delta);
    const bool success = fe_eq(dd, n);         // u/v is a square
    Fe2 Wsel, usel;
    Fe dsel, esel;
    {
        Fe2 z3, W2, u2;
        Fe k, d2_, e2_;
        fe2_const(z3, kZ3);
        fe2_mul(W2, W, z3);
        fe2_mul(u2, u, z3);
        fe_set(k, kCc);
        fe_mul(d2_, delta, k);
        fe_set(k, kCcInv);
        fe_mul(e2_, e, k);
        fe2_sel(Wsel, success, W, W2);
        fe2_sel(usel, success, u, u2);
        fe_sel(dsel, success, delta, d2_);
        fe_sel(esel, success, e, e2_);
    }
    Fe inv2, gamma, a1h, tpow, s, w, ss;
    fe_set(inv2, kInv2);
    fe_add(gamma, Wsel.c0, dsel);
    fe_mul(gamma, gamma, inv2);
    fe_sel(gamma, fe_is_zero(gamma), Wsel.c0, gamma);
    fe_mul(a1h, Wsel.c1, inv2);
    fe_pow_p34(tpow, gamma);
    fe_mul(s, gamma, tpow);
    fe_mul(w, a1h, tpow);
    fe_mul(ss, s, s);
    const bool is_qr = fe_eq(ss, gamma);
    Fe2 rc;  // conj(root)
    {
        Fe nw, ns;
        fe_neg(nw, w);
        fe_neg(ns, s);
        fe_sel(rc.c0, is_qr, s, nw);           // root = (is_qr ? s : -w) + (is_qr ? w : s) i
        fe_sel(rc.c1, is_qr, nw, ns);          // conj: imaginary part negated
    }
    Fe2 ue, y;
    fe2_mul_fe(ue, usel, esel);
    fe2_mul(y, ue, rc);
    {
        Fe2 t3, yt, nz;
        fe2_mul(t3, t2, tt);
        fe2_mul(yt, y, t3);
        fe2_sel(y, success, y, yt);            // second family: y = t^3 sqrt(Z^3 u/v)
        fe2_mul(nz, num, z_t2);
        fe2_sel(num, success, num, nz);        // x = Z t^2 x1
    }
    {
        Fe2 ny;
        fe2_neg(ny, y);
        fe2_sel(y, fe2_sgn0(tt) != fe2_sgn0(y), ny, y);   // math.ts:1262-1263
    }
    X = num;
    fe2_mul(Y, y, den);
    Zc = den;
}

// one field element: 128 uniform bytes -> 288 bytes (X.c0, X.c1, Y.c0, Y.c1, Z.c0, Z.c1; Montgomery residues, big-endian)
SWU_FN void swu_g2_one(const uint8_t* in128, uint8_t* out288) {
    Fe2 tt, X, Y, Zc;
    fe_from_be64(tt.c0, in128);
    fe_from_be64(tt.c1, in128 + 64);
    map_to_curve_g2(X, Y, Zc, tt);
    fe_store_be(out288, X.c0);
    fe_store_be(out288 + 48, X.c1);
    fe_store_be(out288 + 96, Y.c0);
    fe_store_be(out288 + 144, Y.c1);
    fe_store_be(out288 + 192, Zc.c0);
    fe_store_be(out288 + 240, Zc.c1);
}

#if defined(__CUDACC__)
// uniform: n x 256 B (expand_message_xmd output); points: n x 576 B (two points of E' per message)
__global__ void __launch_bounds__(128, 4) swu_g2_kernel(const uint8_t* uniform, uint8_t* points, size_t n_elems) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_elems) return;
    swu_g2_one(uniform + 128 * i, points + 288 * i);
}
#endif

}  // namespace swu
