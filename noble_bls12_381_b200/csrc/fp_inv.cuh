// Modular inversion in Fp by an optimized binary GCD (T. Pornin, "Optimized Binary GCD for Modular Inversion",
// 2020) with a FIXED number of iterations (no data-dependent control flow: all 32 lanes stay in lock-step).
//
// Replaces the reference's extended Euclid (math.ts:134-156) -- any algorithm returns the same canonical
// residue -- and the previous a^(p-2) chain of ~480 dependent micro-ops that serialised the final
// exponentiation (profiles/r1_notes.md).  Cost: 25 outer iterations x (31 branch-free steps on 64-bit
// approximations + four 384x32-bit updates), about 15x fewer instructions than Fermat and a single micro-op.
//
// Input / output are in Montgomery form: out = x^-1 * R mod p for x = a*R (out = 0 for x = 0).
#pragma once
#include <cstdint>
#include "fp_core.cuh"

namespace fpc {

static constexpr int kInvOuter = 25;  // 25 * 31 = 775 >= 2*381 - 1 inner steps

// r = |f| * (neg ? -x : x) accumulated: helper building t += mult * X' over 14 words (mod 2^448)
FPC_DEV void mul_acc_signed(uint32_t* __restrict__ t, const uint32_t* x12, uint32_t mag, bool neg) {
    // X' = neg ? two's complement of x (14 words, sign-extended) : x
    uint64_t carry = 0, borrow_in = neg ? 1 : 0;
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        uint32_t w = i < 12 ? x12[i] : 0u;
        if (neg) {
            uint64_t s = (uint64_t)(uint32_t)(~w) + borrow_in;
            w = (uint32_t)s;
            borrow_in = s >> 32;
        }
        uint64_t p = (uint64_t)w * mag + t[i] + carry;
        t[i] = (uint32_t)p;
        carry = p >> 32;
    }
}

FPC_DEV uint32_t clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)v);
#else
    return v ? (uint32_t)__builtin_clz(v) : 32u;
#endif
}

// out = x^-1 (Montgomery in, Montgomery out); x canonical in [0, p)
FPC_DEV void fp_inv_mont(uint32_t* __restrict__ out, const uint32_t* __restrict__ x) {
    uint32_t a[12], b[12], u[12], v[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { a[i] = x[i]; b[i] = kP1[i]; u[i] = 0; v[i] = 0; }
    u[0] = 1;
#pragma unroll 1
    for (int it = 0; it < kInvOuter; ++it) {
        // ---- 64-bit approximations: top 33 bits and low 31 bits of a and b (n = max(len a, len b, 64))
        uint32_t a2 = a[1], a1 = a[0], a0 = 0, b2 = b[1], b1 = b[0], b0 = 0, ctop = 0;
        bool high = false;  // a word >= 2 of (a | b) is non-zero
#pragma unroll
        for (int i = 11; i >= 2; --i) {
            const uint32_t c = a[i] | b[i];
            const bool take = !high && c != 0;
            a2 = take ? a[i] : a2; a1 = take ? a[i - 1] : a1; a0 = take ? a[i - 2] : a0;
            b2 = take ? b[i] : b2; b1 = take ? b[i - 1] : b1; b0 = take ? b[i - 2] : b0;
            ctop = take ? c : ctop;
            high = high || c != 0;
        }
        const uint32_t lz = high ? clz32(ctop) : 0u;
        uint64_t ah = ((uint64_t)a2 << 32) | a1, bh = ((uint64_t)b2 << 32) | b1;
        if (lz) {
            ah = (ah << lz) | (a0 >> (32 - lz));
            bh = (bh << lz) | (b0 >> (32 - lz));
        }
        uint64_t xa = ((ah >> 31) << 31) | (a[0] & 0x7fffffffu);
        uint64_t xb = ((bh >> 31) << 31) | (b[0] & 0x7fffffffu);
        // ---- 31 branch-free binary GCD steps on the approximations
        int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 1
        for (int j = 0; j < 31; ++j) {
            const bool odd = (xa & 1) != 0;
            const bool swap = odd && xa < xb;
            const uint64_t ta = swap ? xb : xa, tb = swap ? xa : xb;
            const int64_t tf0 = swap ? f1 : f0, tf1 = swap ? f0 : f1, tg0 = swap ? g1 : g0, tg1 = swap ? g0 : g1;
            xa = (odd ? ta - tb : ta) >> 1;
            xb = tb;
            f0 = odd ? tf0 - tf1 : tf0;
            g0 = odd ? tg0 - tg1 : tg0;
            f1 = tf1 << 1;
            g1 = tg1 << 1;
        }
        // ---- (a, b) <- ((f0 a + g0 b) / 2^31, (f1 a + g1 b) / 2^31), made non-negative
        uint32_t na[14], nb[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) { na[i] = 0; nb[i] = 0; }
        const bool sf0 = f0 < 0, sg0 = g0 < 0, sf1 = f1 < 0, sg1 = g1 < 0;
        const uint32_t mf0 = (uint32_t)(sf0 ? -f0 : f0), mg0 = (uint32_t)(sg0 ? -g0 : g0);
        const uint32_t mf1 = (uint32_t)(sf1 ? -f1 : f1), mg1 = (uint32_t)(sg1 ? -g1 : g1);
        mul_acc_signed(na, a, mf0, sf0);
        mul_acc_signed(na, b, mg0, sg0);
        mul_acc_signed(nb, a, mf1, sf1);
        mul_acc_signed(nb, b, mg1, sg1);
        const bool nega = (na[13] >> 31) != 0, negb = (nb[13] >> 31) != 0;
        {   // arithmetic shift right by 31, then absolute value
            uint64_t ca = nega ? 1 : 0, cb = negb ? 1 : 0;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                uint32_t wa = (na[i] >> 31) | (na[i + 1] << 1);
                uint32_t wb = (nb[i] >> 31) | (nb[i + 1] << 1);
                if (nega) { uint64_t s = (uint64_t)(uint32_t)(~wa) + ca; wa = (uint32_t)s; ca = s >> 32; }
                if (negb) { uint64_t s = (uint64_t)(uint32_t)(~wb) + cb; wb = (uint32_t)s; cb = s >> 32; }
                a[i] = wa;
                b[i] = wb;
            }
        }
        // ---- (u, v) <- (f0 u + g0 v, f1 u + g1 v) / 2^32 mod p   (one Montgomery word step each)
        const bool uf0 = sf0 != nega, ug0 = sg0 != nega, uf1 = sf1 != negb, ug1 = sg1 != negb;
        uint32_t nu[12], nv[12];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const bool s0 = k == 0 ? uf0 : uf1, s1 = k == 0 ? ug0 : ug1;
            const uint32_t m0 = k == 0 ? mf0 : mf1, m1 = k == 0 ? mg0 : mg1;
            uint32_t uu[12], vv[12];
            if (s0) neg_raw(uu, u); else copy12(uu, u);   // -u == p - u (mod p), in [0, p]
            if (s1) neg_raw(vv, v); else copy12(vv, v);
            uint32_t t[14];
            uint64_t carry = 0;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                uint64_t p0 = (uint64_t)uu[i] * m0 + (uint32_t)carry;
                uint64_t p1 = (uint64_t)vv[i] * m1 + (uint32_t)p0;
                t[i] = (uint32_t)p1;
                carry = (carry >> 32) + (p0 >> 32) + (p1 >> 32);
            }
            t[12] = (uint32_t)carry;
            t[13] = (uint32_t)(carry >> 32);
            const uint32_t q = t[0] * kN0;
            carry = 0;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                uint64_t p0 = (uint64_t)q * kP1[i] + t[i] + carry;
                t[i] = (uint32_t)p0;
                carry = p0 >> 32;
            }
            uint64_t s = (uint64_t)t[12] + carry;
            t[12] = (uint32_t)s;
            t[13] += (uint32_t)(s >> 32);
            uint32_t* dst = k == 0 ? nu : nv;
#pragma unroll
            for (int i = 0; i < 12; ++i) dst[i] = t[i + 1];   // divide by 2^32 (t[0] == 0); t[13] == 0 by the bound < 2p
            csub_1p(dst);
        }
        copy12(u, nu);
        copy12(v, nv);
    }
    // b == 1 now: x^-1 = v * 2^25 ; out = x^-1 * R^2 = mont_mul(v, 2^25 * R^3 mod p)
    mont_mul(out, v, kInvFix);
}

}  // namespace fpc
