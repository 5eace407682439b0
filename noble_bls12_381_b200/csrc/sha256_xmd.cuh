// expand_message_xmd with SHA-256 (index.ts:207-231, RFC 9380 5.3.1) for len_in_bytes = 256 (hash_to_field for
// G2: count = 2, m = 2, L = 64; index.ts:248-250).  One thread per message; the 21 compressions per 32-byte
// message are negligible next to the field arithmetic that follows.
#pragma once
#include <cstdint>

namespace sha {

#if defined(__CUDACC__)
#define SHA_HD __host__ __device__ __forceinline__
#else
#define SHA_HD static inline
#endif

struct Ctx {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len;
    uint32_t fill;
};

SHA_HD uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

SHA_HD void compress(uint32_t* h, const uint8_t* p) {
    const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64];
    for (int i = 0; i < 16; ++i)
        w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K[i] + w[i];
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

SHA_HD void init(Ctx& c) {
    c.h[0] = 0x6a09e667; c.h[1] = 0xbb67ae85; c.h[2] = 0x3c6ef372; c.h[3] = 0xa54ff53a;
    c.h[4] = 0x510e527f; c.h[5] = 0x9b05688c; c.h[6] = 0x1f83d9ab; c.h[7] = 0x5be0cd19;
    c.len = 0; c.fill = 0;
}

SHA_HD void update(Ctx& c, const uint8_t* p, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        c.buf[c.fill++] = p[i];
        if (c.fill == 64) { compress(c.h, c.buf); c.fill = 0; }
    }
    c.len += n;
}

SHA_HD void final(Ctx& c, uint8_t* out) {
    const uint64_t bits = c.len * 8;
    uint8_t b = 0x80;
    update(c, &b, 1);
    b = 0;
    while (c.fill != 56) update(c, &b, 1);
    uint8_t lenb[8];
    for (int i = 0; i < 8; ++i) lenb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(c, lenb, 8);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(c.h[i] >> 24); out[4 * i + 1] = (uint8_t)(c.h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(c.h[i] >> 8); out[4 * i + 3] = (uint8_t)c.h[i];
    }
}

// uniform = expand_message_xmd(msg, DST, len_in_bytes) (index.ts:207-231), len_in_bytes a multiple of 32 up to 256: 256 for
// hash_to_field of G2 (count 2, m 2, L 64), 128 for G1 (count 2, m 1); dst_prime = DST || I2OSP(len(DST), 1) (already shortened)
SHA_HD void expand_xmd(const uint8_t* msg, uint64_t msg_len, const uint8_t* dst_prime, uint32_t dst_prime_len, uint8_t* out256,
                       uint32_t len_in_bytes = 256) {
    Ctx c;
    uint8_t b0[32], bi[32], tmp[32];
    init(c);
    uint8_t z = 0;
    for (int i = 0; i < 64; ++i) update(c, &z, 1);  // Z_pad
    update(c, msg, msg_len);
    const uint8_t lib[3] = {(uint8_t)(len_in_bytes >> 8), (uint8_t)len_in_bytes, 0x00};  // I2OSP(len_in_bytes, 2) || I2OSP(0, 1)
    update(c, lib, 3);
    update(c, dst_prime, dst_prime_len);
    final(c, b0);
    init(c);
    update(c, b0, 32);
    uint8_t ctr = 1;
    update(c, &ctr, 1);
    update(c, dst_prime, dst_prime_len);
    final(c, bi);
    for (int k = 0; k < 32; ++k) out256[k] = bi[k];
    for (int i = 2; i <= (int)(len_in_bytes / 32); ++i) {
        for (int k = 0; k < 32; ++k) tmp[k] = b0[k] ^ bi[k];
        init(c);
        update(c, tmp, 32);
        ctr = (uint8_t)i;
        update(c, &ctr, 1);
        update(c, dst_prime, dst_prime_len);
        final(c, bi);
        for (int k = 0; k < 32; ++k) out256[32 * (i - 1) + k] = bi[k];
    }
}

}  // namespace sha
