// Fp core for BLS12-381 on sm_100a: 12 x 32-bit limbs, Montgomery form (R = 2^384).
//
// Replaces the reference's `Fp` bigint arithmetic (math.ts:215-291: every op is a heap bigint `*`
// followed by `%`, math.ts:80-83) with register-resident carry chains:
//   * acc_mac   : 768-bit accumulator += a*b       (144 IMAD.WIDE.U32[.X], no reduction)
//   * acc_redc  : Montgomery reduction of the accumulator (12 x (1 IMAD + 12 IMAD.WIDE.U32.X))
//   * add12/sub12/csub_kp: carry-chain add/sub and conditional subtraction of k*p (immediates)
//   * acc_collapse / acc_from_wide / add24 / sub24 / kKP2: unreduced 768-bit sums of products as plain 24-word values,
//     shared between the two outputs of a two-output record (Karatsuba over Fp2 without an extra reduction, vm.cuh)
// "Lazy reduction": tower formulas accumulate many products into one accumulator and reduce once.
//
// The generated primitives also have a portable C++ body (`#else` of __CUDA_ARCH__) used ONLY by the CPU
// emulation tests (tests/emu) to validate interpreter + program logic without a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FPC_DEV __host__ __device__ __forceinline__
#define FPC_CONST static __device__ __constant__ const
#else
#define FPC_DEV static inline
#define FPC_CONST static const
#endif

#include "fp_core_gen.cuh"

namespace fpc {

FPC_DEV void acc_zero(Acc& A) {
#pragma unroll
    for (int i = 0; i < 12; ++i) A.e[i] = 0;
#pragma unroll
    for (int i = 0; i < 11; ++i) A.o[i] = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) A.c[i] = 0;
}

// always 0, but data-dependent on the accumulator (keeps a clock read behind the multiply-accumulate in tracing builds)
FPC_DEV uint32_t acc_dep(const Acc& A) {
    return (uint32_t)(A.e[11] & 0u);
}

FPC_DEV void copy12(uint32_t* r, const uint32_t* a) {
#pragma unroll
    for (int i = 0; i < 12; ++i) r[i] = a[i];
}

FPC_DEV void zero12(uint32_t* r) {
#pragma unroll
    for (int i = 0; i < 12; ++i) r[i] = 0;
}

// bring a value < min(2^rounds * p, 2^384) (rounds <= 4) into [0, p)
FPC_DEV void correct(uint32_t* r, int rounds) {
    if (rounds >= 4) csub_8p(r);
    if (rounds >= 3) csub_4p(r);
    if (rounds >= 2) csub_2p(r);
    if (rounds >= 1) csub_1p(r);
}

// Montgomery product, canonical output: r = a*b/R mod p  (a, b < p)
FPC_DEV void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    Acc A;
    acc_zero(A);
    acc_mac(A, a, b);
    acc_redc(A, r);
    csub_1p(r);
}

// r = a + b mod p, a, b canonical
FPC_DEV void add_mod(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    (void)add12(r, a, b);
    csub_1p(r);
}

// r = a - b mod p, a, b canonical
FPC_DEV void sub_mod(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t t[12];
    uint32_t borrow = sub12(r, a, b);
    (void)add12(t, r, kP1);
#pragma unroll
    for (int i = 0; i < 12; ++i) r[i] = borrow ? t[i] : r[i];
}

}  // namespace fpc
