// Microbenchmark: the two lane mappings of SURVEY section 7 ("measure both") on the Fp12-multiply shape.
//
//   A. lane = item  (what the tower-VM ships): a thread holds a whole Fp element (12 x u32); an Fp12 product is
//      12 output coefficients x (12 multiply-accumulates into one 768-bit accumulator + ONE Montgomery reduction),
//      operands read from shared-memory slots, no cross-lane traffic.
//   B. limb-parallel (north_star's sketch: carries propagated by warp shuffles): 4 lanes hold one Fp element
//      (3 x u32 each, 8 elements per warp); a Montgomery product is a 12-step word-serial CIOS loop with the multiplier
//      word and the quotient word broadcast by shuffles, the one-word right shift done by a shuffle and the pending
//      carries resolved by shuffle rounds at the end.  No lazy reduction is possible without a 24-word cross-lane
//      accumulator, so an Fp12 product is costed the way the reference does it (math.ts:660-678 over 516-527, 407-417):
//      54 Montgomery products + 216 Fp additions/subtractions, operands in registers (no shared-memory traffic is
//      charged), additions without modular correction -- every simplification favours B.
//
// Both are checked against each other (same residues mod p) before timing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I noble_bls12_381_b200/csrc -o lane_mapping lane_mapping.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fp_core.cuh"

// ---------------------------------------------------------------- B: limb-parallel primitives (4 lanes per element)
struct Lp { uint32_t v[3]; };

// r = a*b/2^384 mod p (value < 2p for a, b < 2^381), all 4 lanes of a group call it together
__device__ __forceinline__ Lp lp_montmul(const Lp& a, const Lp& b, const uint32_t (&p)[3], int g) {
    const unsigned full = 0xffffffffu;
    uint32_t t0 = 0, t1 = 0, t2 = 0, c = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const uint32_t bi = __shfl_sync(full, b.v[i % 3], i / 3, 4);
        uint64_t s0 = (uint64_t)a.v[0] * bi + t0;
        uint64_t s1 = (uint64_t)a.v[1] * bi + t1 + (s0 >> 32);
        uint64_t s2 = (uint64_t)a.v[2] * bi + t2 + (s1 >> 32);
        uint64_t t3 = (uint64_t)c + (s2 >> 32);
        const uint32_t low = __shfl_sync(full, (uint32_t)s0, 0, 4);
        const uint32_t m = low * fpc::kN0;
        s0 = (uint64_t)p[0] * m + (uint32_t)s0;
        s1 = (uint64_t)p[1] * m + (uint32_t)s1 + (s0 >> 32);
        s2 = (uint64_t)p[2] * m + (uint32_t)s2 + (s1 >> 32);
        t3 += (s2 >> 32);
        uint32_t in = __shfl_down_sync(full, (uint32_t)s0, 1, 4);
        if (g == 3) in = 0;
        const uint64_t v = (uint64_t)in + t3;
        t0 = (uint32_t)s1;
        t1 = (uint32_t)s2;
        t2 = (uint32_t)v;
        c = (uint32_t)(v >> 32);
    }
    // pending carries: lane g's c belongs to lane g+1's word 0
#pragma unroll
    for (int round = 0; round < 4; ++round) {
        uint32_t in = __shfl_up_sync(full, c, 1, 4);
        if (g == 0) in = 0;
        const uint64_t s0 = (uint64_t)t0 + in;
        const uint64_t s1 = (uint64_t)t1 + (s0 >> 32);
        const uint64_t s2 = (uint64_t)t2 + (s1 >> 32);
        t0 = (uint32_t)s0; t1 = (uint32_t)s1; t2 = (uint32_t)s2;
        c = (uint32_t)(s2 >> 32);
    }
    Lp r; r.v[0] = t0; r.v[1] = t1; r.v[2] = t2;
    return r;
}

// r = a + b (no modular correction), carries resolved across the group
__device__ __forceinline__ Lp lp_add(const Lp& a, const Lp& b, int g) {
    const unsigned full = 0xffffffffu;
    uint64_t s0 = (uint64_t)a.v[0] + b.v[0];
    uint64_t s1 = (uint64_t)a.v[1] + b.v[1] + (s0 >> 32);
    uint64_t s2 = (uint64_t)a.v[2] + b.v[2] + (s1 >> 32);
    uint32_t t0 = (uint32_t)s0, t1 = (uint32_t)s1, t2 = (uint32_t)s2, c = (uint32_t)(s2 >> 32);
#pragma unroll
    for (int round = 0; round < 3; ++round) {
        uint32_t in = __shfl_up_sync(full, c, 1, 4);
        if (g == 0) in = 0;
        s0 = (uint64_t)t0 + in;
        s1 = (uint64_t)t1 + (s0 >> 32);
        s2 = (uint64_t)t2 + (s1 >> 32);
        t0 = (uint32_t)s0; t1 = (uint32_t)s1; t2 = (uint32_t)s2;
        c = (uint32_t)(s2 >> 32);
    }
    Lp r; r.v[0] = t0; r.v[1] = t1; r.v[2] = t2;
    return r;
}

// ---------------------------------------------------------------- correctness: B against A on the same operands
__global__ void check_kernel(uint32_t* mismatches) {
    const uint32_t lane = threadIdx.x & 31, g = lane & 3;
    const uint32_t elem = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    uint32_t x[12], y[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { x[i] = elem * 2654435761u + i * 0x9e3779b9u; y[i] = (elem ^ 0x85ebca6bu) * (2 * i + 3) + 0xc2b2ae35u; }
    x[11] &= 0x0fffffffu; y[11] &= 0x0fffffffu;
    if (elem % 5 == 0) { for (int i = 0; i < 11; ++i) x[i] = 0xffffffffu; }
    if (elem % 7 == 0) { for (int i = 0; i < 11; ++i) y[i] = 0xffffffffu; }
    Lp a, b; uint32_t p[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        a.v[j] = g == 0 ? x[j] : g == 1 ? x[3 + j] : g == 2 ? x[6 + j] : x[9 + j];
        b.v[j] = g == 0 ? y[j] : g == 1 ? y[3 + j] : g == 2 ? y[6 + j] : y[9 + j];
        p[j] = g == 0 ? fpc::kP1[j] : g == 1 ? fpc::kP1[3 + j] : g == 2 ? fpc::kP1[6 + j] : fpc::kP1[9 + j];
    }
    Lp r = lp_montmul(a, b, p, g);
    r = lp_montmul(r, b, p, g);  // second product with a non-canonical operand
    uint32_t got[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) got[k] = __shfl_sync(0xffffffffu, r.v[k % 3], k / 3, 4);
    fpc::correct(got, 2);
    uint32_t want[12], tmp[12];
    fpc::Acc A;
    fpc::acc_zero(A); fpc::acc_mac(A, x, y); fpc::acc_redc(A, tmp); fpc::correct(tmp, 2);
    fpc::acc_zero(A); fpc::acc_mac(A, tmp, y); fpc::acc_redc(A, want); fpc::correct(want, 2);
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 12; ++k) bad |= got[k] != want[k];
    if (bad && g == 0) atomicAdd(mismatches, 1u);
}

// ---------------------------------------------------------------- B timed: 54 products + 216 additions per "Fp12 product"
template <int MINB>
__global__ void __launch_bounds__(256, MINB) lp_fp12_kernel(uint32_t* out, int iters) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    const int g = threadIdx.x & 3;
    uint32_t p[3];
    Lp x[4], y[4];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        p[j] = g == 0 ? fpc::kP1[j] : g == 1 ? fpc::kP1[3 + j] : g == 2 ? fpc::kP1[6 + j] : fpc::kP1[9 + j];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            x[s].v[j] = (t * 2654435761u + j + 17 * s) & (g == 3 && j == 2 ? 0x0fffffffu : 0xffffffffu);
            y[s].v[j] = ((t ^ 0x9e3779b9u) * (j + 3) + s) & (g == 3 && j == 2 ? 0x0fffffffu : 0xffffffffu);
        }
    }
    for (int it = 0; it < iters; ++it) {
        // 54 products in 2 independent chains (instruction-level parallelism as a scheduled tower would have), then
        // 216 additions in 4 independent chains
#pragma unroll 1
        for (int k = 0; k < 54 / 2; ++k) {
            x[0] = lp_montmul(x[0], y[0], p, g);
            x[1] = lp_montmul(x[1], y[1], p, g);
        }
#pragma unroll 1
        for (int k = 0; k < 216 / 4; ++k) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                y[s] = lp_add(y[s], x[s], g);
                y[s].v[2] &= (g == 3 ? 0x0fffffffu : 0xffffffffu);
            }
        }
        x[2] = x[0]; x[3] = x[1];
    }
    uint32_t acc = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc ^= x[s].v[j] ^ y[s].v[j];
    out[t] = acc;
}

// products only (no additions): the pure Montgomery-product rate of mapping B
template <int MINB, int CHAINS>
__global__ void __launch_bounds__(256, MINB) lp_mul_kernel(uint32_t* out, int iters) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    const int g = threadIdx.x & 3;
    uint32_t p[3];
    Lp x[CHAINS], y[CHAINS];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        p[j] = g == 0 ? fpc::kP1[j] : g == 1 ? fpc::kP1[3 + j] : g == 2 ? fpc::kP1[6 + j] : fpc::kP1[9 + j];
#pragma unroll
        for (int s = 0; s < CHAINS; ++s) {
            x[s].v[j] = (t * 2654435761u + j + 17 * s) & (g == 3 && j == 2 ? 0x0fffffffu : 0xffffffffu);
            y[s].v[j] = ((t ^ 0x9e3779b9u) * (j + 3) + s) & (g == 3 && j == 2 ? 0x0fffffffu : 0xffffffffu);
        }
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int s = 0; s < CHAINS; ++s) x[s] = lp_montmul(x[s], y[s], p, g);
    uint32_t acc = 0;
#pragma unroll
    for (int s = 0; s < CHAINS; ++s)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc ^= x[s].v[j];
    out[t] = acc;
}

// ---------------------------------------------------------------- A timed: 12 x (12 MAC from shared-memory slots + reduction + store)
// As in the tower-VM: the warps of a CTA work on the SAME 32 items and split the 12 output coefficients (4 warps x 3), the
// operands a[12], b[12] and the result c[12] are shared-memory slots of the CTA (36 x 1536 B), 4 CTAs = 16 warps per SM.
__global__ void __launch_bounds__(128, 4) item_fp12_kernel(uint32_t* out, int iters) {
    extern __shared__ uint32_t slots[];
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x[12], y[12], r[12];
    for (int s = warp; s < 36; s += 4)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            *reinterpret_cast<uint4*>(slots + s * 384 + q * 128 + lane * 4) =
                make_uint4(t * 2654435761u + s, (t ^ 0x9e3779b9u) * (q + 3), t + 31 * s + q, (t * 7 + s) & 0x0fffffffu);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int o = warp * 3; o < warp * 3 + 3; ++o) {
            fpc::Acc A;
            fpc::acc_zero(A);
#pragma unroll 1
            for (int k = 0; k < 12; ++k) {
                const uint32_t* pa = slots + k * 384;
                const uint32_t* pb = slots + (12 + ((o + 12 - k) % 12)) * 384;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const uint4 v = *reinterpret_cast<const uint4*>(pa + q * 128 + lane * 4);
                    const uint4 w = *reinterpret_cast<const uint4*>(pb + q * 128 + lane * 4);
                    x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
                    y[4 * q] = w.x; y[4 * q + 1] = w.y; y[4 * q + 2] = w.z; y[4 * q + 3] = w.w;
                }
                fpc::acc_mac(A, x, y);
            }
            fpc::acc_redc(A, r);
            fpc::correct(r, 2);
            r[11] &= 0x0fffffffu;
#pragma unroll
            for (int q = 0; q < 3; ++q)
                *reinterpret_cast<uint4*>(slots + (24 + o) * 384 + q * 128 + lane * 4) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        }
        __syncthreads();
        // feed the result back as the next left operand
        for (int s = warp * 3; s < warp * 3 + 3; ++s)
#pragma unroll
            for (int q = 0; q < 3; ++q)
                *reinterpret_cast<uint4*>(slots + s * 384 + q * 128 + lane * 4) = *reinterpret_cast<const uint4*>(slots + (24 + s) * 384 + q * 128 + lane * 4);
        __syncthreads();
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) acc ^= r[i];
    out[t] = acc;
}

template <class F>
static float time_best(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

template <int MINB>
static void run_lp(int sms, uint32_t* d) {
    const int blocks = sms * MINB, iters = 24;
    lp_fp12_kernel<MINB><<<blocks, 256>>>(d, 2);
    const float ms = time_best([&] { lp_fp12_kernel<MINB><<<blocks, 256>>>(d, iters); });
    const double fp12 = (double)blocks * 256 / 4 * iters;
    printf("{\"mapping\": \"limb-parallel (4 lanes/element, shuffle carries)\", \"shape\": \"54 Montgomery products + 216 additions, registers\", "
           "\"warps_per_sm\": %d, \"ms\": %.3f, \"fp12_products_per_s\": %.4g, \"err\": \"%s\"}\n",
           MINB * 8, ms, fp12 / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
}

template <int MINB, int CHAINS>
static void run_lp_mul(int sms, uint32_t* d) {
    const int blocks = sms * MINB, iters = 1024;
    lp_mul_kernel<MINB, CHAINS><<<blocks, 256>>>(d, 2);
    const float ms = time_best([&] { lp_mul_kernel<MINB, CHAINS><<<blocks, 256>>>(d, iters); });
    const double muls = (double)blocks * 256 / 4 * iters * CHAINS;
    printf("{\"mapping\": \"limb-parallel\", \"shape\": \"Montgomery products only, %d independent chains per group\", \"warps_per_sm\": %d, \"ms\": %.3f, "
           "\"fp_products_per_s\": %.4g, \"imad_wide_per_s\": %.4g, \"err\": \"%s\"}\n",
           CHAINS, MINB * 8, ms, muls / (ms * 1e-3), muls * 288 / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint32_t* d; cudaMalloc(&d, (size_t)sms * 8 * 256 * 4 + 64);  // one word per thread of the largest launch
    uint32_t* bad = d + (size_t)sms * 8 * 256;
    cudaMemset(bad, 0, 4);
    check_kernel<<<64, 256>>>(bad);
    uint32_t nbad = 1; cudaMemcpy(&nbad, bad, 4, cudaMemcpyDeviceToHost);
    printf("{\"check\": \"limb-parallel product == lane=item product mod p\", \"elements\": %d, \"mismatches\": %u, \"err\": \"%s\"}\n", 64 * 256 / 4, nbad,
           cudaGetErrorString(cudaGetLastError()));
    if (nbad) return 1;

    {   // A: lane = item at the interpreter's occupancy (16 warps per SM, 128 registers)
        const size_t smem = 36 * 384 * 4;
        cudaFuncSetAttribute(item_fp12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int blocks = sms * 4, iters = 256;
        item_fp12_kernel<<<blocks, 128, smem>>>(d, 2);
        const float ms = time_best([&] { item_fp12_kernel<<<blocks, 128, smem>>>(d, iters); });
        const double fp12 = (double)blocks * 32 * iters;
        printf("{\"mapping\": \"lane = item\", \"shape\": \"12 x (12 multiply-accumulates from shared-memory slots + 1 reduction + store)\", \"warps_per_sm\": 16, "
               "\"ms\": %.3f, \"fp12_products_per_s\": %.4g, \"imad_wide_per_s\": %.4g, \"err\": \"%s\"}\n",
               ms, fp12 / (ms * 1e-3), fp12 * (144.0 * 144 + 12 * 156) / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    }
    run_lp<2>(sms, d);
    run_lp<4>(sms, d);
    run_lp<8>(sms, d);
    run_lp_mul<2, 1>(sms, d);
    run_lp_mul<2, 2>(sms, d);
    run_lp_mul<4, 2>(sms, d);
    run_lp_mul<8, 1>(sms, d);
    run_lp_mul<8, 2>(sms, d);
    return 0;
}
