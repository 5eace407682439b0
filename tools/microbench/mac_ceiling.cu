// Microbenchmark: how close can the multiply-accumulate + Montgomery-reduction structure of the tower-VM get to the
// IMAD.WIDE issue peak at the interpreter's occupancy?  (tuning aid, not part of the library)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I noble_bls12_381_b200/csrc -o mac_ceiling mac_ceiling.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fp_core.cuh"

template <int T, bool SMEM>
__global__ void __launch_bounds__(256, 2) mac_kernel(uint32_t* out, int iters) {
    extern __shared__ uint32_t sm[];
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    uint32_t x[12], y[12], r[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { x[i] = t * 2654435761u + i; y[i] = (t ^ 0x9e3779b9u) * (i + 3); r[i] = 0; }
    x[11] &= 0x0fffffffu; y[11] &= 0x0fffffffu;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* myslots = sm + warp * 4 * 384;  // 4 slots per warp
    if (SMEM) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int q = 0; q < 3; ++q)
                *reinterpret_cast<uint4*>(myslots + s * 384 + q * 128 + lane * 4) = make_uint4(x[4 * q] + s, x[4 * q + 1], y[4 * q + 2], y[4 * q + 3] & 0x0fffffffu);
        __syncwarp();
    }
    for (int it = 0; it < iters; ++it) {
        fpc::Acc A;
        fpc::acc_zero(A);
#pragma unroll 1
        for (int k = 0; k < T; ++k) {
            if (SMEM) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    uint4 v = *reinterpret_cast<const uint4*>(myslots + (k & 3) * 384 + q * 128 + lane * 4);
                    uint4 w = *reinterpret_cast<const uint4*>(myslots + ((k + 1) & 3) * 384 + q * 128 + lane * 4);
                    x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
                    y[4 * q] = w.x; y[4 * q + 1] = w.y; y[4 * q + 2] = w.z; y[4 * q + 3] = w.w;
                }
            }
            fpc::acc_mac(A, x, y);
        }
        fpc::acc_redc(A, r);
        fpc::correct(r, 2);
        if (SMEM) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
                *reinterpret_cast<uint4*>(myslots + (it & 3) * 384 + q * 128 + lane * 4) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) x[i] ^= r[i] & 0x0fffffffu;
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) acc ^= r[i] ^ x[i];
    out[t] = acc;
}

template <int T, bool SMEM>
void run(const char* name, int blocks_per_sm, int sms, uint32_t* d) {
    const int iters = 2048 / T * 2;
    const int blocks = sms * blocks_per_sm;
    const size_t smem = SMEM ? 8 * 4 * 384 * 4 : 0;
    cudaFuncSetAttribute(mac_kernel<T, SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mac_kernel<T, SMEM><<<blocks, 256, smem>>>(d, 8);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        mac_kernel<T, SMEM><<<blocks, 256, smem>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double wide = (double)blocks * 256 * iters * (T * 144.0 + 156.0);
    printf("%-34s T=%2d blocks/SM=%d  %.3f ms  %.2f T IMAD.WIDE/s  (err=%s)\n", name, T, blocks_per_sm, best, wide / (best * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
}

// independent carry chains only (no accumulator dependencies between rows): NCH chains of 6 per thread
template <int NCH>
__global__ void __launch_bounds__(512, 1) chain_kernel(uint32_t* out, int iters) {
    uint64_t acc[NCH][6];
    uint32_t ct[NCH];
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        ct[k] = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[k][i] = ((uint64_t)(t * 2654435761u + k * 97u + i) << 32) | (t + i);
    }
    uint32_t a0 = t | 1, a1 = t ^ 0x9e3779b9u, a2 = t + 77, a3 = ~t, a4 = t * 3, a5 = t * 5 + 1, b = t * 7 + 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) fpc::chain6(acc[k], ct[k], a0, a1, a2, a3, a4, a5, b);
        b += ct[0];
    }
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
        for (int i = 0; i < 6; ++i) x ^= (uint32_t)acc[k][i] ^ (uint32_t)(acc[k][i] >> 32) ^ ct[k];
    out[t] = x;
}

template <int NCH>
void run_chain(int threads, int sms, uint32_t* d) {
    const int iters = 8192;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain_kernel<NCH><<<sms, threads>>>(d, 8);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        chain_kernel<NCH><<<sms, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double wide = (double)sms * threads * iters * NCH * 6.0;
    printf("independent chains x%d, %d warps/SMSP: %.2f T IMAD.WIDE/s\n", NCH, threads / 128, wide / (best * 1e-3) / 1e12);
}

template <int T>
__global__ void __launch_bounds__(512, 1) mac1_kernel(uint32_t* out, int iters) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    uint32_t x[12], y[12], r[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { x[i] = t * 2654435761u + i; y[i] = (t ^ 0x9e3779b9u) * (i + 3); r[i] = 0; }
    x[11] &= 0x0fffffffu; y[11] &= 0x0fffffffu;
    for (int it = 0; it < iters; ++it) {
        fpc::Acc A;
        fpc::acc_zero(A);
#pragma unroll 1
        for (int k = 0; k < T; ++k) fpc::acc_mac(A, x, y);
        fpc::acc_redc(A, r);
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] ^= r[i] & 0x0fffffffu;
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) acc ^= r[i] ^ x[i];
    out[t] = acc;
}

template <int T>
void run_mac1(int threads, int sms, uint32_t* d) {
    const int iters = 1024;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mac1_kernel<T><<<sms, threads>>>(d, 8);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        mac1_kernel<T><<<sms, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double wide = (double)sms * threads * iters * (T * 144.0 + 156.0);
    printf("acc_mac x%d + redc, %d warps/SMSP: %.2f T IMAD.WIDE/s\n", T, threads / 128, wide / (best * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 2 * 256 * 4);
    run<12, false>("registers", 2, p.multiProcessorCount, d);
    run<12, false>("registers", 1, p.multiProcessorCount, d);
    run<4, false>("registers", 2, p.multiProcessorCount, d);
    run<4, false>("registers", 1, p.multiProcessorCount, d);
    run<1, false>("registers", 2, p.multiProcessorCount, d);
    run<12, true>("operands from shared memory", 2, p.multiProcessorCount, d);
    run<4, true>("operands from shared memory", 2, p.multiProcessorCount, d);
    run<4, true>("operands from shared memory", 1, p.multiProcessorCount, d);
    run<1, true>("operands from shared memory", 2, p.multiProcessorCount, d);
    for (int th : {128, 256, 384, 512}) {
        run_chain<1>(th, p.multiProcessorCount, d);
        run_chain<2>(th, p.multiProcessorCount, d);
        run_chain<4>(th, p.multiProcessorCount, d);
        run_mac1<12>(th, p.multiProcessorCount, d);
    }
    return 0;
}
