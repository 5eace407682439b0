// Microbenchmark (planning aid for round 2, not part of the library): can the FP64 pipe carry part of the
// multi-precision products while the 64-bit integer multiplier (IMAD.WIDE, pipe fmaheavy) is saturated?
//
// Kernels (dependent-free streams, 8 independent accumulators per thread, 16 warps/SM like the interpreter and 64 warps/SM):
//   imad  : IMAD.WIDE.U32 chains only             -> T multiply-adds/s (32x32->64: 1024 bit-products each)
//   dfma  : DFMA chains only                      -> T fused multiply-adds/s (exact for 24x24-bit limbs: 576 bit-products)
//   mixed : both in the SAME thread, interleaved  -> do the two pipes overlap inside a warp?
//   split : even warps IMAD.WIDE, odd warps DFMA  -> do they overlap across warps of a sub-partition?
//   ffma  : FFMA chains only (12x12-bit limbs would be exact: 144 bit-products)   -- for completeness
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dual_pipe dual_pipe.cu && ./dual_pipe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kChains = 8;
constexpr int kUnroll = 16;

template <int MODE>  // 0 imad, 1 dfma, 2 mixed, 3 split, 4 ffma
__global__ void __launch_bounds__(256) pipe_kernel(unsigned long long* out, int iters) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    const bool odd_warp = (threadIdx.x >> 5) & 1;
    unsigned long long acc[kChains];
    double dacc[kChains];
    float facc[kChains];
    uint32_t a = t * 2654435761u + 12345u, b = (t ^ 0x9e3779b9u) | 1u;
    double da = (double)(a & 0xffffff), db = (double)(b & 0xffffff);
    float fa = (float)(a & 0xfff), fb = (float)(b & 0xfff);
#pragma unroll
    for (int k = 0; k < kChains; ++k) { acc[k] = k; dacc[k] = (double)k; facc[k] = (float)k; }
    const bool do_i = MODE == 0 || MODE == 2 || (MODE == 3 && !odd_warp);
    const bool do_d = MODE == 1 || MODE == 2 || (MODE == 3 && odd_warp);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
            for (int k = 0; k < kChains; ++k) {
                if (do_i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b + u));
                if (do_d) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(dacc[k]) : "d"(da), "d"(db));
                if (MODE == 4) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(facc[k]) : "f"(fa), "f"(fb));
            }
        }
    }
    unsigned long long r = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) r ^= acc[k] ^ (unsigned long long)__double_as_longlong(dacc[k]) ^ (unsigned long long)__float_as_uint(facc[k]);
    out[t] = r;
}

template <int MODE>
static void run(const char* name, int sms, int blocks_per_sm, int iters, unsigned long long* d) {
    const int blocks = sms * blocks_per_sm, threads = 256;
    pipe_kernel<MODE><<<blocks, threads>>>(d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        pipe_kernel<MODE><<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double per_thread = (double)iters * kUnroll * kChains;
    const double threads_total = (double)blocks * threads;
    // per-kind operation counts
    double n_i = 0, n_d = 0, n_f = 0;
    if (MODE == 0) n_i = per_thread * threads_total;
    if (MODE == 1) n_d = per_thread * threads_total;
    if (MODE == 2) { n_i = per_thread * threads_total; n_d = n_i; }
    if (MODE == 3) { n_i = per_thread * threads_total / 2; n_d = n_i; }
    if (MODE == 4) n_f = per_thread * threads_total;
    printf("%-6s %2d warps/SM: %8.3f ms   IMAD.WIDE %6.2f T/s   DFMA %6.2f T/s   FFMA %6.2f T/s   bit-products %7.0f T/s\n", name, blocks_per_sm * 8, best,
           n_i / best / 1e9, n_d / best / 1e9, n_f / best / 1e9, (n_i * 1024 + n_d * 576 + n_f * 144) / best / 1e9);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    unsigned long long* d = nullptr;
    cudaMalloc(&d, (size_t)sms * 8 * 256 * 8);
    printf("%s, %d SMs\n", p.name, sms);
    for (int bps : {2, 8}) {
        run<0>("imad", sms, bps, 2048, d);
        run<1>("dfma", sms, bps, 2048, d);
        run<2>("mixed", sms, bps, 2048, d);
        run<3>("split", sms, bps, 2048, d);
        run<4>("ffma", sms, bps, 2048, d);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    cudaFree(d);
    return 0;
}
