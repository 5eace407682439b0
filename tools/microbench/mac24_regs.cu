// Register-budget probe for the column-wise accumulator prototype (tools/microbench/mac24_gen.py):
//   python tools/microbench/mac24_gen.py && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -I /tmp/mac24 -o /tmp/mac24/test tools/microbench/mac24_regs.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fp24_gen.cuh"
template <int T>
__global__ void __launch_bounds__(384, 2) k24(uint32_t* out, const uint32_t* in, int iters) {
    extern __shared__ uint32_t sm[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* my = sm + warp * 4 * 384;
    uint32_t r[12];
    for (int it = 0; it < iters; ++it) {
        fpc24::Acc A; fpc24::acc_zero(A);
#pragma unroll 1
        for (int k = 0; k < T; ++k) {
            uint32_t x[12], y[12];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                uint4 v = *reinterpret_cast<const uint4*>(my + (k & 3) * 384 + q * 128 + lane * 4);
                uint4 u = *reinterpret_cast<const uint4*>(my + ((k + 1) & 3) * 384 + q * 128 + lane * 4);
                x[4*q]=v.x; x[4*q+1]=v.y; x[4*q+2]=v.z; x[4*q+3]=v.w;
                y[4*q]=u.x; y[4*q+1]=u.y; y[4*q+2]=u.z; y[4*q+3]=u.w;
            }
            fpc24::acc_mac(A, x, y);
        }
        fpc24::acc_redc(A, r);
#pragma unroll
        for (int q = 0; q < 3; ++q)
            *reinterpret_cast<uint4*>(my + (it & 3) * 384 + q * 128 + lane * 4) = make_uint4(r[4*q], r[4*q+1], r[4*q+2], r[4*q+3]);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = r[0] ^ r[11];
}
template __global__ void k24<4>(uint32_t*, const uint32_t*, int);
int main() { return 0; }
