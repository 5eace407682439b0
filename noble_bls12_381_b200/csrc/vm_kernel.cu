// Tower-VM interpreter kernel for sm_100a.  See vm.cuh for the execution model.
#include <cuda_runtime.h>
#include "vm.cuh"

namespace vm {

// Persistent CTAs: each CTA loops over batches of 32 items (lane = item).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2) vm_kernel(const Launch L) {
    extern __shared__ uint4 smem_raw[];
    uint32_t* slots = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* sconst = slots + (size_t)L.nslots * kSlotWords;
    for (uint32_t i = threadIdx.x; i < L.nconst * 12; i += blockDim.x) sconst[i] = L.consts[i];

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nbatch = (L.n_items + 31) / 32;
    const uint32_t* stream = L.prog + (size_t)warp * L.nrec * kRecWords;

    Ctx c;
    c.slots = slots;
    c.consts = sconst;
    c.far = L.far + (size_t)blockIdx.x * L.nfar * kSlotWords;
    c.nslots = L.nslots;
    c.lane = lane;
    c.buf = L.buf;

    for (uint32_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const uint32_t item = batch * 32 + lane;
        c.store_ok = item < L.n_items;
        c.item = c.store_ok ? item : L.n_items - 1;
        c.batch = batch;
        __syncthreads();  // previous batch fully retired (and constants visible)
        uint32_t next = stream[lane];
        for (uint32_t r = 0; r < L.nrec; ++r) {
            const uint32_t cur = next;
            if (r + 1 < L.nrec) next = stream[(size_t)(r + 1) * kRecWords + lane];  // prefetch
            const uint32_t hdr = __shfl_sync(0xffffffffu, cur, 0);
            const uint32_t aux = __shfl_sync(0xffffffffu, cur, 1);
            if (hdr & H_BAR) __syncthreads();
            exec_record(c, hdr, aux, [&](uint32_t i) { return __shfl_sync(0xffffffffu, cur, i); });
        }
    }
}

template __global__ void vm_kernel<6>(const Launch);
template __global__ void vm_kernel<8>(const Launch);
template __global__ void vm_kernel<12>(const Launch);

}  // namespace vm
