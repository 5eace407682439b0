// Tower-VM interpreter kernel for sm_100a.  See vm.cuh for the execution model.
#include <cuda_runtime.h>
#include <cstdio>
#include "vm.cuh"

#if !defined(BLS381_PACKED_PROGRESS)
#define BLS381_PACKED_PROGRESS 0
#endif

namespace vm {

__device__ __forceinline__ void wait_progress(const volatile uint32_t* progress, uint32_t w, uint32_t need, uint32_t sleep_ns) {
    if (need == 0) return;
    while (progress[w] < need) {
        if (sleep_ns) __nanosleep(sleep_ns);  // back off: a spinning warp steals issue slots and LDS bandwidth
    }
}

// Packed progress counters (WARPS <= 8): eight 16-bit counters in 16 bytes of shared memory, warp w at halfword w,
// i.e. exactly the layout of the four requirement words of a record.  Lane l reads word (l & 3) with ONE LDS.32 (a single
// shared-memory wavefront per poll, like the scalar loop) and compares both halves at once; a warp vote combines the four
// words.  This replaces eight dependent LDS/compare/branch rounds (~500 cycles per record even when nothing has to wait).
// Counters and requirements are < 0x8000, so ((p | 0x8000) - q) has bit 15 set iff p >= q and never borrows across halves.
// (An LDS.128 per lane was tried first: four wavefronts per poll saturate the shared-memory pipe when many warps spin and
// the producers' stores starve -- the kernel live-locks.)
__device__ __forceinline__ bool progress_reached(uint32_t paddr_lane, uint32_t q) {
    uint32_t p;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(p) : "r"(paddr_lane) : "memory");
    const uint32_t m = 0x80008000u;
    return __all_sync(0xffffffffu, (((p | m) - q) & m) == m);
}

// Persistent CTAs: each CTA loops over batches of 32 items (lane = item).  Inside a batch there are NO
// CTA-wide barriers: a warp runs its own record stream and waits only for the progress counters its next
// record names (static dataflow schedule, see vmprog/builder.py).
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) vm_kernel(const Launch L) {
    extern __shared__ uint4 smem_raw[];
    uint32_t* slots = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* sconst = slots + (size_t)L.nslots * kSlotWords;
    volatile uint32_t* progress = sconst + (size_t)L.nconst * 12;  // [16]
    // TMA staging area + its mbarrier (16-byte aligned: slots and constants are multiples of 16 bytes)
    uint8_t* stage = reinterpret_cast<uint8_t*>(sconst + (size_t)L.nconst * 12 + 16 + 4);
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(sconst + (size_t)L.nconst * 12 + 16);
    if (L.stage_bytes && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t phase = 0;
    for (uint32_t i = threadIdx.x; i < L.nconst * 12; i += blockDim.x) sconst[i] = L.consts[i];
    // clock probe: effective SM clock of this launch = cycles / nanoseconds seen by CTA 0 (bench.py reports it:
    // nvidia-smi sampling is too coarse for a 50 ms step)
    unsigned long long probe_c = 0, probe_t = 0;
    const bool probing = L.clk != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (probing) {
        probe_c = (unsigned long long)clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(probe_t));
    }

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nbatch = (L.n_items + 31) / 32;
    const uint32_t* stream = L.prog + (size_t)warp * L.nrec * kRecWords;

    Ctx c;
    c.slots = slots;
    c.consts = sconst;
    {   // 32-bit shared-window addresses held in registers (opaque to the compiler, so they are not re-derived per access)
        const uint32_t sb = (uint32_t)__cvta_generic_to_shared(slots) + lane * 16u;
        const uint32_t cb = (uint32_t)__cvta_generic_to_shared(sconst);
        asm volatile("mov.u32 %0, %1;" : "=r"(c.sbase) : "r"(sb));
        asm volatile("mov.u32 %0, %1;" : "=r"(c.cbase) : "r"(cb));
    }
    uint32_t paddr, paddr_lane;  // packed 16-bit progress counters (first 16 bytes of the progress area); this lane's word
    {
        const uint32_t pa = (uint32_t)__cvta_generic_to_shared(const_cast<uint32_t*>(progress));
        asm volatile("mov.u32 %0, %1;" : "=r"(paddr) : "r"(pa));
        asm volatile("mov.u32 %0, %1;" : "=r"(paddr_lane) : "r"(pa + (lane & 3u) * 4u));
    }
    const uint32_t qidx = 27u + (lane & 3u) + ((lane & 3u) ? 1u : 0u);  // requirement words 27, 29, 30, 31
    c.far = L.far + (size_t)blockIdx.x * L.nfar * kSlotWords;
    c.nslots = L.nslots;
    c.lane = lane;
    c.buf = L.buf;
    c.stage = L.stage_bytes ? stage : nullptr;
    c.stage_off = L.stage_off;

    // Batch assignment: static round-robin, or (L.ticket != nullptr) claimed dynamically from a global counter.  The two
    // CTAs sharing an SM do NOT run at the same speed (measured 6.2 ms vs 10.0 ms per batch: the warp scheduler favours
    // one CTA; a CTA alone needs 5.0 ms), so with the static split the favoured CTAs finish early (profiles/r1_notes.md)
    __shared__ uint32_t next_batch;
    uint32_t round = 0;
    for (uint32_t batch = blockIdx.x; batch < nbatch; ++round) {
        const uint32_t item = batch * 32 + lane;
        c.store_ok = item < L.n_items;
        c.item = c.store_ok ? item : L.n_items - 1;
        c.batch = batch;
        __syncthreads();  // previous batch fully retired (and constants visible)
        if (threadIdx.x < 16) progress[threadIdx.x] = 0;
        if (L.cta_log != nullptr && threadIdx.x == 0) {
            if (round < 16) {
                uint32_t smid;
                unsigned long long t;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                uint32_t* lp = L.cta_log + ((size_t)blockIdx.x * 16 + round) * 4;
                lp[0] = smid; lp[1] = batch; lp[2] = (uint32_t)t;
                if (round > 0) lp[-1] = (uint32_t)t;  // end of the previous round
            }
        }
        if (L.stage_bytes) {
            // stage this batch's wire-format inputs: one elected thread issues 1-D bulk copies (TMA) of the batch's
            // contiguous record block of every staged buffer; everybody then waits on the mbarrier phase
            if (threadIdx.x == 0) {
                const uint32_t items = min(32u, L.n_items - batch * 32u);
                uint32_t total = 0;
#pragma unroll
                for (int b = 0; b < kMaxBuffers; ++b)
                    if (L.stage_off[b] != kNoStage) total += items * L.buf[b].stride;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(total) : "memory");
#pragma unroll
                for (int b = 0; b < kMaxBuffers; ++b) {
                    if (L.stage_off[b] == kNoStage) continue;
                    const uint32_t bytes = items * L.buf[b].stride;
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + L.stage_off[b]);
                    const uint8_t* src = L.buf[b].base + (size_t)batch * 32u * L.buf[b].stride;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
                }
            }
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\t"
                             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                             "selp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
            } while (!done);
            phase ^= 1;
        }
        __syncthreads();
        uint32_t next = stream[lane];
        for (uint32_t r = 0; r < L.nrec; ++r) {
            const uint32_t cur = next;
            if (r + 1 < L.nrec) next = stream[(size_t)(r + 1) * kRecWords + lane];  // prefetch
#if defined(BLS381_VM_TRACE)   // tracing build only (tools/build_trace_lib.sh): per-record clocks of the first CTAs
            const bool tracing = L.trace != nullptr && blockIdx.x < L.trace_ctas && round == 0;
#else
            constexpr bool tracing = false;
#endif
            uint32_t t0 = 0, t1 = 0;
            if (tracing) t0 = (uint32_t)clock64();
            const uint32_t hdr = __shfl_sync(0xffffffffu, cur, 0);
            const uint32_t aux = __shfl_sync(0xffffffffu, cur, 1);
            if (hdr & H_BAR) {  // this record carries progress requirements
                if (WARPS <= 8 && BLS381_PACKED_PROGRESS) {  // eight 16-bit fields: one LDS.32 per lane + a vote
                    const uint32_t q = __shfl_sync(0xffffffffu, cur, qidx);
#if defined(BLS381_PROGRESS_DEBUG)
                    unsigned long long spins = 0;
#endif
                    while (!progress_reached(paddr_lane, q)) {
                        if (L.pad) __nanosleep(L.pad);
#if defined(BLS381_PROGRESS_DEBUG)   // watchdog (debug builds): report the stuck wait and carry on
                        if (++spins > 20000000ull) {
                            if (lane < 4) printf("STUCK cta %u warp %u rec %u lane %u q=%08x\n", blockIdx.x, warp, r, lane, q);
                            break;
                        }
#endif
                    }
                } else if (WARPS <= 8) {  // eight 16-bit fields, one 32-bit counter per warp
                    const uint32_t q0 = __shfl_sync(0xffffffffu, cur, 27), q1 = __shfl_sync(0xffffffffu, cur, 29);
                    const uint32_t q2 = __shfl_sync(0xffffffffu, cur, 30), q3 = __shfl_sync(0xffffffffu, cur, 31);
                    wait_progress(progress, 0, q0 & 0xFFFF, L.pad); wait_progress(progress, 1, q0 >> 16, L.pad);
                    wait_progress(progress, 2, q1 & 0xFFFF, L.pad); wait_progress(progress, 3, q1 >> 16, L.pad);
                    wait_progress(progress, 4, q2 & 0xFFFF, L.pad); wait_progress(progress, 5, q2 >> 16, L.pad);
                    wait_progress(progress, 6, q3 & 0xFFFF, L.pad); wait_progress(progress, 7, q3 >> 16, L.pad);
                } else {           // WARPS fields of 12 bits (up to 10 warps) or 10 bits (12 warps), little-endian over the four words
                    const uint32_t q0 = __shfl_sync(0xffffffffu, cur, 27), q1 = __shfl_sync(0xffffffffu, cur, 29);
                    const uint32_t q2 = __shfl_sync(0xffffffffu, cur, 30), q3 = __shfl_sync(0xffffffffu, cur, 31);
                    const uint64_t lo = ((uint64_t)q1 << 32) | q0, hi = ((uint64_t)q3 << 32) | q2;
                    constexpr int FW = WARPS <= 10 ? 12 : 10;
                    constexpr uint32_t FM = (1u << FW) - 1u;
#pragma unroll
                    for (int k = 0; k < WARPS; ++k) {
                        const int bit = FW * k;
                        uint32_t need;
                        if (bit + FW <= 64) need = (uint32_t)(lo >> bit) & FM;
                        else if (bit >= 64) need = (uint32_t)(hi >> (bit - 64)) & FM;
                        else need = (uint32_t)((lo >> bit) | (hi << (64 - bit))) & FM;
                        wait_progress(progress, k, need, L.pad);
                    }
                }
                __threadfence_block();
                __syncwarp();
            }
            if (tracing) t1 = (uint32_t)clock64();
            uint32_t ph[5] = {0, 0, 0, 0, 0};
#if defined(BLS381_VM_TRACE)
            if (tracing) exec_record<true>(c, hdr, aux, [&](uint32_t i) { return __shfl_sync(0xffffffffu, cur, i); }, ph);
            else
#endif
            exec_record<false>(c, hdr, aux, [&](uint32_t i) { return __shfl_sync(0xffffffffu, cur, i); });
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                if (WARPS <= 8 && BLS381_PACKED_PROGRESS) asm volatile("st.volatile.shared.u16 [%0], %1;" ::"r"(paddr + 2u * warp), "h"((unsigned short)(r + 1)) : "memory");
                else progress[warp] = r + 1;
                if (tracing) {
                    uint32_t* tp = L.trace + (((size_t)blockIdx.x * WARPS + warp) * L.nrec + r) * 8;
                    tp[0] = t0; tp[1] = t1; tp[2] = (uint32_t)clock64();
                    tp[3] = ph[0]; tp[4] = ph[1]; tp[5] = ph[2]; tp[6] = ph[3]; tp[7] = ph[4];
                }
            }
        }
        if (L.ticket != nullptr) {
            if (threadIdx.x == 0) next_batch = gridDim.x + atomicAdd(L.ticket, 1u);
            __syncthreads();
            batch = next_batch;
        } else {
            batch += gridDim.x;
        }
    }
    if (L.cta_log != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0 && blockIdx.x < nbatch) {
            const uint32_t rounds = round;
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (rounds >= 1 && rounds <= 16) L.cta_log[((size_t)blockIdx.x * 16 + rounds - 1) * 4 + 3] = (uint32_t)t;
        }
    }
    if (probing) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        L.clk[0] = (unsigned long long)clock64() - probe_c;
        L.clk[1] = t - probe_t;
    }
}

template __global__ void vm_kernel<2, 8>(const Launch);
template __global__ void vm_kernel<4, 4>(const Launch);
template __global__ void vm_kernel<6, 2>(const Launch);
template __global__ void vm_kernel<6, 3>(const Launch);
template __global__ void vm_kernel<8, 2>(const Launch);
template __global__ void vm_kernel<10, 2>(const Launch);
template __global__ void vm_kernel<12, 2>(const Launch);

}  // namespace vm
