// Tower-VM interpreter kernel for sm_100a.  See vm.cuh for the execution model.
#include <cuda_runtime.h>
#include <cstdio>
#include "vm.cuh"

#if !defined(BLS381_ACQREL)
#define BLS381_ACQREL 1   // progress counters: st.release.cta / ld.acquire.cta (0 = the round-1 volatile + __threadfence_block form)
#endif

namespace vm {

// Progress counters are the ONLY synchronisation inside a batch.  Producer: all lanes store their slot columns, __syncwarp(),
// lane 0 publishes `records completed` with a CTA-scope RELEASE store; consumer: every lane polls with a CTA-scope ACQUIRE
// load before it touches the slots the record names -- ordered by the PTX memory model, not by observation.
__device__ __forceinline__ uint32_t load_progress(uint32_t paddr, uint32_t w) {
    uint32_t v;
#if BLS381_ACQREL
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(paddr + 4u * w) : "memory");
#else
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(paddr + 4u * w) : "memory");
#endif
    return v;
}

__device__ __forceinline__ void publish_progress(uint32_t paddr, uint32_t w, uint32_t done) {
#if BLS381_ACQREL
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(paddr + 4u * w), "r"(done) : "memory");
#else
    __threadfence_block();
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(paddr + 4u * w), "r"(done) : "memory");
#endif
}

__device__ __forceinline__ void wait_progress(uint32_t paddr, uint32_t w, uint32_t need, uint32_t sleep_ns) {
    if (need == 0) return;
    while (load_progress(paddr, w) < need) {
        if (sleep_ns) __nanosleep(sleep_ns);  // back off: a spinning warp steals issue slots and LDS bandwidth
    }
}

// Shared-memory layout of a CTA (dynamic part): slots | constants | progress[16] | mbarrier (16 B) | TMA staging.
// Everything here is a function of the kernel parameters (constant bank): recomputed where it is needed instead of being
// carried in registers across the record loop.
__device__ __forceinline__ uint32_t* smem_consts(uint32_t* slots, const Launch& L) { return slots + (size_t)L.nslots * kSlotWords; }
__device__ __forceinline__ uint32_t* smem_progress(uint32_t* slots, const Launch& L) { return smem_consts(slots, L) + (size_t)L.nconst * 12; }
__device__ __forceinline__ uint8_t* smem_stage(uint32_t* slots, const Launch& L) { return reinterpret_cast<uint8_t*>(smem_progress(slots, L) + 16 + 4); }
__device__ __forceinline__ uint32_t smem_mbar(uint32_t* slots, const Launch& L) { return (uint32_t)__cvta_generic_to_shared(smem_progress(slots, L) + 16); }

// Persistent CTAs: each CTA loops over batches of 32 items (lane = item).  Inside a batch there are NO
// CTA-wide barriers: a warp runs its own record stream and waits only for the progress counters its next
// record names (static dataflow schedule, see vmprog/builder.py).
template <int WARPS, int MINB, int MODE>
__global__ void __launch_bounds__(WARPS * 32, MINB) vm_kernel(const Launch L) {
    static_assert(WARPS <= 8, "eight 16-bit progress-requirement fields");
    extern __shared__ uint4 smem_raw[];
    uint32_t* slots = reinterpret_cast<uint32_t*>(smem_raw);
    __shared__ uint32_t next_batch;
    __shared__ unsigned long long probe[2];   // clock probe of CTA 0: {SM cycles, nanoseconds} at kernel start
    if (threadIdx.x == 0) {
        if (L.stage_bytes) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_mbar(slots, L)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // cold context (shared memory, see vm.cuh); visible after the first __syncthreads of the batch loop
        g_cold.nslots = L.nslots;
        g_cold.far = L.far + (size_t)blockIdx.x * L.nfar * kSlotWords;
        g_cold.consts = smem_consts(slots, L);
        g_cold.stage = L.stage_bytes ? smem_stage(slots, L) : nullptr;
        g_cold.n_items = L.n_items;
#pragma unroll
        for (int b = 0; b < kMaxBuffers; ++b) { g_cold.buf[b] = L.buf[b]; g_cold.stage_off[b] = L.stage_off[b]; }
        // clock probe: effective SM clock of this launch = cycles / nanoseconds seen by CTA 0 (bench.py reports it:
        // nvidia-smi sampling is too coarse for a 50 ms step)
        if (L.clk != nullptr && blockIdx.x == 0) {
            probe[0] = (unsigned long long)clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(probe[1]));
        }
    }
    {
        uint32_t* sconst = smem_consts(slots, L);
        for (uint32_t i = threadIdx.x; i < L.nconst * 12; i += blockDim.x) sconst[i] = L.consts[i];
    }

    const uint32_t lane = threadIdx.x & 31;
    Ctx c;
    c.slots = slots;
    c.cold = nullptr;
    c.lane = lane;
    {   // 32-bit shared-window addresses held in registers (opaque to the compiler, so they are not re-derived per access)
        const uint32_t sb = (uint32_t)__cvta_generic_to_shared(slots) + lane * 16u;
        const uint32_t cb = (uint32_t)__cvta_generic_to_shared(smem_consts(slots, L));
        asm volatile("mov.u32 %0, %1;" : "=r"(c.sbase) : "r"(sb));
        asm volatile("mov.u32 %0, %1;" : "=r"(c.cbase) : "r"(cb));
    }

    // Batch assignment: static round-robin, or (L.ticket != nullptr) claimed dynamically from a global counter.  The two
    // CTAs sharing an SM do NOT run at the same speed (measured 6.2 ms vs 10.0 ms per batch: the warp scheduler favours
    // one CTA; a CTA alone needs 5.0 ms), so with the static split the favoured CTAs finish early (profiles/r1_notes.md)
    const uint32_t nbatch = (L.n_items + 31) / 32;
    uint32_t round = 0;
    for (uint32_t batch = blockIdx.x; batch < nbatch; ++round) {
        __syncthreads();  // previous batch fully retired (and constants visible)
        if (threadIdx.x < 16) asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.cbase + L.nconst * 48u + 4u * threadIdx.x), "r"(0u) : "memory");
        if (threadIdx.x == 0) g_cold.batch = batch;
#if defined(BLS381_VM_TRACE)
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
#endif
        if (L.stage_bytes) {
            // stage this batch's wire-format inputs: one elected thread issues 1-D bulk copies (TMA) of the batch's
            // contiguous record block of every staged buffer; everybody then waits on the mbarrier phase (= round parity)
            const uint32_t mbar = smem_mbar(slots, L);
            if (threadIdx.x == 0) {
                uint8_t* stage = smem_stage(slots, L);
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
                             : "=r"(done) : "r"(mbar), "r"(round & 1u) : "memory");
            } while (!done);
        }
        __syncthreads();
        {
            // The record loop keeps as little as possible in registers (the multiply-accumulate needs ~110 of the 128):
            // a 32-bit word offset into the program, the two record words of this lane, and the three shared-window
            // addresses; the next record is pulled towards L1 with a prefetch instead of being held in registers.
            // (MODE_LEGACY has four registers to spare and holds the next record in two of them: +1.8 % measured.)
            uint32_t soff = (threadIdx.x >> 5) * L.nrec * kRecWords + lane;   // word offset of this lane's first record word
            uint32_t next = 0, next_hi = 0;
            if (MODE == MODE_LEGACY) { next = __ldg(L.prog + soff); next_hi = __ldg(L.prog + soff + 32); }
            for (uint32_t r = 0; r < L.nrec; ++r, soff += kRecWords) {
                const uint32_t* rec = L.prog + soff;
                uint32_t cur, cur_hi;  // record words lane and 32 + lane
                if (MODE == MODE_LEGACY) {
                    cur = next; cur_hi = next_hi;
                    if (r + 1 < L.nrec) { next = __ldg(rec + kRecWords); next_hi = __ldg(rec + kRecWords + 32); }
                } else {
                    cur = __ldg(rec); cur_hi = __ldg(rec + 32);
                    if (r + 1 < L.nrec) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + kRecWords));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + kRecWords + 32));
                    }
                }
#if defined(BLS381_VM_TRACE)   // tracing build only (tools/build_trace_lib.sh): per-record clocks of the first CTAs
                const bool tracing = L.trace != nullptr && blockIdx.x < L.trace_ctas && round == 0;
#else
                constexpr bool tracing = false;
#endif
                uint32_t t0 = 0, t1 = 0;
                if (tracing) t0 = (uint32_t)clock64();
                const uint32_t hdr = __shfl_sync(0xffffffffu, cur, 0);
                const uint32_t aux = __shfl_sync(0xffffffffu, cur, 1);
                if (hdr & H_BAR) {  // this record carries progress requirements: eight 16-bit fields, one 32-bit counter per warp
                    const uint32_t q0 = __shfl_sync(0xffffffffu, cur_hi, kWaitWord - 32), q1 = __shfl_sync(0xffffffffu, cur_hi, kWaitWord - 31);
                    const uint32_t q2 = __shfl_sync(0xffffffffu, cur_hi, kWaitWord - 30), q3 = __shfl_sync(0xffffffffu, cur_hi, kWaitWord - 29);
                    const uint32_t paddr = c.cbase + L.nconst * 48u;   // progress counters follow the constants
                    wait_progress(paddr, 0, q0 & 0xFFFF, L.pad); wait_progress(paddr, 1, q0 >> 16, L.pad);
                    wait_progress(paddr, 2, q1 & 0xFFFF, L.pad); wait_progress(paddr, 3, q1 >> 16, L.pad);
                    wait_progress(paddr, 4, q2 & 0xFFFF, L.pad); wait_progress(paddr, 5, q2 >> 16, L.pad);
                    wait_progress(paddr, 6, q3 & 0xFFFF, L.pad); wait_progress(paddr, 7, q3 >> 16, L.pad);
#if !BLS381_ACQREL
                    __threadfence_block();
#endif
                    __syncwarp();
                }
                if (tracing) t1 = (uint32_t)clock64();
                uint32_t ph[5] = {0, 0, 0, 0, 0};
#if defined(BLS381_VM_TRACE)
                if (tracing) exec_record<true, MODE>(c, hdr, aux, [cur, cur_hi](uint32_t i) { return __shfl_sync(0xffffffffu, i < 32u ? cur : cur_hi, i & 31u); }, ph);
                else
#endif
                exec_record<false, MODE>(c, hdr, aux, [cur, cur_hi](uint32_t i) { return __shfl_sync(0xffffffffu, i < 32u ? cur : cur_hi, i & 31u); });
                __syncwarp();
                if ((threadIdx.x & 31u) == 0) {
                    publish_progress(c.cbase + L.nconst * 48u, threadIdx.x >> 5, r + 1);
#if defined(BLS381_VM_TRACE)
                    if (tracing) {
                        uint32_t* tp = L.trace + (((size_t)blockIdx.x * WARPS + (threadIdx.x >> 5)) * L.nrec + r) * 8;
                        tp[0] = t0; tp[1] = t1; tp[2] = (uint32_t)clock64();
                        tp[3] = ph[0]; tp[4] = ph[1]; tp[5] = ph[2]; tp[6] = ph[3]; tp[7] = ph[4];
                    }
#endif
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
#if defined(BLS381_VM_TRACE)
    if (L.cta_log != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0 && blockIdx.x < nbatch) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (round >= 1 && round <= 16) L.cta_log[((size_t)blockIdx.x * 16 + round - 1) * 4 + 3] = (uint32_t)t;
        }
    }
#endif
    if (L.clk != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        L.clk[0] = (unsigned long long)clock64() - probe[0];
        L.clk[1] = t - probe[1];
    }
}

template __global__ void vm_kernel<2, 8, MODE_LEGACY>(const Launch);
template __global__ void vm_kernel<4, 4, MODE_LEGACY>(const Launch);
template __global__ void vm_kernel<6, 3, MODE_LEGACY>(const Launch);
template __global__ void vm_kernel<8, 2, MODE_LEGACY>(const Launch);
template __global__ void vm_kernel<2, 8, MODE_MAC2>(const Launch);
template __global__ void vm_kernel<4, 4, MODE_MAC2>(const Launch);
template __global__ void vm_kernel<6, 3, MODE_MAC2>(const Launch);
template __global__ void vm_kernel<8, 2, MODE_MAC2>(const Launch);

}  // namespace vm
