// Tower-VM: the B200 execution model for the wide (Fp12) part of the pairing hot path.
//
// One CTA = VM_WARPS warps x 32 lanes processes 32 pairings at once: LANE = PAIRING.  Every field
// element of the batch lives in shared memory as a "slot": 12 limbs x 32 lanes, laid out
// [slot][q=0..2][lane][4 limbs] so that a warp reads a slot with three conflict-free LDS.128.  A warp
// executes one "micro-op" at a time for all 32 pairings in lock-step (warp-uniform control flow, no
// divergence, 100% lane use):
//
//     dst = MontRed( sum_t X_t * Y_t ) + sum_e Z_e          (mod p, canonical)
//
// where each operand is a small signed combination cA*A + cB*B of slots (or a constant) formed in
// registers while loading.  This single op expresses everything the reference's tower does
// (math.ts:403-885: Fp2/Fp6/Fp12 multiply, square, sparse multiply, Frobenius, cyclotomic square, line
// evaluation math.ts:1331-1388) with LAZY REDUCTION: one Montgomery reduction per output coefficient
// instead of one per Fp product.  The per-warp instruction streams ("programs") are produced by the
// host-side builder (noble_bls12_381_b200/vmprog) which also schedules independent micro-ops across
// the warps of a CTA and allocates slots; dependent micro-ops are ordered by per-warp progress counters that a record
// names in its requirement words (static dataflow schedule, no `bar.sync` inside a batch).
//
// This header is shared by the CUDA kernel (vm_kernel.cu) and the CPU emulation used by the tests
// (tests/emu/vm_emu.cpp) -- the emulation exists to validate programs without a GPU and is not part
// of the shipped library.
#pragma once
#include <cstdint>
#include "fp_core.cuh"
#include "fp_inv.cuh"

namespace vm {

static constexpr int kRecWords = 64;       // one micro-op record = 256 bytes
static constexpr int kWaitWord = 60;       // words 60..63: progress requirements
static constexpr int kMaxTerms = 12;
static constexpr int kMaxEntries2 = 24;    // product entries of a two-output record (words 4..51)
static constexpr int kMaxEpi = 2;
static constexpr int kSlotWords = 12 * 32; // 1536 B per slot
static constexpr int kMaxBuffers = 8;
static constexpr int kConstSlots = 64;

// ---- record layout -------------------------------------------------------------------------------
// word 0 (header): [7:0] opcode  [15:8] dst  [19:16] T  [21:20] E  [24:22] ncorr  [25] has_waits
//                  [26] dst_global  [27] reserved  [28] dst_batch (lane 0 stores at item/32)
//                  [29] pad_const (padding lanes yield constant aux[31:24])  [27] dst_word (int32 store)
//                  [30] post_iszero  [31] post_gt_half  (turn the canonical result into a 0/1 flag)
// opcode OP_SEL  : dst = flag ? A : B  (word 2 = flag operand, word 3 = A operand, word 4 = B operand)
// opcode OP_INV  : dst = (operand in word 2)^-1 mod p, Montgomery form (binary GCD, fp_inv.cuh; 0 -> 0)
// opcode OP_BIT  : dst = bit `word 3` (0 = least significant) of the big-endian field of `word 4` bytes at
//                  the GLOBAL operand in word 2
// word 1 (aux)   : [7:0] dst buffer id  [15:8] dst field   [23:16] lane xor mask (for XLANE terms)
//                  [31:24] constant index for pad_const; otherwise post-scale factor k (0/1 = none): the reduced
//                  sum of products is multiplied by k in {2,3,4} before the epilogue operands are added
// words 2..25    : T terms, 2 words each:
//     word A: [7:0] xA  [15:8] xB  [19:16] cA (signed 4-bit)  [23:20] cB  [31:24] xflags
//     word B: [7:0] yA  [15:8] yB  [19:16] cA                [23:20] cB  [31:24] yflags
// words 26, 28   : E epilogue operands (same 1-word operand encoding)
// words 60..63   : eight 16-bit progress requirements (warp 0..7): this record may start only when
//                  warp w has completed at least that many records of its stream (0 = no requirement); CTAs of 10 / 12
//                  warps pack ten 12-bit / twelve 10-bit fields into the same four words
//
// opcode OP_MAC2 : TWO outputs that share products (an Fp2 coefficient pair):
//                      dst0 = ps0 * MontRed( S_P + S_M + S_0 ) + epilogue0
//                      dst1 = ps1 * MontRed( S_P - S_M + S_1 ) + epilogue1
//                  with four groups of product entries: S_M (counted +1 in dst0, -1 in dst1), S_P (+1 in both), S_0, S_1.
//                  S_M and S_P are kept UNREDUCED (768-bit sums of products, 24 words per lane) and are shared by the two
//                  accumulations, so an Fp2 product costs three Fp products (Karatsuba: x0*y0 in S_M, (-x1)*y1 in S_P,
//                  (x0+x1)*(y0+y1) in S_1) and still only one Montgomery reduction per output coefficient (the
//                  reference reduces every product, math.ts:451-462).  The 24-word intermediate lives in the record's own
//                  two destination slots until the first result is stored; no extra shared memory.
//     word 0: [7:0] opcode  [15:8] dst0  [23:16] dst1  [25] has_waits
//     word 1: [4:0] nM  [9:5] nP  [14:10] n0  [19:15] n1 (entries per group)  [21:20] E0  [23:22] E1  [26:24] ncorr0  [29:27] ncorr1
//     word 0 (cont.): [24] H_SINGLE: one result only (group 0 -> dst0) with the flag bits 26..31 of the one-output format
//     word 2: [6:0] K: S_P - S_M + K*p^2 >= 0 (K >= bound of S_M in units of p^2)  [10:8] ps0  [14:12] ps1
//     word 3: H_SINGLE records: [7:0] dst buffer id  [15:8] dst field  [31:24] padding constant
//     words 4..51 : product entries, 2 operand words each, in the order M, P, 0, 1.  An entry whose x word has flag bit 7
//                   (F_FUSE) is ADDED operand-wise to the next entry before the product: x = x_a + x_b, y = y_a + y_b
//                   (an operand word 0 in the second entry = nothing to add): operands of up to four slots
//     words 52,53 : epilogue operands of dst0;  words 54,55 : epilogue operands of dst1
// operand flags  : bit0 CONST  (A/B index the constant table instead of slots)
//                  bit1 GLOBAL (A = buffer id, B = byte offset/16 of a big-endian field; cA = number of top
//                       bits to clear (compression flags), cB = 0: 48-byte field, 1: 32-byte field)
//                  bit2 XLANE  (read the slot column of lane ^ mask)
//                  bit3 SIMPLE (operand is exactly one shared-memory slot, coefficient +1: fast path)
//                  bits 4..6: pre-decoded fast mode for shared-memory slots: 1 = -A, 2 = A+B, 3 = A-B, 4 = -A-B, 5 = 2A
enum : uint32_t { OP_NOP = 0, OP_MAC = 1, OP_SEL = 2, OP_BIT = 3, OP_INV = 4, OP_MAC2 = 5 };
enum : uint32_t { F_CONST = 1, F_GLOBAL = 2, F_XLANE = 4, F_SIMPLE = 8, F_FUSE = 0x80 };
static constexpr uint32_t H_SINGLE = 1u << 24;  // OP_MAC2 only: one result (group 0, dst0); word 3 = destination / padding info
static constexpr uint32_t H_BAR = 1u << 25;
static constexpr uint32_t H_DSTG = 1u << 26;
static constexpr uint32_t H_DSTBATCH = 1u << 28;  // global store by lane 0 only, at index item/32
static constexpr uint32_t H_PADCONST = 1u << 29;  // lanes beyond n_items produce constant aux[31:24] instead
static constexpr uint32_t H_POST_ISZERO = 1u << 30;  // result := (result == 0) ? 1 : 0   (plain integer flag)
static constexpr uint32_t H_POST_GTHALF = 1u << 31;  // result := (result > (p-1)/2) ? 1 : 0
static constexpr uint32_t H_DSTWORD = 1u << 27;      // global store of the low 32 bits as int32 (status arrays)

struct Buffer {
    uint8_t* base;      // device pointer
    uint32_t stride;    // bytes per item
    uint32_t pad;
};

struct Launch {
    const uint32_t* prog;   // [warps][nrec][32]
    const uint32_t* consts; // [nconst][12] Montgomery-form (or plain) constants
    uint32_t nrec;
    uint32_t nconst;
    uint32_t nslots;        // shared-memory slots
    uint32_t nfar;          // far slots (global scratch) per CTA
    uint32_t n_items;
    uint32_t pad;
    uint32_t* far;          // [ctas][nfar][kSlotWords]
    Buffer buf[kMaxBuffers];
    // TMA staging: wire-format input buffers whose 32-item batch record block (32 x stride bytes, contiguous) is
    // bulk-copied into shared memory at the start of every batch; kNoStage = read directly from global memory
    uint32_t stage_off[kMaxBuffers];   // byte offset inside the staging area
    uint32_t stage_bytes;              // total staging bytes (multiple of 16)
    uint32_t trace_ctas;               // tracing (debug): CTAs 0..trace_ctas-1 record 3 clocks per record
    uint32_t* trace;                   // [trace_ctas][warps][nrec][8] (start, after waits, end, 5 phase durations), first batch only
    uint32_t* cta_log;                 // tracing (debug): [grid][16][4] = {smid, batch, start ns, end ns} per CTA and batch round
    uint32_t* ticket;                  // dynamic batch assignment (or null): next unclaimed batch - gridDim.x, zeroed before the launch
    unsigned long long* clk;           // clock probe: CTA 0 stores {SM cycles, nanoseconds} it spent in this launch (or null)
};
static constexpr uint32_t kNoStage = 0xFFFFFFFFu;

// ---- execution context ------------------------------------------------------------------------------------
// Cold: what a record needs only occasionally (far slots, wire-format buffers, padding).  On the device it is ONE struct in
// shared memory per CTA, addressed directly: keeping these ~20 values in registers of every thread is what pushed the
// interpreter loop over its 128-register budget once the two-output records were added.
struct Cold {
    uint32_t nslots;         // shared-memory slots (slot numbers >= nslots are far slots)
    uint32_t pad0;
    uint32_t* far;           // this CTA's far slots
    const uint32_t* consts;  // constant table (generic pointer; the hot path uses Ctx::cbase)
    const uint8_t* stage;    // shared-memory staging area filled by TMA (nullptr: read inputs from global memory)
    uint32_t n_items;
    uint32_t batch;          // index of the 32-item batch being processed
    Buffer buf[kMaxBuffers];
    uint32_t stage_off[kMaxBuffers];
};
struct Ctx {
    // device: 32-bit shared-window addresses, computed ONCE per thread (the generic-pointer path re-derives the
    // shared window base with S2UR/ULEA and re-reads SR_TID before every access, ~5 % of the warp time in the v7 profile)
    uint32_t sbase;          // &slots[0] + lane * 16
    uint32_t cbase;          // &consts[0]
    // CPU emulation only (on the device the lane is threadIdx.x & 31 and everything else is in `g_cold`)
    uint32_t lane;
    uint32_t* slots;
    const Cold* cold;
};
#if defined(__CUDACC__)
__shared__ Cold g_cold;
#endif
#if defined(__CUDA_ARCH__)
#define VM_COLD(c) g_cold
#define VM_LANE(c) (threadIdx.x & 31u)
#else
#define VM_COLD(c) (*(c).cold)
#define VM_LANE(c) ((c).lane)
#endif
#define VM_NSLOTS(c) (VM_COLD(c).nslots)
// lane's item: index in the batch arrays (clamped to n_items - 1 for loads); store_ok: the lane is not padding
FPC_DEV bool ctx_store_ok(const Ctx& c) { return VM_COLD(c).batch * 32u + VM_LANE(c) < VM_COLD(c).n_items; }
FPC_DEV uint32_t ctx_item(const Ctx& c) {
    const uint32_t item = VM_COLD(c).batch * 32u + VM_LANE(c), last = VM_COLD(c).n_items - 1u;
    return item < last ? item : last;
}

#if !defined(BLS381_CACHED_SMEM)
#define BLS381_CACHED_SMEM 1
#endif
#if defined(__CUDA_ARCH__) && BLS381_CACHED_SMEM
#define VM_SMEM_ASM 1
#else
#define VM_SMEM_ASM 0
#endif
#if defined(__CUDA_ARCH__)
FPC_DEV void lds128(uint32_t* r, uint32_t addr) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
FPC_DEV void sts128(uint32_t addr, const uint32_t* r) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
#endif

// start of this lane's record in input buffer `bufid` (TMA-staged copy in shared memory when available)
FPC_DEV const uint8_t* wire_record(const Ctx& c, uint32_t bufid) {
    const Buffer& b = VM_COLD(c).buf[bufid];
    if (VM_COLD(c).stage != nullptr && VM_COLD(c).stage_off[bufid] != kNoStage && ctx_store_ok(c))
        return VM_COLD(c).stage + VM_COLD(c).stage_off[bufid] + (size_t)VM_LANE(c) * b.stride;
    return b.base + (size_t)ctx_item(c) * b.stride;
}

FPC_DEV uint32_t bswap32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0, 0x0123);
#else
    return __builtin_bswap32(v);
#endif
}

// r = 12 limbs of lane `lane`'s column of a slot given by pointer (generic-pointer path: far slots, builds without VM_SMEM_ASM)
FPC_DEV void load_col(uint32_t* r, const uint32_t* s) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        uint4 v = *reinterpret_cast<const uint4*>(s + q * 128);
        r[4 * q + 0] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
    }
#else
    for (int q = 0; q < 3; ++q)
        for (int k = 0; k < 4; ++k) r[4 * q + k] = s[q * 128 + k];
#endif
}

FPC_DEV void store_col(uint32_t* s, const uint32_t* r) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int q = 0; q < 3; ++q)
        *reinterpret_cast<uint4*>(s + q * 128) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
#else
    for (int q = 0; q < 3; ++q)
        for (int k = 0; k < 4; ++k) s[q * 128 + k] = r[4 * q + k];
#endif
}

FPC_DEV void load_slot(uint32_t* r, const Ctx& c, uint32_t slot, uint32_t lane) {
    if (slot < VM_NSLOTS(c)) {
#if VM_SMEM_ASM
        const uint32_t a = c.sbase + slot * (kSlotWords * 4u) + (lane - VM_LANE(c)) * 16u;
#pragma unroll
        for (int q = 0; q < 3; ++q) lds128(r + 4 * q, a + q * 512u);
#else
        load_col(r, c.slots + slot * kSlotWords + lane * 4);
#endif
    } else {
        load_col(r, VM_COLD(c).far + (slot - VM_NSLOTS(c)) * kSlotWords + lane * 4);
    }
}

FPC_DEV void store_slot(const uint32_t* r, const Ctx& c, uint32_t slot) {
    if (slot < VM_NSLOTS(c)) {
#if VM_SMEM_ASM
        const uint32_t a = c.sbase + slot * (kSlotWords * 4u);
#pragma unroll
        for (int q = 0; q < 3; ++q) sts128(a + q * 512u, r + 4 * q);
#else
        store_col(c.slots + slot * kSlotWords + VM_LANE(c) * 4, r);
#endif
    } else {
        store_col(VM_COLD(c).far + (slot - VM_NSLOTS(c)) * kSlotWords + VM_LANE(c) * 4, r);
    }
}

// wire format: big-endian field (48 or 32 bytes) at byte offset 16*off16 -> 12 little-endian limbs
FPC_DEV void load_wire(uint32_t* r, const Ctx& c, uint32_t bufid, uint32_t off16, uint32_t clear_top, uint32_t short32) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(wire_record(c, bufid) + off16 * 16u);
    if (short32) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = bswap32(p[7 - k]);
#pragma unroll
        for (int k = 8; k < 12; ++k) r[k] = 0;
    } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) r[k] = bswap32(p[11 - k]);
        r[11] &= 0xFFFFFFFFu >> clear_top;
    }
}

FPC_DEV void store_wire(const uint32_t* r, const Ctx& c, uint32_t bufid, uint32_t off16, bool per_batch, bool word) {
    if (per_batch ? (VM_LANE(c) != 0) : !ctx_store_ok(c)) return;
    const Buffer& b = VM_COLD(c).buf[bufid];
    const size_t idx = per_batch ? (size_t)(VM_COLD(c).batch) : (size_t)ctx_item(c);
    uint32_t* p = reinterpret_cast<uint32_t*>(b.base + idx * b.stride + off16 * 16u);
    if (word) { p[0] = r[0]; return; }
#pragma unroll
    for (int k = 0; k < 12; ++k) p[11 - k] = bswap32(r[k]);
}

FPC_DEV int sext4(uint32_t v) { return (int)((v & 0xF) ^ 8) - 8; }

// v = k * v for k in 1..4 (no reduction; caller guarantees k*v < 2^384)
FPC_DEV void scale_raw(uint32_t* v, int k) {
    if (k == 1) return;
    uint32_t t[12];
    fpc::copy12(t, v);
    (void)fpc::add12(v, v, v);                 // 2v
    if (k == 3) (void)fpc::add12(v, v, t);     // 3v
    if (k == 4) (void)fpc::add12(v, v, v);     // 4v
}

FPC_DEV void load_one(uint32_t* r, const Ctx& c, uint32_t idx, uint32_t flags, uint32_t xmask) {
    if (flags & F_CONST) {
#if VM_SMEM_ASM
        const uint32_t a = c.cbase + idx * 48u;
#pragma unroll
        for (int q = 0; q < 3; ++q) lds128(r + 4 * q, a + q * 16u);
#else
        const uint32_t* s = VM_COLD(c).consts + idx * 12;
#pragma unroll
        for (int k = 0; k < 12; ++k) r[k] = s[k];
#endif
    } else {
        load_slot(r, c, idx, (flags & F_XLANE) ? (VM_LANE(c) ^ xmask) : VM_LANE(c));
    }
}

FPC_DEV void load_near(uint32_t* r, const Ctx& c, uint32_t slot) {
#if VM_SMEM_ASM
    const uint32_t a = c.sbase + slot * (kSlotWords * 4u);
#pragma unroll
    for (int q = 0; q < 3; ++q) lds128(r + 4 * q, a + q * 512u);
#else
    load_col(r, c.slots + slot * kSlotWords + VM_LANE(c) * 4);
#endif
}

// r = 2*r (no reduction; r < 2^383): funnel shifts, no carry chain
FPC_DEV void shl1_raw(uint32_t* r) {
#pragma unroll
    for (int k = 11; k > 0; --k) r[k] = (r[k] << 1) | (r[k - 1] >> 31);
    r[0] <<= 1;
}

// Pre-decoded operand modes on shared-memory slots (bits 27..30 of the operand word): SIMPLE = +A; fast modes
// 1 = -A, 2 = A+B, 3 = A-B, 4 = -A-B, 5 = 2A.  Everything else takes the general path.
FPC_DEV bool operand_is_near(uint32_t w) { return (w & ((F_SIMPLE << 24) | (0x7u << 28))) != 0; }

// second half of a near operand whose A slot is already in r (load_near issued earlier, see exec_record)
FPC_DEV void finish_near_operand(uint32_t* r, const Ctx& c, uint32_t w) {
    if (w & (F_SIMPLE << 24)) return;
    const uint32_t fast = (w >> 28) & 0x7;
    if (fast == 1) { fpc::neg_raw(r, r); return; }
    if (fast == 5) { shl1_raw(r); return; }
    uint32_t u[12];
    load_near(u, c, (w >> 8) & 0xFF);
    if (fast == 2) { (void)fpc::add12(r, r, u); return; }
    if (fast == 3) { fpc::neg_raw(u, u); (void)fpc::add12(r, r, u); return; }
    (void)fpc::add12(r, r, u);   // fast == 4: -A - B = 2p - (A + B)
    fpc::neg_raw2(r, r);
}

// operand = cA*A + cB*B  (negative coefficients via p - X), unreduced
FPC_DEV void load_operand(uint32_t* r, const Ctx& c, uint32_t w, uint32_t xmask) {
    if (operand_is_near(w)) {
        load_near(r, c, w & 0xFF);
        finish_near_operand(r, c, w);
        return;
    }
    const uint32_t a = w & 0xFF, b = (w >> 8) & 0xFF, flags = w >> 24;
    const int ca = sext4(w >> 16), cb = sext4(w >> 20);
    if (flags & F_GLOBAL) {
        load_wire(r, c, a, b, (w >> 16) & 0xF, (w >> 20) & 0xF);
        return;
    }
    load_one(r, c, a, flags, xmask);
    if (ca < 0) fpc::neg_raw(r, r);
    scale_raw(r, ca < 0 ? -ca : ca);
    if (cb != 0) {
        uint32_t u[12];
        load_one(u, c, b, flags, xmask);
        if (cb < 0) fpc::neg_raw(u, u);
        scale_raw(u, cb < 0 ? -cb : cb);
        (void)fpc::add12(r, r, u);
    }
}

// both operands of a product: the raw A-slot loads of x AND y are issued before either operand is completed, so
// that y's shuffle + shared-memory latency overlaps x's negation / addition chains
FPC_DEV void load_operand_pair(uint32_t* x, uint32_t* y, const Ctx& c, uint32_t wx, uint32_t wy, uint32_t xmask) {
    const bool nx = operand_is_near(wx), ny = operand_is_near(wy);
    if (nx) load_near(x, c, wx & 0xFF);
    if (ny) load_near(y, c, wy & 0xFF);
    if (nx) finish_near_operand(x, c, wx); else load_operand(x, c, wx, xmask);
    if (ny) finish_near_operand(y, c, wy); else load_operand(y, c, wy, xmask);
}

FPC_DEV uint32_t phase_clock() {
#if defined(__CUDA_ARCH__)
    return (uint32_t)clock64();
#else
    return 0;
#endif
}

// ---- result post-processing + store (shared by all record types) ----------------------------------------
// gaux: [7:0] destination buffer id  [15:8] destination field  [31:24] constant index for H_PADCONST
FPC_DEV void post_and_store(const Ctx& c, uint32_t* r, uint32_t hdr, uint32_t gaux, uint32_t dst) {
    if (hdr & (H_POST_ISZERO | H_POST_GTHALF)) {
        uint32_t flag;
        if ((hdr & H_POST_ISZERO) && (hdr & H_POST_GTHALF)) {
            flag = r[0] & 1u;  // parity (sgn0)
        } else if (hdr & H_POST_ISZERO) {
            uint32_t any = 0;
#pragma unroll
            for (int k = 0; k < 12; ++k) any |= r[k];
            flag = any ? 0u : 1u;
        } else {
            uint32_t t[12];
            flag = fpc::sub12(t, fpc::kHalfP, r);  // borrow <=> r > (p-1)/2
        }
        fpc::zero12(r);
        r[0] = flag;
    }
    if ((hdr & H_PADCONST) && !ctx_store_ok(c)) {
        const uint32_t* s = VM_COLD(c).consts + (gaux >> 24) * 12;
#pragma unroll
        for (int k = 0; k < 12; ++k) r[k] = s[k];
    }
    if (hdr & H_DSTG) {
        store_wire(r, c, gaux & 0xFF, (gaux >> 8) & 0xFF, (hdr & H_DSTBATCH) != 0, (hdr & H_DSTWORD) != 0);
    } else {
        store_slot(r, c, dst);
    }
}

// r += operand(w), one slot at a time: at most ONE 12-word temporary is live next to r (a fused product has x, y and this
// temporary in registers -- 36 words like any other product; building the second half with load_operand first would need 48)
FPC_DEV void accumulate_operand(uint32_t* r, const Ctx& c, uint32_t w, uint32_t xmask) {
    uint32_t t[12];
    if (operand_is_near(w)) {
        const uint32_t fast = (w & (F_SIMPLE << 24)) ? 0u : ((w >> 28) & 0x7);
        load_near(t, c, w & 0xFF);
        if (fast == 1 || fast == 4) fpc::neg_raw(t, t);       // -A ...
        if (fast == 5) shl1_raw(t);                            // 2A
        (void)fpc::add12(r, r, t);
        if (fast >= 2 && fast <= 4) {
            load_near(t, c, (w >> 8) & 0xFF);
            if (fast != 2) fpc::neg_raw(t, t);                 // ... - B
            (void)fpc::add12(r, r, t);
        }
        return;
    }
    const uint32_t a = w & 0xFF, b = (w >> 8) & 0xFF, flags = w >> 24;
    if (flags & F_GLOBAL) {
        load_wire(t, c, a, b, (w >> 16) & 0xF, (w >> 20) & 0xF);
        (void)fpc::add12(r, r, t);
        return;
    }
    const int ca = sext4(w >> 16), cb = sext4(w >> 20);
    load_one(t, c, a, flags, xmask);
    if (ca < 0) fpc::neg_raw(t, t);
    scale_raw(t, ca < 0 ? -ca : ca);
    (void)fpc::add12(r, r, t);
    if (cb != 0) {
        load_one(t, c, b, flags, xmask);
        if (cb < 0) fpc::neg_raw(t, t);
        scale_raw(t, cb < 0 ? -cb : cb);
        (void)fpc::add12(r, r, t);
    }
}

// ---- two-output record (OP_MAC2) -------------------------------------------------------------------------
// acc += sum over `n` product entries starting at record word `wi` (advanced); fused entry pairs count as two entries
template <class WordFn>
FPC_DEV void mac_entries(fpc::Acc& A, const Ctx& c, WordFn W, uint32_t& wi, uint32_t n, uint32_t xmask) {
    for (uint32_t t = 0; t < n; ++t) {
        uint32_t x[12], y[12];
        const uint32_t wx = W(wi), wy = W(wi + 1);
        wi += 2;
        load_operand_pair(x, y, c, wx, wy, xmask);
        if (wx & (F_FUSE << 24)) {
            const uint32_t wx2 = W(wi), wy2 = W(wi + 1);
            wi += 2;
            ++t;
            if (wx2) accumulate_operand(x, c, wx2, xmask);
            if (wy2) accumulate_operand(y, c, wy2, xmask);
        }
        fpc::acc_mac(A, x, y);
    }
}

// In programs of the two-output format EVERY multiply-accumulate record uses this layout (a record with one result
// sets H_SINGLE and only has group 0), so that the interpreter contains ONE multiply-accumulate loop body.
template <class WordFn>
FPC_DEV void exec_mac2(const Ctx& c, uint32_t hdr, uint32_t aux, WordFn W) {
    const uint32_t dst0 = (hdr >> 8) & 0xFF, dst1 = (hdr >> 16) & 0xFF;
    const bool shared = (aux & 1023u) != 0;   // any entries in the groups M or P
    constexpr uint32_t xmask = 0;              // cross-lane operands only occur in epilogue-only records (one-output format)
    uint32_t wi = 4;
    // Four accumulation phases (S_M, S_P, dst0, dst1) around ONE multiply-accumulate loop.  Every phase starts its
    // accumulator from the 24-word value `w` (zero, S_P + S_M, or S_P - S_M + K p^2): the 58-register accumulator is
    // never live across a phase boundary, only these 24 words are.
    uint32_t w[24];
#pragma unroll
    for (int k = 0; k < 24; ++k) w[k] = 0;
#pragma unroll 1
    for (uint32_t ph = shared ? ((aux & 31u) ? 0u : 1u) : 2u; ph < 4u; ++ph) {
        fpc::Acc A;
        fpc::acc_from_wide(A, w);
        const uint32_t n = (aux >> (5u * ph)) & 31u;
        mac_entries(A, c, W, wi, n, xmask);
        if (ph == 0u) {            // S_M -> the 24-word scratch formed by the two destination slots
            fpc::acc_collapse(A, w);
            store_slot(w, c, dst0);
            store_slot(w + 12, c, dst1);
#pragma unroll
            for (int k = 0; k < 24; ++k) w[k] = 0;
        } else if (ph == 1u) {     // S_P; combine with S_M
            uint32_t wp[24];
            fpc::acc_collapse(A, wp);
            if (aux & 31u) {
                uint32_t wm[24];
                load_slot(wm, c, dst0, VM_LANE(c));
                load_slot(wm + 12, c, dst1, VM_LANE(c));
                fpc::add24(w, wp, wm);                        // S_P + S_M: start of the dst0 accumulation
                fpc::add24(wp, wp, fpc::kKP2[W(2) & 127u]);   // S_P + K p^2 - S_M >= 0: start of the dst1 accumulation
                fpc::sub24(wp, wp, wm);
            } else {
#pragma unroll
                for (int k = 0; k < 24; ++k) w[k] = wp[k];
            }
            store_slot(wp, c, dst0);
            store_slot(wp + 12, c, dst1);
        } else {
            const uint32_t sh = ph == 2u ? 0u : 1u;   // which output
            uint32_t r[12];
            if (shared || n) {
                fpc::acc_redc(A, r);
                const uint32_t ps = (W(2) >> (8u + 4u * sh)) & 7u;
                if (ps > 1) scale_raw(r, (int)ps);
            } else {
                fpc::zero12(r);
            }
            const uint32_t E = (aux >> (20u + 2u * sh)) & 3u;
            for (uint32_t e = 0; e < E; ++e) {
                uint32_t z[12];
                load_operand(z, c, W(52u + 2u * sh + e), xmask);
                (void)fpc::add12(r, r, z);
            }
            fpc::correct(r, (int)((aux >> (24u + 3u * sh)) & 7u));
            if (ph == 2u) {
                if (hdr & H_SINGLE) {   // one result: flags, padding constant, global / slot store
                    post_and_store(c, r, hdr, W(3), dst0);
                    return;
                }
                if (shared) {   // fetch S_P - S_M + K p^2 before dst0 (half of the scratch) is overwritten
                    load_slot(w, c, dst0, VM_LANE(c));
                    load_slot(w + 12, c, dst1, VM_LANE(c));
                } else {
#pragma unroll
                    for (int k = 0; k < 24; ++k) w[k] = 0;
                }
                store_slot(r, c, dst0);
            } else {
                store_slot(r, c, dst1);
#pragma unroll
                for (int k = 0; k < 24; ++k) w[k] = 0;   // `w` is (re)defined on EVERY path to the loop head: it is dead during
                                                         // the multiply-accumulate instead of being kept in 24 registers
            }
        }
    }
}

// Interpreter flavours: a program is either in the one-output format of round 1 (MODE_LEGACY: OP_MAC records) or in the
// two-output format (MODE_MAC2: every multiply-accumulate is an OP_MAC2 record); each kernel instantiation contains one
// multiply-accumulate body.  MODE_BOTH (CPU emulation) accepts either.
enum : int { MODE_LEGACY = 0, MODE_MAC2 = 1, MODE_BOTH = 2 };

// TRACE: ph[0..4] receive the cycles spent in operand loads, multiply-accumulates, the Montgomery reduction,
// epilogue + correction, and the store (debug tracing only; the normal instantiation has no clock reads)
template <bool TRACE = false, int MODE = MODE_BOTH, class WordFn>
FPC_DEV void exec_record(const Ctx& c, uint32_t hdr, uint32_t aux, WordFn W, uint32_t* ph = nullptr) {
    const uint32_t op = hdr & 0xFF;
    if (op == OP_NOP) return;
    if (MODE != MODE_LEGACY && op == OP_MAC2) { exec_mac2(c, hdr, aux, W); return; }
    const uint32_t dst = (hdr >> 8) & 0xFF;
    const uint32_t T = (hdr >> 16) & 0xF;
    const uint32_t E = (hdr >> 20) & 0x3;
    const int ncorr = (int)((hdr >> 22) & 0x7);
    const uint32_t xmask = (aux >> 16) & 0xFF;

    uint32_t r[12];
    if (op == OP_SEL) {
        uint32_t f[12], b2[12];
        load_operand(f, c, W(2), xmask);
        load_operand(r, c, W(3), xmask);
        load_operand(b2, c, W(4), xmask);
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < 12; ++k) any |= f[k];
#pragma unroll
        for (int k = 0; k < 12; ++k) r[k] = any ? r[k] : b2[k];
    } else if (op == OP_INV) {
        uint32_t xin[12];
        load_operand(xin, c, W(2), xmask);
        fpc::fp_inv_mont(r, xin);
    } else if (op == OP_BIT) {
        const uint32_t w2 = W(2), bit = W(3), nbytes = W(4);
        const uint8_t* p = wire_record(c, w2 & 0xFF) + ((w2 >> 8) & 0xFF) * 16u;
        const uint32_t byte = p[nbytes - 1 - (bit >> 3)];
        fpc::zero12(r);
        r[0] = (byte >> (bit & 7)) & 1u;
    } else if (MODE != MODE_MAC2 && T > 0) {
        fpc::Acc A;
        fpc::acc_zero(A);
        uint32_t nwx = W(2), nwy = W(3);
        for (uint32_t t = 0; t < T; ++t) {
            uint32_t x[12], y[12];
            uint32_t k0 = 0, k1 = 0;
            if (TRACE) k0 = phase_clock();
            const uint32_t wx = nwx, wy = nwy;
            // the operand words of the NEXT product are fetched (warp shuffles) before this product's loads, so their
            // latency is hidden behind the multiply-accumulate (words 26.. of the last iteration are never used as terms)
            nwx = W(4 + 2 * t);
            nwy = W(5 + 2 * t);
            load_operand_pair(x, y, c, wx, wy, xmask);
            if (TRACE) { k1 = phase_clock() + (x[0] & y[0] & 0u); ph[0] += k1 - k0; }
            fpc::acc_mac(A, x, y);
            if (TRACE) ph[1] += phase_clock() + fpc::acc_dep(A) - k1;
        }
        uint32_t k2 = 0;
        if (TRACE) k2 = phase_clock();
        fpc::acc_redc(A, r);
        if (TRACE) ph[2] += phase_clock() + (r[11] & 0u) - k2;
        if (!(hdr & H_PADCONST)) {  // common integer factor of all products, applied once to the reduced sum
            const uint32_t ps = aux >> 24;
            if (ps > 1) scale_raw(r, (int)ps);
        }
    } else {
        fpc::zero12(r);
    }
    uint32_t k3 = 0;
    if (TRACE) k3 = phase_clock();
    for (uint32_t e = 0; e < E; ++e) {
        uint32_t z[12];
        load_operand(z, c, W(26 + 2 * e), xmask);
        (void)fpc::add12(r, r, z);
    }
    fpc::correct(r, ncorr);
    uint32_t k4 = 0;
    if (TRACE) { k4 = phase_clock() + (r[11] & 0u); ph[3] += k4 - k3; }
    post_and_store(c, r, hdr, aux, dst);
    if (TRACE) ph[4] += phase_clock() - k4;
}

}  // namespace vm
