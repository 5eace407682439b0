// CPU emulation of the tower-VM interpreter (noble_bls12_381_b200/csrc/vm.cuh) -- TEST INFRASTRUCTURE.
// Compiles the *same* exec_record()/fp_core source with g++ (portable bodies of the carry chains) so
// that programs produced by the builder, and the interpreter logic itself, can be validated on a box
// without a GPU.  Never linked into the shipped library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../noble_bls12_381_b200/csrc/vm.cuh"
#include "../../noble_bls12_381_b200/csrc/fp_inv.cuh"
#include "../../noble_bls12_381_b200/csrc/swu_g2.cuh"
#include "../../noble_bls12_381_b200/csrc/g1_comb_gen.cuh"
#include "../../noble_bls12_381_b200/csrc/g2_kernels.cuh"

extern "C" {

// r = a*b/R mod p
void emu_mont_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { fpc::mont_mul(r, a, b); }

// r = (sum a_i*b_i)/R, unreduced, then `rounds` correction rounds
void emu_mac_redc(int n, const uint32_t* a, const uint32_t* b, int rounds, uint32_t* r) {
    fpc::Acc A;
    fpc::acc_zero(A);
    for (int i = 0; i < n; ++i) fpc::acc_mac(A, a + 12 * i, b + 12 * i);
    fpc::acc_redc(A, r);
    fpc::correct(r, rounds);
}

// n Montgomery-form inversions (fp_inv.cuh)
void emu_fp_inv(int n, const uint32_t* x, uint32_t* r) { for (int i = 0; i < n; ++i) fpc::fp_inv_mont(r + 12 * i, x + 12 * i); }

// hash_to_field + SWU for G2 (csrc/swu_g2.cuh): n field elements, 128 uniform bytes in, 288 bytes out each
void emu_swu_g2(const uint8_t* in, uint8_t* out, size_t n) { for (size_t i = 0; i < n; ++i) swu::swu_g2_one(in + 128 * i, out + 288 * i); }

// csrc/g2_kernels.cuh: tail of hash-to-curve (576 B -> 192 B affine) and sign (576 B + 32 B digits -> 96 B signature)
void emu_h2g2_tail(const uint8_t* in, uint8_t* out, size_t n) { for (size_t i = 0; i < n; ++i) swu::h2g2_tail_one(in + 576 * i, out + 192 * i); }
void emu_sign_tail(const uint8_t* in, const uint8_t* digits, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) swu::sign_one(in + 576 * i, digits + 32 * i, out + 96 * i);
}

void emu_g1_decompress(const uint8_t* in, uint8_t* out, int32_t* st, size_t n) {
    for (size_t i = 0; i < n; ++i) swu::g1_decompress_one(in + 48 * i, out + 96 * i, st + i);
}
void emu_g1_fixed_base(const uint8_t* sk32, uint8_t* out, int32_t* flags, size_t n) {
    for (size_t i = 0; i < n; ++i) swu::g1_fixed_base_one(sk32 + 32 * i, swu::kG1Comb, out + 96 * i, flags + i);
}
void emu_validate(int g2, const uint8_t* in, int32_t* st, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (g2) swu::g2_validate_one(in + 192 * i, st + i);
        else swu::g1_validate_one(in + 96 * i, st + i);
    }
}
void emu_g2_decompress(const uint8_t* in, uint8_t* out, int32_t* st, size_t n) {
    for (size_t i = 0; i < n; ++i) swu::g2_decompress_one(in + 96 * i, out + 192 * i, st + i);
}
// products / reductions executed by the swu / g2 kernels' source since the last call (host-build counters)
void emu_swu_counters(long* out2) { out2[0] = swu::g_products; out2[1] = swu::g_reductions; swu::g_products = swu::g_reductions = 0; }

void emu_add_mod(const uint32_t* a, const uint32_t* b, uint32_t* r) { fpc::add_mod(r, a, b); }
void emu_sub_mod(const uint32_t* a, const uint32_t* b, uint32_t* r) { fpc::sub_mod(r, a, b); }

// Runs a program on n_items items with the dataflow semantics of the kernel: a warp may execute its next
// record only when the progress counters it names are reached.  `policy` picks which runnable warp goes next
// (0 round-robin, 1 highest index first, 2 pseudo-random) so that missing waits show up as wrong results.
int vm_emu_run(const uint32_t* prog, int warps, int nrec, const uint32_t* consts, int nconst, int nslots,
               int nfar, int n_items, uint8_t** bases, const uint32_t* strides, int nbuf, int policy) {
    vm::Buffer buf[vm::kMaxBuffers];
    memset(buf, 0, sizeof(buf));
    for (int i = 0; i < nbuf && i < vm::kMaxBuffers; ++i) { buf[i].base = bases[i]; buf[i].stride = strides[i]; }
    std::vector<uint32_t> slots((size_t)nslots * vm::kSlotWords), far((size_t)(nfar ? nfar : 1) * vm::kSlotWords);
    int nbatch = (n_items + 31) / 32;
    uint32_t rng = 12345;
    for (int batch = 0; batch < nbatch; ++batch) {
        std::vector<uint32_t> pc(warps, 0);
        for (;;) {
            std::vector<int> runnable;
            bool all_done = true;
            for (int w = 0; w < warps; ++w) {
                if (pc[w] >= (uint32_t)nrec) continue;
                all_done = false;
                const uint32_t* rec = prog + ((size_t)w * nrec + pc[w]) * vm::kRecWords;
                bool ok = true;
                if (rec[0] & vm::H_BAR) {
                    const uint32_t ww[4] = {rec[vm::kWaitWord], rec[vm::kWaitWord + 1], rec[vm::kWaitWord + 2], rec[vm::kWaitWord + 3]};
                    const int fw = warps <= 8 ? 16 : (warps <= 10 ? 12 : 10);  // field width, as in vm_kernel.cu
                    const unsigned __int128 big = ((unsigned __int128)ww[3] << 96) | ((unsigned __int128)ww[2] << 64) |
                                                  ((unsigned __int128)ww[1] << 32) | ww[0];
                    for (int k = 0; k < warps; ++k) {
                        uint32_t need = (uint32_t)(big >> (fw * k)) & ((1u << fw) - 1u);
                        if (pc[k] < need) ok = false;
                    }
                }
                if (ok) runnable.push_back(w);
            }
            if (all_done) break;
            if (runnable.empty()) { fprintf(stderr, "vm_emu: deadlock\n"); return -2; }
            int w;
            if (policy == 0) w = runnable[0];
            else if (policy == 1) w = runnable.back();
            else { rng = rng * 1664525u + 1013904223u; w = runnable[(rng >> 16) % runnable.size()]; }
            const uint32_t* rec = prog + ((size_t)w * nrec + pc[w]) * vm::kRecWords;
            vm::Cold cold;
            memset(&cold, 0, sizeof(cold));
            cold.nslots = (uint32_t)nslots; cold.far = far.data(); cold.consts = consts; cold.stage = nullptr;
            cold.n_items = (uint32_t)n_items; cold.batch = (uint32_t)batch;
            for (int i = 0; i < vm::kMaxBuffers; ++i) cold.buf[i] = buf[i];
            for (uint32_t lane = 0; lane < 32; ++lane) {
                vm::Ctx c;
                c.slots = slots.data(); c.cold = &cold;
                c.lane = lane; c.sbase = 0; c.cbase = 0;
                vm::exec_record<false, vm::MODE_BOTH>(c, rec[0], rec[1], [&](uint32_t i) { return rec[i]; });
            }
            ++pc[w];
        }
    }
    return 0;
}

}  // extern "C"
