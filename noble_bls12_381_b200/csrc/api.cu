// C ABI of the engine (include/bls381_b200.h).  Host-side runtime: program registry, device staging
// pools, launch of the tower-VM kernel, product tree for the multi-Miller entry points.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bls381_b200.h"
#include "vm_kernel.cu"  // single translation unit: the interpreter kernel
#include "sha256_xmd.cuh"
#include "swu_g2.cuh"     // hash_to_field + SWU map for G2 as a plain kernel (the serial square-root chains)
#include "g2_kernels.cuh"  // tail of hash-to-curve and the sign ladder as plain kernels (a thread per item)

namespace {

struct Program {
    uint32_t warps = 0, nrec = 0, nconst = 0, nslots = 0, nfar = 0, staged_mask = 0;
    bool mac2 = false;  // two-output record format (header word 7, bit 31)
    uint32_t* d_prog = nullptr;
    uint32_t* d_consts = nullptr;
};

struct State {
    std::mutex mu;
    bool inited = false;
    int device = -1;
    int sm_count = 0;
    std::map<std::string, Program> programs;
    // far-slot scratch (global memory that stays in L2), ONE BUFFER PER STREAM: launches on different streams can overlap on
    // the device (a persistent grid's tail), launches on one stream cannot
    struct FarBuf { uint32_t* ptr = nullptr; size_t bytes = 0; };
    std::map<cudaStream_t, FarBuf> far;
    // partial products + product-tree ping-pong of the Miller-product entry points, ONE PAIR PER STREAM as well: a `_dev`
    // call returns without synchronising, so a second call on another stream must not reuse its buffers
    struct TreeBuf { uint8_t* ptr[2] = {nullptr, nullptr}; size_t bytes[2] = {0, 0}; };
    std::map<cudaStream_t, TreeBuf> tree;
    static constexpr int kTickets = 4096;
    int dynamic_batches = 1;  // batches claimed from a global counter (+4.9 % at 65536 pairings, profiles/r1_notes.md)
    unsigned long long* d_clk = nullptr;  // clock probe of the last tower-VM launch {cycles, ns}
    // grow-only device staging for the host entry points
    static constexpr int kStages = 11;
    uint8_t* d_stage[kStages] = {};
    size_t stage_bytes[kStages] = {};
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    double last_ms = 0.0;
    std::atomic<uint64_t> launches{0};
    std::string program_dir;
    int force_ctas = 0;  // tuning override (env BLS381_B200_CTAS)
    int sleep_ns = 0;    // back-off of the dataflow poll loop (env BLS381_B200_SLEEP_NS)
    int no_tma = 0;      // disable TMA staging of the inputs (env BLS381_B200_NO_TMA, A/B testing)
    int swu_kernel = 1;  // hash-to-curve front (hash_to_field + SWU) as the hand-written kernel; 0 = all inside the tower-VM program (A/B)
    int g1_kernel = 1;   // G1 key decompression + subgroup check as a hand-written kernel; 0 = tower-VM program g1_decompress (A/B)
    int g2_kernel = 1;  // batches of compressed signatures (fromSignature + assertValidity) as the hand-written per-item kernel; 0 = tower-VM program g2_decompress (A/B)
    int pipeline_copies = 1;  // pairing_batch from host buffers: chunked H2D / kernel / D2H on two streams (0 = one copy in, one launch, one copy out)
    int tail_kernels = 1;  // tail of hash-to-curve / sign ladder as hand-written kernels (needs swu_kernel); 0 = tower-VM programs h2g2_tail / sign_tail (A/B)
    // Miller-product lanes handle this many items (1..4) with shared Fp12 squarings (env BLS381_B200_PAIRS_PER_LANE);
    // verifyBatch at 131072 signatures on a B200: 1.26 / 1.39 / 1.42 / 1.40 M sigs/s for 1 / 2 / 3 / 4
    int pairs_per_lane = 3;
    int log_launches = 0;  // env BLS381_B200_LOG_LAUNCHES: one stderr line per tower-VM launch
};

// One context per device.  Context 0 is the device of bls381_init(); bls381_init_devices() adds one context per further
// device for the in-process multi-GPU entry points.  Every entry point works on the calling thread's current context
// (`g`): context 0, except inside the worker threads of the multi-GPU calls.
constexpr int kMaxDevices = 16;
State g_states[kMaxDevices];
int g_ncontexts = 0;
thread_local State* g_cur = &g_states[0];
#define g (*g_cur)
thread_local std::string g_err;

// serialise the calls of one context and make its device current for the calling thread
struct Enter {
    std::lock_guard<std::mutex> lk;
    Enter() : lk(g.mu) {
        if (g.inited) cudaSetDevice(g.device);
    }
};

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(BLS381_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));      \
    } while (0)

constexpr uint32_t kMagic = 0x4D563242u;

int load_image(const std::string& name, const uint8_t* img, size_t len) {
    if (len < 32) return fail(BLS381_EPROGRAM, "program image too short: " + name);
    uint32_t h[8];
    memcpy(h, img, 32);
    if (h[0] != kMagic || h[1] != 2) return fail(BLS381_EPROGRAM, "bad program magic/version: " + name);
    Program p;
    p.warps = h[2]; p.nrec = h[3]; p.nconst = h[4]; p.nslots = h[5]; p.nfar = h[6]; p.staged_mask = h[7] & 0xFFu; p.mac2 = (h[7] >> 31) != 0;
    const size_t cbytes = (size_t)p.nconst * 48, pbytes = (size_t)p.warps * p.nrec * vm::kRecWords * 4;
    if (len != 32 + cbytes + pbytes) return fail(BLS381_EPROGRAM, "program image size mismatch: " + name);
    if (p.warps != 2 && p.warps != 4 && p.warps != 6 && p.warps != 8) return fail(BLS381_EPROGRAM, "unsupported warp count: " + name);
    if (p.nrec == 0 || p.nrec > 32767u) return fail(BLS381_EPROGRAM, "record count out of range for the progress-requirement fields: " + name);
    CUDA_TRY(cudaMalloc(&p.d_consts, std::max<size_t>(cbytes, 48)));
    CUDA_TRY(cudaMalloc(&p.d_prog, pbytes));
    CUDA_TRY(cudaMemcpy(p.d_consts, img + 32, cbytes, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p.d_prog, img + 32 + cbytes, pbytes, cudaMemcpyHostToDevice));
    auto it = g.programs.find(name);
    if (it != g.programs.end()) {
        cudaFree(it->second.d_consts);
        cudaFree(it->second.d_prog);
    }
    g.programs[name] = p;
    return BLS381_OK;
}

int load_file(const std::string& name) {
    const std::string path = g.program_dir + "/" + name + ".b2vm";
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return fail(BLS381_EPROGRAM, "cannot open " + path + " (run __graft_entry__.build())");
    std::vector<uint8_t> buf;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    size_t rd = fread(buf.data(), 1, (size_t)n, f);
    fclose(f);
    if (rd != (size_t)n) return fail(BLS381_EPROGRAM, "short read " + path);
    return load_image(name, buf.data(), buf.size());
}

int get_program(const char* name, Program** out) {
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    auto it = g.programs.find(name);
    if (it == g.programs.end()) {
        int rc = load_file(name);
        if (rc) return rc;
        it = g.programs.find(name);
    }
    *out = &it->second;
    return BLS381_OK;
}

template <int W, int MINB, int MODE>
int launch_m(const vm::Launch& L, int grid, size_t smem, cudaStream_t s) {
    CUDA_TRY(cudaFuncSetAttribute(vm::vm_kernel<W, MINB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vm::vm_kernel<W, MINB, MODE><<<grid, W * 32, smem, s>>>(L);
    CUDA_TRY(cudaGetLastError());
    return BLS381_OK;
}

template <int W, int MINB>
int launch_w(const vm::Launch& L, bool mac2, int grid, size_t smem, cudaStream_t s) {
    return mac2 ? launch_m<W, MINB, vm::MODE_MAC2>(L, grid, smem, s) : launch_m<W, MINB, vm::MODE_LEGACY>(L, grid, smem, s);
}

int vm_run(const char* name, uint8_t* const* bufs, const uint32_t* strides, int nbuf, size_t n, cudaStream_t s) {
    if (n == 0) return BLS381_OK;
    if (n > 0x7fffffffull) return fail(BLS381_EINVAL, "too many items");
    if (nbuf < 0 || nbuf > vm::kMaxBuffers) return fail(BLS381_EINVAL, "bad buffer count");
    Program* p = nullptr;
    int rc = get_program(name, &p);
    if (rc) return rc;
    const uint32_t nbatch = (uint32_t)((n + 31) / 32);
    // CTAs per SM from the program's own shared memory (slots + constants + progress + mbarrier); the TMA staging of the
    // wire-format inputs then only takes what is left of an SM's share, so that staging never costs a resident CTA
    const size_t smem_base = (size_t)p->nslots * vm::kSlotWords * 4 + (size_t)p->nconst * 48 + 64 + 16;
    const int max_ctas = p->warps == 2 ? 8 : (p->warps == 4 ? 4 : (p->warps == 6 ? 3 : 2));
    int ctas_per_sm = std::max(1, std::min(max_ctas, (int)(232448 / (smem_base + 1024))));
    if (g.force_ctas > 0) ctas_per_sm = std::min(ctas_per_sm, g.force_ctas);
    const size_t per_cta = 232448 / (size_t)ctas_per_sm;
    const size_t stage_budget = std::min<size_t>(per_cta > smem_base + 1024 ? per_cta - smem_base - 1024 : 0, 20480) & ~(size_t)15;
    const int grid = (int)std::min<uint32_t>(nbatch, (uint32_t)(g.sm_count * ctas_per_sm));
    const size_t far_need = (size_t)grid * std::max<uint32_t>(p->nfar, 1) * vm::kSlotWords * 4;
    if (g.far.size() > 64 && !g.far.count(s)) {  // many short-lived caller streams: start over
        CUDA_TRY(cudaDeviceSynchronize());
        for (auto& kv : g.far)
            if (kv.second.ptr) cudaFree(kv.second.ptr);
        g.far.clear();
    }
    State::FarBuf& fb = g.far[s];
    if (far_need > fb.bytes) {
        if (fb.ptr) {
            CUDA_TRY(cudaStreamSynchronize(s));  // earlier launches on this stream may still use the old buffer
            cudaFree(fb.ptr);
        }
        fb.ptr = nullptr;
        fb.bytes = 0;
        CUDA_TRY(cudaMalloc(&fb.ptr, far_need));
        fb.bytes = far_need;
    }
    vm::Launch L;
    memset(&L, 0, sizeof(L));
    L.prog = p->d_prog;
    L.consts = p->d_consts;
    L.nrec = p->nrec;
    L.nconst = p->nconst;
    L.nslots = p->nslots;
    L.nfar = std::max<uint32_t>(p->nfar, 1);
    L.n_items = (uint32_t)n;
    L.pad = (uint32_t)g.sleep_ns;
    L.far = fb.ptr;
    for (int i = 0; i < nbuf; ++i) {
        L.buf[i].base = bufs[i];
        L.buf[i].stride = strides[i];
    }
    // TMA staging plan: every wire-format input buffer the program reads, if its pointer and stride allow 16-byte
    // aligned bulk copies (otherwise that buffer is read directly from global memory)
    uint32_t stage_bytes = 0;
    for (int i = 0; i < vm::kMaxBuffers; ++i) {
        L.stage_off[i] = vm::kNoStage;
        if (!g.no_tma && i < nbuf && (p->staged_mask >> i & 1) && bufs[i] && strides[i] && strides[i] % 16 == 0 &&
            reinterpret_cast<uintptr_t>(bufs[i]) % 16 == 0 && stage_bytes + 32u * strides[i] <= stage_budget) {
            L.stage_off[i] = stage_bytes;
            stage_bytes += 32u * strides[i];
        }
    }
    L.stage_bytes = stage_bytes;
    const size_t smem = (size_t)p->nslots * vm::kSlotWords * 4 + (size_t)p->nconst * 48 + 64 + 16 + stage_bytes;  // slots + constants + progress + mbarrier + staging
    // debug tracing (env BLS381_B200_TRACE=<file prefix>): clocks of every record of the first two CTAs
    uint32_t* d_trace = nullptr;
    uint32_t* d_log = nullptr;
    const char* trace_prefix = getenv("BLS381_B200_TRACE");
    const uint32_t trace_ctas = 2;
    if (trace_prefix && *trace_prefix) {
        CUDA_TRY(cudaMalloc(&d_trace, (size_t)trace_ctas * p->warps * p->nrec * 32));
        CUDA_TRY(cudaMemset(d_trace, 0, (size_t)trace_ctas * p->warps * p->nrec * 32));
        L.trace = d_trace;
        L.trace_ctas = trace_ctas;
        CUDA_TRY(cudaMalloc(&d_log, (size_t)grid * 16 * 16));
        CUDA_TRY(cudaMemset(d_log, 0, (size_t)grid * 16 * 16));
        L.cta_log = d_log;
    }
    L.clk = g.d_clk;
    if (g.dynamic_batches) {  // one counter per launch (a ring: launches on the two internal streams may overlap)
        L.ticket = reinterpret_cast<uint32_t*>(g.d_clk + 2) + (g.launches.load() % State::kTickets);
        CUDA_TRY(cudaMemsetAsync(L.ticket, 0, 4, s));
    }
    g.launches.fetch_add(1);
    if (g.log_launches)  // profiling aid: the i-th vm_kernel launch of an ncu report is the i-th line of this log
        fprintf(stderr, "[vm_run] %s n=%zu grid=%d warps=%u ctas_per_sm=%d nrec=%u slots=%u far=%u\n", name, n, grid, p->warps, ctas_per_sm, p->nrec,
                p->nslots, p->nfar);
    rc = BLS381_EPROGRAM;
    if (p->warps == 2) rc = launch_w<2, 8>(L, p->mac2, grid, smem, s);
    if (p->warps == 4) rc = launch_w<4, 4>(L, p->mac2, grid, smem, s);
    if (p->warps == 6) rc = launch_w<6, 3>(L, p->mac2, grid, smem, s);
    if (p->warps == 8) rc = launch_w<8, 2>(L, p->mac2, grid, smem, s);
    if (d_trace) {
        cudaStreamSynchronize(s);
        std::vector<uint32_t> h((size_t)trace_ctas * p->warps * p->nrec * 8);
        cudaMemcpy(h.data(), d_trace, h.size() * 4, cudaMemcpyDeviceToHost);
        cudaFree(d_trace);
        const std::string path = std::string(trace_prefix) + "_" + name + ".bin";
        if (FILE* f = fopen(path.c_str(), "wb")) {
            uint32_t hdr4[4] = {trace_ctas, p->warps, p->nrec, 8};
            fwrite(hdr4, 4, 4, f);
            fwrite(h.data(), 4, h.size(), f);
            fclose(f);
        }
        std::vector<uint32_t> hl((size_t)grid * 16 * 4);
        cudaMemcpy(hl.data(), d_log, hl.size() * 4, cudaMemcpyDeviceToHost);
        cudaFree(d_log);
        if (FILE* f = fopen((std::string(trace_prefix) + "_" + name + "_ctas.bin").c_str(), "wb")) {
            uint32_t hdr4[4] = {(uint32_t)grid, 16, 4, (uint32_t)n};
            fwrite(hdr4, 4, 4, f);
            fwrite(hl.data(), 4, hl.size(), f);
            fclose(f);
        }
    }
    if (rc != BLS381_EPROGRAM) return rc;
    return fail(BLS381_EPROGRAM, "unsupported warp count");
}

// per-stream scratch `which` (0: partial products, 1: tree ping-pong) of at least `bytes`
int tree_buf(cudaStream_t s, int which, size_t bytes, uint8_t** out) {
    if (g.tree.size() > 64 && !g.tree.count(s)) {  // many short-lived caller streams: start over
        CUDA_TRY(cudaDeviceSynchronize());
        for (auto& kv : g.tree)
            for (int k = 0; k < 2; ++k)
                if (kv.second.ptr[k]) cudaFree(kv.second.ptr[k]);
        g.tree.clear();
    }
    State::TreeBuf& tb = g.tree[s];
    if (bytes > tb.bytes[which]) {
        if (tb.ptr[which]) {
            CUDA_TRY(cudaStreamSynchronize(s));  // earlier launches on this stream may still use the old buffer
            cudaFree(tb.ptr[which]);
        }
        tb.ptr[which] = nullptr;
        tb.bytes[which] = 0;
        const size_t cap = std::max<size_t>(bytes, 1 << 16);
        CUDA_TRY(cudaMalloc(&tb.ptr[which], cap));
        tb.bytes[which] = cap;
    }
    *out = tb.ptr[which];
    return BLS381_OK;
}

int stage(int k, size_t bytes) {
    if (bytes > g.stage_bytes[k]) {
        if (g.d_stage[k]) cudaFree(g.d_stage[k]);
        g.d_stage[k] = nullptr;
        g.stage_bytes[k] = 0;
        size_t cap = std::max<size_t>(bytes, 1 << 20);
        CUDA_TRY(cudaMalloc(&g.d_stage[k], cap));
        g.stage_bytes[k] = cap;
    }
    return BLS381_OK;
}

// product tree: d_partials[count x 576] -> d_out (576 B) ; uses d_tmp as ping-pong
int product_tree(uint8_t* d_a, size_t count, uint8_t* d_b, uint8_t** result, cudaStream_t s) {
    uint8_t* cur = d_a;
    uint8_t* nxt = d_b;
    while (count > 1) {
        uint8_t* bufs[4] = {nullptr, nullptr, nxt, cur};
        uint32_t strides[4] = {0, 0, 576, 576};
        int rc = vm_run("f12_product", bufs, strides, 4, count, s);
        if (rc) return rc;
        count = (count + 31) / 32;
        std::swap(cur, nxt);
    }
    *result = cur;
    return BLS381_OK;
}

// affine generators (math.ts:18-21, 34-45): neutral padding pairs e(G1, G2) * e(-G1, G2) = 1, public-key base point
const uint8_t kG1[96] = {
    0x17, 0xf1, 0xd3, 0xa7, 0x31, 0x97, 0xd7, 0x94, 0x26, 0x95, 0x63, 0x8c, 0x4f, 0xa9, 0xac, 0x0f, 0xc3, 0x68, 0x8c, 0x4f, 0x97, 0x74, 0xb9, 0x05, 0xa1, 0x4e, 0x3a, 0x3f, 0x17, 0x1b, 0xac, 0x58, 0x6c, 0x55, 0xe8, 0x3f, 0xf9, 0x7a, 0x1a, 0xef, 0xfb, 0x3a, 0xf0, 0x0a, 0xdb, 0x22, 0xc6, 0xbb,
    0x08, 0xb3, 0xf4, 0x81, 0xe3, 0xaa, 0xa0, 0xf1, 0xa0, 0x9e, 0x30, 0xed, 0x74, 0x1d, 0x8a, 0xe4, 0xfc, 0xf5, 0xe0, 0x95, 0xd5, 0xd0, 0x0a, 0xf6, 0x00, 0xdb, 0x18, 0xcb, 0x2c, 0x04, 0xb3, 0xed, 0xd0, 0x3c, 0xc7, 0x44, 0xa2, 0x88, 0x8a, 0xe4, 0x0c, 0xaa, 0x23, 0x29, 0x46, 0xc5, 0xe7, 0xe1};
const uint8_t kNegG1[96] = {  // -G1_BASE affine (x, p - y)
    0x17, 0xf1, 0xd3, 0xa7, 0x31, 0x97, 0xd7, 0x94, 0x26, 0x95, 0x63, 0x8c, 0x4f, 0xa9, 0xac, 0x0f, 0xc3, 0x68, 0x8c, 0x4f, 0x97, 0x74, 0xb9, 0x05, 0xa1, 0x4e, 0x3a, 0x3f, 0x17, 0x1b, 0xac, 0x58, 0x6c, 0x55, 0xe8, 0x3f, 0xf9, 0x7a, 0x1a, 0xef, 0xfb, 0x3a, 0xf0, 0x0a, 0xdb, 0x22, 0xc6, 0xbb, 0x11, 0x4d, 0x1d, 0x68, 0x55, 0xd5, 0x45, 0xa8, 0xaa, 0x7d, 0x76, 0xc8, 0xcf, 0x2e, 0x21, 0xf2, 0x67, 0x81, 0x6a, 0xef, 0x1d, 0xb5, 0x07, 0xc9, 0x66, 0x55, 0xb9, 0xd5, 0xca, 0xac, 0x42, 0x36, 0x4e, 0x6f, 0x38, 0xba, 0x0e, 0xcb, 0x75, 0x1b, 0xad, 0x54, 0xdc, 0xd6, 0xb9, 0x39, 0xc2, 0xca};
const uint8_t kG2[192] = {
    0x02, 0x4a, 0xa2, 0xb2, 0xf0, 0x8f, 0x0a, 0x91, 0x26, 0x08, 0x05, 0x27, 0x2d, 0xc5, 0x10, 0x51, 0xc6, 0xe4, 0x7a, 0xd4, 0xfa, 0x40, 0x3b, 0x02, 0xb4, 0x51, 0x0b, 0x64, 0x7a, 0xe3, 0xd1, 0x77, 0x0b, 0xac, 0x03, 0x26, 0xa8, 0x05, 0xbb, 0xef, 0xd4, 0x80, 0x56, 0xc8, 0xc1, 0x21, 0xbd, 0xb8,
    0x13, 0xe0, 0x2b, 0x60, 0x52, 0x71, 0x9f, 0x60, 0x7d, 0xac, 0xd3, 0xa0, 0x88, 0x27, 0x4f, 0x65, 0x59, 0x6b, 0xd0, 0xd0, 0x99, 0x20, 0xb6, 0x1a, 0xb5, 0xda, 0x61, 0xbb, 0xdc, 0x7f, 0x50, 0x49, 0x33, 0x4c, 0xf1, 0x12, 0x13, 0x94, 0x5d, 0x57, 0xe5, 0xac, 0x7d, 0x05, 0x5d, 0x04, 0x2b, 0x7e,
    0x0c, 0xe5, 0xd5, 0x27, 0x72, 0x7d, 0x6e, 0x11, 0x8c, 0xc9, 0xcd, 0xc6, 0xda, 0x2e, 0x35, 0x1a, 0xad, 0xfd, 0x9b, 0xaa, 0x8c, 0xbd, 0xd3, 0xa7, 0x6d, 0x42, 0x9a, 0x69, 0x51, 0x60, 0xd1, 0x2c, 0x92, 0x3a, 0xc9, 0xcc, 0x3b, 0xac, 0xa2, 0x89, 0xe1, 0x93, 0x54, 0x86, 0x08, 0xb8, 0x28, 0x01,
    0x06, 0x06, 0xc4, 0xa0, 0x2e, 0xa7, 0x34, 0xcc, 0x32, 0xac, 0xd2, 0xb0, 0x2b, 0xc2, 0x8b, 0x99, 0xcb, 0x3e, 0x28, 0x7e, 0x85, 0xa7, 0x63, 0xaf, 0x26, 0x74, 0x92, 0xab, 0x57, 0x2e, 0x99, 0xab, 0x3f, 0x37, 0x0d, 0x27, 0x5c, 0xec, 0x1d, 0xa1, 0xaa, 0xa9, 0x07, 0x5f, 0xf0, 0x5f, 0x79, 0xbe};

constexpr size_t kMillerPad = 4;  // writable slack (items) a caller that allows in-place padding provides behind its arrays

// prod_i millerLoop(P_i, Q_i) [+ final exponentiation] of n affine pairs in device memory -> 576 B at d_out.
// k consecutive items per lane share the Fp12 squarings of the Miller loop (programs miller_product2/3/4: the product of
// the individual Miller loops, bit for bit).  `writable_slack`: the arrays have kMillerPad writable items behind the n-th;
// the batch is then padded to a multiple of k with neutral pairs (G1, G2), (-G1, G2) -- one launch instead of a second,
// single-CTA, latency-bound launch (~2 ms) for the n mod k last items.
int miller_product_dev(const uint8_t* d_g1, const uint8_t* d_g2, size_t n, int fe, uint8_t* d_out, cudaStream_t s,
                       bool writable_slack = false) {
    if (n == 0) return fail(BLS381_EINVAL, "empty batch");
    size_t k = (size_t)std::min(std::max(g.pairs_per_lane, 1), 4);
    int rc;
    // Padding changes the UN-exponentiated product by a factor the final exponentiation kills, so it is only used when the
    // final exponentiation follows (fe); a shard of a multi-GPU batch (fe = 0: the 576 bytes are exchanged and must not
    // depend on the sharding) instead picks a k that divides its item count when one of 3, 4, 2 does.
    if (!fe && k >= 2 && n % k != 0) {
        for (size_t cand : {(size_t)3, (size_t)4, (size_t)2})
            if (n % cand == 0) { k = cand; break; }
    }
    if (fe && writable_slack && k >= 2 && n % k != 0) {
        // an even number of padding items (pairs that cancel) up to the next multiple of k: k = 2: +2; k = 3: +2 / +4; k = 4: +2
        size_t pad = 0;
        while ((n + pad) % k != 0 || pad % 2 != 0) ++pad;
        if (pad <= kMillerPad) {
            uint8_t* w1 = const_cast<uint8_t*>(d_g1) + n * 96;
            uint8_t* w2 = const_cast<uint8_t*>(d_g2) + n * 192;
            for (size_t j = 0; j < pad; ++j) {
                CUDA_TRY(cudaMemcpyAsync(w1 + j * 96, (j & 1) ? kNegG1 : kG1, 96, cudaMemcpyHostToDevice, s));
                CUDA_TRY(cudaMemcpyAsync(w2 + j * 192, kG2, 192, cudaMemcpyHostToDevice, s));
            }
            n += pad;
        }
    }
    const size_t nk = k >= 2 ? n / k : 0, n1 = n - k * nk;
    const size_t nbk = (nk + 31) / 32, nb1 = (n1 + 31) / 32, nb = nbk + nb1;
    uint8_t *part = nullptr, *tree = nullptr;
    if ((rc = tree_buf(s, 0, nb * 576, &part))) return rc;
    if ((rc = tree_buf(s, 1, ((nb + 31) / 32) * 576 + 576, &tree))) return rc;
    // The n mod k last items (no padding possible: un-exponentiated product wanted, or read-only caller arrays) run as a
    // single-CTA launch on the second stream, issued BEFORE the bulk launch: it takes one CTA slot for ~2 ms while the
    // persistent bulk grid fills the others, instead of a serial, latency-bound tail behind it.
    const bool side = n1 && nk && s != g.stream2;
    if (n1) {
        cudaStream_t s1 = side ? g.stream2 : s;
        if (side) {
            CUDA_TRY(cudaEventRecord(g.ev_fork, s));
            CUDA_TRY(cudaStreamWaitEvent(g.stream2, g.ev_fork, 0));
        }
        uint8_t* bufs[3] = {const_cast<uint8_t*>(d_g1) + k * nk * 96, const_cast<uint8_t*>(d_g2) + k * nk * 192, part + nbk * 576};
        uint32_t strides[3] = {96, 192, 576};
        if ((rc = vm_run("miller_product", bufs, strides, 3, n1, s1))) return rc;
        if (side) CUDA_TRY(cudaEventRecord(g.ev_join, g.stream2));
    }
    if (nk) {
        static const char* const kProg[5] = {nullptr, nullptr, "miller_product2", "miller_product3", "miller_product4"};
        uint8_t* bufs[3] = {const_cast<uint8_t*>(d_g1), const_cast<uint8_t*>(d_g2), part};
        uint32_t strides[3] = {(uint32_t)(96 * k), (uint32_t)(192 * k), 576};
        if ((rc = vm_run(kProg[k], bufs, strides, 3, nk, s))) return rc;
    }
    if (side) CUDA_TRY(cudaStreamWaitEvent(s, g.ev_join, 0));
    uint8_t* res = nullptr;
    if ((rc = product_tree(part, nb, tree, &res, s))) return rc;
    if (fe) {
        uint8_t* b2[4] = {nullptr, nullptr, d_out, res};
        uint32_t st2[4] = {0, 0, 576, 576};
        return vm_run("final_exp", b2, st2, 4, 1, s);
    }
    CUDA_TRY(cudaMemcpyAsync(d_out, res, 576, cudaMemcpyDeviceToDevice, s));
    return BLS381_OK;
}

// msg_off must start at 0 and be non-decreasing (a bad offset would make the SHA-256 kernel read past the staged buffer)
int check_offsets(const uint64_t* off, size_t n) {
    if (!off) return fail(BLS381_EINVAL, "null message offsets");
    if (off[0] != 0) return fail(BLS381_EINVAL, "msg_off[0] must be 0");
    for (size_t i = 0; i < n; ++i)
        if (off[i + 1] < off[i]) return fail(BLS381_EINVAL, "msg_off must be non-decreasing");
    return BLS381_OK;
}

// ---- ingest: decompression, hash-to-curve, verifyBatch ------------------------------------------------
__global__ void xmd_kernel(const uint8_t* msgs, const uint64_t* off, size_t n, const uint8_t* dst_prime,
                           uint32_t dst_prime_len, uint8_t* out, uint32_t len_in_bytes = 256) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    sha::expand_xmd(msgs + off[i], off[i + 1] - off[i], dst_prime, dst_prime_len, out + (size_t)len_in_bytes * i, len_in_bytes);
}

// DST' = DST || len(DST)  (oversize DSTs are hashed first, index.ts:214)
int make_dst_prime(const uint8_t* dst, size_t dst_len, std::vector<uint8_t>& out) {
    out.clear();
    if (dst_len > 255) {
        sha::Ctx c;
        sha::init(c);
        const char* pre = "H2C-OVERSIZE-DST-";
        sha::update(c, reinterpret_cast<const uint8_t*>(pre), 17);
        sha::update(c, dst, dst_len);
        out.resize(32);
        sha::final(c, out.data());
    } else {
        out.assign(dst, dst + dst_len);
    }
    out.push_back((uint8_t)out.size());
    return BLS381_OK;
}

int run3(const char* prog, uint8_t* in, uint32_t in_stride, uint8_t* out, uint32_t out_stride, int32_t* d_status,
         size_t n, cudaStream_t s) {
    uint8_t* bufs[6] = {in, nullptr, out, nullptr, nullptr, reinterpret_cast<uint8_t*>(d_status)};
    uint32_t strides[6] = {in_stride, 0, out_stride, 0, 0, 4};
    return vm_run(prog, bufs, strides, 6, n, s);
}

// PointG1.fromHex (48-byte keys) + assertValidity: n x 48 B -> n x 96 B affine + status   (index.ts:298-327, 383-388)
int g1_decompress_dev(uint8_t* d_in, uint8_t* d_out, int32_t* d_status, size_t n, cudaStream_t s) {
    if (!g.g1_kernel) return run3("g1_decompress", d_in, 48, d_out, 96, d_status, n, s);
    swu::g1_decompress_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_in, d_out, d_status, n);
    CUDA_TRY(cudaGetLastError());
    g.launches.fetch_add(1);
    return BLS381_OK;
}

// PointG2.fromSignature (96-byte signatures) + assertValidity: n x 96 B -> n x 192 B affine + status   (index.ts:500-530, 633-638)
// Small batches keep the tower-VM program: its four warps work on one item set, a thread of the kernel runs the item alone
// (1.8 ms against ~5 ms for a single signature).
int g2_decompress_dev(uint8_t* d_in, uint8_t* d_out, int32_t* d_status, size_t n, cudaStream_t s) {
    if (!g.g2_kernel || n < 256) return run3("g2_decompress", d_in, 96, d_out, 192, d_status, n, s);
    swu::g2_decompress_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_in, d_out, d_status, n);
    CUDA_TRY(cudaGetLastError());
    g.launches.fetch_add(1);
    return BLS381_OK;
}

// g.d_stage[4] (n x 256 uniform bytes) -> g.d_stage[10] (n x 576 B: the two points of E' per message), csrc/swu_g2.cuh
int swu_points(size_t n, cudaStream_t s) {
    int rc;
    if ((rc = stage(10, n * 576))) return rc;
    swu::swu_g2_kernel<<<(unsigned)((2 * n + 127) / 128), 128, 0, s>>>(g.d_stage[4], g.d_stage[10], 2 * n);
    CUDA_TRY(cudaGetLastError());
    g.launches.fetch_add(1);
    return BLS381_OK;
}

// msgs (device, packed) -> n x 192 B affine H(m) at d_out
int hash_to_g2_dev(const uint8_t* d_msgs, const uint64_t* d_off, size_t n, const uint8_t* dst, size_t dst_len,
                   uint8_t* d_out, cudaStream_t s) {
    std::vector<uint8_t> dp;
    make_dst_prime(dst, dst_len, dp);
    int rc;
    if ((rc = stage(4, n * 256)) || (rc = stage(5, 512))) return rc;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[5], dp.data(), dp.size(), cudaMemcpyHostToDevice, s));
    xmd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_msgs, d_off, n, g.d_stage[5], (uint32_t)dp.size(), g.d_stage[4]);
    CUDA_TRY(cudaGetLastError());
    if (!g.swu_kernel) return run3("hash_to_g2", g.d_stage[4], 256, d_out, 192, nullptr, n, s);
    if ((rc = swu_points(n, s))) return rc;
    if (!g.tail_kernels) return run3("h2g2_tail", g.d_stage[10], 576, d_out, 192, nullptr, n, s);
    swu::h2g2_tail_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g.d_stage[10], d_out, n);
    CUDA_TRY(cudaGetLastError());
    g.launches.fetch_add(1);
    return BLS381_OK;
}

// OR the compression flag bits (index.ts:26-28) into byte 0 of each compressed body
__global__ void apply_flags_kernel(uint8_t* body, uint32_t stride, const int32_t* flags, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t f = flags[i];
    uint8_t* p = body + i * stride;
    if (f & 2) {  // point at infinity: 0xc0 00 .. 00
        for (uint32_t k = 0; k < stride; ++k) p[k] = 0;
        p[0] = 0xC0;
    } else {
        p[0] |= 0x80 | ((f & 1) << 5);
    }
}

// sum of n affine points (+ status words: INFINITY = skip) -> compressed point at d_out (1 record)
int aggregate_dev(bool g2, uint8_t* d_affine, int32_t* d_status, size_t n, uint8_t* d_out, cudaStream_t s) {
    const uint32_t fe = g2 ? 96 : 48;       // bytes per coordinate
    const uint32_t proj = 3 * fe;           // projective record
    int rc;
    size_t nb = (n + 31) / 32;
    if ((rc = stage(2, nb * proj + proj)) || (rc = stage(3, ((nb + 31) / 32) * proj + proj)) || (rc = stage(5, 512))) return rc;
    {
        uint8_t* bufs[3] = {d_affine, reinterpret_cast<uint8_t*>(d_status), g.d_stage[2]};
        uint32_t strides[3] = {2 * fe, 4, proj};
        if ((rc = vm_run(g2 ? "g2_sum_affine" : "g1_sum_affine", bufs, strides, 3, n, s))) return rc;
    }
    uint8_t *cur = g.d_stage[2], *nxt = g.d_stage[3];
    size_t count = nb;
    while (count > 1) {
        uint8_t* bufs[3] = {cur, nullptr, nxt};
        uint32_t strides[3] = {proj, 0, proj};
        if ((rc = vm_run(g2 ? "g2_sum_proj" : "g1_sum_proj", bufs, strides, 3, count, s))) return rc;
        count = (count + 31) / 32;
        std::swap(cur, nxt);
    }
    int32_t* d_flags = reinterpret_cast<int32_t*>(g.d_stage[5]);
    uint8_t* bufs[6] = {cur, nullptr, d_out, nullptr, nullptr, reinterpret_cast<uint8_t*>(d_flags)};
    uint32_t strides[6] = {proj, 0, fe, 0, 0, 4};
    if ((rc = vm_run(g2 ? "g2_compress" : "g1_compress", bufs, strides, 6, 1, s))) return rc;
    apply_flags_kernel<<<1, 32, 0, s>>>(d_out, fe, d_flags, 1);
    CUDA_TRY(cudaGetLastError());
    return BLS381_OK;
}

// ---- IMAD.WIDE issue-rate microbenchmark -------------------------------------------------------
__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t* out, int iters, unsigned long long* clk) {
    unsigned long long probe_c = 0, probe_t = 0;
    const bool probing = clk != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (probing) {
        probe_c = (unsigned long long)clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(probe_t));
    }
    uint64_t acc[4][6];
    uint32_t ct[4] = {0, 0, 0, 0};
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[k][i] = ((uint64_t)(t * 2654435761u + k * 97u + i) << 32) | (t + i);
    uint32_t a0 = t | 1, a1 = t ^ 0x9e3779b9u, a2 = t + 77, a3 = ~t, a4 = t * 3, a5 = t * 5 + 1, b = t * 7 + 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) fpc::chain6(acc[k], ct[k], a0, a1, a2, a3, a4, a5, b);
        b += ct[0];
    }
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 6; ++i) x ^= (uint32_t)acc[k][i] ^ (uint32_t)(acc[k][i] >> 32);
    out[t] = x ^ ct[1] ^ ct[2] ^ ct[3];
    if (probing) {
        unsigned long long tt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
        clk[0] = (unsigned long long)clock64() - probe_c;
        clk[1] = tt - probe_t;
    }
}


// ---- device contexts -------------------------------------------------------------------------------------------
int init_context(int device, const char* program_dir) {   // caller holds g.mu; g = the context to initialise
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(BLS381_ENODEV, "no CUDA device: this engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(BLS381_EINVAL, "bad device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    g.device = device;
    g.sm_count = prop.multiProcessorCount;
    if (program_dir && *program_dir) {
        g.program_dir = program_dir;
    } else {
        Dl_info info;
        if (dladdr((void*)&init_context, &info) && info.dli_fname) {
            std::string p = info.dli_fname;
            size_t k = p.find_last_of('/');
            g.program_dir = (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/programs";
        } else {
            g.program_dir = "programs";
        }
    }
    if (const char* e = getenv("BLS381_B200_CTAS")) g.force_ctas = atoi(e);
    if (const char* e = getenv("BLS381_B200_SLEEP_NS")) g.sleep_ns = atoi(e);
    if (const char* e = getenv("BLS381_B200_NO_TMA")) g.no_tma = atoi(e);
    if (const char* e = getenv("BLS381_B200_SWU_KERNEL")) g.swu_kernel = atoi(e);
    if (const char* e = getenv("BLS381_B200_TAIL_KERNELS")) g.tail_kernels = atoi(e);
    if (const char* e = getenv("BLS381_B200_G2_KERNEL")) g.g2_kernel = atoi(e);
    if (const char* e = getenv("BLS381_B200_G1_KERNEL")) g.g1_kernel = atoi(e);
    if (const char* e = getenv("BLS381_B200_PAIRS_PER_LANE")) g.pairs_per_lane = atoi(e);
    if (const char* e = getenv("BLS381_B200_DYNAMIC")) g.dynamic_batches = atoi(e);  // 0 = static round-robin batches
    if (const char* e = getenv("BLS381_B200_LOG_LAUNCHES")) g.log_launches = atoi(e);
    CUDA_TRY(cudaMalloc(&g.d_clk, 16 + 4 * State::kTickets));  // {cycles, ns} clock probe + batch ticket counters
    CUDA_TRY(cudaMemset(g.d_clk, 0, 16 + 4 * State::kTickets));
    CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&g.stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&g.ev0));
    CUDA_TRY(cudaEventCreate(&g.ev1));
    CUDA_TRY(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&g.ev_join, cudaEventDisableTiming));
    g.inited = true;
    return BLS381_OK;
}

void shutdown_context() {   // caller holds g.mu
    if (!g.inited) return;
    cudaSetDevice(g.device);
    cudaDeviceSynchronize();
    for (auto& kv : g.programs) {
        cudaFree(kv.second.d_consts);
        cudaFree(kv.second.d_prog);
    }
    g.programs.clear();
    for (auto& kv : g.far)
        if (kv.second.ptr) cudaFree(kv.second.ptr);
    g.far.clear();
    for (auto& kv : g.tree)
        for (int k = 0; k < 2; ++k)
            if (kv.second.ptr[k]) cudaFree(kv.second.ptr[k]);
    g.tree.clear();
    if (g.d_clk) cudaFree(g.d_clk);
    g.d_clk = nullptr;
    for (int k = 0; k < State::kStages; ++k) {
        if (g.d_stage[k]) cudaFree(g.d_stage[k]);
        g.d_stage[k] = nullptr;
        g.stage_bytes[k] = 0;
    }
    cudaEventDestroy(g.ev0);
    cudaEventDestroy(g.ev1);
    cudaEventDestroy(g.ev_fork);
    cudaEventDestroy(g.ev_join);
    cudaStreamDestroy(g.stream);
    cudaStreamDestroy(g.stream2);
    g.inited = false;
}

// ---- NCCL, loaded at run time (dlopen): the library itself has no link-time dependency on it, and a host process that
// already carries its own copy (PyTorch bundles one) is not disturbed.  Only the handful of entry points the single
// exchange step of a sharded verifyBatch needs.
typedef struct ncclComm* nccl_comm_t;
struct Nccl {
    void* handle = nullptr;
    int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    nccl_comm_t comms[kMaxDevices] = {};
    int ncomms = 0;  // number of devices the communicators were created for
    bool ready = false;
    std::string why;  // why NCCL is not used (peer copies instead)
};
Nccl g_nccl;
constexpr int kNcclUint8 = 1;  // ncclUint8 (nccl.h: ncclInt8 = 0, ncclUint8 = 1)

void nccl_load() {
    if (g_nccl.handle) return;
    const char* names[] = {getenv("BLS381_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) { g_nccl.why = "libnccl.so.2 not found"; return; }
    g_nccl.CommInitAll = (int (*)(nccl_comm_t*, int, const int*))dlsym(g_nccl.handle, "ncclCommInitAll");
    g_nccl.CommDestroy = (int (*)(nccl_comm_t))dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t))dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.GroupStart = (int (*)())dlsym(g_nccl.handle, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(g_nccl.handle, "ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.handle, "ncclGetErrorString");
    if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        g_nccl.why = "NCCL symbols missing";
        dlclose(g_nccl.handle);
        g_nccl.handle = nullptr;
    }
}

void multi_shutdown() {
    if (g_nccl.ready) {
        for (int j = 0; j < g_nccl.ncomms; ++j)
            if (g_nccl.comms[j]) g_nccl.CommDestroy(g_nccl.comms[j]);
        for (int j = 0; j < kMaxDevices; ++j) g_nccl.comms[j] = nullptr;
        g_nccl.ncomms = 0;
        g_nccl.ready = false;
    }
}

std::mutex g_multi_mu;  // one multi-device call at a time (the partials live in per-context staging buffers)

}  // namespace

extern "C" {

int bls381_init(int device, const char* program_dir) {
    Enter enter_;
    if (g.inited) {
        if (g.device != device) return fail(BLS381_EINVAL, "already initialised on CUDA device " + std::to_string(g.device));
        return BLS381_OK;
    }
    int rc = init_context(device, program_dir);
    if (rc == BLS381_OK && g_cur == &g_states[0]) g_ncontexts = std::max(g_ncontexts, 1);
    return rc;
}

int bls381_shutdown(void) {
    multi_shutdown();
    for (int j = kMaxDevices - 1; j >= 0; --j) {
        State* prev = g_cur;
        g_cur = &g_states[j];
        {
            Enter enter_;
            shutdown_context();
        }
        g_cur = prev;
    }
    g_ncontexts = 0;
    return BLS381_OK;
}

const char* bls381_last_error(void) { return g_err.c_str(); }
int bls381_sm_count(void) { return g.inited ? g.sm_count : 0; }
uint64_t bls381_launch_count(void) { return g.launches.load(); }
double bls381_last_kernel_ms(void) { return g.last_ms; }

int bls381_set_option(const char* name, int value) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!name) return fail(BLS381_EINVAL, "null argument");
    const std::string n(name);
    if (n == "dynamic_batches") g.dynamic_batches = value != 0;
    else if (n == "ctas_per_sm") g.force_ctas = value;
    else if (n == "poll_sleep_ns") g.sleep_ns = value;
    else if (n == "no_tma") g.no_tma = value != 0;
    else if (n == "swu_kernel") g.swu_kernel = value != 0;
    else if (n == "tail_kernels") g.tail_kernels = value != 0;
    else if (n == "g1_kernel") g.g1_kernel = value != 0;
    else if (n == "g2_kernel") g.g2_kernel = value != 0;
    else if (n == "pipeline_copies") g.pipeline_copies = value != 0;
    else if (n == "pairs_per_lane") g.pairs_per_lane = value;
    else return fail(BLS381_EINVAL, "unknown option: " + n);
    return BLS381_OK;
}

double bls381_last_kernel_sm_mhz(void) {
    if (!g.inited || !g.d_clk) return 0.0;
    unsigned long long h[2] = {0, 0};
    if (cudaDeviceSynchronize() != cudaSuccess) return 0.0;
    if (cudaMemcpy(h, g.d_clk, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess || h[1] == 0) return 0.0;
    return (double)h[0] / (double)h[1] * 1000.0;
}

int bls381_vm_load(const char* name, const uint8_t* image, size_t len) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!name || !image) return fail(BLS381_EINVAL, "null argument");
    return load_image(name, image, len);
}

int bls381_vm_run_dev(const char* program, uint8_t* const* d_bufs, const uint32_t* strides, int nbuf,
                      size_t n_items, void* cuda_stream) {
    Enter enter_;
    if (!program || !d_bufs || !strides) return fail(BLS381_EINVAL, "null argument");
    return vm_run(program, d_bufs, strides, nbuf, n_items, (cudaStream_t)cuda_stream);
}

int bls381_pairing_batch_dev(const uint8_t* d_g1, const uint8_t* d_g2, size_t n, int with_final_exp,
                             uint8_t* d_out, void* cuda_stream) {
    Enter enter_;
    if (!d_g1 || !d_g2 || !d_out) return fail(BLS381_EINVAL, "null argument");
    uint8_t* bufs[3] = {const_cast<uint8_t*>(d_g1), const_cast<uint8_t*>(d_g2), d_out};
    uint32_t strides[3] = {96, 192, 576};
    return vm_run(with_final_exp ? "pairing" : "miller", bufs, strides, 3, n, (cudaStream_t)cuda_stream);
}

// per-item status of pairing() in the reference's order of checks (index.ts:716-718): infinity of either point (affine
// (0, 0), what toAffine() gives for ZERO, math.ts:955), then P.assertValidity(), then Q.assertValidity(); the output of a
// failing item is zeroed (the reference throws, nothing is returned)
__global__ void pairing_status_kernel(const uint8_t* g1, const uint8_t* g2, const int32_t* st1, const int32_t* st2, int32_t* st,
                                      uint8_t* out576, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* a = reinterpret_cast<const uint4*>(g1 + 96 * i);
    const uint4* c = reinterpret_cast<const uint4*>(g2 + 192 * i);
    uint32_t any1 = 0, any2 = 0;
    for (int k = 0; k < 6; ++k) any1 |= a[k].x | a[k].y | a[k].z | a[k].w;
    for (int k = 0; k < 12; ++k) any2 |= c[k].x | c[k].y | c[k].z | c[k].w;
    int32_t v = BLS381_ST_OK;
    if (!any1 || !any2) v = BLS381_ST_INFINITY;
    else if (st1[i] != BLS381_ST_OK) v = st1[i];
    else if (st2[i] != BLS381_ST_OK) v = st2[i];
    st[i] = v;
    if (v != BLS381_ST_OK) {
        uint4* o = reinterpret_cast<uint4*>(out576 + 576 * i);
        for (int k = 0; k < 36; ++k) o[k] = make_uint4(0, 0, 0, 0);
    }
}

int bls381_pairing_batch(const uint8_t* g1, const uint8_t* g2, size_t n, int with_final_exp, uint8_t* out,
                         int32_t* status) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!g1 || !g2 || !out) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    int rc;
    if ((rc = stage(0, n * 96)) || (rc = stage(1, n * 192)) || (rc = stage(2, n * 576))) return rc;
    if (status && (rc = stage(6, 3 * n * 4))) return rc;
    // Large batches without validity checks: three chunks on the two internal streams, so that only the first chunk's
    // H2D copy and the last chunk's D2H copy are exposed.  The kernels of consecutive chunks overlap: the persistent CTAs
    // of the next chunk become resident as those of the previous one run out of batches, so there is still ONE tail.
    const size_t round_items = (size_t)g.sm_count * 2 * 32;   // one batch per resident CTA of the 8-warp programs
    if (!status && g.pipeline_copies && n >= 5 * round_items) {
        const char* prog = with_final_exp ? "pairing" : "miller";
        const size_t c0 = round_items, c2 = round_items, c1 = n - c0 - c2;
        const size_t off[3] = {0, c0, c0 + c1}, len[3] = {c0, c1, c2};
        cudaStream_t st[3] = {g.stream, g.stream2, g.stream};
        CUDA_TRY(cudaEventRecord(g.ev_fork, g.stream));             // stream2 starts after everything queued so far
        CUDA_TRY(cudaStreamWaitEvent(g.stream2, g.ev_fork, 0));
        for (int c = 0; c < 3; ++c) {
            CUDA_TRY(cudaMemcpyAsync(g.d_stage[0] + off[c] * 96, g1 + off[c] * 96, len[c] * 96, cudaMemcpyHostToDevice, st[c]));
            CUDA_TRY(cudaMemcpyAsync(g.d_stage[1] + off[c] * 192, g2 + off[c] * 192, len[c] * 192, cudaMemcpyHostToDevice, st[c]));
            if (c == 0) CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
            uint8_t* cb[3] = {g.d_stage[0] + off[c] * 96, g.d_stage[1] + off[c] * 192, g.d_stage[2] + off[c] * 576};
            uint32_t cs[3] = {96, 192, 576};
            if ((rc = vm_run(prog, cb, cs, 3, len[c], st[c]))) return rc;
            if (c == 1) CUDA_TRY(cudaEventRecord(g.ev_join, g.stream2));
            if (c == 2) {
                CUDA_TRY(cudaStreamWaitEvent(g.stream, g.ev_join, 0));   // kernel time = first kernel start .. all kernels done
                CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
            }
            CUDA_TRY(cudaMemcpyAsync(out + off[c] * 576, g.d_stage[2] + off[c] * 576, len[c] * 576, cudaMemcpyDeviceToHost, st[c]));
        }
        CUDA_TRY(cudaStreamSynchronize(g.stream2));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, g.ev0, g.ev1);
        g.last_ms = ms;
        return BLS381_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], g1, n * 96, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[1], g2, n * 192, cudaMemcpyHostToDevice, g.stream));
    uint8_t* bufs[3] = {g.d_stage[0], g.d_stage[1], g.d_stage[2]};
    uint32_t strides[3] = {96, 192, 576};
    CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
    if ((rc = vm_run(with_final_exp ? "pairing" : "miller", bufs, strides, 3, n, g.stream))) return rc;
    if (status) {  // P.assertValidity(), Q.assertValidity() (index.ts:717-718); status == NULL: the caller vouches for the points
        int32_t* d_st = reinterpret_cast<int32_t*>(g.d_stage[6]);
        if ((rc = run3("g1_validate", g.d_stage[0], 96, nullptr, 0, d_st + n, n, g.stream))) return rc;
        if ((rc = run3("g2_validate", g.d_stage[1], 192, nullptr, 0, d_st + 2 * n, n, g.stream))) return rc;
        pairing_status_kernel<<<(unsigned)((n + 127) / 128), 128, 0, g.stream>>>(g.d_stage[0], g.d_stage[1], d_st + n, d_st + 2 * n, d_st,
                                                                                  g.d_stage[2], n);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
    CUDA_TRY(cudaMemcpyAsync(out, g.d_stage[2], n * 576, cudaMemcpyDeviceToHost, g.stream));
    if (status) CUDA_TRY(cudaMemcpyAsync(status, g.d_stage[6], n * 4, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_ms = ms;
    return BLS381_OK;
}

int bls381_final_exp_batch_dev(const uint8_t* d_in, size_t n, uint8_t* d_out, void* cuda_stream) {
    Enter enter_;
    if (!d_in || !d_out) return fail(BLS381_EINVAL, "null argument");
    uint8_t* bufs[4] = {nullptr, nullptr, d_out, const_cast<uint8_t*>(d_in)};
    uint32_t strides[4] = {0, 0, 576, 576};
    return vm_run("final_exp", bufs, strides, 4, n, (cudaStream_t)cuda_stream);
}

int bls381_final_exp_batch(const uint8_t* in, size_t n, uint8_t* out) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!in || !out) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    int rc;
    if ((rc = stage(0, n * 576)) || (rc = stage(2, n * 576))) return rc;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], in, n * 576, cudaMemcpyHostToDevice, g.stream));
    uint8_t* bufs[4] = {nullptr, nullptr, g.d_stage[2], g.d_stage[0]};
    uint32_t strides[4] = {0, 0, 576, 576};
    CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
    if ((rc = vm_run("final_exp", bufs, strides, 4, n, g.stream))) return rc;
    CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
    CUDA_TRY(cudaMemcpyAsync(out, g.d_stage[2], n * 576, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_ms = ms;
    return BLS381_OK;
}

int bls381_miller_product_dev(const uint8_t* d_g1, const uint8_t* d_g2, size_t n, int with_final_exp,
                              uint8_t* d_out, void* cuda_stream) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!d_g1 || !d_g2 || !d_out) return fail(BLS381_EINVAL, "null argument");
    return miller_product_dev(d_g1, d_g2, n, with_final_exp, d_out, (cudaStream_t)cuda_stream);
}

int bls381_miller_product(const uint8_t* g1, const uint8_t* g2, size_t n, int with_final_exp, uint8_t* out) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!g1 || !g2 || !out) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return fail(BLS381_EINVAL, "empty batch");
    int rc;
    if ((rc = stage(0, (n + kMillerPad) * 96 + 576)) || (rc = stage(1, (n + kMillerPad) * 192))) return rc;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0] + 576, g1, n * 96, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[1], g2, n * 192, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
    if ((rc = miller_product_dev(g.d_stage[0] + 576, g.d_stage[1], n, with_final_exp, g.d_stage[0], g.stream, true))) return rc;
    CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
    CUDA_TRY(cudaMemcpyAsync(out, g.d_stage[0], 576, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_ms = ms;
    return BLS381_OK;
}

int bls381_g1_decompress_batch(const uint8_t* in48, size_t n, uint8_t* out96, int32_t* status) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!in48 || !out96 || !status) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    int rc;
    if ((rc = stage(0, n * 48)) || (rc = stage(2, n * 96)) || (rc = stage(6, n * 4))) return rc;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], in48, n * 48, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
    if ((rc = g1_decompress_dev(g.d_stage[0], g.d_stage[2], (int32_t*)g.d_stage[6], n, g.stream))) return rc;
    CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
    CUDA_TRY(cudaMemcpyAsync(out96, g.d_stage[2], n * 96, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaMemcpyAsync(status, g.d_stage[6], n * 4, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_ms = ms;
    return BLS381_OK;
}

int bls381_g2_decompress_batch(const uint8_t* in96, size_t n, uint8_t* out192, int32_t* status) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!in96 || !out192 || !status) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    int rc;
    if ((rc = stage(0, n * 96)) || (rc = stage(2, n * 192)) || (rc = stage(6, n * 4))) return rc;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], in96, n * 96, cudaMemcpyHostToDevice, g.stream));
    if ((rc = g2_decompress_dev(g.d_stage[0], g.d_stage[2], (int32_t*)g.d_stage[6], n, g.stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out192, g.d_stage[2], n * 192, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaMemcpyAsync(status, g.d_stage[6], n * 4, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return BLS381_OK;
}

int bls381_hash_to_g2_batch(const uint8_t* msgs, const uint64_t* msg_off, size_t n, const uint8_t* dst, size_t dst_len,
                            uint8_t* out192) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!msg_off || !dst || !out192) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    int rc;
    if ((rc = check_offsets(msg_off, n))) return rc;
    const size_t mbytes = msg_off[n];
    if (!msgs && mbytes) return fail(BLS381_EINVAL, "null message buffer");
    if ((rc = stage(0, mbytes + 16)) || (rc = stage(1, (n + 1) * 8)) || (rc = stage(2, n * 192))) return rc;
    if (mbytes) CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], msgs, mbytes, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[1], msg_off, (n + 1) * 8, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
    if ((rc = hash_to_g2_dev(g.d_stage[0], (const uint64_t*)g.d_stage[1], n, dst, dst_len, g.d_stage[2], g.stream))) return rc;
    CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
    CUDA_TRY(cudaMemcpyAsync(out192, g.d_stage[2], n * 192, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_ms = ms;
    return BLS381_OK;
}

// PointG1.hashToCurve(msg, {DST})                                              replaces index.ts:331-339
// (hash_to_field with m = 1: 128 uniform bytes; map_to_curve_simple_swu_3mod4 math.ts:1270-1313; 11-isogeny math.ts:1327,
//  1612-1790; clearCofactor index.ts:401-405).  out96: affine H(m_i).
int bls381_hash_to_g1_batch(const uint8_t* msgs, const uint64_t* msg_off, size_t n, const uint8_t* dst, size_t dst_len,
                            uint8_t* out96) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!msg_off || !dst || !out96) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    int rc;
    if ((rc = check_offsets(msg_off, n))) return rc;
    const size_t mbytes = msg_off[n];
    if (!msgs && mbytes) return fail(BLS381_EINVAL, "null message buffer");
    if ((rc = stage(0, mbytes + 16)) || (rc = stage(1, (n + 1) * 8)) || (rc = stage(2, n * 96)) || (rc = stage(4, n * 128)) || (rc = stage(5, 512))) return rc;
    cudaStream_t s = g.stream;
    std::vector<uint8_t> dp;
    make_dst_prime(dst, dst_len, dp);
    if (mbytes) CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], msgs, mbytes, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[1], msg_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[5], dp.data(), dp.size(), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(g.ev0, s));
    xmd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g.d_stage[0], (const uint64_t*)g.d_stage[1], n, g.d_stage[5], (uint32_t)dp.size(),
                                                       g.d_stage[4], 128u);
    CUDA_TRY(cudaGetLastError());
    if ((rc = run3("hash_to_g1", g.d_stage[4], 128, g.d_stage[2], 96, nullptr, n, s))) return rc;
    CUDA_TRY(cudaEventRecord(g.ev1, s));
    CUDA_TRY(cudaMemcpyAsync(out96, g.d_stage[2], n * 96, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_ms = ms;
    return BLS381_OK;
}

static int verdict_of(const uint8_t* res576, const int32_t* status, size_t n);

// Shared body of verifyBatch: product of the Miller loops e(pk_i, H(m_i)) of n items [times e(-G1, sig) when
// sig96 != NULL], optionally followed by the final exponentiation.  Result: 576 bytes at `out` (host), or left in the
// device buffer g.d_stage[8] when out == NULL.
static int verify_partial(const uint8_t* sig96, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* pks48,
                          size_t n, const uint8_t* dst, size_t dst_len, int with_final_exp, uint8_t* out, int32_t* status) {
    int rc;
    if (n && (rc = check_offsets(msg_off, n))) return rc;
    const size_t mbytes = n ? msg_off[n] : 0;
    if (!msgs && mbytes) return fail(BLS381_EINVAL, "null message buffer");
    const size_t np = n + (sig96 ? 1 : 0);
    if (np == 0) return fail(BLS381_EINVAL, "empty batch");
    // staging: 0 msgs, 1 offsets, 7 pks(48) + sig(96), 8 result + g1 pairs, 9 g2 pairs, 6 status (n+1)
    if ((rc = stage(0, mbytes + 16)) || (rc = stage(1, (n + 1) * 8)) || (rc = stage(7, n * 48 + 96)) ||
        (rc = stage(8, (n + 1 + kMillerPad) * 96 + 576)) || (rc = stage(9, (n + 1 + kMillerPad) * 192)) || (rc = stage(6, (n + 1) * 4)))
        return rc;
    cudaStream_t s = g.stream;
    uint8_t* d_g1 = g.d_stage[8] + 576;  // first 576 B: result
    uint8_t* d_g2 = g.d_stage[9];
    int32_t* d_st = (int32_t*)g.d_stage[6];
    if (n) {
        if (mbytes) CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], msgs, mbytes, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(g.d_stage[1], msg_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(g.d_stage[7], pks48, n * 48, cudaMemcpyHostToDevice, s));
    }
    if (sig96) CUDA_TRY(cudaMemcpyAsync(g.d_stage[7] + n * 48, sig96, 96, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(g.ev0, s));
    if (sig96) {
        // normP2(signature) (index.ts:799) and the pairing(G1.negate(), sig) term (index.ts:814): a single-item,
        // latency-bound launch (no far slots) -> runs on a second stream underneath the bulk stages
        CUDA_TRY(cudaEventRecord(g.ev_fork, s));
        CUDA_TRY(cudaStreamWaitEvent(g.stream2, g.ev_fork, 0));
        if ((rc = run3("g2_decompress", g.d_stage[7] + n * 48, 96, d_g2 + n * 192, 192, d_st + n, 1, g.stream2))) return rc;
        CUDA_TRY(cudaMemcpyAsync(d_g1 + n * 96, kNegG1, 96, cudaMemcpyHostToDevice, g.stream2));
        CUDA_TRY(cudaEventRecord(g.ev_join, g.stream2));
    }
    if (n) {
        // publicKeys.map(normP1)  (index.ts:801)
        if ((rc = g1_decompress_dev(g.d_stage[7], d_g1, d_st, n, s))) return rc;
        // messages.map(normP2Hash)  (index.ts:800)
        if ((rc = hash_to_g2_dev(g.d_stage[0], (const uint64_t*)g.d_stage[1], n, dst, dst_len, d_g2, s))) return rc;
    }
    if (sig96) CUDA_TRY(cudaStreamWaitEvent(s, g.ev_join, 0));
    // product of the Miller loops (+ one final exponentiation)  (index.ts:812-816)
    if ((rc = miller_product_dev(d_g1, d_g2, np, with_final_exp, g.d_stage[8], s, true))) return rc;
    CUDA_TRY(cudaEventRecord(g.ev1, s));
    if (out) CUDA_TRY(cudaMemcpyAsync(out, g.d_stage[8], 576, cudaMemcpyDeviceToHost, s));  // NULL: stays in g.d_stage[8]
    CUDA_TRY(cudaMemcpyAsync(status, d_st, np * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_ms = ms;
    return BLS381_OK;
}

int bls381_verify_batch(const uint8_t* sig96, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* pks48,
                        size_t n, const uint8_t* dst, size_t dst_len, int* verdict, int32_t* status) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!sig96 || !msg_off || !pks48 || !dst || !verdict || !status) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return fail(BLS381_EINVAL, "Expected non-empty messages array");
    uint8_t res[576];
    int rc = verify_partial(sig96, msgs, msg_off, pks48, n, dst, dst_len, 1, res, status);
    if (rc) return rc;
    *verdict = verdict_of(res, status, n);
    return BLS381_OK;
}

int bls381_verify_batch_partial(const uint8_t* sig96_or_null, const uint8_t* msgs, const uint64_t* msg_off,
                                const uint8_t* pks48, size_t n, const uint8_t* dst, size_t dst_len, uint8_t* out_fp12,
                                int32_t* status) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!dst || !out_fp12 || !status || (n && (!msg_off || !pks48))) return fail(BLS381_EINVAL, "null argument");
    return verify_partial(sig96_or_null, msgs, msg_off, pks48, n, dst, dst_len, 0, out_fp12, status);
}

// verdict of a verifyBatch from the exponentiated product and the status words (index.ts:716, 802-820)
static int verdict_of(const uint8_t* res576, const int32_t* status, size_t n) {
    bool one = res576[47] == 1;
    for (int i = 0; i < 576 && one; ++i)
        if (i != 47 && res576[i] != 0) one = false;
    // reference semantics: decoding errors throw (reported as verdict -1 + status codes); an infinity public key
    // or signature makes pairing() throw inside the try block => false
    int v = one ? 1 : 0;
    for (size_t i = 0; i <= n; ++i) {
        if (status[i] == BLS381_ST_INFINITY) { if (v > 0) v = 0; }
        else if (status[i] != BLS381_ST_OK) v = -1;
    }
    return v;
}

// ---- in-process multi-GPU: one context + one host thread per device, items sharded by index, ONE all-gather of
// the device-resident 576-byte partial products (NCCL over NVLink; peer copies if NCCL cannot be loaded), product + final
// exponentiation on the first device.  No host bounce of the partials.
// Same as bls381_verify_batch_partial, the 576-byte product stays on the device (d_out_fp12: device pointer): the partials
// of a multi-process run are all-gathered device to device (torch.distributed / NCCL) without touching the host.
int bls381_verify_batch_partial_dev(const uint8_t* sig96_or_null, const uint8_t* msgs, const uint64_t* msg_off,
                                    const uint8_t* pks48, size_t n, const uint8_t* dst, size_t dst_len, uint8_t* d_out_fp12,
                                    int32_t* status) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!dst || !d_out_fp12 || !status || (n && (!msg_off || !pks48))) return fail(BLS381_EINVAL, "null argument");
    int rc = verify_partial(sig96_or_null, msgs, msg_off, pks48, n, dst, dst_len, 0, nullptr, status);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_out_fp12, g.d_stage[8], 576, cudaMemcpyDeviceToDevice, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return BLS381_OK;
}

// prod_i f_i (+ finalExponentiate) on device pointers, asynchronous on the caller's stream
int bls381_fp12_product_dev(const uint8_t* d_in_fp12, size_t n, int with_final_exp, uint8_t* d_out_fp12, void* cuda_stream) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!d_in_fp12 || !d_out_fp12) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return fail(BLS381_EINVAL, "empty batch");
    cudaStream_t s = (cudaStream_t)cuda_stream;
    int rc;
    uint8_t *a = nullptr, *b = nullptr;
    if ((rc = tree_buf(s, 0, n * 576 + 576, &a)) || (rc = tree_buf(s, 1, ((n + 31) / 32) * 576 + 576, &b))) return rc;
    CUDA_TRY(cudaMemcpyAsync(a, d_in_fp12, n * 576, cudaMemcpyDeviceToDevice, s));
    uint8_t* res = nullptr;
    if ((rc = product_tree(a, n, b, &res, s))) return rc;
    if (with_final_exp) {
        uint8_t* b2[4] = {nullptr, nullptr, d_out_fp12, res};
        uint32_t st2[4] = {0, 0, 576, 576};
        return vm_run("final_exp", b2, st2, 4, 1, s);
    }
    CUDA_TRY(cudaMemcpyAsync(d_out_fp12, res, 576, cudaMemcpyDeviceToDevice, s));
    return BLS381_OK;
}

int bls381_init_devices(uint32_t device_mask, const char* program_dir) {
    if (device_mask == 0) return fail(BLS381_EINVAL, "empty device mask");
    int devs[kMaxDevices], nd = 0;
    for (int d = 0; d < 32 && nd < kMaxDevices; ++d)
        if (device_mask >> d & 1) devs[nd++] = d;
    for (int j = 0; j < nd; ++j) {
        State* prev = g_cur;
        g_cur = &g_states[j];
        int rc;
        {
            Enter enter_;
            if (g.inited && g.device != devs[j]) rc = fail(BLS381_EINVAL, "context already bound to another CUDA device");
            else rc = g.inited ? BLS381_OK : init_context(devs[j], program_dir);
        }
        g_cur = prev;
        if (rc) return rc;
    }
    std::lock_guard<std::mutex> multi_lock(g_multi_mu);
    g_ncontexts = std::max(g_ncontexts, nd);
    if (g_ncontexts > 1 && g_nccl.ncomms != g_ncontexts) {
        // the communicators span exactly the current set of contexts: rebuilt when devices are added
        multi_shutdown();
        nccl_load();
        int all[kMaxDevices];
        for (int j = 0; j < g_ncontexts; ++j) all[j] = g_states[j].device;
        if (g_nccl.handle) {
            const int rc = g_nccl.CommInitAll(g_nccl.comms, g_ncontexts, all);
            if (rc == 0) {
                g_nccl.ready = true;
                g_nccl.ncomms = g_ncontexts;
            } else {
                g_nccl.why = std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
            }
        }
        if (!g_nccl.ready) {  // fallback transport: peer-to-peer copies into the first device
            for (int j = 1; j < g_ncontexts; ++j) {
                cudaSetDevice(all[0]);
                cudaDeviceEnablePeerAccess(all[j], 0);
                cudaGetLastError();
            }
        }
    }
    cudaSetDevice(g_states[0].device);
    return BLS381_OK;
}

int bls381_device_count(void) { return g_ncontexts; }
const char* bls381_multi_transport(void) { return g_nccl.ready ? "nccl" : (g_ncontexts > 1 ? "peer-copy" : "single-device"); }

int bls381_verify_batch_multi(const uint8_t* sig96, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* pks48,
                              size_t n, const uint8_t* dst, size_t dst_len, int* verdict, int32_t* status) {
    std::lock_guard<std::mutex> multi_lock(g_multi_mu);
    const int W = g_ncontexts;
    if (W < 1 || !g_states[0].inited) return fail(BLS381_ENOINIT, "bls381_init_devices() has not been called");
    if (!sig96 || !msg_off || !pks48 || !dst || !verdict || !status) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return fail(BLS381_EINVAL, "Expected non-empty messages array");
    if (W == 1 || n < (size_t)W) return bls381_verify_batch(sig96, msgs, msg_off, pks48, n, dst, dst_len, verdict, status);
    int rc0 = check_offsets(msg_off, n);
    if (rc0) return rc0;
    // shard j = items [n j / W, n (j + 1) / W); the e(-G1, sig) term goes with shard 0.  Status layout as in the
    // single-device call: n public keys, then the signature -- shard 0 writes its keys and (via a bounce word) the signature.
    std::vector<int> rcs(W, 0);
    std::vector<std::string> errs(W);
    std::vector<std::vector<uint64_t>> offs(W);
    std::vector<std::vector<int32_t>> sts(W);
    std::vector<std::thread> th;
    for (int j = 0; j < W; ++j) {
        th.emplace_back([&, j]() {
            g_cur = &g_states[j];
            Enter enter_;
            const size_t lo = n * (size_t)j / W, hi = n * (size_t)(j + 1) / W, cnt = hi - lo;
            offs[j].resize(cnt + 1);
            for (size_t i = 0; i <= cnt; ++i) offs[j][i] = msg_off[lo + i] - msg_off[lo];
            sts[j].assign(cnt + 1, 0);
            rcs[j] = verify_partial(j == 0 ? sig96 : nullptr, msgs ? msgs + msg_off[lo] : nullptr, offs[j].data(), pks48 + 48 * lo, cnt, dst,
                                    dst_len, 0, nullptr, sts[j].data());
            if (rcs[j]) errs[j] = g_err;
        });
    }
    for (auto& t : th) t.join();
    for (int j = 0; j < W; ++j)
        if (rcs[j]) return fail(rcs[j], "device " + std::to_string(g_states[j].device) + ": " + errs[j]);
    for (int j = 0; j < W; ++j) {
        const size_t lo = n * (size_t)j / W, cnt = n * (size_t)(j + 1) / W - lo;
        memcpy(status + lo, sts[j].data(), cnt * sizeof(int32_t));
        if (j == 0) status[n] = sts[0][cnt];
    }
    // the single exchange step: all-gather of W x 576 bytes, device to device
    State& g0 = g_states[0];
    uint8_t* gathered = nullptr;
    {
        g_cur = &g_states[0];
        Enter enter_;
        int rc;
        if ((rc = stage(2, (size_t)W * 576 + 576)) || (rc = stage(3, 2 * 576)) || (rc = stage(5, 1024))) { g_cur = &g_states[0]; return rc; }
        gathered = g0.d_stage[2];
    }
    if (g_nccl.ready) {
        std::vector<uint8_t*> recv(W);
        for (int j = 0; j < W; ++j) {
            g_cur = &g_states[j];
            Enter enter_;
            int rc = stage(2, (size_t)W * 576 + 576);
            if (rc) { g_cur = &g_states[0]; return rc; }
            recv[j] = g.d_stage[2];
        }
        g_cur = &g_states[0];
        int nrc = g_nccl.GroupStart();
        for (int j = 0; j < W && nrc == 0; ++j) {
            cudaSetDevice(g_states[j].device);
            nrc = g_nccl.AllGather(g_states[j].d_stage[8], recv[j], 576, kNcclUint8, g_nccl.comms[j], g_states[j].stream);
        }
        if (nrc == 0) nrc = g_nccl.GroupEnd();
        if (nrc != 0) return fail(BLS381_ECUDA, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "error"));
        for (int j = 1; j < W; ++j) {  // the other devices only take part in the collective; wait for them
            cudaSetDevice(g_states[j].device);
            CUDA_TRY(cudaStreamSynchronize(g_states[j].stream));
        }
        cudaSetDevice(g0.device);
    } else {
        cudaSetDevice(g0.device);
        for (int j = 0; j < W; ++j)
            CUDA_TRY(cudaMemcpyPeerAsync(gathered + (size_t)j * 576, g0.device, g_states[j].d_stage[8], g_states[j].device, 576, g0.stream));
    }
    // product of the W partials + ONE final exponentiation on the first device (index.ts:815-816)
    g_cur = &g_states[0];
    Enter enter_;
    uint8_t* res = nullptr;
    int rc;
    if ((rc = product_tree(gathered, (size_t)W, g0.d_stage[3], &res, g0.stream))) return rc;
    uint8_t* b2[4] = {nullptr, nullptr, g0.d_stage[5], res};
    uint32_t st2[4] = {0, 0, 576, 576};
    if ((rc = vm_run("final_exp", b2, st2, 4, 1, g0.stream))) return rc;
    uint8_t out576[576];
    CUDA_TRY(cudaMemcpyAsync(out576, g0.d_stage[5], 576, cudaMemcpyDeviceToHost, g0.stream));
    CUDA_TRY(cudaStreamSynchronize(g0.stream));
    *verdict = verdict_of(out576, status, n);
    return BLS381_OK;
}

// Secret scalars are prepared WITHOUT secret-dependent branches or variable-time divisions (the reference's own ladder is
// constant-time, math.ts:1061-1078): the reduction mod r is two masked subtractions (k < 2^256 < 3r), the base-z digits
// come from a bit-serial restoring division with masked updates.
__device__ __forceinline__ void load_be256(unsigned long long* w, const uint8_t* p) {  // w[3] most significant
    for (int k = 0; k < 4; ++k) {
        unsigned long long v = 0;
        for (int b = 0; b < 8; ++b) v = (v << 8) | p[8 * (3 - k) + b];
        w[k] = v;
    }
}
__device__ __forceinline__ void store_be256(uint8_t* p, const unsigned long long* w) {
    for (int k = 0; k < 4; ++k)
        for (int b = 0; b < 8; ++b) p[8 * (3 - k) + b] = (uint8_t)(w[k] >> (8 * (7 - b)));
}
__device__ __forceinline__ void reduce_mod_r(unsigned long long* w) {
    const unsigned long long r[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
    for (int rep = 0; rep < 2; ++rep) {
        unsigned long long t[4], borrow = 0;
        for (int k = 0; k < 4; ++k) {
            const unsigned long long a = w[k], s1 = a - r[k], s2 = s1 - borrow;
            borrow = (unsigned long long)(a < r[k]) | (unsigned long long)(s1 < borrow);
            t[k] = s2;
        }
        const unsigned long long keep = 0ull - borrow;  // all ones: w < r, keep w
        for (int k = 0; k < 4; ++k) w[k] = (w[k] & keep) | (t[k] & ~keep);
    }
}

// sign: the secret scalar as four base-z digits (z = |x| = 0xd201000000010000; k < r < z^4), written in place as
// a_3 || a_2 || a_1 || a_0 (8 bytes big-endian each), the input format of the `sign` program's joint ladder over
// psi^i(H(m)) (vmprog/curves.py: g2_mul_secret_gls4).  Scalars >= r are reduced first ([k]P = [k mod r]P on G2).
__global__ void base_z_digits_kernel(uint8_t* sk32, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t* p = sk32 + 32 * i;
    unsigned long long w[4];
    load_be256(w, p);
    reduce_mod_r(w);
    const unsigned long long z = 0xd201000000010000ull;
    unsigned long long digit[4];
    for (int d = 0; d < 4; ++d) {  // w = z * q + digit, q -> w
        unsigned long long rem = 0, q[4] = {0, 0, 0, 0};
        for (int bit = 255; bit >= 0; --bit) {
            const unsigned long long top = rem >> 63;                       // rem < z < 2^64: (rem << 1 | b) has 65 bits
            rem = (rem << 1) | ((w[bit >> 6] >> (bit & 63)) & 1ull);
            const unsigned long long ge = top | (unsigned long long)(rem >= z);
            rem -= z & (0ull - ge);
            q[bit >> 6] |= ge << (bit & 63);
        }
        digit[d] = rem;
        for (int k = 0; k < 4; ++k) w[k] = q[k];
    }
    for (int d = 0; d < 4; ++d)
        for (int b = 0; b < 8; ++b) p[8 * (3 - d) + b] = (uint8_t)(digit[d] >> (8 * (7 - b)));
}

// getPublicKey: scalars reduced mod r in place (normalizePrivKey, index.ts:269-279; a zero result stays zero and yields the
// encoding of the point at infinity, the host wrapper rejects it like the reference)
__global__ void reduce_scalars_kernel(uint8_t* sk32, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long w[4];
    load_be256(w, sk32 + 32 * i);
    reduce_mod_r(w);
    store_be256(sk32 + 32 * i, w);
}

// PointG1.toHex(true) (index.ts:359-371) from affine x || y + the point-at-infinity flag word of the scalar multiplication
__global__ void g1_compress_affine_kernel(const uint8_t* aff96, const int32_t* flags, uint8_t* out48, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    static const uint8_t half[48] = {  // (p - 1) / 2: (y * 2) / P == 1  <=>  y > (p - 1) / 2
        0x0d, 0x00, 0x88, 0xf5, 0x1c, 0xbf, 0xf3, 0x4d, 0x25, 0x8d, 0xd3, 0xdb, 0x21, 0xa5, 0xd6, 0x6b, 0xb2, 0x3b, 0xa5, 0xc2, 0x79, 0xc2, 0x89, 0x5f,
        0xb3, 0x98, 0x69, 0x50, 0x7b, 0x58, 0x7b, 0x12, 0x0f, 0x55, 0xff, 0xff, 0x58, 0xa9, 0xff, 0xff, 0xdc, 0xff, 0x7f, 0xff, 0xff, 0xff, 0xd5, 0x55};
    const uint8_t* a = aff96 + 96 * i;
    uint8_t* o = out48 + 48 * i;
    if (flags[i] & 2) {
        for (int k = 0; k < 48; ++k) o[k] = 0;
        o[0] = 0xC0;
        return;
    }
    int gt = 0, decided = 0;
    for (int k = 0; k < 48; ++k) {
        const int y = a[48 + k], h = half[k];
        gt |= (!decided) & (y > h);
        decided |= (y != h);
    }
    for (int k = 0; k < 48; ++k) o[k] = a[k];
    o[0] |= (uint8_t)(0x80 | (gt << 5));
}

int bls381_sign_batch(const uint8_t* sks32, const uint8_t* msgs, const uint64_t* msg_off, size_t n, const uint8_t* dst,
                      size_t dst_len, uint8_t* out_sig96) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!sks32 || !msg_off || !dst || !out_sig96) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    int rc;
    if ((rc = check_offsets(msg_off, n))) return rc;
    const size_t mbytes = msg_off[n];
    if (!msgs && mbytes) return fail(BLS381_EINVAL, "null message buffer");
    if ((rc = stage(0, mbytes + 16)) || (rc = stage(1, (n + 1) * 8)) || (rc = stage(7, n * 32)) || (rc = stage(2, n * 96)) ||
        (rc = stage(6, n * 4)) || (rc = stage(4, n * 256)) || (rc = stage(5, 512)))
        return rc;
    cudaStream_t s = g.stream;
    std::vector<uint8_t> dp;
    make_dst_prime(dst, dst_len, dp);
    // (A chunked copy / compute pipeline like the one of bls381_pairing_batch was measured here and is NOT used: cutting the
    // per-item kernels into chunks of two waves costs 5 % in drained waves, and kernels of different kinds sharing an SM on
    // two streams cost 7 %; the copies are 2 % of the call.  profiles/r2_notes.md)
    if (mbytes) CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], msgs, mbytes, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[1], msg_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[7], sks32, n * 32, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[5], dp.data(), dp.size(), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(g.ev0, s));
    xmd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g.d_stage[0], (const uint64_t*)g.d_stage[1], n, g.d_stage[5],
                                                       (uint32_t)dp.size(), g.d_stage[4]);
    CUDA_TRY(cudaGetLastError());
    base_z_digits_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g.d_stage[7], n);
    CUDA_TRY(cudaGetLastError());
    bool flags_applied = false;
    if (g.swu_kernel && g.tail_kernels) {
        if ((rc = swu_points(n, s))) return rc;
        swu::sign_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g.d_stage[10], g.d_stage[7], g.d_stage[2], n);
        CUDA_TRY(cudaGetLastError());
        g.launches.fetch_add(1);
        flags_applied = true;  // the kernel writes the finished signature (flag bits included)
    } else if (g.swu_kernel) {
        if ((rc = swu_points(n, s))) return rc;
        uint8_t* bufs[6] = {g.d_stage[10], g.d_stage[7], g.d_stage[2], nullptr, nullptr, g.d_stage[6]};
        uint32_t strides[6] = {576, 32, 96, 0, 0, 4};
        if ((rc = vm_run("sign_tail", bufs, strides, 6, n, s))) return rc;
    } else {
        uint8_t* bufs[6] = {g.d_stage[4], g.d_stage[7], g.d_stage[2], nullptr, nullptr, g.d_stage[6]};
        uint32_t strides[6] = {256, 32, 96, 0, 0, 4};
        if ((rc = vm_run("sign", bufs, strides, 6, n, s))) return rc;
    }
    if (!flags_applied) {
        apply_flags_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g.d_stage[2], 96, (const int32_t*)g.d_stage[6], n);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(g.ev1, s));
    CUDA_TRY(cudaMemcpyAsync(out_sig96, g.d_stage[2], n * 96, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemsetAsync(g.d_stage[7], 0, n * 32, s));  // the scalars / digits do not stay in the pooled staging buffer
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_ms = ms;
    return BLS381_OK;
}

// getPublicKey(privateKey) for n keys                                          replaces index.ts:738-740, 351-353
// (the reference's fixed-base wNAF gives the same point as any scalar multiplication of the generator)
int bls381_get_public_key_batch(const uint8_t* sks32, size_t n, uint8_t* out48) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!sks32 || !out48) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    int rc;
    if ((rc = stage(0, 96)) || (rc = stage(7, n * 32)) || (rc = stage(2, n * 96)) || (rc = stage(6, n * 4)) || (rc = stage(3, n * 48))) return rc;
    cudaStream_t s = g.stream;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], kG1, 96, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[7], sks32, n * 32, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(g.ev0, s));
    reduce_scalars_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g.d_stage[7], n);
    CUDA_TRY(cudaGetLastError());
    uint8_t* bufs[6] = {g.d_stage[0], g.d_stage[7], g.d_stage[2], nullptr, nullptr, g.d_stage[6]};
    uint32_t strides[6] = {0, 32, 96, 0, 0, 4};  // stride 0: every item multiplies the generator
    if ((rc = vm_run("g1_scalar_mul", bufs, strides, 6, n, s))) return rc;
    g1_compress_affine_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g.d_stage[2], (const int32_t*)g.d_stage[6], g.d_stage[3], n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(g.ev1, s));
    CUDA_TRY(cudaMemcpyAsync(out48, g.d_stage[3], n * 48, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemsetAsync(g.d_stage[7], 0, n * 32, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_ms = ms;
    return BLS381_OK;
}

static int aggregate_host(bool g2, const uint8_t* in, size_t n, uint8_t* out, int32_t* status) {
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!in || !out || !status) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return fail(BLS381_EINVAL, "Expected non-empty array");
    const uint32_t cb = g2 ? 96 : 48;
    int rc;
    if ((rc = stage(0, n * cb)) || (rc = stage(8, n * 2 * cb)) || (rc = stage(6, n * 4)) || (rc = stage(9, 256))) return rc;
    cudaStream_t s = g.stream;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], in, n * cb, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(g.ev0, s));
    if ((rc = g2 ? g2_decompress_dev(g.d_stage[0], g.d_stage[8], (int32_t*)g.d_stage[6], n, s)
                 : g1_decompress_dev(g.d_stage[0], g.d_stage[8], (int32_t*)g.d_stage[6], n, s)))
        return rc;
    if ((rc = aggregate_dev(g2, g.d_stage[8], (int32_t*)g.d_stage[6], n, g.d_stage[9], s))) return rc;
    CUDA_TRY(cudaEventRecord(g.ev1, s));
    CUDA_TRY(cudaMemcpyAsync(out, g.d_stage[9], cb, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(status, g.d_stage[6], n * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_ms = ms;
    return BLS381_OK;
}

int bls381_aggregate_g1(const uint8_t* pks48, size_t n, uint8_t* out48, int32_t* status) {
    Enter enter_;
    return aggregate_host(false, pks48, n, out48, status);
}

int bls381_aggregate_g2(const uint8_t* sigs96, size_t n, uint8_t* out96, int32_t* status) {
    Enter enter_;
    return aggregate_host(true, sigs96, n, out96, status);
}

int bls381_fp12_product(const uint8_t* in_fp12, size_t n, int with_final_exp, uint8_t* out_fp12) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!in_fp12 || !out_fp12) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return fail(BLS381_EINVAL, "empty batch");
    int rc;
    if ((rc = stage(2, n * 576 + 576)) || (rc = stage(3, ((n + 31) / 32) * 576 + 576)) || (rc = stage(8, 576))) return rc;
    cudaStream_t s = g.stream;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[2], in_fp12, n * 576, cudaMemcpyHostToDevice, s));
    uint8_t* res = nullptr;
    if ((rc = product_tree(g.d_stage[2], n, g.d_stage[3], &res, s))) return rc;
    if (with_final_exp) {
        uint8_t* b2[4] = {nullptr, nullptr, g.d_stage[8], res};
        uint32_t st2[4] = {0, 0, 576, 576};
        if ((rc = vm_run("final_exp", b2, st2, 4, 1, s))) return rc;
        res = g.d_stage[8];
    }
    CUDA_TRY(cudaMemcpyAsync(out_fp12, res, 576, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return BLS381_OK;
}

// ---- wire formats (index.ts:298-381, 532-631) on the device -----------------------------------------------------
// y > (p - 1) / 2 for a 48-byte big-endian canonical value: the reference's `(y * 2) / P` flag
__device__ __forceinline__ int be48_gt_half(const uint8_t* y) {
    static const uint8_t half[48] = {
        0x0d, 0x00, 0x88, 0xf5, 0x1c, 0xbf, 0xf3, 0x4d, 0x25, 0x8d, 0xd3, 0xdb, 0x21, 0xa5, 0xd6, 0x6b, 0xb2, 0x3b, 0xa5, 0xc2, 0x79, 0xc2, 0x89, 0x5f,
        0xb3, 0x98, 0x69, 0x50, 0x7b, 0x58, 0x7b, 0x12, 0x0f, 0x55, 0xff, 0xff, 0x58, 0xa9, 0xff, 0xff, 0xdc, 0xff, 0x7f, 0xff, 0xff, 0xff, 0xd5, 0x55};
    int gt = 0, decided = 0;
    for (int k = 0; k < 48; ++k) {
        const int a = y[k], h = half[k];
        gt |= (!decided) & (a > h);
        decided |= (a != h);
    }
    return gt;
}
__device__ __forceinline__ int be_is_zero(const uint8_t* p, int n) {
    int any = 0;
    for (int k = 0; k < n; ++k) any |= p[k];
    return any == 0;
}

// PointG1#toHex (index.ts:359-381) / PointG2#toHex (index.ts:604-631) from affine C-ABI coordinates; the affine image (0, 0)
// of ZERO (math.ts:955) encodes the point at infinity.  g2 = 0: in 96 B -> out 48 / 96 B; g2 = 1: in 192 B -> out 96 / 192 B
__global__ void encode_points_kernel(const uint8_t* in, uint8_t* out, size_t n, int g2, int compressed) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int fe = 48, inb = g2 ? 192 : 96;
    const int outb = g2 ? (compressed ? 96 : 192) : (compressed ? 48 : 96);
    const uint8_t* a = in + (size_t)inb * i;
    uint8_t* o = out + (size_t)outb * i;
    if (be_is_zero(a, inb)) {
        for (int k = 0; k < outb; ++k) o[k] = 0;
        o[0] = compressed ? 0xC0 : 0x40;
        return;
    }
    if (!g2) {
        for (int k = 0; k < (compressed ? fe : 2 * fe); ++k) o[k] = a[k];
        if (compressed) o[0] |= (uint8_t)(0x80 | (be48_gt_half(a + fe) << 5));
        return;
    }
    // C-ABI order x.c0 x.c1 y.c0 y.c1 ; wire order x.c1 x.c0 [y.c1 y.c0]
    const uint8_t *x0 = a, *x1 = a + fe, *y0 = a + 2 * fe, *y1 = a + 3 * fe;
    for (int k = 0; k < fe; ++k) { o[k] = x1[k]; o[fe + k] = x0[k]; }
    if (compressed) {
        const int flag = be_is_zero(y1, fe) ? be48_gt_half(y0) : be48_gt_half(y1);
        o[0] |= (uint8_t)(0x80 | (flag << 5));
    } else {
        for (int k = 0; k < fe; ++k) { o[2 * fe + k] = y1[k]; o[3 * fe + k] = y0[k]; }
    }
}

// flag byte of an uncompressed encoding (index.ts:317, 533-538, 565-568): overrides the status of the validity program
__global__ void uncompressed_flags_kernel(const uint8_t* in, int32_t* status, uint8_t* out, size_t n, int g2) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int inb = g2 ? 192 : 96;
    const uint8_t b0 = in[(size_t)inb * i];
    int32_t v = status[i];
    if (g2) {
        const uint8_t m = b0 & 0xE0;
        if (m == 0x20 || m == 0x60 || m == 0xE0 || (m & 0x80)) v = BLS381_ST_BAD_ENCODING;  // 'Invalid encoding flag' / not the uncompressed form
        else if (b0 & 0x40) v = BLS381_ST_INFINITY;
    } else if (b0 & 0x40) {
        v = BLS381_ST_INFINITY;
    }
    status[i] = v;
    if (v == BLS381_ST_INFINITY || v == BLS381_ST_BAD_ENCODING)
        for (int k = 0; k < inb; ++k) out[(size_t)inb * i + k] = 0;
}

static int encode_host(bool g2, const uint8_t* in, size_t n, int compressed, uint8_t* out) {
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!in || !out) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    const size_t inb = g2 ? 192 : 96, outb = g2 ? (compressed ? 96 : 192) : (compressed ? 48 : 96);
    int rc;
    if ((rc = stage(0, n * inb)) || (rc = stage(2, n * outb))) return rc;
    cudaStream_t s = g.stream;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], in, n * inb, cudaMemcpyHostToDevice, s));
    encode_points_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g.d_stage[0], g.d_stage[2], n, g2 ? 1 : 0, compressed ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, g.d_stage[2], n * outb, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return BLS381_OK;
}

int bls381_g1_encode_batch(const uint8_t* g1_affine, size_t n, int compressed, uint8_t* out) {
    Enter enter_;
    return encode_host(false, g1_affine, n, compressed, out);
}

int bls381_g2_encode_batch(const uint8_t* g2_affine, size_t n, int compressed, uint8_t* out) {
    Enter enter_;
    return encode_host(true, g2_affine, n, compressed, out);
}

static int from_uncompressed_host(bool g2, const uint8_t* in, size_t n, uint8_t* out, int32_t* status) {
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!in || !out || !status) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    const uint32_t ab = g2 ? 192 : 96;
    int rc;
    if ((rc = stage(0, n * ab)) || (rc = stage(2, n * ab)) || (rc = stage(6, n * 4))) return rc;
    cudaStream_t s = g.stream;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], in, n * ab, cudaMemcpyHostToDevice, s));
    if ((rc = run3(g2 ? "g2_from_uncompressed" : "g1_from_uncompressed", g.d_stage[0], ab, g.d_stage[2], ab, (int32_t*)g.d_stage[6], n, s))) return rc;
    uncompressed_flags_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g.d_stage[0], (int32_t*)g.d_stage[6], g.d_stage[2], n, g2 ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, g.d_stage[2], n * ab, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(status, g.d_stage[6], n * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return BLS381_OK;
}

int bls381_g1_from_uncompressed_batch(const uint8_t* in96, size_t n, uint8_t* out96, int32_t* status) {
    Enter enter_;
    return from_uncompressed_host(false, in96, n, out96, status);
}

int bls381_g2_from_uncompressed_batch(const uint8_t* in192, size_t n, uint8_t* out192, int32_t* status) {
    Enter enter_;
    return from_uncompressed_host(true, in192, n, out192, status);
}

static int validate_host(bool g2, const uint8_t* in, size_t n, int32_t* status) {
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!in || !status) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    const uint32_t ab = g2 ? 192 : 96;
    int rc;
    if ((rc = stage(0, n * ab)) || (rc = stage(6, n * 4))) return rc;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], in, n * ab, cudaMemcpyHostToDevice, g.stream));
    if ((rc = run3(g2 ? "g2_validate" : "g1_validate", g.d_stage[0], ab, nullptr, 0, (int32_t*)g.d_stage[6], n, g.stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(status, g.d_stage[6], n * 4, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return BLS381_OK;
}

int bls381_g1_validate_batch(const uint8_t* g1_affine, size_t n, int32_t* status) {
    Enter enter_;
    return validate_host(false, g1_affine, n, status);
}

int bls381_g2_validate_batch(const uint8_t* g2_affine, size_t n, int32_t* status) {
    Enter enter_;
    return validate_host(true, g2_affine, n, status);
}

static int scalar_mul_host(bool g2, const uint8_t* pts, const uint8_t* scalars32, size_t n, uint8_t* out, int32_t* flags) {
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!pts || !scalars32 || !out || !flags) return fail(BLS381_EINVAL, "null argument");
    if (n == 0) return BLS381_OK;
    const uint32_t ab = g2 ? 192 : 96;
    int rc;
    if ((rc = stage(0, n * ab)) || (rc = stage(7, n * 32)) || (rc = stage(2, n * ab)) || (rc = stage(6, n * 4))) return rc;
    cudaStream_t s = g.stream;
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[0], pts, n * ab, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(g.d_stage[7], scalars32, n * 32, cudaMemcpyHostToDevice, s));
    uint8_t* bufs[6] = {g.d_stage[0], g.d_stage[7], g.d_stage[2], nullptr, nullptr, g.d_stage[6]};
    uint32_t strides[6] = {ab, 32, ab, 0, 0, 4};
    if ((rc = vm_run(g2 ? "g2_scalar_mul" : "g1_scalar_mul", bufs, strides, 6, n, s))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, g.d_stage[2], n * ab, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(flags, g.d_stage[6], n * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemsetAsync(g.d_stage[7], 0, n * 32, s));  // scalars may be secret keys (sign(PointG2, key))
    CUDA_TRY(cudaStreamSynchronize(s));
    return BLS381_OK;
}

int bls381_g2_scalar_mul_batch(const uint8_t* g2_affine, const uint8_t* scalars32, size_t n, uint8_t* out192, int32_t* flags) {
    Enter enter_;
    return scalar_mul_host(true, g2_affine, scalars32, n, out192, flags);
}

int bls381_g1_scalar_mul_batch(const uint8_t* g1_affine, const uint8_t* scalars32, size_t n, uint8_t* out96, int32_t* flags) {
    Enter enter_;
    return scalar_mul_host(false, g1_affine, scalars32, n, out96, flags);
}

int bls381_imad_peak(double* imad_per_second) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!imad_per_second) return fail(BLS381_EINVAL, "null argument");
    const int blocks = g.sm_count * 8, threads = 256, iters = 4096;
    uint32_t* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)blocks * threads * 4));
    imad_peak_kernel<<<blocks, threads, 0, g.stream>>>(d, 64, nullptr);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
        imad_peak_kernel<<<blocks, threads, 0, g.stream>>>(d, iters, nullptr);
        CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, g.ev0, g.ev1);
        const double ops = (double)blocks * threads * iters * 24.0;
        best = std::max(best, ops / (ms * 1e-3));
    }
    cudaFree(d);
    *imad_per_second = best;
    return BLS381_OK;
}

int bls381_imad_peak_sustained(double seconds, double* imad_per_second, double* sm_mhz) {
    Enter enter_;
    if (!g.inited) return fail(BLS381_ENOINIT, "bls381_init() has not been called");
    if (!imad_per_second || !(seconds > 0) || seconds > 10) return fail(BLS381_EINVAL, "bad argument");
    const int blocks = g.sm_count * 8, threads = 256, iters = 4096;
    const double ops = (double)blocks * threads * iters * 24.0;
    uint32_t* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)blocks * threads * 4));
    // back-to-back launches for `seconds`; the rate is taken over the second half, when clocks and power have settled
    const int chunk = 16;
    double elapsed = 0, rate = 0;
    while (elapsed < seconds) {
        CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
        for (int k = 0; k < chunk; ++k) imad_peak_kernel<<<blocks, threads, 0, g.stream>>>(d, iters, g.d_clk);
        CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, g.ev0, g.ev1);
        elapsed += ms * 1e-3;
        if (elapsed >= seconds * 0.5) rate = rate == 0 ? ops * chunk / (ms * 1e-3) : std::min(rate, ops * chunk / (ms * 1e-3));
    }
    cudaFree(d);
    *imad_per_second = rate;
    if (sm_mhz) {
        unsigned long long h[2] = {0, 0};
        CUDA_TRY(cudaMemcpy(h, g.d_clk, sizeof(h), cudaMemcpyDeviceToHost));
        *sm_mhz = h[1] ? (double)h[0] / (double)h[1] * 1000.0 : 0.0;
    }
    return BLS381_OK;
}

}  // extern "C"
