#!/usr/bin/env python3
"""bench.py -- pairings/sec and verifyBatch / sign sigs/sec of the B200 engine on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n ITEMS] [--impl reference] [--full-parity]

Headline (`metric`, `value`): config 2, 65 536 independent pairings per GPU.  One "step" = one pass of the hot path (Miller
loop + final exponentiation) over the batch.  Inputs follow SURVEY 8d: the reference's 1000 kilic pairs (i G1, i G2) as
prefix, then P_i = a_i G1, Q_i = b_i G2 with scalars from a SHA-256 counter-mode PRNG (seed 0xB200 + rank).  `value` =
whole-job pairings/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` = the same
through the C-ABI host entry point with pinned host buffers (H2D + kernel + D2H inside the timed region).  `parity` = EVERY
output of the timed batch byte-compared with the C oracle (n_ok / n), the prefix also with the reference's fixture file.
Further sections of the same JSON line:
    verify_batch         config 3 (262 144 signatures per GPU; 8 GPUs = config 5, weak scaling) incl. negative controls at size
    verify_batch_strong  config 5 as a fixed 2 097 152-signature batch over the N GPUs (strong scaling) + the bit-for-bit check
                         "sharded result == single-GPU result" on the un-exponentiated 576 bytes
    sign                 config 4 (1 048 576 signatures, one GPU): KAT prefix + strided sample (or --full-parity: all) vs the C oracle
    table                the rows of BASELINE.md section 3
Multi-GPU: one process per GPU (torchrun); independent pairings shard by index with no collective; verifyBatch exchanges the
device-resident 576-byte partial products with ONE NCCL all-gather (no host bounce).  tools/bench_multi.py measures the
in-process multi-GPU entry point (bls381_verify_batch_multi) the Node addon uses.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP_MUL_PER_PAIRING = 15.4e3  # 14 960 Fp-mul + 1 inversion
IMAD_PER_FP_MUL = 288  # 12x12 product + 12x12 Montgomery reduction, 32x32->64 multiply-adds
IMAD_PER_PAIRING = FP_MUL_PER_PAIRING * IMAD_PER_FP_MUL  # ~4.44 M
BYTES_PER_PAIRING = 288 + 576  # algorithmic HBM bytes


def issued_imad_per_item(program_path, wide_per_product=144):
    """IMAD.WIDE the interpreter executes per item for a tower-VM program image: products x 144 + Montgomery reductions
    x 156, counted from the records (256 bytes each).  One-output records: opcode 1, T = bits 16..19 of word 0.
    Two-output records: opcode 5, entries per group in word 1 (4 x 5 bits; a fused entry pair is ONE product and shows as
    two entries whose first x word has flag bit 31), one reduction per result (one for the H_SINGLE form, bit 24)."""
    import struct
    raw = open(program_path, "rb").read()
    _magic, _ver, warps, nrec, nconst, _ns, _nf, _st = struct.unpack_from("<8I", raw, 0)
    base = 32 + nconst * 48
    products = reductions = 0
    for k in range(warps * nrec):
        w = struct.unpack_from("<64I", raw, base + k * 256)
        op = w[0] & 0xFF
        if op == 1:
            t = (w[0] >> 16) & 0xF
            products += t
            reductions += 1 if t else 0
        elif op == 5:
            n = [(w[1] >> (5 * g)) & 31 for g in range(4)]
            ent = sum(n)
            fused = sum(1 for e in range(ent) if w[4 + 2 * e] >> 31)
            products += ent - fused
            single = (w[0] >> 24) & 1
            shared = (n[0] + n[1]) > 0
            reductions += 1 if single else (2 if shared else (1 if n[2] else 0) + (1 if n[3] else 0))
    return products * wide_per_product + reductions * 156, products, reductions


# 384-bit products / Montgomery reductions per message of the hand-written per-item kernels (csrc/swu_g2.cuh,
# csrc/g2_kernels.cuh), counted by the host build of the same source (tests/test_vm_ingest_emu.py pins these numbers to the
# counters); a product = 144, a reduction = 156 IMAD.WIDE.  The Fp inversion of the affine conversion is not included
# (the tower-VM image counts leave its inversion record out as well).
KERNEL_COUNTS = {"swu_g2_kernel": (2030, 1964), "h2g2_tail_kernel": (3842, 2217), "sign_kernel": (8382, 4738),
                 "g1_decompress_kernel": (1632, 1446)}


def kernel_imad_per_item(name):
    p, r = KERNEL_COUNTS[name]
    return p * 144 + r * 156, p, r


class ClockSampler:
    """One long-lived `nvidia-smi -lms 50` process (spawning a fresh nvidia-smi per sample takes longer than a step);
    samples are time-stamped so that only those inside the timed region are summarised."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", os.environ.get("LOCAL_RANK", "0"), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) >= 7:
                self.samples.append((time.time(), f))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1, kernel_mhz=None):
        """median SM clock / reasons over the samples taken in [t0, t1] (the timed regions)"""
        inside = [f for (t, f) in self.samples if t0 <= t <= t1]
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(inside)}
        if kernel_mhz:
            out["sm_mhz_in_kernel"] = kernel_mhz
            out["note"] = "sm_mhz_in_kernel = SM cycles / nanoseconds counted by CTA 0 of the last timed launch (clock64 vs globaltimer): an exact in-kernel clock next to the coarse nvidia-smi samples"
        if not self.samples:
            out["reasons"] = ["nvidia-smi unavailable"]
            return out
        use = inside or [f for (_, f) in self.samples]
        sm = sorted(int(float(f[0])) for f in use if f[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for f in use:
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        pw = [float(f[2]) for f in use if f[2].replace(".", "").isdigit()]
        out.update({"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                    "sm_max_mhz": int(float(use[0][1])) if use[0][1] else None, "power_w_max": max(pw) if pw else None,
                    "reasons": sorted(reasons)})
        return out


def cpu_baseline(total: int, g1=None, g2=None):
    """The reference's algorithm on all host cores: the plain-C port of math.ts/index.ts (oracle/c, 64-bit limbs,
    the reference's own Karatsuba tower formulas; pinned to the reference's golden vectors), one thread per core,
    on `total` pairings of the bench workload (the first `total` pairs of the GPU arm's batch when given: kilic prefix +
    random pairs; else (i G1, i G2)).  Returns (pairings/s, cores, n)."""
    from noble_bls12_381_b200 import synth
    from oracle import c_oracle
    cores = os.cpu_count() or 1
    if g1 is None or len(g1) < 96 * total:
        g1, g2 = synth.multiples_wire(total)
    else:
        g1, g2 = g1[: 96 * total], g2[: 192 * total]
    w = min(total, cores * 32)
    c_oracle.pairing_batch(g1[: 96 * w], g2[: 192 * w], w, True, cores)  # warm-up: tables + CPU clocks
    wall = None
    for _ in range(2):  # best of two passes (the first seconds after an idle period run at a reduced CPU clock)
        t = time.perf_counter()
        out = c_oracle.pairing_batch(g1, g2, total, True, cores)
        dt = time.perf_counter() - t
        wall = dt if wall is None else min(wall, dt)
    gold_path = os.path.join(ROOT, "tests", "golden", "pairing_kilic_1000.bin")
    if os.path.exists(gold_path):
        k = min(total, 1000)
        assert out[: 576 * k] == open(gold_path, "rb").read()[: 576 * k], "CPU baseline output differs from the reference fixtures"
    return total / wall, cores, total


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; node/tsc are absent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores * 1600, 4096)  # ~3 s of all-core CPU work per pass, two passes per step
    for _ in range(args.warmup and 1):
        cpu_baseline(cores)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, cores, n = cpu_baseline(per_step)
        vals.append(v)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "pairings/sec", "value": v, "unit": "pairings/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "config 2: independent pairings e(i*G1, i*G2), Miller loop + final exponentiation",
                   "items_per_step": n},
        "cpu_baseline": {"value": v, "unit": "pairings/s", "cores": cores, "kind": "port",
                         "sample": f"{n} pairings per step ({n // cores} per core), plain-C port of math.ts/index.ts with the reference's own tower formulas, one thread per core (node/tsc absent on this image; index.ts:719 implies ~43 pairings/s/core for the TypeScript original)"},
        "e2e": {"value": v, "unit": "pairings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def _oracle_threads(world):
    return max(1, (os.cpu_count() or 1) // max(world, 1))


def sign_section(eng, n, args):
    """BASELINE config 4: n signatures (hash-to-curve G2 + constant-time scalar multiplication + compression) on one GPU,
    through the C ABI with host buffers.  Inputs: the reference's 559 `priv:msg:sig` KATs as prefix, then seeded
    random keys / 32-byte messages.  Parity: the KAT prefix against the fixture file and a strided sample (or, with
    --full-parity, every signature) against the C oracle."""
    import ctypes
    from noble_bls12_381_b200 import synth
    dst = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"
    kats = [l.split(":") for l in open(os.path.join(ROOT, "tests", "golden", "sign_g2_vectors.txt")).read().split("\n") if l]
    t0 = time.time()
    sks, msgs = synth.signing_inputs(n, 0x5167, kats)
    packed, off = eng._pack(msgs)
    t_gen = time.time() - t0
    out = ctypes.create_string_buffer(96 * n)
    w = min(n, 4096)
    eng.sign_batch(sks[: 32 * w], msgs[:w], dst)  # warm-up (program load)
    best = None
    reps = 3
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = eng.lib.bls381_sign_batch(sks, packed, off, n, dst, len(dst), out)
        dt = time.perf_counter() - t0
        assert rc == 0, eng.lib.bls381_last_error()
        best = dt if best is None else min(best, dt)
        kernel_ms = eng.last_kernel_ms()
    sigs = out.raw
    nk = min(n, len(kats))
    kat_ok = sum(sigs[96 * i: 96 * i + 96].hex() == kats[i][2].strip().lower() for i in range(nk))
    # C oracle (checker): every signature with --full-parity, else a strided sample
    from oracle import c_oracle
    if args.full_parity:
        idx = list(range(n))
    else:
        k = min(n, args.sign_check)
        idx = sorted(set([(j * n) // k for j in range(k)] + list(range(min(n, 64)))))
    t0 = time.time()
    ref = c_oracle.sign_batch(b"".join(sks[32 * i: 32 * i + 32] for i in idx), [msgs[i] for i in idx], dst)
    t_ref = time.time() - t0
    n_ok = sum(ref[96 * j: 96 * j + 96] == sigs[96 * i: 96 * i + 96] for j, i in enumerate(idx))
    cores = os.cpu_count() or 1
    return {
        "metric": "sign sigs/sec (config 4: hash_to_curve G2 + privKey * H(m) + toSignature, host buffers through the C ABI)",
        "value": n / best, "unit": "sigs/s", "n": n, "ms": best * 1e3, "kernel_ms": kernel_ms, "kernel_value": n / (kernel_ms * 1e-3),
        "parity": {"kats_ok": kat_ok, "kats": nk, "n_ok": n_ok, "n": len(idx), "checked": "all" if args.full_parity else "strided sample + first 64",
                   "checker": "oracle/c (pinned on the 559 reference KATs)"},
        "cpu_baseline": {"value": len(idx) / t_ref, "unit": "sigs/s", "cores": cores, "kind": "port", "per_core": len(idx) / t_ref / cores,
                         "sample": f"{len(idx)} signatures of the same batch on {cores} host threads (plain-C port of index.ts:746-752)"},
        "input_generation_s": t_gen,
    }


def verify_section(eng, n, rank, world, dist, torch, tag, controls=True, check_single_gpu=False):
    """verifyBatch on a (world x n)-item batch sharded by index, n items per GPU (BASELINE config 3 at world = 1 with
    n = 262 144; config 5 at world = 8).  SURVEY 8d inputs: sk_i seeded, msg_i 32 distinct bytes, pk_i = getPublicKey(sk_i),
    sig_i = sign(msg_i, sk_i), agg = aggregateSignatures(sigs) -- all produced by the engine (the C oracle checks a sample).
    Timed: host buffers in, verdict out.  world = 1: ONE C-ABI call (bls381_verify_batch).  world > 1: every rank runs
    bls381_verify_batch_partial_dev on its shard, ONE all-gather of the device-resident W x 576-byte partial products
    (NCCL), product + final exponentiation on every rank (bls381_fp12_product_dev): no host bounce."""
    import ctypes
    import numpy as np
    from noble_bls12_381_b200 import dist as bdist
    from noble_bls12_381_b200 import synth
    dst = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"
    seed = 0xBA7C4
    t0 = time.time()
    sks, msgs = synth.signing_inputs(n, seed, None, first=rank * n)   # global item index = rank * n + i
    packed, off = eng._pack(msgs)
    t_gen = time.time() - t0
    pks = eng.get_public_key_batch(sks)
    pk_ms = eng.last_kernel_ms()
    eng.sign_batch(sks[: 32 * 64], msgs[:64], dst)  # warm-up (program load)
    sigs = eng.sign_batch(sks, msgs, dst)
    agg_s = None
    for _ in range(2):  # best of two: the first full-size call also grows the library's staging / scratch buffers
        t0 = time.perf_counter()
        agg, st = eng.aggregate_g2(sigs, n)
        dt = time.perf_counter() - t0
        agg_s = dt if agg_s is None else min(agg_s, dt)
    t0 = time.perf_counter()
    agg_pk, _ = eng.aggregate_g1(pks[: 48 * min(n, 2048)], min(n, 2048))  # test/benchmark.js:94 aggregatePublicKeys/2048
    agg_pk_s = time.perf_counter() - t0
    # sample check of the generated inputs against the C oracle
    from oracle import c_oracle
    k = min(n, 256)
    gen_ok = (c_oracle.get_public_key_batch(sks[: 32 * k]) == pks[: 48 * k]) and (c_oracle.sign_batch(sks[: 32 * k], msgs[:k], dst) == sigs[: 96 * k])
    if world > 1:  # global aggregate signature = sum of the per-rank aggregates (set-up, untimed)
        t = torch.frombuffer(bytearray(agg), dtype=torch.uint8).cuda()
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        agg, _ = eng.aggregate_g2(b"".join(bytes(o.cpu().numpy().tobytes()) for o in outs), world)
    st_buf = (ctypes.c_int32 * (n + 1))()
    sig_arg = agg if rank == 0 else None
    np_ = n + (1 if rank == 0 else 0)
    verdict = ctypes.c_int(0)
    stream = torch.cuda.current_stream()
    # device-side exchange buffers: 576-byte partial + one status-level byte, padded to 640
    mine = torch.zeros(640, dtype=torch.uint8, device="cuda")
    gathered = torch.zeros(640 * world, dtype=torch.uint8, device="cuda")
    parts = torch.zeros(576 * world, dtype=torch.uint8, device="cuda")
    result = torch.zeros(576, dtype=torch.uint8, device="cuda")
    lvl_host = torch.zeros(64, dtype=torch.uint8).pin_memory()

    def run(msg_buf=packed, pk_buf=pks, sig=agg):
        """-> verdict (1 true, 0 false, -1 the reference would throw)"""
        if world == 1:  # the fused single-GPU entry point: one C-ABI call, one synchronisation
            rc = eng.lib.bls381_verify_batch(sig, msg_buf, off, pk_buf, n, dst, len(dst), ctypes.byref(verdict), st_buf)
            assert rc == 0, eng.lib.bls381_last_error()
            return verdict.value
        rc = eng.lib.bls381_verify_batch_partial_dev(sig if rank == 0 else None, msg_buf, off, pk_buf, n, dst, len(dst), mine.data_ptr(), st_buf)
        assert rc == 0, eng.lib.bls381_last_error()
        lvl_host[0] = bdist._level(np.frombuffer(st_buf, dtype=np.int32, count=np_))
        mine[576:640].copy_(lvl_host, non_blocking=True)
        dist.all_gather_into_tensor(gathered, mine)  # the single exchange step: W x 640 bytes, device to device (NCCL)
        g2d = gathered.view(world, 640)
        parts.view(world, 576).copy_(g2d[:, :576])
        rc = eng.lib.bls381_fp12_product_dev(parts.data_ptr(), world, 1, result.data_ptr(), stream.cuda_stream)
        assert rc == 0, eng.lib.bls381_last_error()
        lvl = int(g2d[:, 576].max().item())
        res = bytes(result.cpu().numpy().tobytes())
        return -1 if lvl == 2 else (0 if lvl == 1 else (1 if res == bdist.FP12_ONE else 0))

    ok = run() == 1  # warm-up + correctness
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        ok = (run() == 1) and ok
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt[0])
    out = {
        "metric": "verifyBatch sigs/sec (end to end, host buffers, sharded by index, one all-gather of W x 576 B device to device)",
        "config": tag, "value": world * n / dt, "unit": "sigs/s", "sigs_per_gpu": n, "sigs_total": world * n, "n_gpus": world,
        "ms": dt * 1e3, "reps": reps, "verdict_true": bool(ok), "inputs_match_c_oracle_sample": bool(gen_ok),
        "aggregate_signatures": {"value": n / agg_s, "unit": "sigs/s per GPU (decompress + validate + tree sum)", "n": n},
        "aggregate_public_keys_2048": {"ms": agg_pk_s * 1e3, "note": "test/benchmark.js:94; latency-bound single batch"},
        "get_public_key": {"value": n / (pk_ms * 1e-3), "unit": "keys/s per GPU (kernel)"},
        "input_generation_s": t_gen,
    }
    if controls:
        # negative controls AT CONFIG SIZE on this rank's shard: every control is one more full call
        ctl = {}
        bad = bytearray(packed)
        bad[32 * (n // 2)] ^= 1
        ctl["flipped_message_byte"] = run(msg_buf=bytes(bad))
        sw = bytearray(pks)
        sw[0:48], sw[48 * (n - 1): 48 * n] = pks[48 * (n - 1): 48 * n], pks[0:48]
        ctl["swapped_public_keys"] = run(pk_buf=bytes(sw))
        inf = bytearray(pks)
        inf[48 * (n // 3): 48 * (n // 3) + 48] = bytes([0xC0]) + bytes(47)
        ctl["infinity_public_key"] = run(pk_buf=bytes(inf))
        # an off-subgroup key is covered by the test suite (tests/test_gpu_ingest.py searches one);
        # here: a key whose x has no square root -> the reference throws ("Invalid compressed G1 point")
        junk = bytearray(pks)
        junk[48 * (n // 5): 48 * (n // 5) + 48] = bytes([0x80]) + bytes(46) + b"\x07"  # x = 7: 7^3 + 4 is not a square
        ctl["undecodable_public_key"] = run(pk_buf=bytes(junk))
        if world > 1:
            t = torch.tensor([ctl[k] for k in sorted(ctl)], dtype=torch.int32, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ctl = {k: int(v) for k, v in zip(sorted(ctl), t.tolist())}
        ctl["expected"] = {"flipped_message_byte": 0, "swapped_public_keys": 0, "infinity_public_key": 0, "undecodable_public_key": -1}
        ctl["ok"] = all(ctl[k] == v for k, v in ctl["expected"].items())
        out["negative_controls"] = ctl
    if check_single_gpu and world > 1:
        # config 5: the sharded result equals the single-GPU result on the same data bit for bit -- compared on the
        # UN-exponentiated 576 bytes: product of the W gathered partials vs ONE GPU running all world x n items
        rc = eng.lib.bls381_verify_batch_partial_dev(sig_arg, packed, off, pks, n, dst, len(dst), mine.data_ptr(), st_buf)
        assert rc == 0
        dist.all_gather_into_tensor(gathered, mine)
        parts.view(world, 576).copy_(gathered.view(world, 640)[:, :576])
        raw = torch.zeros(576, dtype=torch.uint8, device="cuda")
        assert eng.lib.bls381_fp12_product_dev(parts.data_ptr(), world, 0, raw.data_ptr(), stream.cuda_stream) == 0
        sharded = bytes(raw.cpu().numpy().tobytes())
        same = None
        if rank == 0:
            all_sks, all_msgs = synth.signing_inputs(world * n, seed, None, first=0)
            all_pks = eng.get_public_key_batch(all_sks)
            whole, st_all = eng.verify_batch_partial(agg, all_msgs, all_pks, dst)
            same = (whole == sharded) and not np.any(st_all)
        out["sharded_equals_single_gpu_bit_for_bit"] = same
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=65536, help="pairings per step per GPU")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--program-dir", default=None, help="alternative tower-VM program directory (tuning)")
    ap.add_argument("--wide-per-product", type=int, default=144, help="IMAD.WIDE per 384x384 product of the built Fp core")
    ap.add_argument("--verify-n", type=int, default=262144, help="signatures per GPU for the verifyBatch section: BASELINE config 3 (262 144 on one GPU; x8 GPUs = config 5); 0 = skip")
    ap.add_argument("--sign-n", type=int, default=1048576, help="signatures of the sign section (BASELINE config 4, one GPU); 0 = skip")
    ap.add_argument("--sign-check", type=int, default=4096, help="signatures of the sign section compared with the C oracle (strided sample)")
    ap.add_argument("--strong-total", type=int, default=2097152, help="total signatures of the strong-scaling verifyBatch row (config 5: fixed batch over 1/2/4/8 GPUs); 0 = skip")
    ap.add_argument("--full-parity", action="store_true", help="compare EVERY signature of the sign section with the C oracle (minutes of CPU time)")
    ap.add_argument("--parity-n", type=int, default=-1, help="pairings compared with the C oracle per GPU (-1 = all)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import noble_bls12_381_b200 as bls
    from noble_bls12_381_b200 import synth
    eng = bls.Engine(local_rank, args.program_dir)
    n = args.n
    # inputs: the reference's 1000 kilic pairs (i G1, i G2) as prefix, then P_i = a_i G1, Q_i = b_i G2
    # with a_i, b_i from a SHA-256 counter-mode PRNG (seed 0xB200 + rank); every rank processes its own n items
    t_gen = time.time()
    g1, g2 = synth.random_pairs_wire(eng, n, seed=0xB200 + rank, prefix=1000)
    t_gen = time.time() - t_gen
    h1 = torch.frombuffer(bytearray(g1), dtype=torch.uint8).pin_memory()
    h2 = torch.frombuffer(bytearray(g2), dtype=torch.uint8).pin_memory()
    hout = torch.empty(576 * n, dtype=torch.uint8).pin_memory()
    d1, d2 = h1.cuda(), h2.cuda()
    dout = torch.zeros(576 * n, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def step_resident():
        eng.pairing_batch_dev(d1.data_ptr(), d2.data_ptr(), n, True, dout.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()

    sampler = ClockSampler()
    time.sleep(0.15)
    launches0 = eng.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_region0 = time.time()
    ev[0].record(stream)
    for i in range(args.steps):
        step_resident()
        ev[i + 1].record(stream)
    barrier()
    kernel_mhz = eng.last_kernel_sm_mhz()
    launches = eng.launch_count() - launches0
    ms_total = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    resident_out = bytes(dout.cpu().numpy().tobytes())

    # end-to-end through the C-ABI host entry point (pinned host buffers in, host buffers out, copies inside the timed region)
    e2e_steps = max(3, args.steps)
    eng.lib.bls381_pairing_batch(h1.data_ptr(), h2.data_ptr(), n, 1, hout.data_ptr(), None)  # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rc = eng.lib.bls381_pairing_batch(h1.data_ptr(), h2.data_ptr(), n, 1, hout.data_ptr(), None)
        assert rc == 0
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t_region1 = time.time()
    sampler.stop()
    # the same with the reference's validity checks of every point (index.ts:717-718) on the device
    st_arr = torch.empty(n, dtype=torch.int32).pin_memory()
    eng.lib.bls381_pairing_batch(h1.data_ptr(), h2.data_ptr(), n, 1, hout.data_ptr(), st_arr.data_ptr())
    t0 = time.perf_counter()
    for _ in range(2):
        rc = eng.lib.bls381_pairing_batch(h1.data_ptr(), h2.data_ptr(), n, 1, hout.data_ptr(), st_arr.data_ptr())
        assert rc == 0
    e2e_checked_s = (time.perf_counter() - t0) / 2
    statuses_ok = bool((st_arr == 0).all().item())

    # ---- parity: EVERY output of the timed batch against the C oracle (checker; its threads are split over the ranks),
    # the kilic prefix additionally against the reference's fixture file
    e2e_out = bytes(hout.numpy().tobytes())
    gold_path = os.path.join(ROOT, "tests", "golden", "pairing_kilic_1000.bin")
    k = min(n, 1000)
    fixtures_ok = (e2e_out[: 576 * k] == open(gold_path, "rb").read()[: 576 * k]) if os.path.exists(gold_path) else None
    from oracle import c_oracle
    pn = n if args.parity_n < 0 else min(n, args.parity_n)
    t0 = time.time()
    ref = c_oracle.pairing_batch(g1[: 96 * pn], g2[: 192 * pn], pn, True, _oracle_threads(world))
    t_parity = time.time() - t0
    n_ok = sum(ref[576 * i: 576 * i + 576] == e2e_out[576 * i: 576 * i + 576] == resident_out[576 * i: 576 * i + 576] for i in range(pn))

    t = torch.tensor([ms_total, e2e_s, e2e_checked_s, float(n_ok), float(pn)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t[:3].clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t[3:].clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        t = torch.cat([tm, ts])
    ms_total, e2e_s, e2e_checked_s, n_ok_all, pn_all = [float(x) for x in t]
    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3)
    e2e_value = world * n / e2e_s

    vb = vs = sg = None
    if args.verify_n > 0:
        vb = verify_section(eng, args.verify_n, rank, world, dist if world > 1 else None, torch,
                            "config 3 (262 144 per GPU)" if world == 1 else "config 3 sized shards, weak scaling (config 5 at 8 GPUs)",
                            controls=True, check_single_gpu=False)
    if args.strong_total > 0 and args.strong_total // world != args.verify_n:
        vs = verify_section(eng, args.strong_total // world, rank, world, dist if world > 1 else None, torch,
                            "config 5: %d signatures in total, strong scaling" % args.strong_total, controls=False, check_single_gpu=(world > 1))
    elif args.strong_total > 0 and vb is not None:
        vs = dict(vb, config="config 5: %d signatures in total, strong scaling (same run as the weak row at this GPU count)" % args.strong_total)
        if world > 1:
            chk = verify_section(eng, min(args.verify_n, 32768), rank, world, dist, torch, "bit-for-bit check", controls=False, check_single_gpu=True)
            vs["sharded_equals_single_gpu_bit_for_bit"] = chk.get("sharded_equals_single_gpu_bit_for_bit")
            vs["sharded_equals_single_gpu_note"] = "checked on %d signatures per GPU" % min(args.verify_n, 32768)
    if args.sign_n > 0 and world == 1:
        sg = sign_section(eng, args.sign_n, args)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        imad_peak = eng.imad_peak()
        imad_sustained, imad_sustained_mhz = eng.imad_peak_sustained(1.0)
        prog_dir = args.program_dir or os.path.join(ROOT, "noble_bls12_381_b200", "programs")

        def issued(name):
            return issued_imad_per_item(os.path.join(prog_dir, name + ".b2vm"), args.wide_per_product)

        IMAD_ISSUED_PER_PAIRING, n_products, n_reductions = issued("pairing")
        per_gpu = n / (ms_per_step * 1e-3)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": "pairings/sec", "value": value, "unit": "pairings/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "config 2: %d independent pairings per GPU on random (a_i G1, b_i G2) (SHA-256 counter-mode scalars, seed 0xB200 + rank; the reference's 1000 kilic pairs as prefix), Miller loop + final exponentiation, every output byte-compared" % n,
                       "items_per_step_per_gpu": n, "parallelism": "shard-by-index x%d, no collective" % world,
                       "cache": "inputs+outputs+VM scratch (%.0f MB) are re-streamed every step; compute-bound kernel (0.9 KB/pairing), no L2 flush needed" % ((288 + 576) * n / 1e6)},
            "parity": {"n_ok": int(n_ok_all), "n": int(pn_all), "checker": "oracle/c pairing (pinned on the reference fixtures), all ranks",
                       "kilic_prefix_vs_reference_fixture_file": fixtures_ok, "seconds": t_parity},
            "parity_first_1000_vs_reference_fixtures": fixtures_ok,
            "kernel_ms_per_step": kernel_ms,
            "roofline": {
                "bound": "imad", "achieved": per_gpu * IMAD_PER_PAIRING / 1e12, "peak": imad_peak / 1e12, "unit": "TIMAD/s",
                "frac": per_gpu * IMAD_PER_PAIRING / imad_peak,
                "peak_sustained": imad_sustained / 1e12, "peak_sustained_sm_mhz": imad_sustained_mhz,
                "frac_sustained": per_gpu * IMAD_PER_PAIRING / imad_sustained,
                "issued_per_pairing": {"imad_wide": IMAD_ISSUED_PER_PAIRING, "products": n_products, "reductions": n_reductions},
                "issued": per_gpu * IMAD_ISSUED_PER_PAIRING / 1e12, "issued_frac_sustained": per_gpu * IMAD_ISSUED_PER_PAIRING / imad_sustained,
                "traffic": 4776448, "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE vm_kernel launch over 9472 pairings (ncu --set full, profiles/r2_final2_ncu_pairing_sign.txt): the 2.7 MB algorithmic bytes plus the program image; results are still in L2 when the kernel ends.  The kernel is compute-bound: the HBM fraction is ~1e-4",
                "note": "integer-multiply pipe bound (SURVEY 8d): achieved = pairings/s/GPU x 15.4k Fp-mul x 288 IMAD; peak = IMAD.WIDE.U32 issue rate measured live on this GPU in a 3 ms burst (bls381_imad_peak, boost clock); peak_sustained = the same microbenchmark back to back for 1 s (power-settled clock, the fair denominator for a step this long); issued = the multiply-adds the kernel actually executes (lazy-reduction program: more products, fewer reductions)",
                "hbm": {"achieved": per_gpu * BYTES_PER_PAIRING / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": per_gpu * BYTES_PER_PAIRING / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            },
            "e2e": {"value": e2e_value, "unit": "pairings/s", "h2d_bytes_per_step": 288 * n, "d2h_bytes_per_step": 576 * n, "steps": e2e_steps},
            "e2e_with_validity_checks": {"value": world * n / e2e_checked_s, "unit": "pairings/s", "all_status_ok": statuses_ok,
                                         "note": "bls381_pairing_batch with a status array: P.assertValidity() + Q.assertValidity() (index.ts:717-718) of every point on the device"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(t_region0, t_region1, kernel_mhz),
            "input_generation_s": t_gen,
        }
        if vb is not None:
            line["verify_batch"] = vb
        if vs is not None:
            line["verify_batch_strong"] = vs
        if sg is not None:
            line["sign"] = sg
        # per-section issued-multiply rooflines (issued IMAD.WIDE of the program images / kernel time)
        sect = {}
        if sg is not None:
            # sign = swu_g2_kernel (hash_to_field + SWU) + sign_kernel (tail of hash-to-curve, ladder, toSignature)
            kw, kp, kr = kernel_imad_per_item("swu_g2_kernel")
            sw, sp, sr = kernel_imad_per_item("sign_kernel")
            sect["sign"] = {"imad_wide_per_item": kw + sw, "products": kp + sp, "reductions": kr + sr,
                            "kernels": {"swu_g2_kernel": {"imad_wide_per_item": kw}, "sign_kernel": {"imad_wide_per_item": sw}},
                            "issued_frac_sustained": sg["kernel_value"] * (kw + sw) / imad_sustained}
        line["roofline_sections"] = sect
        cpu = None
        if not args.no_cpu_baseline:
            v, cores, cnt = cpu_baseline(max((os.cpu_count() or 1) * 1600, 4096), g1, g2)  # ~10 s of all-core CPU work in total
            cpu = {"value": v, "unit": "pairings/s", "cores": cores, "kind": "port", "per_core": v / cores,
                   "sample": f"{cnt} pairings of the same workload ({cnt // cores} per core), plain-C port of math.ts/index.ts (oracle/c, reference's Karatsuba tower formulas), one thread per core, outputs checked against the reference fixtures; the TypeScript original is ~43 pairings/s/core by its own comment (index.ts:719)"}
            line["cpu_baseline"] = cpu
        # BASELINE.md section 3 table shape
        rows = [{"config": "2: independent pairings", "N": n * world, "GPUs": world, "rate": value, "unit": "pairings/s",
                 "imad_per_s": per_gpu * world * IMAD_PER_PAIRING, "frac_of_measured_imad_peak": per_gpu * IMAD_PER_PAIRING / imad_peak,
                 "hbm_gbs": per_gpu * world * BYTES_PER_PAIRING / 1e9, "bit_exact": "%d/%d" % (int(n_ok_all), int(pn_all)),
                 "cpu_baseline": None if cpu is None else {"cores": cpu["cores"], "rate": cpu["value"], "per_core": cpu["per_core"]},
                 "speed_up": None if cpu is None else value / cpu["value"]}]
        if vb is not None:
            rows.append({"config": "3/5 weak: verifyBatch", "N": vb["sigs_total"], "GPUs": world, "rate": vb["value"], "unit": "sigs/s",
                         "bit_exact": "verdict %s, controls %s" % (vb["verdict_true"], (vb.get("negative_controls") or {}).get("ok"))})
        if vs is not None:
            rows.append({"config": "5 strong: verifyBatch", "N": vs["sigs_total"], "GPUs": world, "rate": vs["value"], "unit": "sigs/s",
                         "bit_exact": "verdict %s, sharded == single GPU: %s" % (vs["verdict_true"], vs.get("sharded_equals_single_gpu_bit_for_bit"))})
        if sg is not None:
            rows.append({"config": "4: sign", "N": sg["n"], "GPUs": 1, "rate": sg["value"], "unit": "sigs/s",
                         "bit_exact": "%d/%d KATs, %d/%d vs oracle" % (sg["parity"]["kats_ok"], sg["parity"]["kats"], sg["parity"]["n_ok"], sg["parity"]["n"]),
                         "cpu_baseline": {"cores": sg["cpu_baseline"]["cores"], "rate": sg["cpu_baseline"]["value"], "per_core": sg["cpu_baseline"]["per_core"]},
                         "speed_up": sg["value"] / sg["cpu_baseline"]["value"]})
        line["table"] = rows
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
