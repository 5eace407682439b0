"""CPU-emulation parity tests of the ingest programs (decompression + validity, hash-to-curve) vs the oracle."""
import random
import struct

import pytest

from noble_bls12_381_b200.vmprog import compile as vmcompile
from noble_bls12_381_b200.vmprog import curves
from oracle import noble_oracle as O
from tests.emu import emu
import os

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _g1_cases():
    rng = random.Random(2)
    g1c = open(os.path.join(GOLDEN, "zkcrypto_g1_compressed.dat"), "rb").read()
    items = [g1c[48 * i : 48 * i + 48] for i in (1, 2, 3, 50, 999)]
    items += [O.g1_to_hex(O.pt_multiply_unsafe(O.G1, O.G1_BASE, rng.randrange(1, O.R_ORDER)), True) for _ in range(4)]
    items.append(g1c[0:48])  # infinity
    x = 5
    while O.fp_sqrt((x**3 + 4) % O.P) is not None:
        x += 1
    items.append((x + (1 << 383)).to_bytes(48, "big"))  # no square root
    x = 1
    while True:
        x += 1
        y = O.fp_sqrt((x**3 + 4) % O.P)
        if y is not None and not O.g1_is_torsion_free((x, y, 1)):
            break
    items.append((x + (1 << 383)).to_bytes(48, "big"))  # on curve, outside the subgroup
    items.append((x + (1 << 383) + (1 << 381)).to_bytes(48, "big"))
    items += _g1_noncanonical(g1c)
    return items


def _g1_noncanonical(g1c):
    """Encodings the reference accepts although they are not canonical (index.ts:304-314): an x coordinate >= p in the
    381 value bits (`new Fp` reduces it), a missing compression bit (never looked at), an infinity flag with a
    non-zero body (returns ZERO before the body is read)."""
    out = []
    found = 0
    for i in range(1, 1000):
        raw = g1c[48 * i: 48 * i + 48]
        v = int.from_bytes(raw, "big")
        x = v % (1 << 381)
        if x + O.P < (1 << 381):
            out.append((v + O.P).to_bytes(48, "big"))              # x + p, same flags
            found += 1
            if found == 3:
                break
    assert found == 3
    v = int.from_bytes(g1c[48 * 7: 48 * 8], "big")
    out.append((v - (1 << 383)).to_bytes(48, "big"))                # compression bit cleared
    out.append((v | (1 << 382)).to_bytes(48, "big"))                # infinity flag + the body of 7*G
    out.append(((1 << 383) | (1 << 382) | (1 << 381) | 12345).to_bytes(48, "big"))  # infinity flag, sign flag, junk
    return out


def g1_expected(item):
    try:
        p = O.g1_from_hex(item)
        if O.pt_is_zero(O.G1, p):
            return curves.ST_INFINITY, None
        return 0, p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")
    except O.OracleError as e:
        msg = e.args[0]
        return (curves.ST_BAD_ENCODING if "compressed" in msg else curves.ST_NOT_IN_SUBGROUP), None


def _g2_cases():
    rng = random.Random(3)
    g2c = open(os.path.join(GOLDEN, "zkcrypto_g2_compressed.dat"), "rb").read()
    items = [g2c[96 * i : 96 * i + 96] for i in (1, 2, 7, 500)]
    items += [O.g2_to_signature(O.pt_multiply_unsafe(O.G2, O.G2_BASE, rng.randrange(1, O.R_ORDER))) for _ in range(3)]
    items.append(g2c[0:96])
    xx = (3, 7)
    while O.fp2_sqrt(O.fp2_add(O.fp2_pow(xx, 3), O.B2)) is not None:
        xx = (xx[0] + 1, xx[1])
    items.append((xx[1] + (1 << 383)).to_bytes(48, "big") + xx[0].to_bytes(48, "big"))
    xx = (1, 1)
    while True:
        xx = (xx[0] + 1, xx[1])
        y = O.fp2_sqrt(O.fp2_add(O.fp2_pow(xx, 3), O.B2))
        if y is not None and not O.g2_is_torsion_free((xx, y, O.FP2_ONE)):
            break
    items.append((xx[1] + (1 << 383)).to_bytes(48, "big") + xx[0].to_bytes(48, "big"))
    items.append((xx[1] + (1 << 383) + (1 << 381)).to_bytes(48, "big") + xx[0].to_bytes(48, "big"))
    items += _g2_noncanonical(g2c)
    return items


def _g2_noncanonical(g2c):
    """Non-canonical signatures the reference accepts (index.ts:506-514): z2 >= p (48 full bytes, reduced by `new Fp`),
    z1 mod 2^381 >= p, a missing compression bit, an infinity flag with a non-zero body."""
    out = []
    for i in (3, 11):
        raw = g2c[96 * i: 96 * i + 96]
        z1, z2 = int.from_bytes(raw[:48], "big"), int.from_bytes(raw[48:], "big")
        out.append(z1.to_bytes(48, "big") + (z2 + O.P).to_bytes(48, "big"))          # real part + p
        out.append(z1.to_bytes(48, "big") + (z2 + 2 * O.P).to_bytes(48, "big"))      # real part + 2p (< 2^384)
    found = 0
    for i in range(1, 1000):
        raw = g2c[96 * i: 96 * i + 96]
        z1 = int.from_bytes(raw[:48], "big")
        if z1 % (1 << 381) + O.P < (1 << 381):
            out.append((z1 + O.P).to_bytes(48, "big") + raw[48:])                    # imaginary part + p, same flags
            found += 1
            if found == 2:
                break
    assert found == 2
    raw = g2c[96 * 5: 96 * 6]
    z1 = int.from_bytes(raw[:48], "big")
    out.append((z1 - (1 << 383)).to_bytes(48, "big") + raw[48:])                     # compression bit cleared
    out.append((z1 | (1 << 382)).to_bytes(48, "big") + raw[48:])                     # infinity flag + body
    return out


def g2_expected(item):
    try:
        p = O.g2_from_signature(item)
        if O.pt_is_zero(O.G2, p):
            return curves.ST_INFINITY, None
        (x0, x1), (y0, y1) = p[0], p[1]
        return 0, b"".join(v.to_bytes(48, "big") for v in (x0, x1, y0, y1))
    except O.OracleError as e:
        msg = e.args[0]
        return (curves.ST_NO_SQRT if "square root" in msg else curves.ST_NOT_IN_SUBGROUP), None


def test_g1_decompress_program():
    b = vmcompile.compile_program("g1_decompress")
    items = _g1_cases()
    n = len(items)
    out, st = bytearray(96 * n), bytearray(4 * n)
    emu.run_program(b, {0: (bytearray(b"".join(items)), 48), 2: (out, 96), 5: (st, 4)}, n)
    status = struct.unpack("<%di" % n, st)
    for i, it in enumerate(items):
        exp_st, exp = g1_expected(it)
        assert status[i] == exp_st, i
        if exp is not None:
            assert bytes(out[96 * i : 96 * i + 96]) == exp, i


def test_g1_decompress_kernel_source():
    """csrc/g2_kernels.cuh g1_decompress_one (host build of the kernel's source) on the same cases as the tower-VM program:
    zkcrypto vectors, infinity, no square root, outside the subgroup, non-canonical encodings."""
    import ctypes
    items = _g1_cases()
    n = len(items)
    out = (ctypes.c_uint8 * (96 * n))()
    st = (ctypes.c_int32 * n)()
    emu.lib().emu_g1_decompress(b"".join(items), out, st, ctypes.c_size_t(n))
    out = bytes(out)
    for i, it in enumerate(items):
        exp_st, exp = g1_expected(it)
        assert st[i] == exp_st, i
        if exp is not None:
            assert out[96 * i : 96 * i + 96] == exp, i


def test_g2_decompress_program():
    b = vmcompile.compile_program("g2_decompress")
    items = _g2_cases()
    n = len(items)
    out, st = bytearray(192 * n), bytearray(4 * n)
    emu.run_program(b, {0: (bytearray(b"".join(items)), 96), 2: (out, 192), 5: (st, 4)}, n)
    status = struct.unpack("<%di" % n, st)
    for i, it in enumerate(items):
        exp_st, exp = g2_expected(it)
        assert status[i] == exp_st, i
        if exp is not None:
            assert bytes(out[192 * i : 192 * i + 192]) == exp, i


def test_fixed_base_table_is_reproducible_and_correct(tmp_path):
    """csrc/g1_comb_gen.cuh is what tools/gen_g1_comb.py writes, and its entries are (d * 16^w) G1 by the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "noble_bls12_381_b200", "csrc", "g1_comb_gen.cuh")
    out = tmp_path / "comb.cuh"
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "gen_g1_comb.py"), str(out)])
    assert out.read_text() == open(hdr).read()
    rows = [l for l in open(hdr) if l.startswith("    0x")]
    assert len(rows) == 64 * 15
    rinv = pow(1 << 384, -1, O.P)
    for w, d in ((0, 1), (0, 15), (1, 1), (31, 9), (63, 15)):
        words = [int(t.strip().rstrip("u"), 16) for t in rows[w * 15 + d - 1].split("//")[0].split(",") if t.strip()]
        x = sum(v << (32 * i) for i, v in enumerate(words[:12])) * rinv % O.P
        y = sum(v << (32 * i) for i, v in enumerate(words[12:24])) * rinv % O.P
        assert (x, y) == tuple(O.pt_to_affine(O.G1, O.pt_multiply_unsafe(O.G1, O.G1_BASE, d * 16**w % O.R_ORDER))[:2])


def test_g1_fixed_base_kernel_source():
    """getPublicKey through the fixed-base table (csrc/g2_kernels.cuh g1_fixed_base_one, table from tools/gen_g1_comb.py):
    edge scalars (single nibbles, zero nibbles, all-ones nibbles, r - 1, 0) and random ones against the oracle."""
    import ctypes
    rng = random.Random(11)
    ks = [1, 2, 15, 16, 17, 0xF0, 1 << 252, (1 << 252) + 1, O.R_ORDER - 1, 0, 0x0F0F0F0F0F0F0F0F, int("f" * 62, 16)]
    ks += [rng.randrange(1, O.R_ORDER) for _ in range(12)]
    n = len(ks)
    out = (ctypes.c_uint8 * (96 * n))()
    fl = (ctypes.c_int32 * n)()
    emu.lib().emu_g1_fixed_base(b"".join(k.to_bytes(32, "big") for k in ks), out, fl, ctypes.c_size_t(n))
    out = bytes(out)
    for i, k in enumerate(ks):
        if k % O.R_ORDER == 0:
            assert fl[i] == 2, i
            continue
        x, y = O.pt_to_affine(O.G1, O.pt_multiply_unsafe(O.G1, O.G1_BASE, k))[:2]
        assert fl[i] == 0 and out[96 * i : 96 * i + 96] == x.to_bytes(48, "big") + y.to_bytes(48, "big"), i


def test_g2_decompress_kernel_source():
    """csrc/g2_kernels.cuh g2_decompress_one (host build of the kernel's source: complex-method square root, sign rule,
    psi subgroup test) on the same cases as the tower-VM program."""
    import ctypes
    items = _g2_cases()
    n = len(items)
    out = (ctypes.c_uint8 * (192 * n))()
    st = (ctypes.c_int32 * n)()
    emu.lib().emu_g2_decompress(b"".join(items), out, st, ctypes.c_size_t(n))
    out = bytes(out)
    for i, it in enumerate(items):
        exp_st, exp = g2_expected(it)
        assert st[i] == exp_st, i
        if exp is not None:
            assert out[192 * i : 192 * i + 192] == exp, i


def test_hash_to_g2_program():
    b = vmcompile.compile_program("hash_to_g2")
    msgs = [b"", b"abc", bytes(range(32)), b"x" * 100, bytes.fromhex("d2")]
    n = len(msgs)
    inb = bytearray(b"".join(O.expand_message_xmd(m, O.DEFAULT_DST, 256) for m in msgs))
    out = bytearray(192 * n)
    emu.run_program(b, {0: (inb, 256), 2: (out, 192)}, n)
    for i, m in enumerate(msgs):
        (x0, x1), (y0, y1) = O.pt_to_affine(O.G2, O.g2_hash_to_curve(m))
        assert bytes(out[192 * i : 192 * i + 192]) == b"".join(v.to_bytes(48, "big") for v in (x0, x1, y0, y1)), i


def _swu_points(uniform: bytes) -> bytearray:
    """hash_to_field + SWU of csrc/swu_g2.cuh (host build of the kernel's source): n x 256 uniform bytes -> n x 576 bytes."""
    import ctypes
    n2 = len(uniform) // 128
    out = (ctypes.c_uint8 * (288 * n2))()
    emu.lib().emu_swu_g2(bytes(uniform), out, ctypes.c_size_t(n2))
    return bytearray(out)


def test_swu_kernel_source_against_the_oracle_map():
    """map_to_curve_simple_swu_9mod16 (math.ts:1220-1267) of the hand-written kernel: both candidate families, t = 0 (the
    exceptional denominator), inputs >= p in the 64-byte chunks."""
    rng = random.Random(5)
    raws = [bytes(128), O.P.to_bytes(64, "big") * 2, b"\xff" * 128] + [bytes(rng.getrandbits(8) for _ in range(128)) for _ in range(21)]
    out = _swu_points(b"".join(raws))
    rinv = pow(1 << 384, -1, O.P)
    fam = set()
    for i, raw in enumerate(raws):
        t = (int.from_bytes(raw[:64], "big") % O.P, int.from_bytes(raw[64:], "big") % O.P)
        f = [int.from_bytes(out[288 * i + 48 * k: 288 * i + 48 * k + 48], "big") for k in range(6)]
        assert all(v < O.P for v in f)
        f = [v * rinv % O.P for v in f]
        zi = O.fp2_inv((f[4], f[5]))
        x, y = O.fp2_mul((f[0], f[1]), zi), O.fp2_mul((f[2], f[3]), zi)
        assert (x, y) == O.map_to_curve_simple_swu_9mod16(t), i
        fam.add(O.fp2_sqrt(O.fp2_add(O.fp2_add(O.fp2_mul(O.fp2_sqr(x), x), O.fp2_mul((0, 240), x)), (1012, 1012))) is not None)
    assert fam == {True}  # every output is on E'


def test_h2g2_tail_program_behind_the_swu_kernel():
    """PointG2.hashToCurve as the library runs it: xmd -> SWU kernel -> `h2g2_tail` program."""
    b = vmcompile.compile_program("h2g2_tail")
    msgs = [b"", b"abc", bytes(range(32)), b"x" * 100, bytes.fromhex("d2"), b"abcdef0123456789", b"q128_" + b"q" * 128]
    n = len(msgs)
    pts = _swu_points(b"".join(O.expand_message_xmd(m, O.DEFAULT_DST, 256) for m in msgs))
    out = bytearray(192 * n)
    emu.run_program(b, {0: (pts, 576), 2: (out, 192)}, n)
    for i, m in enumerate(msgs):
        (x0, x1), (y0, y1) = O.pt_to_affine(O.G2, O.g2_hash_to_curve(m))
        assert bytes(out[192 * i : 192 * i + 192]) == b"".join(v.to_bytes(48, "big") for v in (x0, x1, y0, y1)), i


def _scalars():
    rng = random.Random(11)
    r = O.R_ORDER
    return [1, 2, 15, 16, 17, r - 1, r, (1 << 254) + 1, (1 << 255) - 19 if (1 << 255) - 19 <= r else r - 2, 0xF0F0F0F0 << 200] + \
        [rng.randrange(1, r) for _ in range(4)]


def test_g2_scalar_mul_program_windowed():
    """Fixed-window constant-time scalar multiplication vs the oracle's ladder (math.ts:1061-1078), edge scalars."""
    b = vmcompile.compile_program("g2_scalar_mul")
    ks = _scalars()
    n = len(ks)
    q = O.pt_to_affine(O.G2, O.pt_multiply_unsafe(O.G2, O.G2_BASE, 0xC0FFEE))
    qb = b"".join(v.to_bytes(48, "big") for v in (q[0][0], q[0][1], q[1][0], q[1][1]))
    out, st = bytearray(192 * n), bytearray(4 * n)
    emu.run_program(b, {0: (bytearray(qb * n), 192), 1: (bytearray(b"".join(k.to_bytes(32, "big") for k in ks)), 32),
                        2: (out, 192), 5: (st, 4)}, n)
    status = struct.unpack("<%di" % n, st)
    qp = (q[0], q[1], O.FP2_ONE)
    for i, k in enumerate(ks):
        e = O.pt_multiply_unsafe(O.G2, qp, k)
        if O.pt_is_zero(O.G2, e):
            assert status[i] == 2, i
            continue
        (x0, x1), (y0, y1) = O.pt_to_affine(O.G2, e)
        assert status[i] == 0, i
        assert bytes(out[192 * i : 192 * i + 192]) == b"".join(v.to_bytes(48, "big") for v in (x0, x1, y0, y1)), i


def test_g1_scalar_mul_program_windowed():
    b = vmcompile.compile_program("g1_scalar_mul")
    ks = _scalars()
    n = len(ks)
    q = O.pt_to_affine(O.G1, O.pt_multiply_unsafe(O.G1, O.G1_BASE, 0xBEEF))
    qb = q[0].to_bytes(48, "big") + q[1].to_bytes(48, "big")
    out, st = bytearray(96 * n), bytearray(4 * n)
    emu.run_program(b, {0: (bytearray(qb * n), 96), 1: (bytearray(b"".join(k.to_bytes(32, "big") for k in ks)), 32),
                        2: (out, 96), 5: (st, 4)}, n)
    status = struct.unpack("<%di" % n, st)
    for i, k in enumerate(ks):
        e = O.pt_multiply_unsafe(O.G1, (q[0], q[1], 1), k)
        if O.pt_is_zero(O.G1, e):
            assert status[i] == 2, i
            continue
        x, y = O.pt_to_affine(O.G1, e)
        assert status[i] == 0, i
        assert bytes(out[96 * i : 96 * i + 96]) == x.to_bytes(48, "big") + y.to_bytes(48, "big"), i


def test_sign_program_against_reference_kats():
    """sign = hash-to-curve + windowed sk*H(m) + compression, on the first reference KATs (index.test.ts:287-293)."""
    lines = [l.split(":") for l in open(os.path.join(GOLDEN, "sign_g2_vectors.txt")).read().split("\n") if l][:6]
    b = vmcompile.compile_program("sign")
    n = len(lines)
    xmd = bytearray(b"".join(O.expand_message_xmd(bytes.fromhex(m), O.DEFAULT_DST, 256) for _, m, _ in lines))
    z = 0xD201000000010000  # the sign program takes the scalar as four base-|x| digits (api.cu: base_z_digits_kernel)

    def digits(k):
        a = [(k // z**i) % z for i in range(4)]
        assert k < z**4
        return b"".join(a[i].to_bytes(8, "big") for i in (3, 2, 1, 0))

    sks = bytearray(b"".join(digits(int(sk, 16) % O.R_ORDER) for sk, _, _ in lines))
    out, fl = bytearray(96 * n), bytearray(4 * n)
    emu.run_program(b, {0: (xmd, 256), 1: (sks, 32), 2: (out, 96), 5: (fl, 4)}, n)
    flags = struct.unpack("<%di" % n, fl)
    for i, (_, _, sig) in enumerate(lines):
        body = bytearray(out[96 * i : 96 * i + 96])
        body[0] |= 0x80 | (0x40 if flags[i] & 2 else 0) | (0x20 if flags[i] & 1 else 0)
        assert bytes(body).hex() == sig.strip().lower(), i


def test_sign_tail_program_against_reference_kats():
    """sign as the library runs it: xmd -> SWU kernel -> `sign_tail` (tail of hash-to-curve + ladder + compression)."""
    lines = [l.split(":") for l in open(os.path.join(GOLDEN, "sign_g2_vectors.txt")).read().split("\n") if l][6:12]
    b = vmcompile.compile_program("sign_tail")
    n = len(lines)
    pts = _swu_points(b"".join(O.expand_message_xmd(bytes.fromhex(m), O.DEFAULT_DST, 256) for _, m, _ in lines))
    z = 0xD201000000010000

    def digits(k):
        a = [(k // z**i) % z for i in range(4)]
        return b"".join(a[i].to_bytes(8, "big") for i in (3, 2, 1, 0))

    sks = bytearray(b"".join(digits(int(sk, 16) % O.R_ORDER) for sk, _, _ in lines))
    out, fl = bytearray(96 * n), bytearray(4 * n)
    emu.run_program(b, {0: (pts, 576), 1: (sks, 32), 2: (out, 96), 5: (fl, 4)}, n)
    flags = struct.unpack("<%di" % n, fl)
    for i, (_, _, sig) in enumerate(lines):
        body = bytearray(out[96 * i : 96 * i + 96])
        body[0] |= 0x80 | (0x40 if flags[i] & 2 else 0) | (0x20 if flags[i] & 1 else 0)
        assert bytes(body).hex() == sig.strip().lower(), i


def test_tail_kernels_source_against_oracle_and_kats():
    """csrc/g2_kernels.cuh (host build of the kernels' source): tail of hash-to-curve vs the oracle's hashToCurve, and the
    sign ladder + toSignature vs the reference's KATs (index.test.ts:287-293) incl. edge scalars vs the oracle."""
    import ctypes
    lib = emu.lib()
    msgs = [b"", b"abc", bytes(range(32)), b"x" * 100, b"abcdef0123456789"]
    n = len(msgs)
    pts = _swu_points(b"".join(O.expand_message_xmd(m, O.DEFAULT_DST, 256) for m in msgs))
    out = (ctypes.c_uint8 * (192 * n))()
    lib.emu_h2g2_tail(bytes(pts), out, ctypes.c_size_t(n))
    out = bytes(out)
    for i, m in enumerate(msgs):
        (x0, x1), (y0, y1) = O.pt_to_affine(O.G2, O.g2_hash_to_curve(m))
        assert out[192 * i : 192 * i + 192] == b"".join(v.to_bytes(48, "big") for v in (x0, x1, y0, y1)), i
    lines = [l.split(":") for l in open(os.path.join(GOLDEN, "sign_g2_vectors.txt")).read().split("\n") if l][12:24]
    z = 0xD201000000010000

    def digits(k):
        a = [(k // z**i) % z for i in range(4)]
        return b"".join(a[i].to_bytes(8, "big") for i in (3, 2, 1, 0))

    cases = [(int(sk, 16) % O.R_ORDER, bytes.fromhex(m), sig.strip().lower()) for sk, m, sig in lines]
    for k in (1, 2, z - 1, z, z + 1, z * z, z**3 - 1, O.R_ORDER - 1):  # digit patterns with zero / single / full digits
        m = b"edge%d" % (k % 1000)
        cases.append((k, m, O.g2_to_signature(O.pt_multiply_unsafe(O.G2, O.g2_hash_to_curve(m), k)).hex()))
    n = len(cases)
    pts = _swu_points(b"".join(O.expand_message_xmd(m, O.DEFAULT_DST, 256) for _, m, _ in cases))
    sig = (ctypes.c_uint8 * (96 * n))()
    lib.emu_sign_tail(bytes(pts), b"".join(digits(k) for k, _, _ in cases), sig, ctypes.c_size_t(n))
    sig = bytes(sig)
    for i, (_, _, want) in enumerate(cases):
        assert sig[96 * i : 96 * i + 96].hex() == want, i


def test_sign_ladder_digit_patterns_against_the_c_oracle():
    """The Jacobian sign ladder of csrc/g2_kernels.cuh (host build): scalars whose base-|x| digits have long leading-zero runs,
    zero digits, single bits and all-ones patterns -- the cases in which the accumulator stays at infinity, starts late, or
    adds nothing for many steps -- plus random scalars, against the C oracle's sign()."""
    import ctypes
    from oracle import c_oracle as C
    rng = random.Random(2024)
    z = 0xD201000000010000
    ks = [1 << 63, (1 << 63) + 1, z * ((1 << 63) | 1), z**2 * 3 + 1, z**3 * 5, z**3 + z, (1 << 32) - 1, 1 << 200,
          (z - 1) * (1 + z + z * z), z**3 * (1 << 60) + (1 << 62), 0x5555555555555555, 0xAAAAAAAAAAAAAAAA * z]
    ks = [k % O.R_ORDER for k in ks if k % O.R_ORDER] + [rng.randrange(1, O.R_ORDER) for _ in range(40)]
    msgs = [b"pattern %d" % i for i in range(len(ks))]

    def digits(k):
        a = [(k // z**i) % z for i in range(4)]
        return b"".join(a[i].to_bytes(8, "big") for i in (3, 2, 1, 0))

    pts = _swu_points(b"".join(O.expand_message_xmd(m, O.DEFAULT_DST, 256) for m in msgs))
    sig = (ctypes.c_uint8 * (96 * len(ks)))()
    emu.lib().emu_sign_tail(bytes(pts), b"".join(digits(k) for k in ks), sig, ctypes.c_size_t(len(ks)))
    want = C.sign_batch(b"".join(k.to_bytes(32, "big") for k in ks), msgs, O.DEFAULT_DST)
    assert bytes(sig) == want


def test_kernel_multiply_counts_used_by_the_bench():
    """bench.py's issued-multiply figures of the hand-written kernels are the counters of the host build of their source."""
    import ctypes
    import bench
    lib = emu.lib()
    c = (ctypes.c_long * 2)()
    uni = O.expand_message_xmd(b"count me", O.DEFAULT_DST, 256)
    pts = (ctypes.c_uint8 * 576)()
    out = (ctypes.c_uint8 * 192)()
    sig = (ctypes.c_uint8 * 96)()
    lib.emu_swu_counters(c)
    lib.emu_swu_g2(uni, pts, ctypes.c_size_t(2))
    lib.emu_swu_counters(c)
    assert (c[0], c[1]) == bench.KERNEL_COUNTS["swu_g2_kernel"]
    lib.emu_h2g2_tail(pts, out, ctypes.c_size_t(1))
    lib.emu_swu_counters(c)
    assert (c[0], c[1]) == bench.KERNEL_COUNTS["h2g2_tail_kernel"]
    lib.emu_sign_tail(pts, bytes(31) + b"\x05", sig, ctypes.c_size_t(1))
    lib.emu_swu_counters(c)
    assert (c[0], c[1]) == bench.KERNEL_COUNTS["sign_kernel"]
    g1c = open(os.path.join(GOLDEN, "zkcrypto_g1_compressed.dat"), "rb").read()
    o96, st = (ctypes.c_uint8 * 96)(), (ctypes.c_int32 * 1)()
    lib.emu_g1_decompress(g1c[48:96], o96, st, ctypes.c_size_t(1))
    lib.emu_swu_counters(c)
    assert (c[0], c[1]) == bench.KERNEL_COUNTS["g1_decompress_kernel"]


def test_validate_programs_edge_points():
    """assertValidity (index.ts:383-388 / 633-638) through the fixed-scalar multiplication with Jacobian doubling runs:
    subgroup points, on-curve points outside the subgroup, a point of the curve's small-order part (a multiple of r of a
    curve point: its multiples hit the point at infinity in the middle of the ladder) and off-curve points."""
    r = O.R_ORDER
    # ---- G1
    b1 = vmcompile.compile_program("g1_validate")
    pts, exp = [], []
    for k in (1, 2, 0xABCDEF):
        pts.append(O.pt_to_affine(O.G1, O.pt_multiply_unsafe(O.G1, O.G1_BASE, k))); exp.append(0)
    x = 1
    found = 0
    while found < 2:
        x += 1
        y = O.fp_sqrt((x**3 + 4) % O.P)
        if y is None:
            continue
        p = (x, y, 1)
        if O.g1_is_torsion_free(p):
            continue
        found += 1
        pts.append((x, y)); exp.append(curves.ST_NOT_IN_SUBGROUP)
        s = O.pt_multiply_unsafe(O.G1, p, r)  # lands in the cofactor part: small order
        if not O.pt_is_zero(O.G1, s):
            pts.append(O.pt_to_affine(O.G1, s)); exp.append(curves.ST_NOT_IN_SUBGROUP)
        h1 = 0x396C8C005555E1568C00AAAB0000AAAB  # G1 cofactor = 3 * 11^2 * 10177^2 * 859267^2 * 52437899^2
        for q in (3, 11):  # points of order 3 / 11: the ladder's running multiple IS the point at infinity several times
            sq = O.pt_multiply_unsafe(O.G1, s, h1 // q)  # s = [r]p (scalars above r are refused by the oracle, as by the reference)
            if not O.pt_is_zero(O.G1, sq):
                assert O.pt_is_zero(O.G1, O.pt_multiply_unsafe(O.G1, sq, q))
                pts.append(O.pt_to_affine(O.G1, sq)); exp.append(curves.ST_NOT_IN_SUBGROUP)
    pts.append((5, 7)); exp.append(curves.ST_NOT_ON_CURVE)
    n = len(pts)
    st = bytearray(4 * n)
    raw1 = b"".join(a.to_bytes(48, "big") + c.to_bytes(48, "big") for a, c in pts)
    emu.run_program(b1, {0: (bytearray(raw1), 96), 5: (st, 4)}, n)
    assert list(struct.unpack("<%di" % n, st)) == exp
    # the per-item kernel's source (csrc/g2_kernels.cuh g1_validate_one), host build, same cases + a coordinate >= p
    import ctypes
    raw1 += (pts[0][0] + O.P).to_bytes(48, "big") + pts[0][1].to_bytes(48, "big")
    kst = (ctypes.c_int32 * (n + 1))()
    emu.lib().emu_validate(0, raw1, kst, ctypes.c_size_t(n + 1))
    assert list(kst) == exp + [0]
    # ---- G2
    b2 = vmcompile.compile_program("g2_validate")
    pts, exp = [], []
    for k in (1, 3, 0x123456789):
        pts.append(O.pt_to_affine(O.G2, O.pt_multiply_unsafe(O.G2, O.G2_BASE, k))); exp.append(0)
    xx = (1, 1)
    found = 0
    while found < 2:
        xx = (xx[0] + 1, xx[1])
        y = O.fp2_sqrt(O.fp2_add(O.fp2_pow(xx, 3), O.B2))
        if y is None:
            continue
        p = (xx, y, O.FP2_ONE)
        if O.g2_is_torsion_free(p):
            continue
        found += 1
        pts.append((xx, y)); exp.append(curves.ST_NOT_IN_SUBGROUP)
        s = O.pt_multiply_unsafe(O.G2, p, r)
        if not O.pt_is_zero(O.G2, s):
            pts.append(O.pt_to_affine(O.G2, s)); exp.append(curves.ST_NOT_IN_SUBGROUP)
    pts.append(((5, 6), (7, 8))); exp.append(curves.ST_NOT_ON_CURVE)
    n = len(pts)
    st = bytearray(4 * n)
    raw = b"".join(b"".join(v.to_bytes(48, "big") for v in (a[0], a[1], c[0], c[1])) for a, c in pts)
    emu.run_program(b2, {0: (bytearray(raw), 192), 5: (st, 4)}, n)
    assert list(struct.unpack("<%di" % n, st)) == exp
    kst = (ctypes.c_int32 * n)()
    emu.lib().emu_validate(1, raw, kst, ctypes.c_size_t(n))   # g2_validate_one, host build
    assert list(kst) == exp


def test_from_uncompressed_programs():
    """index.ts:315-325 / 565-579 on the emulator: coordinates as they are (reduced mod p), on-curve and subgroup checks."""
    g1u = open(os.path.join(GOLDEN, "zkcrypto_g1_uncompressed.dat"), "rb").read()
    g2u = open(os.path.join(GOLDEN, "zkcrypto_g2_uncompressed.dat"), "rb").read()
    p5 = g1u[96 * 5: 96 * 6]
    x, y = int.from_bytes(p5[:48], "big"), int.from_bytes(p5[48:], "big")
    items = [p5, g1u[96 * 999: 96 * 1000], p5[:95] + bytes([p5[95] ^ 1]), (x + O.P).to_bytes(48, "big") + (y + 2 * O.P).to_bytes(48, "big")]
    n = len(items)
    b = vmcompile.compile_program("g1_from_uncompressed")
    out, st = bytearray(96 * n), bytearray(4 * n)
    emu.run_program(b, {0: (bytearray(b"".join(items)), 96), 2: (out, 96), 5: (st, 4)}, n)
    assert list(struct.unpack("<%di" % n, st)) == [0, 0, curves.ST_NOT_ON_CURVE, 0]
    assert bytes(out[:96]) == p5 and bytes(out[96 * 3:]) == p5 and bytes(out[96:192]) == items[1]
    q = g2u[192 * 7: 192 * 8]
    items = [q, g2u[192 * 500: 192 * 501], q[:191] + bytes([q[191] ^ 1])]
    n = len(items)
    b = vmcompile.compile_program("g2_from_uncompressed")
    out, st = bytearray(192 * n), bytearray(4 * n)
    emu.run_program(b, {0: (bytearray(b"".join(items)), 192), 2: (out, 192), 5: (st, 4)}, n)
    assert list(struct.unpack("<%di" % n, st)) == [0, 0, curves.ST_NOT_ON_CURVE]
    assert bytes(out[:192]) == q[48:96] + q[:48] + q[144:192] + q[96:144]


def test_hash_to_g1_program():
    """PointG1.hashToCurve on the emulator against the reference's kilic / RFC random-oracle vectors (hashToCurve.test.ts)."""
    import json
    d = json.load(open(os.path.join(GOLDEN, "hash_to_curve.json")))
    b = vmcompile.compile_program("hash_to_g1")
    for key in ("g1_kilic_ro", "g1_rfc_ro"):
        dst = d[key]["dst"].encode("latin1")
        vecs = d[key]["vectors"]
        n = len(vecs)
        uniform = bytearray(b"".join(O.expand_message_xmd(v["msg"].encode("latin1"), dst, 128) for v in vecs))
        out = bytearray(96 * n)
        emu.run_program(b, {0: (uniform, 128), 2: (out, 96)}, n)
        for i, v in enumerate(vecs):
            assert bytes(out[96 * i: 96 * i + 96]).hex() == v["expected"], (key, i)
