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
    return items


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
    return items


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
