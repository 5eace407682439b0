"""ctypes wrapper of the plain-C oracle port (oracle/c/bls381_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbls381_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "c", f) for f in ("bls381_oracle.c", "bls381_curves.c")]
        if not os.path.exists(_SO) or any(os.path.getmtime(src) > os.path.getmtime(_SO) for src in srcs):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_init()
        _lib.oracle_curves_init()
    return _lib


def pairing_batch(g1: bytes, g2: bytes, n: int, with_fe: bool = True, threads: int = 0) -> bytes:
    out = ctypes.create_string_buffer(576 * n)
    lib().oracle_pairing_batch(g1, g2, ctypes.c_size_t(n), int(with_fe), out, threads or (os.cpu_count() or 1))
    return out.raw


def final_exp(f12: bytes) -> bytes:
    out = ctypes.create_string_buffer(576)
    lib().oracle_final_exp(f12, out)
    return out.raw


def fp12_mul(a: bytes, b: bytes) -> bytes:
    out = ctypes.create_string_buffer(576)
    lib().oracle_fp12_mul(a, b, out)
    return out.raw


def _pack(msgs):
    off = [0]
    for m in msgs:
        off.append(off[-1] + len(m))
    return b"".join(msgs), (ctypes.c_uint64 * len(off))(*off)


def _threads(threads):
    return threads or (os.cpu_count() or 1)


def sign_batch(sks32: bytes, msgs, dst: bytes, threads: int = 0) -> bytes:
    """sign(msg_i, sk_i) (index.ts:746-752) for 32-byte big-endian keys: n x 96 B compressed signatures."""
    n = len(msgs)
    data, off = _pack(msgs)
    out = ctypes.create_string_buffer(96 * n)
    bad = lib().oracle_sign_batch(sks32, data, off, ctypes.c_size_t(n), dst, ctypes.c_size_t(len(dst)), out, _threads(threads))
    if bad:
        raise ValueError(f"{bad} invalid private keys / hash-to-curve failures")
    return out.raw


def hash_to_g2_batch(msgs, dst: bytes, threads: int = 0) -> bytes:
    """PointG2.hashToCurve(msg_i).toAffine() (index.ts:481-490): n x 192 B (x.c0, x.c1, y.c0, y.c1)."""
    n = len(msgs)
    data, off = _pack(msgs)
    out = ctypes.create_string_buffer(192 * n)
    bad = lib().oracle_hash_to_g2_batch(data, off, ctypes.c_size_t(n), dst, ctypes.c_size_t(len(dst)), out, _threads(threads))
    if bad:
        raise ValueError(f"{bad} hash-to-curve failures")
    return out.raw


def get_public_key_batch(sks32: bytes, threads: int = 0) -> bytes:
    """getPublicKey(sk_i) (index.ts:738-740): n x 48 B compressed public keys."""
    n = len(sks32) // 32
    out = ctypes.create_string_buffer(48 * n)
    bad = lib().oracle_get_public_key_batch(sks32, ctypes.c_size_t(n), out, _threads(threads))
    if bad:
        raise ValueError(f"{bad} invalid private keys")
    return out.raw


def scalar_mul_bases_batch(a32: bytes, b32: bytes, threads: int = 0):
    """(a_i * G1, b_i * G2) as affine wire bytes (n x 96 B, n x 192 B): inputs of the random-pair configuration."""
    n = len(a32) // 32
    g1 = ctypes.create_string_buffer(96 * n)
    g2 = ctypes.create_string_buffer(192 * n)
    bad = lib().oracle_scalar_mul_bases_batch(a32, b32, ctypes.c_size_t(n), g1, g2, _threads(threads))
    if bad:
        raise ValueError(f"{bad} zero scalars")
    return g1.raw, g2.raw
