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
        src = os.path.join(_HERE, "c", "bls381_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_init()
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
