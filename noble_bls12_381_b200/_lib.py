"""ctypes binding of the C ABI (include/bls381_b200.h).  There is no CPU fallback: if the CUDA library
or a CUDA device is missing, every call raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BLS381_B200_LIB") or os.path.join(_HERE, "libbls381_b200.so")  # env override: tuning builds


class EngineError(RuntimeError):
    pass


_u8p = ctypes.POINTER(ctypes.c_uint8)


def _load():
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first; there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    lib.bls381_init.argtypes = [ctypes.c_int, ctypes.c_char_p]
    lib.bls381_last_error.restype = ctypes.c_char_p
    lib.bls381_launch_count.restype = ctypes.c_uint64
    lib.bls381_last_kernel_ms.restype = ctypes.c_double
    lib.bls381_last_kernel_sm_mhz.restype = ctypes.c_double
    lib.bls381_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    lib.bls381_imad_peak_sustained.argtypes = [ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    for name in ("bls381_pairing_batch",):
        getattr(lib, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_pairing_batch_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_final_exp_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.bls381_final_exp_batch_dev.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_miller_product.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    lib.bls381_miller_product_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_vm_run_dev.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint32), ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
    lib.bls381_vm_load.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
    lib.bls381_imad_peak.argtypes = [ctypes.POINTER(ctypes.c_double)]
    lib.bls381_g1_decompress_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_g2_decompress_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_hash_to_g2_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.bls381_hash_to_g1_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.bls381_sign_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.bls381_aggregate_g1.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_aggregate_g2.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_fp12_product.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    lib.bls381_g1_validate_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.bls381_g2_validate_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.bls381_g1_scalar_mul_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_g2_scalar_mul_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_get_public_key_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.bls381_g1_encode_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    lib.bls381_g2_encode_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    lib.bls381_g1_from_uncompressed_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_g2_from_uncompressed_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_verify_batch_partial_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_fp12_product_dev.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_init_devices.argtypes = [ctypes.c_uint32, ctypes.c_char_p]
    lib.bls381_multi_transport.restype = ctypes.c_char_p
    lib.bls381_verify_batch_multi.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
    lib.bls381_verify_batch_partial.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.bls381_verify_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
    return lib


EXPORTS = [
    "bls381_init", "bls381_shutdown", "bls381_last_error", "bls381_sm_count",
    "bls381_pairing_batch", "bls381_pairing_batch_dev", "bls381_final_exp_batch", "bls381_final_exp_batch_dev",
    "bls381_miller_product", "bls381_miller_product_dev", "bls381_vm_run_dev", "bls381_vm_load",
    "bls381_launch_count", "bls381_imad_peak", "bls381_last_kernel_ms", "bls381_last_kernel_sm_mhz",
    "bls381_imad_peak_sustained", "bls381_set_option",
    "bls381_g1_decompress_batch", "bls381_g2_decompress_batch", "bls381_hash_to_g2_batch", "bls381_verify_batch",
    "bls381_sign_batch", "bls381_aggregate_g1", "bls381_aggregate_g2", "bls381_fp12_product",
    "bls381_g1_validate_batch", "bls381_g2_validate_batch", "bls381_g2_scalar_mul_batch", "bls381_g1_scalar_mul_batch", "bls381_verify_batch_partial",
    "bls381_hash_to_g1_batch", "bls381_g1_encode_batch", "bls381_g2_encode_batch", "bls381_g1_from_uncompressed_batch", "bls381_g2_from_uncompressed_batch",
    "bls381_get_public_key_batch", "bls381_verify_batch_partial_dev", "bls381_fp12_product_dev", "bls381_init_devices",
    "bls381_device_count", "bls381_multi_transport", "bls381_verify_batch_multi",
]


class Engine:
    """Thin object over the C ABI; one per process (the library holds one device context)."""

    def __init__(self, device: int = 0, program_dir: str | None = None):
        self.lib = _load()
        rc = self.lib.bls381_init(device, program_dir.encode() if program_dir else None)
        self._check(rc)

    def _check(self, rc):
        if rc != 0:
            raise EngineError(f"bls381 error {rc}: {self.lib.bls381_last_error().decode()}")

    # ---- host-buffer entry points -------------------------------------------------------------
    def pairing_batch(self, g1: bytes, g2: bytes, n: int, with_final_exp: bool = True) -> bytes:
        assert len(g1) == 96 * n and len(g2) == 192 * n
        out = ctypes.create_string_buffer(576 * n)
        self._check(self.lib.bls381_pairing_batch(g1, g2, n, int(with_final_exp), out, None))
        return out.raw

    def final_exp_batch(self, f12: bytes, n: int) -> bytes:
        assert len(f12) == 576 * n
        out = ctypes.create_string_buffer(576 * n)
        self._check(self.lib.bls381_final_exp_batch(f12, n, out))
        return out.raw

    def miller_product(self, g1: bytes, g2: bytes, n: int, with_final_exp: bool = True) -> bytes:
        assert len(g1) == 96 * n and len(g2) == 192 * n
        out = ctypes.create_string_buffer(576)
        self._check(self.lib.bls381_miller_product(g1, g2, n, int(with_final_exp), out))
        return out.raw

    # ---- ingest / verification ---------------------------------------------------------------------
    @staticmethod
    def _pack(msgs):
        """messages -> (concatenated bytes, n + 1 byte offsets as a ctypes uint64 array); vectorised: a Python loop
        over 262 144 messages costs more than the device needs to hash them"""
        import numpy as np
        n = len(msgs)
        off = np.zeros(n + 1, dtype=np.uint64)
        if n:
            np.cumsum(np.fromiter(map(len, msgs), dtype=np.uint64, count=n), out=off[1:])
        return b"".join(msgs), (ctypes.c_uint64 * (n + 1)).from_buffer(off)

    def g1_decompress_batch(self, keys48: bytes, n: int):
        out = ctypes.create_string_buffer(96 * n)
        st = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_g1_decompress_batch(keys48, n, out, st))
        return out.raw, list(st)

    def g2_decompress_batch(self, sigs96: bytes, n: int):
        out = ctypes.create_string_buffer(192 * n)
        st = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_g2_decompress_batch(sigs96, n, out, st))
        return out.raw, list(st)

    def hash_to_g2_batch(self, msgs, dst: bytes) -> bytes:
        packed, off = self._pack(msgs)
        out = ctypes.create_string_buffer(192 * len(msgs))
        self._check(self.lib.bls381_hash_to_g2_batch(packed, off, len(msgs), dst, len(dst), out))
        return out.raw

    def hash_to_g1_batch(self, msgs, dst: bytes) -> bytes:
        packed, off = self._pack(msgs)
        out = ctypes.create_string_buffer(96 * len(msgs))
        self._check(self.lib.bls381_hash_to_g1_batch(packed, off, len(msgs), dst, len(dst), out))
        return out.raw

    def verify_batch(self, sig96: bytes, msgs, pks48: bytes, dst: bytes):
        """-> (verdict in {1, 0, -1}, status list of n + 1 codes)"""
        n = len(msgs)
        packed, off = self._pack(msgs)
        v = ctypes.c_int(0)
        st = (ctypes.c_int32 * (n + 1))()
        self._check(self.lib.bls381_verify_batch(sig96, packed, off, pks48, n, dst, len(dst), ctypes.byref(v), st))
        return v.value, list(st)

    def verify_batch_partial(self, sig96, msgs, pks48: bytes, dst: bytes):
        """-> (576-byte un-exponentiated product of this shard, status list)"""
        n = len(msgs)
        packed, off = self._pack(msgs)
        np_ = n + (1 if sig96 is not None else 0)
        out = ctypes.create_string_buffer(576)
        st = (ctypes.c_int32 * max(np_, 1))()
        self._check(self.lib.bls381_verify_batch_partial(sig96, packed, off, pks48, n, dst, len(dst), out, st))
        import numpy as np
        return out.raw, np.frombuffer(st, dtype=np.int32, count=np_).copy()

    def sign_batch(self, sks32: bytes, msgs, dst: bytes) -> bytes:
        n = len(msgs)
        assert len(sks32) == 32 * n
        packed, off = self._pack(msgs)
        out = ctypes.create_string_buffer(96 * n)
        self._check(self.lib.bls381_sign_batch(sks32, packed, off, n, dst, len(dst), out))
        return out.raw

    def g1_encode_batch(self, g1: bytes, n: int, compressed: bool) -> bytes:
        out = ctypes.create_string_buffer((48 if compressed else 96) * n)
        self._check(self.lib.bls381_g1_encode_batch(g1, n, int(compressed), out))
        return out.raw

    def g2_encode_batch(self, g2: bytes, n: int, compressed: bool) -> bytes:
        out = ctypes.create_string_buffer((96 if compressed else 192) * n)
        self._check(self.lib.bls381_g2_encode_batch(g2, n, int(compressed), out))
        return out.raw

    def g1_from_uncompressed_batch(self, in96: bytes, n: int):
        out = ctypes.create_string_buffer(96 * n)
        st = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_g1_from_uncompressed_batch(in96, n, out, st))
        return out.raw, list(st)

    def g2_from_uncompressed_batch(self, in192: bytes, n: int):
        out = ctypes.create_string_buffer(192 * n)
        st = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_g2_from_uncompressed_batch(in192, n, out, st))
        return out.raw, list(st)

    def get_public_key_batch(self, sks32: bytes) -> bytes:
        """getPublicKey for n 32-byte big-endian scalars -> n x 48 B compressed public keys"""
        n = len(sks32) // 32
        assert len(sks32) == 32 * n
        out = ctypes.create_string_buffer(48 * n)
        self._check(self.lib.bls381_get_public_key_batch(sks32, n, out))
        return out.raw

    def pairing_batch_checked(self, g1: bytes, g2: bytes, n: int, with_final_exp: bool = True):
        """pairing() with the reference's checks (index.ts:716-718) -> (n x 576 B, status list)"""
        assert len(g1) == 96 * n and len(g2) == 192 * n
        out = ctypes.create_string_buffer(576 * n)
        st = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_pairing_batch(g1, g2, n, int(with_final_exp), out, st))
        return out.raw, list(st)

    def init_devices(self, mask: int) -> int:
        """one context per CUDA device in `mask` for verify_batch_multi; returns the number of contexts"""
        self._check(self.lib.bls381_init_devices(mask, None))
        return int(self.lib.bls381_device_count())

    def multi_transport(self) -> str:
        return self.lib.bls381_multi_transport().decode()

    def verify_batch_multi(self, sig96: bytes, msgs, pks48: bytes, dst: bytes):
        """verifyBatch sharded over every initialised device -> (verdict, status list of n + 1 codes)"""
        n = len(msgs)
        packed, off = self._pack(msgs)
        v = ctypes.c_int(0)
        st = (ctypes.c_int32 * (n + 1))()
        self._check(self.lib.bls381_verify_batch_multi(sig96, packed, off, pks48, n, dst, len(dst), ctypes.byref(v), st))
        return v.value, list(st)

    def aggregate_g1(self, pks48: bytes, n: int):
        out = ctypes.create_string_buffer(48)
        st = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_aggregate_g1(pks48, n, out, st))
        return out.raw, list(st)

    def aggregate_g2(self, sigs96: bytes, n: int):
        out = ctypes.create_string_buffer(96)
        st = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_aggregate_g2(sigs96, n, out, st))
        return out.raw, list(st)

    def g1_validate_batch(self, g1: bytes, n: int):
        st = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_g1_validate_batch(g1, n, st))
        return list(st)

    def g2_validate_batch(self, g2: bytes, n: int):
        st = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_g2_validate_batch(g2, n, st))
        return list(st)

    def g2_scalar_mul_batch(self, g2: bytes, scalars32: bytes, n: int):
        out = ctypes.create_string_buffer(192 * n)
        fl = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_g2_scalar_mul_batch(g2, scalars32, n, out, fl))
        return out.raw, list(fl)

    def g1_scalar_mul_batch(self, g1: bytes, scalars32: bytes, n: int):
        out = ctypes.create_string_buffer(96 * n)
        fl = (ctypes.c_int32 * n)()
        self._check(self.lib.bls381_g1_scalar_mul_batch(g1, scalars32, n, out, fl))
        return out.raw, list(fl)

    def fp12_product(self, f12: bytes, n: int, with_final_exp: bool = False) -> bytes:
        out = ctypes.create_string_buffer(576)
        self._check(self.lib.bls381_fp12_product(f12, n, int(with_final_exp), out))
        return out.raw

    # ---- device-pointer entry points (ints = CUDA device addresses, e.g. torch tensor.data_ptr()) ----
    def pairing_batch_dev(self, d_g1: int, d_g2: int, n: int, with_final_exp: bool, d_out: int, stream: int = 0):
        self._check(self.lib.bls381_pairing_batch_dev(d_g1, d_g2, n, int(with_final_exp), d_out, stream))

    def final_exp_batch_dev(self, d_in: int, n: int, d_out: int, stream: int = 0):
        self._check(self.lib.bls381_final_exp_batch_dev(d_in, n, d_out, stream))

    def miller_product_dev(self, d_g1: int, d_g2: int, n: int, with_final_exp: bool, d_out: int, stream: int = 0):
        self._check(self.lib.bls381_miller_product_dev(d_g1, d_g2, n, int(with_final_exp), d_out, stream))

    def vm_load(self, name: str, image: bytes):
        self._check(self.lib.bls381_vm_load(name.encode(), image, len(image)))

    def vm_run_dev(self, program: str, bufs, strides, n_items: int, stream: int = 0):
        nb = len(bufs)
        arr = (ctypes.c_void_p * nb)(*bufs)
        st = (ctypes.c_uint32 * nb)(*strides)
        self._check(self.lib.bls381_vm_run_dev(program.encode(), arr, st, nb, n_items, stream))

    # ---- measurement aids ----------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.bls381_launch_count())

    def last_kernel_ms(self) -> float:
        return float(self.lib.bls381_last_kernel_ms())

    def set_option(self, name: str, value: int) -> None:
        """engine tuning knob (see include/bls381_b200.h); never changes results"""
        self._check(self.lib.bls381_set_option(name.encode(), int(value)))

    def last_kernel_sm_mhz(self) -> float:
        """effective SM clock during the last tower-VM launch (clock64 / globaltimer of CTA 0)"""
        return float(self.lib.bls381_last_kernel_sm_mhz())

    def imad_peak_sustained(self, seconds: float = 1.0):
        """(multiply-adds/s, SM MHz) of the IMAD.WIDE microbenchmark run back to back for `seconds`"""
        v, m = ctypes.c_double(0), ctypes.c_double(0)
        self._check(self.lib.bls381_imad_peak_sustained(seconds, ctypes.byref(v), ctypes.byref(m)))
        return v.value, m.value

    def sm_count(self) -> int:
        return int(self.lib.bls381_sm_count())

    def imad_peak(self) -> float:
        v = ctypes.c_double(0)
        self._check(self.lib.bls381_imad_peak(ctypes.byref(v)))
        return v.value


_ENGINE = None


def engine(device: int = 0) -> Engine:
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine(device)
    return _ENGINE
