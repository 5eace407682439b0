"""The C-ABI shared library loads on a box without a GPU and exports every symbol include/*.h declares;
without a CUDA device it fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "noble_bls12_381_b200", "libbls381_b200.so")


def _declared():
    h = open(os.path.join(ROOT, "include", "bls381_b200.h")).read()
    return sorted(set(re.findall(r"\b(bls381_[a-z0-9_]+)\s*\(", h)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(LIB)


def test_all_declared_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n


def test_binding_lists_match_header():
    from noble_bls12_381_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib.bls381_last_error.restype = ctypes.c_char_p
    rc = lib.bls381_init(0, None)
    assert rc != 0
    assert b"no CUDA device" in lib.bls381_last_error() or rc == -3
    out = ctypes.create_string_buffer(576)
    assert lib.bls381_pairing_batch(b"\0" * 96, b"\0" * 192, ctypes.c_size_t(1), 1, out, None) != 0
