"""B200-native batched BLS12-381 pairing engine behind noble-bls12-381's API surface.

`engine()` returns the ctypes binding of the C ABI (include/bls381_b200.h); `noble_bls12_381_b200.api`
mirrors the reference's TypeScript exports (pairing, verify, verifyBatch, aggregate*, sign, ...).
"""
from ._lib import Engine, EngineError, engine  # noqa: F401
