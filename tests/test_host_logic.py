"""CPU tests of small host-side helpers (no GPU, no CUDA calls)."""
import ctypes

import numpy as np

from noble_bls12_381_b200 import dist as bdist
from noble_bls12_381_b200._lib import Engine


def test_pack_messages_offsets():
    """Vectorised message packing == the obvious prefix sums, for ragged, empty and zero-message inputs."""
    for msgs in ([], [b""], [b"a"], [b"", b"abc", b"", bytes(range(100)), b"x" * 7], [bytes([i % 251]) * (i % 17) for i in range(1000)]):
        packed, off = Engine._pack(msgs)
        assert isinstance(off, ctypes.Array) and len(off) == len(msgs) + 1
        want = [0]
        for m in msgs:
            want.append(want[-1] + len(m))
        assert list(off) == want
        assert packed == b"".join(msgs)
        for i, m in enumerate(msgs):
            assert packed[off[i]: off[i + 1]] == m


def test_status_level_matches_reference_semantics():
    """0 = all items fine, 1 = some item is the point at infinity (verifyBatch -> false, index.ts:812-820), 2 = some item is
    malformed (verifyBatch rejects, index.ts:799-801); malformed wins over infinity; list and ndarray inputs agree."""
    cases = [([], 0), ([0, 0, 0], 0), ([0, 1, 0], 1), ([0, 2, 0], 2), ([1, 3, 0], 2), ([5], 2), ([1, 1], 1)]
    for st, want in cases:
        assert bdist._level(st) == want
        assert bdist._level(np.array(st, dtype=np.int32)) == want


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 8, 9, 262144, 2097152 + 3):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
