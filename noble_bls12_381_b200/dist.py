"""Multi-GPU sharding of verifyBatch (BASELINE.json config 5).

One process per GPU (`torch.distributed`, NCCL on GPUs, gloo in the CPU tests).  Items are sharded by index in
contiguous blocks; every rank runs decompress + hash-to-curve + Miller loops + the in-GPU product tree on its
shard and ends with ONE 576-byte partial product in Fp12 (rank 0 also folds in the (-G1, signature) term).  The
only exchange step is an all-gather of those W x 576 bytes; each rank then multiplies the W partials and applies
one final exponentiation.  NCCL has no modular-product reduction operator, so all-gather + local product IS the
all-reduce for this monoid; Fp12 multiplication is exact, hence the verdict and the 576 result bytes are
identical for every world size and every shard order (tests/test_dist_gloo.py).
"""
from __future__ import annotations

FP12_ONE = (1).to_bytes(48, "big") + bytes(576 - 48)


def shard_range(n: int, rank: int, world: int):
    """Contiguous block [lo, hi) of item indices owned by `rank` (balanced to within one item)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class EngineBackend:
    """The device engine as the per-rank worker."""

    def __init__(self, eng):
        self.eng = eng

    def partial(self, sig96, msgs, pks48: bytes, dst: bytes):
        """-> (576-byte product of Miller loops of this shard [and of (-G1, sig) when sig96 is given], statuses)"""
        return self.eng.verify_batch_partial(sig96, msgs, pks48, dst)

    def combine(self, partials, with_final_exp=True) -> bytes:
        return self.eng.fp12_product(b"".join(partials), len(partials), with_final_exp)


def _level(statuses):
    """Severity of a shard's per-item status words: 0 = all OK, 1 = some item is the point at infinity (verdict false,
    index.ts:812-820), 2 = some item is malformed (the reference rejects, index.ts:799-801).  Vectorised: a shard holds
    hundreds of thousands of items."""
    import numpy as np
    a = np.asarray(statuses, dtype=np.int64)
    if a.size == 0:
        return 0
    if np.any((a != 0) & (a != 1)):
        return 2
    return 1 if np.any(a == 1) else 0


def verify_batch_sharded(backend, sig96: bytes, msgs, pks48_list, dst: bytes, dist=None, device=None):
    """verifyBatch (index.ts:792-821) over the ranks of `dist` (None = single process).
    Returns (verdict in {1, 0, -1}, 576-byte exponentiated product) with the semantics of bls381_verify_batch."""
    on = dist is not None and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if on else 0
    world = dist.get_world_size() if on else 1
    lo, hi = shard_range(len(msgs), rank, world)
    sig = sig96 if rank == 0 else None
    if hi > lo or sig is not None:
        partial, st = backend.partial(sig, msgs[lo:hi], b"".join(pks48_list[lo:hi]), dst)
        lvl = _level(st)
    else:
        partial, lvl = FP12_ONE, 0
    partials = [partial]
    if on:
        import torch
        t = torch.frombuffer(bytearray(partial) + bytearray([lvl, 0, 0, 0]), dtype=torch.uint8)
        if device is not None:
            t = t.to(device)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)  # the single exchange step: W x 580 bytes
        raw = [bytes(o.cpu().numpy().tobytes()) for o in outs]
        partials = [r[:576] for r in raw]
        lvl = max(r[576] for r in raw)
    result = backend.combine(partials, True)
    if lvl == 2:
        return -1, result
    if lvl == 1:
        return 0, result
    return (1 if result == FP12_ONE else 0), result
