"""Multi-process CPU test of the verifyBatch sharding logic (noble_bls12_381_b200/dist.py): world_size 2 and 3
over the gloo backend.  The per-rank worker is an oracle-backed stand-in for the device engine (the test checks
partitioning, the single all-gather exchange and the combine step, not the kernels); results must be identical to
the single-process result byte for byte."""
import os
import socket

import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class OracleBackend:
    def __init__(self):
        from oracle import noble_oracle as O
        self.O = O

    def partial(self, sig96, msgs, pks48: bytes, dst: bytes):
        O = self.O
        f = O.FP12_ONE
        st = []
        for i, m in enumerate(msgs):
            pk = O.g1_from_hex(pks48[48 * i : 48 * i + 48])
            if O.pt_is_zero(O.G1, pk):
                st.append(1)
                continue
            st.append(0)
            f = O.fp12_mul(f, O.g1_miller_loop(pk, O.g2_hash_to_curve(m, dst)))
        if sig96 is not None:
            s = O.g2_from_signature(sig96)
            st.append(0)
            f = O.fp12_mul(f, O.g1_miller_loop(O.pt_negate(O.G1, O.G1_BASE), s))
        return O.fp12_to_bytes(f), st

    def combine(self, partials, with_final_exp=True):
        O = self.O
        f = O.FP12_ONE
        for p in partials:
            f = O.fp12_mul(f, O.fp12_from_bytes(p))
        if with_final_exp:
            f = O.fp12_final_exponentiate(f)
        return O.fp12_to_bytes(f)


def _batch(n):
    from oracle import noble_oracle as O
    vec = [l.split(":") for l in open(os.path.join(GOLDEN, "sign_g2_vectors.txt")).read().strip().split("\n")]
    vec = [v for v in vec if v[1]][:n]
    msgs = [bytes.fromhex(v[1]) for v in vec]
    sigs = [bytes.fromhex(v[2]) for v in vec]
    pks = [O.get_public_key(v[0].rjust(64, "0")) for v in vec]
    return msgs, O.aggregate_signatures(sigs), pks


def _worker(rank, world, port, n, tamper, q):
    import torch.distributed as dist
    from noble_bls12_381_b200 import dist as bdist
    from oracle import noble_oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    msgs, agg, pks = _batch(n)
    if tamper:
        msgs[1] = msgs[1] + b"!"
    v, res = bdist.verify_batch_sharded(OracleBackend(), agg, msgs, pks, O.DEFAULT_DST, dist=dist)
    q.put((rank, v, res))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_verify_batch_matches_single_process(world):
    import torch.multiprocessing as mp
    from noble_bls12_381_b200 import dist as bdist
    from oracle import noble_oracle as O
    n = 5
    msgs, agg, pks = _batch(n)
    v1, res1 = bdist.verify_batch_sharded(OracleBackend(), agg, msgs, pks, O.DEFAULT_DST)
    assert v1 == 1 and res1 == bdist.FP12_ONE
    for tamper in (False, True):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n, tamper, q)) for r in range(world)]
        for p in procs:
            p.start()
        out = [q.get(timeout=300) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
        assert all(o[1] == (0 if tamper else 1) for o in out)
        assert len({o[2] for o in out}) == 1  # every rank holds the same 576 bytes
        if not tamper:
            assert out[0][2] == res1


def test_shard_ranges_cover_everything():
    from noble_bls12_381_b200.dist import shard_range
    for n in (0, 1, 5, 8, 1000, 2097152):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
