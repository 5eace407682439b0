"""Synthetic workload generation for the bench / tests: (i*G1, i*G2) affine pairs in wire format.

Host-side Python-int group law (homogeneous projective add, then one batched inversion); independent of
the oracle.  i = 1..n reproduces exactly the inputs of the reference's 1000 kilic fixtures
(test/deterministic.test.ts:34-46) as a prefix of any batch.
"""
from __future__ import annotations

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
GX = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
GY = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G2X = (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
       0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E)
G2Y = (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
       0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE)


class _F1:
    one = 1
    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def mul(a, b): return a * b % P
    @staticmethod
    def inv(a): return pow(a, -1, P)


class _F2:
    one = (1, 0)
    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
    @staticmethod
    def mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
    @staticmethod
    def inv(a):
        f = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
        return (a[0] * f % P, (-a[1]) * f % P)


def _madd(F, p, q):
    """projective p + affine-like q=(x,y,1), p != +-q, neither infinity (add-1998-cmo-2)."""
    X1, Y1, Z1 = p
    X2, Y2 = q
    U = F.sub(F.mul(Y2, Z1), Y1)
    V = F.sub(F.mul(X2, Z1), X1)
    VV = F.mul(V, V)
    VVV = F.mul(VV, V)
    R = F.mul(VV, X1)
    A = F.sub(F.sub(F.mul(F.mul(U, U), Z1), VVV), F.add(R, R))
    return (F.mul(V, A), F.sub(F.mul(U, F.sub(R, A)), F.mul(VVV, Y1)), F.mul(VVV, Z1))


def _dbl(F, p, three):
    X, Y, Z = p
    W = F.mul(F.mul(X, X), three)
    S = F.mul(Y, Z)
    B = F.mul(F.mul(X, Y), S)
    B4 = F.add(F.add(B, B), F.add(B, B))
    H = F.sub(F.mul(W, W), F.add(B4, B4))
    SS = F.mul(S, S)
    X3 = F.mul(F.add(H, H), S)
    YY = F.mul(Y, Y)
    YY8 = F.add(F.add(F.add(YY, YY), F.add(YY, YY)), F.add(F.add(YY, YY), F.add(YY, YY)))
    Y3 = F.sub(F.mul(W, F.sub(B4, H)), F.mul(YY8, SS))
    S3 = F.mul(SS, S)
    Z3 = F.add(F.add(F.add(S3, S3), F.add(S3, S3)), F.add(F.add(S3, S3), F.add(S3, S3)))
    return (X3, Y3, Z3)


def _multiples(F, g, n, three):
    """[1*g, 2*g, ..., n*g] as affine pairs."""
    pts = [(g[0], g[1], F.one)]
    if n >= 2:
        pts.append(_dbl(F, pts[0], three))
    for _ in range(2, n):
        pts.append(_madd(F, pts[-1], g))
    # batched inversion of the z's
    pref = [F.one]
    for p in pts:
        pref.append(F.mul(pref[-1], p[2]))
    inv = F.inv(pref[-1])
    out = [None] * len(pts)
    for i in range(len(pts) - 1, -1, -1):
        zi = F.mul(inv, pref[i])
        inv = F.mul(inv, pts[i][2])
        out[i] = (F.mul(pts[i][0], zi), F.mul(pts[i][1], zi))
    return out[:n]


def multiples_wire(n: int):
    """(g1_bytes n x 96, g2_bytes n x 192) for (i*G1, i*G2), i = 1..n, in the C-ABI wire format."""
    g1 = _multiples(_F1, (GX, GY), n, 3)
    g2 = _multiples(_F2, (G2X, G2Y), n, (3, 0))
    b1 = b"".join(x.to_bytes(48, "big") + y.to_bytes(48, "big") for x, y in g1)
    b2 = b"".join(b"".join(c.to_bytes(48, "big") for c in (x[0], x[1], y[0], y[1])) for x, y in g2)
    return b1, b2


def g1_multiples_wire(n: int):
    """g1_bytes n x 96 for i*G1, i = 1..n (the public keys of the secret keys 1..n)."""
    g1 = _multiples(_F1, (GX, GY), n, 3)
    return b"".join(x.to_bytes(48, "big") + y.to_bytes(48, "big") for x, y in g1)


# ---- seeded random workloads ------------------------------------------------------------------
R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def random_scalars(n: int, seed: int, first: int = 0):
    """(a, b): two n x 32-byte big-endian scalar arrays from a SHA-256 counter-mode PRNG: item i (counted from `first`) gets
    SHA-256(seed || tag || i) mod r, a zero remapped to 1."""
    import hashlib
    s = seed.to_bytes(8, "big")

    def gen(tag):
        out = bytearray()
        for i in range(first, first + n):
            v = int.from_bytes(hashlib.sha256(s + tag + i.to_bytes(8, "big")).digest(), "big") % R_ORDER
            out += (v or 1).to_bytes(32, "big")
        return bytes(out)

    return gen(b"a"), gen(b"b")


def random_pairs_wire(eng, n: int, seed: int = 0xB200, prefix: int = 0):
    """(g1 n x 96 B, g2 n x 192 B): the first `prefix` pairs are (i G1, i G2), i = 1..prefix (the reference's kilic fixtures,
    test/deterministic.test.ts:34-46), the rest P_i = a_i G1, Q_i = b_i G2 with the scalars of random_scalars(seed).  The
    random multiples are computed by the ENGINE's own scalar multiplication (the checker -- oracle/c -- only ever sees the
    resulting bytes)."""
    prefix = min(prefix, n)
    g1p, g2p = multiples_wire(prefix) if prefix else (b"", b"")
    m = n - prefix
    if m == 0:
        return g1p, g2p
    a, b = random_scalars(m, seed, prefix)
    base1 = (GX.to_bytes(48, "big") + GY.to_bytes(48, "big")) * m
    base2 = b"".join(c.to_bytes(48, "big") for c in (G2X[0], G2X[1], G2Y[0], G2Y[1])) * m
    g1, f1 = eng.g1_scalar_mul_batch(base1, a, m)
    g2, f2 = eng.g2_scalar_mul_batch(base2, b, m)
    assert not any(f1) and not any(f2)
    return g1p + g1, g2p + g2


def signing_inputs(n: int, seed: int, kats=None, first: int = 0):
    """(sks n x 32 B, [messages]): optional `kats` = [(sk_hex, msg_hex, sig_hex)] lines of the reference's
    test/bls12-381-g2-test-vectors.txt as prefix, then sk_i = SHA-256(seed || "sk" || i) (the engine reduces mod r),
    msg_i = SHA-256(seed || "msg" || i) (32 bytes, distinct)."""
    import hashlib
    s = seed.to_bytes(8, "big")
    sks = bytearray()
    msgs = []
    for sk, m, _ in (kats or [])[:n]:
        sks += int(sk, 16).to_bytes(32, "big")
        msgs.append(bytes.fromhex(m))
    for i in range(first + len(msgs), first + n):
        sks += hashlib.sha256(s + b"sk" + i.to_bytes(8, "big")).digest()
        msgs.append(hashlib.sha256(s + b"msg" + i.to_bytes(8, "big")).digest())
    return bytes(sks), msgs
