"""CPU oracle: a restatement of paulmillr/noble-bls12-381 (v1.4.0, e02d93f) on Python ints.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
it, and only as the checker.  The product path (``noble_bls12_381_b200``) never routes through this
module and fails loudly when its CUDA library is missing.

Parity: PINNED.  ``tests/test_oracle_golden.py`` checks this module against the reference's own
fixtures (test/go_pairing_vectors/pairing.json 1000/1000, test/pairing.test.ts:46-96 KATs,
test/bls12-381-g2-test-vectors.txt 559 sign KATs, test/zkcrypto/converted.json encodings,
test/hashToCurve.test.ts RFC vectors) which are committed under ``tests/golden/``.

Every function cites the reference file:line (relative to /root/reference) it restates.  The
algorithms (formulas, evaluation order, exponent of the final exponentiation, error order) follow the
reference literally; only the data representation differs (ints / tuples instead of classes):

    Fp   = int in [0, P)
    Fp2  = (c0, c1)                       math.ts:403
    Fp6  = (c0, c1, c2)   of Fp2          math.ts:554
    Fp12 = (c0, c1)       of Fp6          math.ts:705
    point = (x, y, z) homogeneous projective, x = X/Z (math.ts:892-901)
"""
from __future__ import annotations

import hashlib

# ----------------------------------------------------------------------------- constants (math.ts:10-51)
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
H1 = 0x396C8C005555E1568C00AAAB0000AAAB
GX = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
GY = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G2X = (
    0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
    0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
)
G2Y = (
    0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
    0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE,
)
B1 = 4  # math.ts:23
B2 = (4, 4)  # math.ts:46
X_PARAM = 0xD201000000010000  # math.ts:48  (|x|; the curve parameter is -x)
BLS_X_LEN = X_PARAM.bit_length()  # math.ts:53 == 64
P2_ORDER = P * P  # Fp2.ORDER = CURVE.P2 + 1 ... (math.ts:29-32 defines P2 = p^2 - 1; ORDER used as p^2-1)
FP2_ORDER = P**2 - 1  # math.ts:404 (Fp2.ORDER = CURVE.P2 = p^2 - 1)

POW_2_381 = 1 << 381  # index.ts:26
POW_2_382 = 1 << 382  # index.ts:27
POW_2_383 = 1 << 383  # index.ts:28

DEFAULT_DST = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"  # index.ts:64


class OracleError(Exception):
    """Mirrors `throw new Error(msg)` in the reference; .args[0] is the reference's message."""


# ----------------------------------------------------------------------------- Fp (math.ts:215-291)
def fp_inv(a: int) -> int:
    # math.ts:134-156 extended Euclid; result is the canonical residue, so pow(-1) is identical.
    a %= P
    if a == 0:
        raise OracleError("invert: expected positive integers, got n=0 mod=" + str(P))
    return pow(a, -1, P)


def fp_sqrt(a: int):
    # math.ts:260-264  root = a^((p+1)/4); undefined if root^2 != a
    root = pow(a, (P + 1) // 4, P)
    if root * root % P != a % P:
        return None
    return root


# ----------------------------------------------------------------------------- Fp2 (math.ts:403-550)
FP2_ZERO = (0, 0)
FP2_ONE = (1, 0)


def fp2_add(a, b):  # math.ts:440-444
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fp2_sub(a, b):  # math.ts:445-449
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fp2_neg(a):  # math.ts:431-434
    return ((-a[0]) % P, (-a[1]) % P)


def fp2_mul(a, b):  # math.ts:451-462 (Karatsuba)
    t1 = a[0] * b[0]
    t2 = a[1] * b[1]
    return ((t1 - t2) % P, ((a[0] + a[1]) * (b[0] + b[1]) - (t1 + t2)) % P)


def fp2_mul_scalar(a, k: int):  # math.ts:453-455 multiply(bigint)
    return (a[0] * k % P, a[1] * k % P)


def fp2_sqr(a):  # math.ts:477-484
    c0, c1 = a
    return ((c0 + c1) * (c0 - c1) % P, (c0 + c0) * c1 % P)


def fp2_mul_by_nonresidue(a):  # math.ts:471-475  * (1 + u)
    return ((a[0] - a[1]) % P, (a[0] + a[1]) % P)


def fp2_inv(a):  # math.ts:522-526
    factor = fp_inv((a[0] * a[0] + a[1] * a[1]) % P)
    return (factor * a[0] % P, factor * (-a[1]) % P)


def fp2_div_scalar(a, k: int):  # math.ts:466-469 div(bigint)
    return fp2_mul_scalar(a, fp_inv(k))


def fp2_div(a, b):  # math.ts:466-469
    return fp2_mul(a, fp2_inv(b))


def fp2_mul_by_b(a):  # math.ts:532-539  * 4(1+u)
    t0 = a[0] * 4
    t1 = a[1] * 4
    return ((t0 - t1) % P, (t0 + t1) % P)


def fp2_pow(a, n: int):  # math.ts:388-400,463-465 powMod_FQP
    if n == 0:
        return FP2_ONE
    if n == 1:
        return a
    p = FP2_ONE
    d = a
    while n > 0:
        if n & 1:
            p = fp2_mul(p, d)
        n >>= 1
        d = fp2_sqr(d)
    return p


def fp2_is_zero(a):
    return a[0] == 0 and a[1] == 0


def fp2_frobenius(a, power: int):  # math.ts:529-531
    return (a[0], a[1] * FP2_FROBENIUS_COEFFICIENTS[power % 2] % P)


# math.ts:1415-1424
_rv1 = 0x6AF0E0437FF400B6831E36D6BD17FFE48395DABC2D3435E77F76E17009241C5EE67992F72EC05F4C81084FBEDE3CC09
_ev1 = 0x699BE3B8C6870965E5BF892AD5D2CC7B0E85A117402DFD83B7F4A947E02D978498255A2AAEC0AC627B5AFBDF1BF1C90
_ev2 = 0x8157CD83046453F5DD0972B6E3949E4288020B5B8A9CC99CA07E27089A2CE2436D965026ADAD3EF7BABA37F2183E9B5
_ev3 = 0xAB1C2FFDD6C253CA155231EB3E71BA044FD562F6F72BC5BAD5EC46A0B7A3B0247CF08CE6C6317F40EDBC653A72DEE17
_ev4 = 0xAA404866706722864480885D68AD0CCAC1967C7544B447873CC37E0181271E006DF72162A3D3E0287BF597FBF7F8FC1

FP2_FROBENIUS_COEFFICIENTS = (1, P - 1)  # math.ts:1428-1431


def _fp2t(pair):
    return (pair[0] % P, pair[1] % P)


# math.ts:1436-1445
FP2_ROOTS_OF_UNITY = tuple(
    _fp2t(t)
    for t in (
        (1, 0),
        (_rv1, -_rv1),
        (0, 1),
        (_rv1, _rv1),
        (-1, 0),
        (-_rv1, _rv1),
        (0, -1),
        (-_rv1, -_rv1),
    )
)
# math.ts:1447-1452
FP2_ETAS = tuple(_fp2t(t) for t in ((_ev1, _ev2), (-_ev2, _ev1), (_ev3, _ev4), (-_ev4, _ev3)))


def fp2_sqrt(a):  # math.ts:486-507
    candidate = fp2_pow(a, (FP2_ORDER + 8) // 16)
    check = fp2_div(fp2_sqr(candidate), a)
    R = FP2_ROOTS_OF_UNITY
    divisor = None
    for r in (R[0], R[2], R[4], R[6]):
        if r == check:
            divisor = r
            break
    if divisor is None:
        return None
    index = R.index(divisor)
    root = R[index // 2]
    x1 = fp2_div(candidate, root)
    x2 = fp2_neg(x1)
    re1, im1 = x1
    re2, im2 = x2
    if im1 > im2 or (im1 == im2 and re1 > re2):
        return x1
    return x2


# ----------------------------------------------------------------------------- Fp6 (math.ts:554-700)
FP6_ZERO = (FP2_ZERO, FP2_ZERO, FP2_ZERO)
FP6_ONE = (FP2_ONE, FP2_ZERO, FP2_ZERO)


def fp6_add(a, b):  # math.ts:590-594
    return (fp2_add(a[0], b[0]), fp2_add(a[1], b[1]), fp2_add(a[2], b[2]))


def fp6_sub(a, b):  # math.ts:595-599
    return (fp2_sub(a[0], b[0]), fp2_sub(a[1], b[1]), fp2_sub(a[2], b[2]))


def fp6_neg(a):  # math.ts:578-581
    return (fp2_neg(a[0]), fp2_neg(a[1]), fp2_neg(a[2]))


def fp6_mul(a, b):  # math.ts:601-618
    c0, c1, c2 = a
    r0, r1, r2 = b
    t0 = fp2_mul(c0, r0)
    t1 = fp2_mul(c1, r1)
    t2 = fp2_mul(c2, r2)
    return (
        fp2_add(t0, fp2_mul_by_nonresidue(fp2_sub(fp2_mul(fp2_add(c1, c2), fp2_add(r1, r2)), fp2_add(t1, t2)))),
        fp2_add(fp2_sub(fp2_mul(fp2_add(c0, c1), fp2_add(r0, r1)), fp2_add(t0, t1)), fp2_mul_by_nonresidue(t2)),
        fp2_add(t1, fp2_sub(fp2_mul(fp2_add(c0, c2), fp2_add(r0, r2)), fp2_add(t0, t2))),
    )


def fp6_mul_by_nonresidue(a):  # math.ts:627-629  * v
    return (fp2_mul_by_nonresidue(a[2]), a[0], a[1])


def fp6_mul_by_1(a, b1):  # math.ts:631-637
    return (fp2_mul_by_nonresidue(fp2_mul(a[2], b1)), fp2_mul(a[0], b1), fp2_mul(a[1], b1))


def fp6_mul_by_01(a, b0, b1):  # math.ts:639-651
    c0, c1, c2 = a
    t0 = fp2_mul(c0, b0)
    t1 = fp2_mul(c1, b1)
    return (
        fp2_add(fp2_mul_by_nonresidue(fp2_sub(fp2_mul(fp2_add(c1, c2), b1), t1)), t0),
        fp2_sub(fp2_sub(fp2_mul(fp2_add(b0, b1), fp2_add(c0, c1)), t0), t1),
        fp2_add(fp2_sub(fp2_mul(fp2_add(c0, c2), b0), t0), t1),
    )


def fp6_mul_by_fp2(a, k):  # math.ts:653-656
    return (fp2_mul(a[0], k), fp2_mul(a[1], k), fp2_mul(a[2], k))


def fp6_sqr(a):  # math.ts:658-670
    c0, c1, c2 = a
    t0 = fp2_sqr(c0)
    t1 = fp2_mul_scalar(fp2_mul(c0, c1), 2)
    t3 = fp2_mul_scalar(fp2_mul(c1, c2), 2)
    t4 = fp2_sqr(c2)
    return (
        fp2_add(fp2_mul_by_nonresidue(t3), t0),
        fp2_add(fp2_mul_by_nonresidue(t4), t1),
        fp2_sub(fp2_sub(fp2_add(fp2_add(t1, fp2_sqr(fp2_add(fp2_sub(c0, c1), c2))), t3), t0), t4),
    )


def fp6_inv(a):  # math.ts:672-680
    c0, c1, c2 = a
    t0 = fp2_sub(fp2_sqr(c0), fp2_mul_by_nonresidue(fp2_mul(c2, c1)))
    t1 = fp2_sub(fp2_mul_by_nonresidue(fp2_sqr(c2)), fp2_mul(c0, c1))
    t2 = fp2_sub(fp2_sqr(c1), fp2_mul(c0, c2))
    t4 = fp2_inv(
        fp2_add(fp2_mul_by_nonresidue(fp2_add(fp2_mul(c2, t1), fp2_mul(c1, t2))), fp2_mul(c0, t0))
    )
    return (fp2_mul(t4, t0), fp2_mul(t4, t1), fp2_mul(t4, t2))


def fp6_frobenius(a, power: int):  # math.ts:682-688
    return (
        fp2_frobenius(a[0], power),
        fp2_mul(fp2_frobenius(a[1], power), FP6_FROBENIUS_COEFFICIENTS_1[power % 6]),
        fp2_mul(fp2_frobenius(a[2], power), FP6_FROBENIUS_COEFFICIENTS_2[power % 6]),
    )


# ----------------------------------------------------------------------------- Fp12 (math.ts:705-885)
FP12_ZERO = (FP6_ZERO, FP6_ZERO)
FP12_ONE = (FP6_ONE, FP6_ZERO)


def fp12_from_twelve(t):  # math.ts:709-714
    t = [v % P for v in t]
    return (
        ((t[0], t[1]), (t[2], t[3]), (t[4], t[5])),
        ((t[6], t[7]), (t[8], t[9]), (t[10], t[11])),
    )


def fp12_flat(f):
    """Flat order of math.ts:709-714 / toBytes math.ts:882-884."""
    return [c for six in f for two in six for c in two]


def fp12_mul(a, b):  # math.ts:748-759
    c0, c1 = a
    r0, r1 = b
    t1 = fp6_mul(c0, r0)
    t2 = fp6_mul(c1, r1)
    return (
        fp6_add(t1, fp6_mul_by_nonresidue(t2)),
        fp6_sub(fp6_mul(fp6_add(c0, c1), fp6_add(r0, r1)), fp6_add(t1, t2)),
    )


def fp12_mul_by_014(a, o0, o1, o4):  # math.ts:768-777
    c0, c1 = a
    t0 = fp6_mul_by_01(c0, o0, o1)
    t1 = fp6_mul_by_1(c1, o4)
    return (
        fp6_add(fp6_mul_by_nonresidue(t1), t0),
        fp6_sub(fp6_sub(fp6_mul_by_01(fp6_add(c1, c0), o0, fp2_add(o1, o4)), t0), t1),
    )


def fp12_sqr(a):  # math.ts:783-791
    c0, c1 = a
    ab = fp6_mul(c0, c1)
    return (
        fp6_sub(
            fp6_sub(fp6_mul(fp6_add(fp6_mul_by_nonresidue(c1), c0), fp6_add(c0, c1)), ab),
            fp6_mul_by_nonresidue(ab),
        ),
        fp6_add(ab, ab),
    )


def fp12_inv(a):  # math.ts:793-797
    c0, c1 = a
    t = fp6_inv(fp6_sub(fp6_sqr(c0), fp6_mul_by_nonresidue(fp6_sqr(c1))))
    return (fp6_mul(c0, t), fp6_neg(fp6_mul(c1, t)))


def fp12_conjugate(a):  # math.ts:799-801
    return (a[0], fp6_neg(a[1]))


def fp12_frobenius(a, power: int):  # math.ts:804-809
    r0 = fp6_frobenius(a[0], power)
    c0, c1, c2 = fp6_frobenius(a[1], power)
    coeff = FP12_FROBENIUS_COEFFICIENTS[power % 12]
    return (r0, (fp2_mul(c0, coeff), fp2_mul(c1, coeff), fp2_mul(c2, coeff)))


def _fp4_square(a, b):  # math.ts:811-818
    a2 = fp2_sqr(a)
    b2 = fp2_sqr(b)
    return (
        fp2_add(fp2_mul_by_nonresidue(b2), a2),
        fp2_sub(fp2_sub(fp2_sqr(fp2_add(a, b)), a2), b2),
    )


def fp12_cyclotomic_square(f):  # math.ts:824-843
    (c0c0, c0c1, c0c2), (c1c0, c1c1, c1c2) = f
    t3, t4 = _fp4_square(c0c0, c1c1)
    t5, t6 = _fp4_square(c1c0, c0c2)
    t7, t8 = _fp4_square(c0c1, c1c2)
    t9 = fp2_mul_by_nonresidue(t8)
    return (
        (
            fp2_add(fp2_mul_scalar(fp2_sub(t3, c0c0), 2), t3),
            fp2_add(fp2_mul_scalar(fp2_sub(t5, c0c1), 2), t5),
            fp2_add(fp2_mul_scalar(fp2_sub(t7, c0c2), 2), t7),
        ),
        (
            fp2_add(fp2_mul_scalar(fp2_add(t9, c1c0), 2), t9),
            fp2_add(fp2_mul_scalar(fp2_add(t4, c1c1), 2), t4),
            fp2_add(fp2_mul_scalar(fp2_add(t6, c1c2), 2), t6),
        ),
    )


def fp12_cyclotomic_exp(f, n: int):  # math.ts:845-852 (all BLS_X_LEN bits, starting from ONE)
    z = FP12_ONE
    for i in range(BLS_X_LEN - 1, -1, -1):
        z = fp12_cyclotomic_square(z)
        if (n >> i) & 1:
            z = fp12_mul(z, f)
    return z


def fp12_final_exponentiate(f, taps=None):  # math.ts:856-874  == f^(3*(p^12-1)/r)
    x = X_PARAM
    t0 = fp12_mul(fp12_frobenius(f, 6), fp12_inv(f))  # :859 frobeniusMap(6).div(this)
    t1 = fp12_mul(fp12_frobenius(t0, 2), t0)  # :861
    t2 = fp12_conjugate(fp12_cyclotomic_exp(t1, x))  # :862
    t3 = fp12_mul(fp12_conjugate(fp12_cyclotomic_square(t1)), t2)  # :863
    t4 = fp12_conjugate(fp12_cyclotomic_exp(t3, x))  # :864
    t5 = fp12_conjugate(fp12_cyclotomic_exp(t4, x))  # :865
    t6 = fp12_mul(fp12_conjugate(fp12_cyclotomic_exp(t5, x)), fp12_cyclotomic_square(t2))  # :866
    t7 = fp12_conjugate(fp12_cyclotomic_exp(t6, x))  # :867
    t2_t5_pow_q2 = fp12_frobenius(fp12_mul(t2, t5), 2)  # :868
    t4_t1_pow_q3 = fp12_frobenius(fp12_mul(t4, t1), 3)  # :869
    t6_t1c_pow_q1 = fp12_frobenius(fp12_mul(t6, fp12_conjugate(t1)), 1)  # :870
    t7_t3c_t1 = fp12_mul(fp12_mul(t7, fp12_conjugate(t3)), t1)  # :871
    if taps is not None:
        taps.update(t0=t0, t1=t1, t2=t2, t3=t3, t4=t4, t5=t5, t6=t6, t7=t7)
    return fp12_mul(fp12_mul(fp12_mul(t2_t5_pow_q2, t4_t1_pow_q3), t6_t1c_pow_q1), t7_t3c_t1)  # :873


def fp12_pow(a, n: int):  # math.ts:388-400,760-762
    if n == 0:
        return FP12_ONE
    if n == 1:
        return a
    p = FP12_ONE
    d = a
    while n > 0:
        if n & 1:
            p = fp12_mul(p, d)
        n >>= 1
        d = fp12_sqr(d)
    return p


def fp12_to_bytes(f) -> bytes:  # math.ts:882-884 (+ :697-699, :547-549, :288-290)
    return b"".join(c.to_bytes(48, "big") for c in fp12_flat(f))


def fp12_from_bytes(b: bytes):  # math.ts:875-881
    if len(b) != 576:
        raise OracleError(f"fromBytes wrong length={len(b)}")
    return fp12_from_twelve([int.from_bytes(b[i * 48 : (i + 1) * 48], "big") for i in range(12)])


# ----------------------------------------------------------------------------- Frobenius tables
# math.ts:1454-1543.  Re-derived from xi = 1+u (and asserted against the reference's literals in
# tests/test_oracle_golden.py::test_frobenius_tables) rather than pasted:
#   FP6_FROBENIUS_COEFFICIENTS_1[k] = xi^((p^k-1)/3), _2[k] = xi^(2(p^k-1)/3), FP12[k] = xi^((p^k-1)/6)
def _derive_frobenius():
    xi = (1, 1)
    c1 = tuple(fp2_pow(xi, (P**k - 1) // 3) for k in range(6))
    c2 = tuple(fp2_pow(xi, 2 * (P**k - 1) // 3) for k in range(6))
    c12 = tuple(fp2_pow(xi, (P**k - 1) // 6) for k in range(12))
    return c1, c2, c12


FP6_FROBENIUS_COEFFICIENTS_1, FP6_FROBENIUS_COEFFICIENTS_2, FP12_FROBENIUS_COEFFICIENTS = _derive_frobenius()


# ----------------------------------------------------------------------------- generic field dispatch
class _FpOps:
    """Field<T> interface (math.ts:65-76) for T = Fp (ints)."""

    ZERO = 0
    ONE = 1
    MAX_BITS = P.bit_length()

    add = staticmethod(lambda a, b: (a + b) % P)
    sub = staticmethod(lambda a, b: (a - b) % P)
    mul = staticmethod(lambda a, b: a * b % P)
    muls = staticmethod(lambda a, k: a * k % P)  # multiply(bigint)
    sqr = staticmethod(lambda a: a * a % P)
    neg = staticmethod(lambda a: (-a) % P)
    inv = staticmethod(fp_inv)
    is_zero = staticmethod(lambda a: a == 0)

    @staticmethod
    def pow(a, n):
        return pow(a, n, P)


class _Fp2Ops:
    """Field<T> interface for T = Fp2."""

    ZERO = FP2_ZERO
    ONE = FP2_ONE
    MAX_BITS = FP2_ORDER.bit_length()  # math.ts:405

    add = staticmethod(fp2_add)
    sub = staticmethod(fp2_sub)
    mul = staticmethod(fp2_mul)
    muls = staticmethod(fp2_mul_scalar)
    sqr = staticmethod(fp2_sqr)
    neg = staticmethod(fp2_neg)
    inv = staticmethod(fp2_inv)
    is_zero = staticmethod(fp2_is_zero)
    pow = staticmethod(fp2_pow)


# ----------------------------------------------------------------------------- ProjectivePoint (math.ts:893-1168)
def pt_is_zero(F, p):  # math.ts:903-905
    return F.is_zero(p[2])


def pt_zero(F):  # math.ts:910-912  (1, 1, 0)
    return (F.ONE, F.ONE, F.ZERO)


def pt_equals(F, a, b):  # math.ts:915-927
    xe = F.mul(a[0], b[2]) == F.mul(b[0], a[2])
    ye = F.mul(a[1], b[2]) == F.mul(b[1], a[2])
    return xe and ye


def pt_negate(F, p):  # math.ts:929-931
    return (p[0], F.neg(p[1]), p[2])


def pt_to_affine(F, p, inv_z=None):  # math.ts:949-958
    x, y, z = p
    is0 = pt_is_zero(F, p)
    if inv_z is None:
        inv_z = x if is0 else F.inv(z)
    ax = F.mul(x, inv_z)
    ay = F.mul(y, inv_z)
    if is0:
        return (F.ZERO, F.ZERO)
    if F.is_zero(inv_z):
        raise OracleError("Invalid inverted z")
    return (ax, ay)


def pt_double(F, p):  # math.ts:974-989  dbl-1998-cmo-2
    x, y, z = p
    W = F.muls(F.mul(x, x), 3)
    S = F.mul(y, z)
    SS = F.mul(S, S)
    SSS = F.mul(SS, S)
    B = F.mul(F.mul(x, y), S)
    H = F.sub(F.mul(W, W), F.muls(B, 8))
    X3 = F.muls(F.mul(H, S), 2)
    Y3 = F.sub(F.mul(W, F.sub(F.muls(B, 4), H)), F.mul(F.muls(F.mul(y, y), 8), SS))
    Z3 = F.muls(SSS, 8)
    return (X3, Y3, Z3)


def pt_add(F, p1, p2):  # math.ts:993-1025  add-1998-cmo-2
    if pt_is_zero(F, p1):
        return p2
    if pt_is_zero(F, p2):
        return p1
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    U1 = F.mul(Y2, Z1)
    U2 = F.mul(Y1, Z2)
    V1 = F.mul(X2, Z1)
    V2 = F.mul(X1, Z2)
    if V1 == V2 and U1 == U2:
        return pt_double(F, p1)
    if V1 == V2:
        return pt_zero(F)
    U = F.sub(U1, U2)
    V = F.sub(V1, V2)
    VV = F.mul(V, V)
    VVV = F.mul(VV, V)
    V2VV = F.mul(V2, VV)
    W = F.mul(Z1, Z2)
    A = F.sub(F.sub(F.mul(F.mul(U, U), W), VVV), F.muls(V2VV, 2))
    X3 = F.mul(V, A)
    Y3 = F.sub(F.mul(U, F.sub(V2VV, A)), F.mul(VVV, U2))
    Z3 = F.mul(VVV, W)
    return (X3, Y3, Z3)


def pt_subtract(F, a, b):  # math.ts:1027-1033
    return pt_add(F, a, pt_negate(F, b))


def _validate_scalar(n: int) -> int:  # math.ts:1035-1043
    if not isinstance(n, int) or n <= 0 or n > R_ORDER:
        raise OracleError(f"Point#multiply: invalid scalar, expected positive integer < CURVE.r. Got: {n}")
    return n


def pt_multiply_unsafe(F, p, scalar: int):  # math.ts:1048-1058
    n = _validate_scalar(scalar)
    point = pt_zero(F)
    d = p
    while n > 0:
        if n & 1:
            point = pt_add(F, point, d)
        d = pt_double(F, d)
        n >>= 1
    return point


def pt_multiply(F, p, scalar: int):  # math.ts:1061-1078 constant-time ladder over bits of Fp.ORDER
    n = _validate_scalar(scalar)
    point = pt_zero(F)
    fake = pt_zero(F)
    d = p
    bits = P
    while bits > 0:
        if n & 1:
            point = pt_add(F, point, d)
        else:
            fake = pt_add(F, fake, d)
        d = pt_double(F, d)
        n >>= 1
        bits >>= 1
    return point


# ----------------------------------------------------------------------------- hash-to-curve helpers (math.ts:1179-1325)
def sgn0_fp2(x):  # math.ts:1179-1185
    x0, x1 = x
    sign_0 = x0 % 2
    zero_0 = x0 == 0
    sign_1 = x1 % 2
    return int(bool(sign_0 or (zero_0 and sign_1)))


P_MINUS_9_DIV_16 = (P**2 - 9) // 16  # math.ts:1191


def sqrt_div_fp2(u, v):  # math.ts:1195-1214
    v7 = fp2_pow(v, 7)
    uv7 = fp2_mul(u, v7)
    uv15 = fp2_mul(uv7, fp2_mul(v7, v))
    gamma = fp2_mul(fp2_pow(uv15, P_MINUS_9_DIV_16), uv7)
    success = False
    result = gamma
    for root in FP2_ROOTS_OF_UNITY[:4]:
        candidate = fp2_mul(root, gamma)
        if fp2_is_zero(fp2_sub(fp2_mul(fp2_pow(candidate, 2), v), u)) and not success:
            success = True
            result = candidate
    return success, result


def map_to_curve_simple_swu_9mod16(t):  # math.ts:1220-1267
    iso_3_a = (0, 240)
    iso_3_b = (1012, 1012)
    iso_3_z = ((-2) % P, (-1) % P)
    t = (t[0] % P, t[1] % P)
    t2 = fp2_pow(t, 2)
    iso_3_z_t2 = fp2_mul(iso_3_z, t2)
    ztzt = fp2_add(iso_3_z_t2, fp2_pow(iso_3_z_t2, 2))
    denominator = fp2_neg(fp2_mul(iso_3_a, ztzt))
    numerator = fp2_mul(iso_3_b, fp2_add(ztzt, FP2_ONE))
    if fp2_is_zero(denominator):
        denominator = fp2_mul(iso_3_z, iso_3_a)
    v = fp2_pow(denominator, 3)
    u = fp2_add(
        fp2_add(fp2_pow(numerator, 3), fp2_mul(fp2_mul(iso_3_a, numerator), fp2_pow(denominator, 2))),
        fp2_mul(iso_3_b, v),
    )
    success, sqrt_candidate_or_gamma = sqrt_div_fp2(u, v)
    y = None
    if success:
        y = sqrt_candidate_or_gamma
    sqrt_candidate_x1 = fp2_mul(sqrt_candidate_or_gamma, fp2_pow(t, 3))
    u = fp2_mul(fp2_pow(iso_3_z_t2, 3), u)
    success2 = False
    for eta in FP2_ETAS:
        eta_sqrt_candidate = fp2_mul(eta, sqrt_candidate_x1)
        temp = fp2_sub(fp2_mul(fp2_pow(eta_sqrt_candidate, 2), v), u)
        if fp2_is_zero(temp) and not success and not success2:
            y = eta_sqrt_candidate
            success2 = True
    if not success and not success2:
        raise OracleError("Hash to Curve - Optimized SWU failure")
    if success2:
        numerator = fp2_mul(numerator, iso_3_z_t2)
    if sgn0_fp2(t) != sgn0_fp2(y):
        y = fp2_neg(y)
    return fp2_div(numerator, denominator), y


# math.ts:1547-1610 3-isogeny coefficient tables (xnum, xden, ynum, yden), high-degree coefficient first
_K = 0x11560BF17BAA99BC32126FCED787C88F984F87ADF7AE0C7F9A208C6B4F20A4181472AAA9CB8D555526A9FFFFFFFFC71E
ISO3_XNUM = (
    (0x171D6541FA38CCFAED6DEA691F5FB614CB14B4E7F4E810AA22D6108F142B85757098E38D0F671C7188E2AAAAAAAA5ED1, 0),
    (
        0x11560BF17BAA99BC32126FCED787C88F984F87ADF7AE0C7F9A208C6B4F20A4181472AAA9CB8D555526A9FFFFFFFFC71E,
        0x8AB05F8BDD54CDE190937E76BC3E447CC27C3D6FBD7063FCD104635A790520C0A395554E5C6AAAA9354FFFFFFFFE38D,
    ),
    (0, 0x11560BF17BAA99BC32126FCED787C88F984F87ADF7AE0C7F9A208C6B4F20A4181472AAA9CB8D555526A9FFFFFFFFC71A),
    (
        0x5C759507E8E333EBB5B7A9A47D7ED8532C52D39FD3A042A88B58423C50AE15D5C2638E343D9C71C6238AAAAAAAA97D6,
        0x5C759507E8E333EBB5B7A9A47D7ED8532C52D39FD3A042A88B58423C50AE15D5C2638E343D9C71C6238AAAAAAAA97D6,
    ),
)
ISO3_XDEN = (
    (0, 0),
    (1, 0),
    (0xC, P - 12),
    (0, P - 72),
)
ISO3_YNUM = (
    (0x124C9AD43B6CF79BFBF7043DE3811AD0761B0F37A1E26286B0E977C69AA274524E79097A56DC4BD9E1B371C71C718B10, 0),
    (
        0x11560BF17BAA99BC32126FCED787C88F984F87ADF7AE0C7F9A208C6B4F20A4181472AAA9CB8D555526A9FFFFFFFFC71C,
        0x8AB05F8BDD54CDE190937E76BC3E447CC27C3D6FBD7063FCD104635A790520C0A395554E5C6AAAA9354FFFFFFFFE38F,
    ),
    (0, 0x5C759507E8E333EBB5B7A9A47D7ED8532C52D39FD3A042A88B58423C50AE15D5C2638E343D9C71C6238AAAAAAAA97BE),
    (
        0x1530477C7AB4113B59A4C18B076D11930F7DA5D4A07F649BF54439D87D27E500FC8C25EBF8C92F6812CFC71C71C6D706,
        0x1530477C7AB4113B59A4C18B076D11930F7DA5D4A07F649BF54439D87D27E500FC8C25EBF8C92F6812CFC71C71C6D706,
    ),
)
ISO3_YDEN = (
    (1, 0),
    (0x12, P - 18),
    (0, P - 216),
    (P - 432, P - 432),
)


def isogeny_map_g2(x, y):  # math.ts:1315-1325
    def horner(coeffs):
        acc = coeffs[0]
        for c in coeffs[1:]:
            acc = fp2_add(fp2_mul(acc, x), c)
        return acc

    x_num, x_den, y_num, y_den = (horner(c) for c in (ISO3_XNUM, ISO3_XDEN, ISO3_YNUM, ISO3_YDEN))
    return fp2_div(x_num, x_den), fp2_mul(y, fp2_div(y_num, y_den))


# ----------------------------------------------------------------------------- pairing core (math.ts:1331-1388)
def calc_pairing_precomputes(qx, qy):  # math.ts:1331-1371
    Qx, Qy, Qz = qx, qy, FP2_ONE
    Rx, Ry, Rz = Qx, Qy, Qz
    ell = []
    for i in range(BLS_X_LEN - 2, -1, -1):
        t0 = fp2_sqr(Ry)
        t1 = fp2_sqr(Rz)
        t2 = fp2_mul_by_b(fp2_mul_scalar(t1, 3))
        t3 = fp2_mul_scalar(t2, 3)
        t4 = fp2_sub(fp2_sub(fp2_sqr(fp2_add(Ry, Rz)), t1), t0)
        ell.append((fp2_sub(t2, t0), fp2_mul_scalar(fp2_sqr(Rx), 3), fp2_neg(t4)))
        Rx = fp2_div_scalar(fp2_mul(fp2_mul(fp2_sub(t0, t3), Rx), Ry), 2)
        Ry = fp2_sub(fp2_sqr(fp2_div_scalar(fp2_add(t0, t3), 2)), fp2_mul_scalar(fp2_sqr(t2), 3))
        Rz = fp2_mul(t0, t4)
        if (X_PARAM >> i) & 1:
            t0 = fp2_sub(Ry, fp2_mul(Qy, Rz))
            t1 = fp2_sub(Rx, fp2_mul(Qx, Rz))
            ell.append((fp2_sub(fp2_mul(t0, Qx), fp2_mul(t1, Qy)), fp2_neg(t0), t1))
            t2 = fp2_sqr(t1)
            t3 = fp2_mul(t2, t1)
            t4 = fp2_mul(t2, Rx)
            t5 = fp2_add(fp2_sub(t3, fp2_mul_scalar(t4, 2)), fp2_mul(fp2_sqr(t0), Rz))
            Rx = fp2_mul(t1, t5)
            Ry = fp2_sub(fp2_mul(fp2_sub(t4, t5), t0), fp2_mul(t3, Ry))
            Rz = fp2_mul(Rz, t3)
    return ell


def miller_loop(ell, g1):  # math.ts:1373-1388
    Px, Py = g1
    f12 = FP12_ONE
    j = 0
    for i in range(BLS_X_LEN - 2, -1, -1):
        E = ell[j]
        f12 = fp12_mul_by_014(f12, E[0], fp2_mul_scalar(E[1], Px), fp2_mul_scalar(E[2], Py))
        if (X_PARAM >> i) & 1:
            j += 1
            F = ell[j]
            f12 = fp12_mul_by_014(f12, F[0], fp2_mul_scalar(F[1], Px), fp2_mul_scalar(F[2], Py))
        if i != 0:
            f12 = fp12_sqr(f12)
        j += 1
    return fp12_conjugate(f12)


# psi (math.ts:1390-1403).  The reference routes through Fp12 (untwist-Frobenius-twist); algebraically
# psi(x, y) = (conj(x) * PSI_CX, conj(y) * PSI_CY) with two fixed Fp2 constants.  Both routes are
# implemented; tests assert they agree.
def _fp12_mul_by_fp2(f, k):  # math.ts:779-781
    return (fp6_mul_by_fp2(f[0], k), fp6_mul_by_fp2(f[1], k))


_UT_ROOT = (FP2_ZERO, FP2_ONE, FP2_ZERO)  # math.ts:1390
_WSQ = (_UT_ROOT, FP6_ZERO)  # math.ts:1391
_WCU = (FP6_ZERO, _UT_ROOT)  # math.ts:1392
_WSQ_INV = fp12_inv(_WSQ)  # math.ts:1393 (genInvertBatch == individual inverses)
_WCU_INV = fp12_inv(_WCU)


def psi_via_fp12(x, y):  # math.ts:1398-1403 (literal)
    x2 = fp12_mul(fp12_frobenius(_fp12_mul_by_fp2(_WSQ_INV, x), 1), _WSQ)[0][0]
    y2 = fp12_mul(fp12_frobenius(_fp12_mul_by_fp2(_WCU_INV, y), 1), _WCU)[0][0]
    return x2, y2


PSI_CX, PSI_CY = psi_via_fp12(FP2_ONE, FP2_ONE)


def psi(x, y):
    return fp2_mul(fp2_frobenius(x, 1), PSI_CX), fp2_mul(fp2_frobenius(y, 1), PSI_CY)


PSI2_C1 = 0x1A0111EA397FE699EC02408663D4DE85AA0D857D89759AD4897D29650FB85F9B409427EB4F49FFFD8BFD00000000AAAC  # math.ts:1411


def psi2(x, y):  # math.ts:1406-1408
    return fp2_mul_scalar(x, PSI2_C1), fp2_neg(y)


# ----------------------------------------------------------------------------- PointG1 (index.ts:287-462)
G1 = _FpOps
G2 = _Fp2Ops
G1_BASE = (GX, GY, 1)
G1_ZERO = (1, 1, 0)
G2_BASE = (G2X, G2Y, FP2_ONE)
G2_ZERO = (FP2_ONE, FP2_ONE, FP2_ZERO)


def g1_is_on_curve(p):  # index.ts:408-414
    x, y, z = p
    left = (pow(y, 2, P) * z - pow(x, 3, P)) % P
    right = B1 * pow(z, 3, P) % P
    return (left - right) % P == 0


_CUBIC_ROOT = 0x5F19672FDF76CE51BA69C6076A0F77EADDB3A93BE6F89688DE17D813620A00022E01FFFFFFFEFFFE  # index.ts:426-427


def g1_is_torsion_free(p):  # index.ts:444-448
    xP = pt_negate(G1, pt_multiply_unsafe(G1, p, X_PARAM))  # mulCurveX :432-434
    u2P = pt_multiply_unsafe(G1, xP, X_PARAM)  # mulCurveMinusX :436-438
    phi = (p[0] * _CUBIC_ROOT % P, p[1], p[2])  # :425-429
    return pt_equals(G1, u2P, phi)


def g1_assert_validity(p):  # index.ts:383-388
    if pt_is_zero(G1, p):
        return p
    if not g1_is_on_curve(p):
        raise OracleError("Invalid G1 point: not on curve Fp")
    if not g1_is_torsion_free(p):
        raise OracleError("Invalid G1 point: must be of prime-order subgroup")
    return p


def g1_from_hex(b: bytes):  # index.ts:298-327
    if len(b) == 48:
        v = int.from_bytes(b, "big")
        bflag = (v % POW_2_383) // POW_2_382
        if bflag == 1:
            return G1_ZERO
        x = (v % POW_2_381) % P
        right = (pow(x, 3, P) + B1) % P
        y = fp_sqrt(right)
        if y is None:
            raise OracleError("Invalid compressed G1 point")
        aflag = (v % POW_2_382) // POW_2_381
        if (y * 2) // P != aflag:
            y = (-y) % P
        point = (x, y, 1)
    elif len(b) == 96:
        if b[0] & (1 << 6):
            return G1_ZERO
        x = int.from_bytes(b[:48], "big")
        y = int.from_bytes(b[48:], "big")
        point = (x % P, y % P, 1)
    else:
        raise OracleError("Invalid point G1, expected 48/96 bytes")
    g1_assert_validity(point)
    return point


def g1_to_hex(p, compressed=False) -> bytes:  # index.ts:359-381 (bytes, i.e. toRawBytes :355-357)
    g1_assert_validity(p)
    if compressed:
        if pt_is_zero(G1, p):
            v = POW_2_383 + POW_2_382
        else:
            x, y = pt_to_affine(G1, p)
            flag = (y * 2) // P
            v = x + flag * POW_2_381 + POW_2_383
        return v.to_bytes(48, "big")
    if pt_is_zero(G1, p):
        return bytes([0x40]) + bytes(95)
    x, y = pt_to_affine(G1, p)
    return x.to_bytes(48, "big") + y.to_bytes(48, "big")


def g1_from_private_key(sk) -> tuple:  # index.ts:351-353 (wNAF result == plain scalar multiple)
    return pt_multiply_unsafe(G1, G1_BASE, normalize_priv_key(sk))


# ----------------------------------------------------------------------------- PointG2 (index.ts:466-712)
def g2_is_on_curve(p):  # index.ts:675-681
    x, y, z = p
    left = fp2_sub(fp2_mul(fp2_pow(y, 2), z), fp2_pow(x, 3))
    right = fp2_mul(B2, fp2_pow(z, 3))
    return fp2_is_zero(fp2_sub(left, right))


def g2_mul_curve_x(p):  # index.ts:651-653  [-x]P
    return pt_negate(G2, pt_multiply_unsafe(G2, p, X_PARAM))


def g2_psi(p):  # index.ts:641-643
    x, y = psi(*pt_to_affine(G2, p))
    return (x, y, FP2_ONE)


def g2_psi2(p):  # index.ts:646-648
    x, y = psi2(*pt_to_affine(G2, p))
    return (x, y, FP2_ONE)


def g2_is_torsion_free(p):  # index.ts:688-690
    return pt_equals(G2, g2_mul_curve_x(p), g2_psi(p))


def g2_assert_validity(p):  # index.ts:633-638
    if pt_is_zero(G2, p):
        return p
    if not g2_is_on_curve(p):
        raise OracleError("Invalid G2 point: not on curve Fp2")
    if not g2_is_torsion_free(p):
        raise OracleError("Invalid G2 point: must be of prime-order subgroup")
    return p


def g2_clear_cofactor(p):  # index.ts:659-672
    t1 = g2_mul_curve_x(p)
    t2 = g2_psi(p)
    t3 = pt_double(G2, p)
    t3 = g2_psi2(t3)
    t3 = pt_subtract(G2, t3, t2)
    t2 = pt_add(G2, t1, t2)
    t2 = g2_mul_curve_x(t2)
    t3 = pt_add(G2, t3, t2)
    t3 = pt_subtract(G2, t3, t1)
    return pt_subtract(G2, t3, p)


def g2_from_signature(b: bytes):  # index.ts:500-530
    half = len(b) // 2
    if len(b) % 2 or (half != 48 and half != 96):
        raise OracleError("Invalid compressed signature length, must be 96 or 192")
    z1 = int.from_bytes(b[:half], "big")
    z2 = int.from_bytes(b[half:], "big")
    bflag1 = (z1 % POW_2_383) // POW_2_382
    if bflag1 == 1:
        return G2_ZERO
    x1 = (z1 % POW_2_381) % P
    x2 = z2 % P
    x = (x2, x1)
    y2 = fp2_add(fp2_pow(x, 3), B2)
    y = fp2_sqrt(y2)
    if y is None:
        raise OracleError("Failed to find a square root")
    y0, y1 = y
    aflag1 = (z1 % POW_2_382) // POW_2_381
    is_greater = y1 > 0 and (y1 * 2) // P != aflag1
    is_zero = y1 == 0 and (y0 * 2) // P != aflag1
    if is_greater or is_zero:
        y = fp2_mul_scalar(y, -1)
    point = (x, y, FP2_ONE)
    g2_assert_validity(point)
    return point


def g2_to_signature(p) -> bytes:  # index.ts:586-598
    if pt_equals(G2, p, G2_ZERO):
        return (POW_2_383 + POW_2_382).to_bytes(48, "big") + bytes(48)
    (x0, x1), (y0, y1) = pt_to_affine(G2, p)
    tmp = y1 * 2 if y1 > 0 else y0 * 2
    aflag1 = tmp // P
    z1 = x1 + aflag1 * POW_2_381 + POW_2_383
    z2 = x0
    return z1.to_bytes(48, "big") + z2.to_bytes(48, "big")


def g2_from_hex(b: bytes):  # index.ts:532-580
    b = bytearray(b)
    m_byte = b[0] & 0xE0
    if m_byte in (0x20, 0x60, 0xE0):
        raise OracleError(f"Invalid encoding flag: {m_byte}")
    bitC = m_byte & 0x80
    bitI = m_byte & 0x40
    bitS = m_byte & 0x20
    if len(b) == 96 and bitC:
        b[0] &= 0x1F
        if bitI:
            # index.ts:549 (reduce((p, c) => p !== 0 ? c + 1 : c, 0) > 0)
            acc = 0
            for c in b:
                acc = c + 1 if acc != 0 else c
            if acc > 0:
                raise OracleError("Invalid compressed G2 point")
            return G2_ZERO
        x_1 = int.from_bytes(b[:48], "big")
        x_0 = int.from_bytes(b[48:], "big")
        x = (x_0 % P, x_1 % P)
        right = fp2_add(fp2_pow(x, 3), B2)
        y = fp2_sqrt(right)
        if y is None:
            raise OracleError("Invalid compressed G2 point")
        Y_bit = (y[0] * 2) // P if y[1] == 0 else (1 if (y[1] * 2) // P else 0)
        y = y if (bitS > 0 and Y_bit > 0) else fp2_neg(y)
        return (x, y, FP2_ONE)  # NB: reference returns here WITHOUT assertValidity (index.ts:562)
    elif len(b) == 192 and not bitC:
        if b[0] & (1 << 6):
            return G2_ZERO
        x1 = int.from_bytes(b[0:48], "big")
        x0 = int.from_bytes(b[48:96], "big")
        y1 = int.from_bytes(b[96:144], "big")
        y0 = int.from_bytes(b[144:192], "big")
        point = ((x0 % P, x1 % P), (y0 % P, y1 % P), FP2_ONE)
    else:
        raise OracleError("Invalid point G2, expected 96/192 bytes")
    g2_assert_validity(point)
    return point


def g2_to_hex(p, compressed=False) -> bytes:  # index.ts:604-631
    g2_assert_validity(p)
    if compressed:
        x_1 = 0
        x_0 = 0
        if pt_is_zero(G2, p):
            x_1 = POW_2_383 + POW_2_382
        else:
            x, y = pt_to_affine(G2, p)
            flag = (y[0] * 2) // P if y[1] == 0 else (1 if (y[1] * 2) // P else 0)
            x_1 = x[1] + flag * POW_2_381 + POW_2_383
            x_0 = x[0]
        return x_1.to_bytes(48, "big") + x_0.to_bytes(48, "big")
    if pt_equals(G2, p, G2_ZERO):
        return bytes([0x40]) + bytes(191)
    (x0, x1), (y0, y1) = pt_to_affine(G2, p)
    return b"".join(v.to_bytes(48, "big") for v in (x1, x0, y1, y0))


# ----------------------------------------------------------------------------- hashing (index.ts:207-267)
def _i2osp(value: int, length: int) -> bytes:  # index.ts:185-195
    if value < 0 or value >= 1 << (8 * length):
        raise OracleError(f"bad I2OSP call: value={value} length={length}")
    return value.to_bytes(length, "big")


def _strxor(a: bytes, b: bytes) -> bytes:  # index.ts:197-203
    return bytes(x ^ y for x, y in zip(a, b))


def expand_message_xmd(msg: bytes, dst: bytes, len_in_bytes: int, hash_name: str = "sha256") -> bytes:
    # index.ts:207-231
    H = lambda m: hashlib.new(hash_name, m).digest()
    if len(dst) > 255:
        dst = H(b"H2C-OVERSIZE-DST-" + dst)
    b_in_bytes = hashlib.new(hash_name).digest_size
    r_in_bytes = b_in_bytes * 2
    ell = -(-len_in_bytes // b_in_bytes)
    if ell > 255:
        raise OracleError("Invalid xmd length")
    dst_prime = dst + _i2osp(len(dst), 1)
    z_pad = _i2osp(0, r_in_bytes)
    l_i_b_str = _i2osp(len_in_bytes, 2)
    b = [None] * (ell + 1)
    b_0 = H(z_pad + msg + l_i_b_str + _i2osp(0, 1) + dst_prime)
    b[0] = H(b_0 + _i2osp(1, 1) + dst_prime)
    for i in range(1, ell + 1):  # index.ts:225 (`i <= ell`: one unused extra block)
        b[i] = H(_strxor(b_0, b[i - 1]) + _i2osp(i + 1, 1) + dst_prime)
    return b"".join(b)[:len_in_bytes]


def hash_to_field(msg: bytes, count: int, dst: bytes = DEFAULT_DST, p: int = P, m: int = 2, k: int = 128,
                  expand: bool = True, hash_name: str = "sha256"):
    # index.ts:240-267
    log2p = p.bit_length()
    L = -(-(log2p + k) // 8)
    len_in_bytes = count * m * L
    prb = msg
    if expand:
        prb = expand_message_xmd(msg, dst, len_in_bytes, hash_name)
    u = []
    for i in range(count):
        e = []
        for j in range(m):
            off = L * (j + i * m)
            tv = prb[off : off + L]
            e.append(int.from_bytes(tv, "big") % p)
        u.append(e)
    return u


def g2_hash_to_curve(msg: bytes, dst: bytes = DEFAULT_DST):  # index.ts:481-490
    u = hash_to_field(msg, 2, dst)
    x0, y0 = map_to_curve_simple_swu_9mod16(u[0])
    x1, y1 = map_to_curve_simple_swu_9mod16(u[1])
    x2, y2 = pt_to_affine(G2, pt_add(G2, (x0, y0, FP2_ONE), (x1, y1, FP2_ONE)))
    x3, y3 = isogeny_map_g2(x2, y2)
    return g2_clear_cofactor((x3, y3, FP2_ONE))


# ----------------------------------------------------------------------------- hash to curve G1 (math.ts:1270-1313, 1612-1790)
from .iso11_table import ISO11_XNUM, ISO11_XDEN, ISO11_YNUM, ISO11_YDEN  # noqa: E402

_SWU_G1_A = 0x144698A3B8E9433D693A02C96D4982B0EA985383EE66A8D8E8981AEFD881AC98936F8DA0E0F97F5CF428082D584C1D  # math.ts:1271-1273
_SWU_G1_B = 0x12E2908D11688030018B12E8753EEE3B2016C1F0F24F4070A0B9C14FCEF35EF55A23215A316CEAA5D1CC48E98E172BE0  # math.ts:1274-1276
_SWU_G1_Z = 11  # math.ts:1277


def map_to_curve_simple_swu_3mod4(u: int):  # math.ts:1270-1313
    A, B, Z = _SWU_G1_A, _SWU_G1_B, _SWU_G1_Z
    u %= P
    c1 = (P - 3) // 4
    c2 = fp_sqrt(pow((-Z) % P, 3, P))  # sqrt((-Z)^3): a static value, the root always exists
    tv1 = u * u % P
    tv3 = Z * tv1 % P
    x_den = (tv3 * tv3 + tv3) % P
    x_num1 = (x_den + 1) * B % P
    x_num2 = tv3 * x_num1 % P
    x_den = (-A) * x_den % P
    if x_den == 0:
        x_den = A * Z % P
    tv2 = x_den * x_den % P
    gxd = tv2 * x_den % P
    tv2 = A * tv2 % P
    gx1 = (x_num1 * x_num1 + tv2) * x_num1 % P
    tv2 = B * gxd % P
    gx1 = (gx1 + tv2) % P
    tv2 = gx1 * gxd % P
    tv4 = gxd * gxd % P * tv2 % P
    y1 = pow(tv4, c1, P) * tv2 % P
    y2 = y1 * c2 % P * tv1 % P * u % P
    if y1 * y1 % P * gxd % P == gx1:
        x_num, y_pos = x_num1, y1
    else:
        x_num, y_pos = x_num2, y2
    y = y_pos if (u % 2) == (y_pos % 2) else (-y_pos) % P  # sgn0_m_eq_1, math.ts:1187-1189
    return x_num * fp_inv(x_den) % P, y


def isogeny_map_g1(x: int, y: int):  # math.ts:1306-1313, 1327 with ISOGENY_COEFFICIENTS_G1
    def horner(coeffs):
        acc = coeffs[0]
        for c in coeffs[1:]:
            acc = (acc * x + c) % P
        return acc

    x_num, x_den, y_num, y_den = (horner(c) for c in (ISO11_XNUM, ISO11_XDEN, ISO11_YNUM, ISO11_YDEN))
    return x_num * fp_inv(x_den) % P, y * (y_num * fp_inv(y_den) % P) % P


def g1_clear_cofactor(p):  # index.ts:401-405:  [|x|]P + P
    return pt_add(G1, pt_multiply_unsafe(G1, p, X_PARAM), p)


def g1_hash_to_curve(msg: bytes, dst: bytes = DEFAULT_DST):  # index.ts:331-339
    (u0,), (u1,) = hash_to_field(msg, 2, dst, m=1)
    x0, y0 = map_to_curve_simple_swu_3mod4(u0)
    x1, y1 = map_to_curve_simple_swu_3mod4(u1)
    x2, y2 = pt_to_affine(G1, pt_add(G1, (x0, y0, 1), (x1, y1, 1)))
    x3, y3 = isogeny_map_g1(x2, y2)
    return g1_clear_cofactor((x3, y3, 1))


def g1_encode_to_curve(msg: bytes, dst: bytes = DEFAULT_DST):  # index.ts:341-350
    ((u0,),) = hash_to_field(msg, 1, dst, m=1)
    x0, y0 = map_to_curve_simple_swu_3mod4(u0)
    x1, y1 = isogeny_map_g1(x0, y0)
    return g1_clear_cofactor((x1, y1, 1))


def g2_encode_to_curve(msg: bytes, dst: bytes = DEFAULT_DST):  # index.ts:491-497
    u = hash_to_field(msg, 1, dst)
    x0, y0 = map_to_curve_simple_swu_9mod16(u[0])
    x1, y1 = isogeny_map_g2(x0, y0)
    return g2_clear_cofactor((x1, y1, FP2_ONE))


def normalize_priv_key(key) -> int:  # index.ts:269-279
    if isinstance(key, (bytes, bytearray)) and len(key) == 32:
        n = int.from_bytes(key, "big")
    elif isinstance(key, str) and len(key) == 64:
        n = int(key, 16)
    elif isinstance(key, int) and not isinstance(key, bool) and key > 0:
        n = key
    else:
        raise TypeError("Expected valid private key")
    n %= R_ORDER
    if not (0 < n < R_ORDER):
        raise OracleError("Private key must be 0 < key < CURVE.r")
    return n


# ----------------------------------------------------------------------------- BLS API (index.ts:715-821)
def g1_miller_loop(p, q):  # index.ts:395-397 + 707-711
    return miller_loop(calc_pairing_precomputes(*pt_to_affine(G2, q)), pt_to_affine(G1, p))


def pairing(p, q, with_final_exponent: bool = True):  # index.ts:715-722
    if pt_is_zero(G1, p) or pt_is_zero(G2, q):
        raise OracleError("No pairings at point of Infinity")
    g1_assert_validity(p)
    g2_assert_validity(q)
    looped = g1_miller_loop(p, q)
    return fp12_final_exponentiate(looped) if with_final_exponent else looped


def _norm_p1(point):  # index.ts:726-728
    return point if isinstance(point, tuple) else g1_from_hex(bytes(point))


def _norm_p2(point):  # index.ts:729-731
    return point if isinstance(point, tuple) else g2_from_signature(bytes(point))


def _norm_p2_hash(point, dst=DEFAULT_DST):  # index.ts:732-734
    return point if isinstance(point, tuple) else g2_hash_to_curve(bytes(point), dst)


def get_public_key(sk) -> bytes:  # index.ts:738-740
    return g1_to_hex(g1_from_private_key(sk), True)


def sign(message, sk, dst=DEFAULT_DST):  # index.ts:746-752
    msg_point = _norm_p2_hash(message, dst)
    g2_assert_validity(msg_point)
    sig_point = pt_multiply(G2, msg_point, normalize_priv_key(sk))
    if isinstance(message, tuple):
        return sig_point
    return g2_to_signature(sig_point)


def verify(signature, message, public_key, dst=DEFAULT_DST) -> bool:  # index.ts:756-767
    Pk = _norm_p1(public_key)
    Hm = _norm_p2_hash(message, dst)
    S = _norm_p2(signature)
    ePHm = pairing(pt_negate(G1, Pk), Hm, False)
    eGS = pairing(G1_BASE, S, False)
    exp = fp12_final_exponentiate(fp12_mul(eGS, ePHm))
    return exp == FP12_ONE


def aggregate_public_keys(public_keys):  # index.ts:773-778
    if not len(public_keys):
        raise OracleError("Expected non-empty array")
    agg = G1_ZERO
    for pk in map(_norm_p1, public_keys):
        agg = pt_add(G1, agg, pk)
    if isinstance(public_keys[0], tuple):
        return g1_assert_validity(agg)
    return g1_to_hex(agg, True)


def aggregate_signatures(signatures):  # index.ts:783-788
    if not len(signatures):
        raise OracleError("Expected non-empty array")
    agg = G2_ZERO
    for s in map(_norm_p2, signatures):
        agg = pt_add(G2, agg, s)
    if isinstance(signatures[0], tuple):
        return g2_assert_validity(agg)
    return g2_to_signature(agg)


def verify_batch(signature, messages, public_keys, dst=DEFAULT_DST) -> bool:  # index.ts:792-821
    if not len(messages):
        raise OracleError("Expected non-empty messages array")
    if len(public_keys) != len(messages):
        raise OracleError("Pubkey count should equal msg count")
    sig = _norm_p2(signature)
    n_messages = [_norm_p2_hash(m, dst) for m in messages]
    n_public_keys = [_norm_p1(pk) for pk in public_keys]
    try:
        paired = []
        # `new Set(nMessages)` dedups by OBJECT IDENTITY (index.ts:804): points hashed from bytes are
        # always distinct objects; only a PointG2 object passed twice is grouped.
        seen = []
        for m in n_messages:
            if not any(m is s for s in seen):
                seen.append(m)
        for message in seen:
            group = G1_ZERO
            for i, sub in enumerate(n_messages):
                if sub is message:
                    group = pt_add(G1, group, n_public_keys[i])
            paired.append(pairing(group, message, False))
        paired.append(pairing(pt_negate(G1, G1_BASE), sig, False))
        product = FP12_ONE
        for f in paired:
            product = fp12_mul(product, f)
        exp = fp12_final_exponentiate(product)
        return exp == FP12_ONE
    except OracleError:
        return False
