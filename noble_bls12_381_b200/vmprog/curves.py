"""G1 / G2 group operations, (de)compression, validity checks and hash-to-curve as tower-VM programs.

Restates (as branch-free dataflow over the 32 lanes of a CTA) the host-side bigint code of the reference:
  PointG1.fromHex index.ts:298-327, PointG2.fromSignature index.ts:500-530, assertValidity index.ts:383-388 /
  633-638 (isOnCurve :408-414 / :675-681, isTorsionFree :444-448 / :688-690), psi / psi2 math.ts:1398-1408,
  clearCofactor index.ts:659-672, map_to_curve_simple_swu_9mod16 math.ts:1220-1267, isogenyMapG2
  math.ts:1315-1325, ProjectivePoint add/double math.ts:974-1025, multiply math.ts:1061-1078.

Group law: the reference uses the incomplete add-1998-cmo-2 / dbl-1998-cmo-2 formulas plus explicit special
cases (math.ts:1000-1013).  On the device every point operation uses the COMPLETE projective formulas for
y^2 = x^3 + b (Renes-Costello-Batina 2015, a = 0), which compute the same group element for every input
(doubling, inverse, infinity) without branches; all observable results are either affine/compressed bytes or
projective-equality predicates, so they are bit-identical to the reference's.
Flags are plain 0/1 integers in slots; `select` picks per lane.
"""
from __future__ import annotations

from .builder import Builder, Lin, Quad, Val, P
from .tower import E2, Tower, X_PARAM, _mat, e2_zero

R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

# status codes (include/bls381_b200.h)
ST_OK, ST_INFINITY, ST_NOT_ON_CURVE, ST_NOT_IN_SUBGROUP, ST_BAD_ENCODING, ST_NO_SQRT = 0, 1, 2, 3, 4, 5


class FpF:
    """Field adaptor: Fp elements are Lin/Quad expressions."""

    def __init__(self, t: Tower):
        self.t, self.b = t, t.b

    def const(self, v):
        return self.t.fp_const(v)

    def zero(self):
        return Lin()

    def one(self):
        return self.t.fp_const(1)

    def mul(self, a, c):
        return a * c

    def sqr(self, a):
        return a * a

    def scale(self, a, k):
        return a * k

    def m(self, a):
        return _mat(self.b, a)

    def mul_b3(self, a):  # 3*b = 12
        return self.m(a) * self.t.fp_const(12)

    def mul_b(self, a):  # b = 4 (small scaling)
        return a * 4

    def is_zero(self, a) -> Val:
        return self.b.is_zero(a)

    def select(self, f: Val, a, c):
        return Lin.of(self.b.select(f, self.m(a), self.m(c)))

    def coeffs(self, a):
        return [a]


class Fp2F:
    """Field adaptor: Fp2 elements are E2."""

    def __init__(self, t: Tower):
        self.t, self.b = t, t.b

    def const(self, pair):
        return self.t.e2_const(pair)

    def zero(self):
        return e2_zero()

    def one(self):
        return E2(self.t.fp_const(1), Lin())

    def mul(self, a, c):
        return a * c

    def sqr(self, a):
        return a.sqr()

    def scale(self, a, k):
        return a.scale(k)

    def m(self, a):
        return a.m(self.b)

    def mul_b3(self, a):  # 3*b = 12(1+u)
        a = a.m(self.b)
        return a.mul_xi() * self.t.fp_const(12)

    def mul_b(self, a):  # b = 4(1+u)
        return a.mul_xi().scale(4)

    def is_zero(self, a) -> Val:
        a = a.m(self.b) if not (isinstance(a.c0, Lin) and isinstance(a.c1, Lin)) else a
        return self.b.flag_and(self.b.is_zero(a.c0), self.b.is_zero(a.c1))

    def select(self, f: Val, a, c):
        a, c = a.m(self.b), c.m(self.b)
        return E2(Lin.of(self.b.select(f, a.c0, c.c0)), Lin.of(self.b.select(f, a.c1, c.c1)))

    def coeffs(self, a):
        return [a.c0, a.c1]


SECRET_WINDOW = 3  # window of the constant-time scalar multiplication; measured on a B200 (sign, 65536 items): w=3 55.7 ms, w=4 60.8 ms (tools/ab_programs.py + tools/ab_sign.py)


class Curve:
    """Projective points (X, Y, Z) over a field adaptor F with complete formulas."""

    def __init__(self, F):
        self.F = F
        self.b = F.b

    def neg(self, p):
        return (p[0], -p[1], p[2])

    def m(self, p):
        return tuple(self.F.m(c) for c in p)

    def add(self, p, q):
        """Complete addition (RCB15 Alg. 7, a = 0).  p, q materialised."""
        F = self.F
        X1, Y1, Z1 = p
        X2, Y2, Z2 = q
        A = F.m(F.mul(X1, X2))
        B = F.m(F.mul(Y1, Y2))
        C = F.m(F.mul(Z1, Z2))
        D = F.m(F.mul(X1, Y2) + F.mul(X2, Y1))
        E = F.m(F.mul(Y1, Z2) + F.mul(Y2, Z1))
        Fv = F.m(F.mul(X1, Z2) + F.mul(X2, Z1))
        bC = F.m(F.mul_b3(C))
        bF = F.m(F.mul_b3(Fv))
        A3 = F.scale(A, 3)
        X3 = F.mul(D, B - bC) - F.mul(E, bF)
        Y3 = F.mul(B + bC, B - bC) + F.mul(A3, bF)
        Z3 = F.mul(E, B + bC) + F.mul(A3, D)
        return (F.m(X3), F.m(Y3), F.m(Z3))

    def dbl(self, p):
        """Complete doubling (RCB15 Alg. 9, a = 0)."""
        F = self.F
        X, Y, Z = p
        YY = F.m(F.sqr(Y))
        ZZ = F.m(F.sqr(Z))
        XY = F.m(F.mul(X, Y))
        YZ = F.m(F.mul(Y, Z))
        bZZ = F.m(F.mul_b3(ZZ))
        # X3 = 2 XY (YY - 3 bZZ) ; Y3 = (YY + bZZ)(YY - 3 bZZ) + 8 YY bZZ ; Z3 = 8 YY YZ
        t = YY - F.scale(bZZ, 3)
        X3 = F.mul(F.scale(XY, 2), t)
        Y3 = F.mul(YY + bZZ, t) + F.mul(F.scale(YY, 4), F.scale(bZZ, 2))
        Z3 = F.mul(F.scale(YY, 4), F.scale(YZ, 2))
        return (F.m(X3), F.m(Y3), F.m(Z3))

    # ---- Jacobian doublings for long runs of zero bits of a public scalar ---------------------------------------
    # (x, y) = (X/Z^2, Y/Z^3).  dbl-2009-l for a = 0: 24 products in 12 micro-ops over Fp2 instead of the 30 / 16 of the
    # complete homogeneous doubling.  The formulas hold for every point of an odd-order curve, including the point at
    # infinity in the form (X, Y, 0) with X^3 = Y^2 != 0 (it maps to (X^4, X^6, 0)); additions stay complete (RCB15).
    def to_jacobian(self, p):
        F = self.F
        X, Y, Z = p
        inf = F.is_zero(Z)
        ZZ = F.m(F.sqr(Z))
        one = self._mat_point((F.one(),))[0]
        return (F.select(inf, one, F.m(F.mul(X, Z))), F.select(inf, one, F.m(F.mul(Y, ZZ))), Z)

    def from_jacobian(self, p):
        F = self.F
        X, Y, Z = p
        ZZ = F.m(F.sqr(Z))
        return (F.m(F.mul(X, Z)), Y, F.m(F.mul(ZZ, Z)))

    def dbl_jacobian(self, p):
        F = self.F
        X, Y, Z = p
        A = F.m(F.sqr(X))
        B = F.m(F.sqr(Y))
        Z3 = F.m(F.mul(F.scale(Y, 2), Z))
        D = F.m(F.mul(F.scale(X, 4), B))
        A3 = F.scale(A, 3)
        X3 = F.m(F.sqr(A3) - F.scale(D, 2))
        Y3 = F.m(F.mul(A3, D - X3) - F.scale(F.sqr(B), 8))
        return (X3, Y3, Z3)

    MIN_JACOBIAN_RUN = 4  # shorter runs of doublings do not pay for the two coordinate conversions

    def mul_fixed(self, p, k: int):
        """[k]P for a public constant k (MSB-first double-and-add; complete additions cover every case, runs of at least
        MIN_JACOBIAN_RUN doublings are done in Jacobian coordinates)."""
        acc = p
        bits = [(k >> i) & 1 for i in range(k.bit_length() - 2, -1, -1)]
        i = 0
        while i < len(bits):
            run = 1  # doublings up to and including the next set bit (or the end)
            while bits[i + run - 1] == 0 and i + run < len(bits):
                run += 1
            if run >= self.MIN_JACOBIAN_RUN:
                j = self.to_jacobian(acc)
                for _ in range(run):
                    j = self.dbl_jacobian(j)
                acc = self.from_jacobian(j)
            else:
                for _ in range(run):
                    acc = self.dbl(acc)
            if bits[i + run - 1]:
                acc = self.add(acc, p)
            i += run
        return acc

    def mul_secret(self, p, nbits: int, bit_fn, window: int = None):
        """[k]P for a per-lane secret k < 2^nbits, constant-time: fixed windows of `window` bits, a table of
        0*P .. (2^w - 1)*P, per digit `window` doublings, a select tree over the digit's bit flags (every lane
        touches every table entry) and ONE complete addition.  The reference's ladder (math.ts:1061-1078) does a
        double and an add per bit; the group element -- hence every serialised byte -- is the same.
        `bit_fn(i)` creates the flag for bit i (MSB first); it is fenced behind the previous digit so that the
        flags are produced just in time instead of occupying slots from the start of the program."""
        F, b = self.F, self.b
        window = window or SECRET_WINDOW
        p = self._mat_point(p)
        tab = [self._mat_point((F.zero(), F.one(), F.zero())), p]
        for k in range(2, 1 << window):
            tab.append(self.dbl(tab[k // 2]) if k % 2 == 0 else self.add(tab[k - 1], p))
        ndig = (nbits + window - 1) // window
        acc = None
        for d in range(ndig - 1, -1, -1):
            if acc is not None:
                for _ in range(window):
                    acc = self.dbl(acc)
                fence = [c.t and list(c.t)[0].op for c in F.coeffs(acc[0]) if isinstance(c, Lin) and c.t]
                b.set_after([op for op in fence if op is not None])
            flags = [bit_fn(window * d + i) for i in range(window)]
            b.set_after([])
            level = tab
            for f in flags:  # bit i of the digit halves the candidate list
                level = [tuple(F.select(f, hi, lo) for hi, lo in zip(level[2 * j + 1], level[2 * j])) for j in range(len(level) // 2)]
            acc = level[0] if acc is None else self.add(acc, level[0])
        return acc

    def _mat_point(self, p):
        """Materialise a point with possibly constant coordinates into slots (copy ops)."""
        F, b = self.F, self.b
        out = []
        for c in p:
            cs = []
            for x in F.coeffs(c):
                if isinstance(x, Lin) and x.is_zero():
                    x = Lin.of(b.const_raw(0))
                cs.append(Lin.of(b.mat(x)))
            out.append(cs[0] if len(cs) == 1 else E2(cs[0], cs[1]))
        return tuple(out)

    def eq(self, p, q) -> Val:
        """Projective equality (math.ts:915-927)."""
        F, b = self.F, self.b
        xe = F.is_zero(F.mul(p[0], q[2]) - F.mul(q[0], p[2]))
        ye = F.is_zero(F.mul(p[1], q[2]) - F.mul(q[1], p[2]))
        return b.flag_and(xe, ye)

    def is_on_curve(self, p) -> Val:
        """y^2 z - x^3 - b z^3 == 0  (index.ts:408-414 / 675-681)."""
        F = self.F
        X, Y, Z = p
        XX = F.m(F.sqr(X))
        YY = F.m(F.sqr(Y))
        ZZ = F.m(F.sqr(Z))
        return F.is_zero(F.mul(YY, Z) - F.mul(XX, X) - F.mul(F.mul_b(ZZ), Z))


# ------------------------------------------------------------------------------------------ exponent chains
def pow_fixed(mul, sqr, m, one, a, e: int, window: int = 4):
    """a^e for a public exponent with a fixed window; `m` materialises, `one` is used only if e == 0."""
    if e == 0:
        return one
    tab = [None, a]
    for k in range(2, 1 << window):
        tab.append(m(mul(tab[k - 1], a)))
    digits = []
    while e:
        digits.append(e & ((1 << window) - 1))
        e >>= window
    digits.reverse()
    acc = tab[digits[0]]
    for d in digits[1:]:
        for _ in range(window):
            acc = m(sqr(acc))
        if d:
            acc = m(mul(acc, tab[d]))
    return acc


class Ingest:
    """Builds the ingest programs on top of a Tower."""

    def __init__(self, t: Tower):
        self.t, self.b = t, t.b
        self.F1, self.F2 = FpF(t), Fp2F(t)
        self.G1, self.G2 = Curve(self.F1), Curve(self.F2)
        # psi constants (math.ts:1390-1403 reduce to conj(x)*cx, conj(y)*cy)
        from .tower import _fp2_pow, _fp2_mul
        xi = (1, 1)
        inv = lambda a: _fp2_pow(a, P * P - 2)
        self.psi_cx = inv(_fp2_pow(xi, (P - 1) // 3))
        self.psi_cy = inv(_fp2_pow(xi, (P - 1) // 2))
        self.psi2_c1 = 0x1A0111EA397FE699EC02408663D4DE85AA0D857D89759AD4897D29650FB85F9B409427EB4F49FFFD8BFD00000000AAAC
        self.cubic_root = 0x5F19672FDF76CE51BA69C6076A0F77EADDB3A93BE6F89688DE17D813620A00022E01FFFFFFFEFFFE

    # ---- field helpers ---------------------------------------------------------------------------
    def fp_sqrt_candidate(self, a: Lin) -> Lin:  # a^((p+1)/4)   math.ts:260-264
        b = self.b
        a = Lin.of(b.mat(a))
        return pow_fixed(lambda x, y: x * y, lambda x: x * x, lambda q: Lin.of(b.mat(q)), self.t.fp_const(1), a, (P + 1) // 4)

    def fp2_pow(self, a: E2, e: int) -> E2:
        """a^e in Fp2 for a public exponent.  Exponents longer than p are split with the Frobenius map:
        e = e1*p + e0 and a^p = conj(a) (math.ts:529-531), so a^e = conj(a)^e1 * a^e0 is ONE joint
        double-base chain over 2-bit digit pairs -- half the squarings of the plain 758-bit chains the
        reference runs for (p^2-9)/16 and (p^2+8)/16 (math.ts:1200, 489); same field element, same bytes."""
        b = self.b
        a = a.m(b)
        if e.bit_length() <= P.bit_length():
            return pow_fixed(lambda x, y: x * y, lambda x: x.sqr(), lambda q: q.m(b), self.F2.one(), a, e)
        e1, e0 = divmod(e, P)
        assert e1 < P
        pw = [None, a, a.sqr().m(b)]
        pw.append((pw[2] * a).m(b))
        tab = {}  # (i, j) -> a^i * conj(a)^j
        for i in range(4):
            for j in range(4):
                if i == 0 and j == 0:
                    continue
                if j == 0:
                    tab[(i, j)] = pw[i]
                elif i == 0:
                    tab[(i, j)] = pw[j].conj()
                else:
                    tab[(i, j)] = (pw[i] * pw[j].conj()).m(b)
        ndig = (max(e0.bit_length(), e1.bit_length()) + 1) // 2
        acc = None
        for d in range(ndig - 1, -1, -1):
            i, j = (e0 >> (2 * d)) & 3, (e1 >> (2 * d)) & 3
            if acc is not None:
                acc = acc.sqr().m(b).sqr().m(b)
            if i or j:
                acc = tab[(i, j)] if acc is None else (acc * tab[(i, j)]).m(b)
        return acc

    def fp2_eq(self, a: E2, c: E2) -> Val:
        return self.F2.is_zero(a - c)

    def flag_const(self, v: int) -> Lin:
        return Lin.of(self.b.const_raw(v))

    def status_chain(self, conds) -> Val:
        """conds = [(flag, code), ...] in priority order (first true wins); else ST_OK."""
        b = self.b
        st = self.flag_const(ST_OK)
        for flag, code in reversed(conds):
            st = Lin.of(b.select(flag, self.flag_const(code), st))
        return b.mat(st)

    # ---- endomorphisms ---------------------------------------------------------------------------
    def g2_psi(self, p):  # projective form of math.ts:1398-1403
        X, Y, Z = p
        return self.G2.m((self.t.mul_const2(X.conj(), self.psi_cx), self.t.mul_const2(Y.conj(), self.psi_cy), Z.conj()))

    def g2_psi2(self, p):  # math.ts:1406-1408
        X, Y, Z = p
        return self.G2.m((X * self.t.fp_const(self.psi2_c1), -Y, Z))

    def g1_is_torsion_free(self, p) -> Val:  # index.ts:444-448:  [x]([x]P) == -... : u2P = x^2 P == phi(P)
        xP = self.G1.neg(self.G1.mul_fixed(p, X_PARAM))  # mulCurveX = -(|x| P)
        u2P = self.G1.mul_fixed(self.G1.m(xP), X_PARAM)  # mulCurveMinusX
        phi = self.G1.m((p[0] * self.t.fp_const(self.cubic_root), p[1], p[2]))
        return self.G1.eq(u2P, phi)

    def g2_is_torsion_free(self, p) -> Val:  # index.ts:688-690:  [-x]P == psi(P)
        xP = self.G2.m(self.G2.neg(self.G2.mul_fixed(p, X_PARAM)))
        return self.G2.eq(xP, self.g2_psi(p))

    def g2_clear_cofactor(self, p):  # index.ts:659-672
        G = self.G2
        t1 = G.m(G.neg(G.mul_fixed(p, X_PARAM)))  # [-x]P
        t2 = self.g2_psi(p)
        t3 = self.g2_psi2(G.dbl(p))
        t3 = G.add(t3, G.m(G.neg(t2)))
        t2 = G.add(t1, t2)
        t2 = G.m(G.neg(G.mul_fixed(t2, X_PARAM)))
        t3 = G.add(t3, t2)
        t3 = G.add(t3, G.m(G.neg(t1)))
        return G.add(t3, G.m(G.neg(p)))

    def g2_mul_secret_gls4(self, p, bit_fn):
        """[k]P for P in G2 (the order-r subgroup) and a per-lane secret k < z^4, z = |x|, given by its four base-z digits
        a_0..a_3 (64 bits each; `bit_fn(64*i + j)` = bit j of a_i).  On G2 the untwist-Frobenius-twist endomorphism acts as
        psi(P) = [x]P = -[z]P (that is the reference's own subgroup test, index.ts:688-690), so
            [k]P = a_0 Q_0 + a_1 Q_1 + a_2 Q_2 + a_3 Q_3,   Q_i = (-1)^i psi^i(P),
        one joint 64-step ladder: a double, a 16-way select over the table of subset sums (every lane touches every entry)
        and ONE complete addition per step -- 64 doublings instead of the 255 of a plain ladder (math.ts:1061-1078 does 381);
        the group element, hence every serialised byte, is the same."""
        G, F, b = self.G2, self.F2, self.b
        p = G._mat_point(p)
        q1 = G.m(G.neg(self.g2_psi(p)))
        q2 = self.g2_psi2(p)
        q3 = G.m(G.neg(self.g2_psi(q2)))
        tab = [G._mat_point((F.zero(), F.one(), F.zero())), p]
        for q in (q1, q2, q3):  # tab[idx] = sum of Q_i over the set bits of idx
            base = len(tab)
            tab.append(q)
            for k in range(1, base):
                tab.append(G.add(q, tab[k]))
        acc = None
        for j in range(63, -1, -1):
            if acc is not None:
                acc = G.dbl(acc)
                fence = [c.t and list(c.t)[0].op for c in F.coeffs(acc[0]) if isinstance(c, Lin) and c.t]
                b.set_after([op for op in fence if op is not None])
            flags = [bit_fn(64 * i + j) for i in range(4)]
            b.set_after([])
            level = tab
            for f in flags:
                level = [tuple(F.select(f, hi, lo) for hi, lo in zip(level[2 * t + 1], level[2 * t])) for t in range(len(level) // 2)]
            acc = level[0] if acc is None else G.add(acc, level[0])
        return acc

    # ---- affine conversion ---------------------------------------------------------------------------
    def g1_to_affine(self, p):
        zi = self.t.fp_inv(p[2])
        return Lin.of(self.b.mat(p[0] * zi)), Lin.of(self.b.mat(p[1] * zi))

    def g2_to_affine(self, p):
        zi = self.t.fp2_inv(p[2]).m(self.b)
        return (p[0] * zi).m(self.b), (p[1] * zi).m(self.b)


# ------------------------------------------------------------------------------------------ programs
# Buffer conventions for the ingest programs (api.cu):
#   g1_decompress : 0 = in  n x 48 B compressed      2 = out n x 96 B affine    5 = status n x int32
#   g2_decompress : 0 = in  n x 96 B signature form  2 = out n x 192 B affine   5 = status n x int32
BUF_IN, BUF_OUT, BUF_STATUS = 0, 2, 5

# eighth roots of unity / etas used by the reference's Fp2 sqrt and SWU (math.ts:1415-1452)
_RV1 = 0x6AF0E0437FF400B6831E36D6BD17FFE48395DABC2D3435E77F76E17009241C5EE67992F72EC05F4C81084FBEDE3CC09
_EV = (
    0x699BE3B8C6870965E5BF892AD5D2CC7B0E85A117402DFD83B7F4A947E02D978498255A2AAEC0AC627B5AFBDF1BF1C90,
    0x8157CD83046453F5DD0972B6E3949E4288020B5B8A9CC99CA07E27089A2CE2436D965026ADAD3EF7BABA37F2183E9B5,
    0xAB1C2FFDD6C253CA155231EB3E71BA044FD562F6F72BC5BAD5EC46A0B7A3B0247CF08CE6C6317F40EDBC653A72DEE17,
    0xAA404866706722864480885D68AD0CCAC1967C7544B447873CC37E0181271E006DF72162A3D3E0287BF597FBF7F8FC1,
)
ROOTS_OF_UNITY_POS = [(1, 0), (_RV1, (-_RV1) % P), (0, 1), (_RV1, _RV1)]
ETAS = [(_EV[0], _EV[1]), ((-_EV[1]) % P, _EV[0]), (_EV[2], _EV[3]), ((-_EV[3]) % P, _EV[2])]


def _fp2_inv_const(a):
    from .tower import _fp2_pow
    return _fp2_pow(a, P * P - 2)


def build_g1_decompress(warps=8) -> Builder:
    """PointG1.fromHex for 48-byte compressed keys (index.ts:301-315, 325) incl. assertValidity."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    flag_inf = b.bit(BUF_IN, 0, 48, 382)
    flag_sign = b.bit(BUF_IN, 0, 48, 381)
    x = Lin.of(b.inp_bytes(BUF_IN, 0, 48, clear_top=3))
    x2 = Lin.of(b.mat(x * x))
    right = Lin.of(b.mat(x2 * x + t.fp_const(4)))  # y^2 = x^3 + b
    cand = ig.fp_sqrt_candidate(right)
    no_sqrt = b.flag_not(b.is_zero(cand * cand - right))
    flip = b.flag_xor(b.gt_half(cand), flag_sign)  # (y*2)/P != aflag
    y = Lin.of(b.select(flip, -cand, cand))
    pt = (x, y, Lin.of(b.mat(t.fp_const(1))))
    not_sub = b.flag_not(ig.g1_is_torsion_free(pt))
    st = ig.status_chain([(flag_inf, ST_INFINITY), (no_sqrt, ST_BAD_ENCODING), (not_sub, ST_NOT_IN_SUBGROUP)])
    b.out(x, BUF_OUT, 0)
    b.out(y, BUF_OUT, 1)
    b.out_word(st, BUF_STATUS)
    return b


def _fp2_sqrt_any(ig: Ingest, a: E2):
    """Some square root of `a` (or garbage) and the flag `found`; candidates per math.ts:486-501."""
    b, t = ig.b, ig.t
    a = a.m(b)
    c = ig.fp2_pow(a, (P * P + 8) // 16)
    y = None
    found = None
    for root in ROOTS_OF_UNITY_POS:
        yk = t.mul_const2(c, _fp2_inv_const(root)).m(b)
        ok = ig.fp2_eq(yk.sqr(), a)
        if y is None:
            y, found = yk, ok
        else:
            y = ig.F2.select(ok, yk, y)
            found = b.flag_or(ok, found)
    return y, found


def build_g2_decompress(warps=8) -> Builder:
    """PointG2.fromSignature for 96-byte compressed signatures (index.ts:500-530) incl. assertValidity."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    flag_inf = b.bit(BUF_IN, 0, 48, 382)
    aflag = b.bit(BUF_IN, 0, 48, 381)
    x1 = Lin.of(b.inp_bytes(BUF_IN, 0, 48, clear_top=3))   # z1 mod 2^381  -> imaginary part
    x0 = Lin.of(b.inp_bytes(BUF_IN, 48, 48))               # z2            -> real part
    x = E2(x0, x1)
    xx = x.sqr().m(b)
    y2 = (xx * x + E2(t.fp_const(4), t.fp_const(4))).m(b)   # x^3 + 4(1+u)
    y, found = _fp2_sqrt_any(ig, y2)
    y = y.m(b)
    y1_zero = b.is_zero(y.c1)
    g1 = b.gt_half(y.c1)
    g0 = b.gt_half(y.c0)
    # isGreater = y1 > 0 && (y1*2)/P != aflag ; isZero = y1 == 0 && (y0*2)/P != aflag     (index.ts:522-526)
    is_greater = b.flag_and(b.flag_not(y1_zero), b.flag_xor(g1, aflag))
    is_zero = b.flag_and(y1_zero, b.flag_xor(g0, aflag))
    flip = b.flag_or(is_greater, is_zero)
    y = ig.F2.select(flip, -y, y)
    pt = (x, y, E2(Lin.of(b.mat(t.fp_const(1))), Lin.of(b.mat(Lin.of(b.const_raw(0))))))
    not_sub = b.flag_not(ig.g2_is_torsion_free(pt))
    st = ig.status_chain([(flag_inf, ST_INFINITY), (b.flag_not(found), ST_NO_SQRT), (not_sub, ST_NOT_IN_SUBGROUP)])
    b.out(x.c0, BUF_OUT, 0)
    b.out(x.c1, BUF_OUT, 1)
    b.out(y.c0, BUF_OUT, 2)
    b.out(y.c1, BUF_OUT, 3)
    b.out_word(st, BUF_STATUS)
    return b


PROGRAMS = {
    "g1_decompress": build_g1_decompress,
    "g2_decompress": build_g2_decompress,
}


# ------------------------------------------------------------------------------------------ hash to G2
# 3-isogeny coefficients E' -> E (math.ts:1547-1610), leading coefficient first
_ISO3 = {
    "xnum": [
        (0x171D6541FA38CCFAED6DEA691F5FB614CB14B4E7F4E810AA22D6108F142B85757098E38D0F671C7188E2AAAAAAAA5ED1, 0),
        (0x11560BF17BAA99BC32126FCED787C88F984F87ADF7AE0C7F9A208C6B4F20A4181472AAA9CB8D555526A9FFFFFFFFC71E,
         0x8AB05F8BDD54CDE190937E76BC3E447CC27C3D6FBD7063FCD104635A790520C0A395554E5C6AAAA9354FFFFFFFFE38D),
        (0, 0x11560BF17BAA99BC32126FCED787C88F984F87ADF7AE0C7F9A208C6B4F20A4181472AAA9CB8D555526A9FFFFFFFFC71A),
        (0x5C759507E8E333EBB5B7A9A47D7ED8532C52D39FD3A042A88B58423C50AE15D5C2638E343D9C71C6238AAAAAAAA97D6,
         0x5C759507E8E333EBB5B7A9A47D7ED8532C52D39FD3A042A88B58423C50AE15D5C2638E343D9C71C6238AAAAAAAA97D6),
    ],
    "xden": [(0, 0), (1, 0), (12, P - 12), (0, P - 72)],
    "ynum": [
        (0x124C9AD43B6CF79BFBF7043DE3811AD0761B0F37A1E26286B0E977C69AA274524E79097A56DC4BD9E1B371C71C718B10, 0),
        (0x11560BF17BAA99BC32126FCED787C88F984F87ADF7AE0C7F9A208C6B4F20A4181472AAA9CB8D555526A9FFFFFFFFC71C,
         0x8AB05F8BDD54CDE190937E76BC3E447CC27C3D6FBD7063FCD104635A790520C0A395554E5C6AAAA9354FFFFFFFFE38F),
        (0, 0x5C759507E8E333EBB5B7A9A47D7ED8532C52D39FD3A042A88B58423C50AE15D5C2638E343D9C71C6238AAAAAAAA97BE),
        (0x1530477C7AB4113B59A4C18B076D11930F7DA5D4A07F649BF54439D87D27E500FC8C25EBF8C92F6812CFC71C71C6D706,
         0x1530477C7AB4113B59A4C18B076D11930F7DA5D4A07F649BF54439D87D27E500FC8C25EBF8C92F6812CFC71C71C6D706),
    ],
    "yden": [(1, 0), (18, P - 18), (0, P - 216), (P - 432, P - 432)],
}


def _sgn0_fp2(ig: Ingest, x: E2) -> Val:  # math.ts:1179-1185
    b = ig.b
    x = x.m(b)
    s0 = b.parity(x.c0)
    z0 = b.is_zero(x.c0)
    s1 = b.parity(x.c1)
    return b.flag_or(s0, b.flag_and(z0, s1))


def _swu_g2(ig: Ingest, tt: E2):
    """map_to_curve_simple_swu_9mod16 (math.ts:1220-1267) returning the point of E' as (N : y*D : D)."""
    b, t, F = ig.b, ig.t, ig.F2
    A = (0, 240)
    B = (1012, 1012)
    Z = ((-2) % P, (-1) % P)
    tt = tt.m(b)
    t2 = tt.sqr().m(b)
    z_t2 = t.mul_const2(t2, Z).m(b)
    ztzt = (z_t2 + z_t2.sqr()).m(b)
    den = (-t.mul_const2(ztzt, A)).m(b)
    num = t.mul_const2(ztzt + E2(t.fp_const(1), Lin()), B).m(b)
    den_zero = F.is_zero(den)
    den = F.select(den_zero, t.e2_const(_fp2_mul_c(Z, A)).m(b) if False else _const_e2(ig, _fp2_mul_c(Z, A)), den)
    d2 = den.sqr().m(b)
    v = (d2 * den).m(b)
    n2 = num.sqr().m(b)
    u = (n2 * num + t.mul_const2((num * d2).m(b), A) + t.mul_const2(v, B)).m(b)
    # sqrt_div_fp2(u, v)   (math.ts:1195-1214)
    v2 = v.sqr().m(b)
    v4 = v2.sqr().m(b)
    v7 = ((v4 * v2).m(b) * v).m(b)
    uv7 = (u * v7).m(b)
    uv15 = (uv7 * (v7 * v).m(b)).m(b)
    gamma = (ig.fp2_pow(uv15, (P * P - 9) // 16) * uv7).m(b)
    y = gamma
    success = None
    for root in ROOTS_OF_UNITY_POS:
        cand = t.mul_const2(gamma, root).m(b)
        ok = F.is_zero((cand.sqr().m(b) * v).m(b) - u)
        y = F.select(ok, cand, y)
        success = ok if success is None else b.flag_or(ok, success)
    # second candidate family (x1 = Z t^2 x0)
    t3 = (t2 * tt).m(b)
    sc_x1 = (gamma * t3).m(b)
    zt2_3 = ((z_t2.sqr().m(b)) * z_t2).m(b)
    u2 = (zt2_3 * u).m(b)
    y2 = sc_x1
    for eta in ETAS:
        cand = t.mul_const2(sc_x1, eta).m(b)
        ok = F.is_zero((cand.sqr().m(b) * v).m(b) - u2)
        y2 = F.select(ok, cand, y2)
    y = F.select(success, y, y2)
    num = F.select(success, num, (num * z_t2).m(b))
    flip = b.flag_xor(_sgn0_fp2(ig, tt), _sgn0_fp2(ig, y))
    y = F.select(flip, -y, y)
    return (num, (y * den).m(b), den)


def _fp2_mul_c(a, c):
    return ((a[0] * c[0] - a[1] * c[1]) % P, (a[0] * c[1] + a[1] * c[0]) % P)


def _const_e2(ig: Ingest, pair) -> E2:
    """Fp2 constant copied into slots (select needs slot operands of one kind)."""
    b, t = ig.b, ig.t
    return E2(Lin.of(b.mat(t.fp_const(pair[0]) if pair[0] % P else Lin.of(b.const_raw(0)))),
              Lin.of(b.mat(t.fp_const(pair[1]) if pair[1] % P else Lin.of(b.const_raw(0)))))


def _add_generic(ig: Ingest, p, q):
    """add-1998-cmo-2 as coded in math.ts:1008-1024 (valid on the isogenous curve E' where a != 0; the
    reference's special cases P == +-Q cannot be reached by two independent SWU outputs in practice)."""
    b, F = ig.b, ig.F2
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    U1 = (Y2 * Z1).m(b)
    U2 = (Y1 * Z2).m(b)
    V1 = (X2 * Z1).m(b)
    V2 = (X1 * Z2).m(b)
    U = U1 - U2
    V = V1 - V2
    VV = V.sqr().m(b) if False else (V.m(b)).sqr().m(b)
    Vm = V.m(b)
    Um = U.m(b)
    VVV = (VV * Vm).m(b)
    V2VV = (V2 * VV).m(b)
    W = (Z1 * Z2).m(b)
    UU = Um.sqr().m(b)
    Av = ((UU * W) - VVV - V2VV.scale(2)).m(b)
    X3 = (Vm * Av).m(b)
    Y3 = (Um * (V2VV - Av) - VVV * U2).m(b)
    Z3 = (VVV * W).m(b)
    return (X3, Y3, Z3)


def _isogeny3_projective(ig: Ingest, p):
    """isogenyMapG2 (math.ts:1315-1325) on a projective point of E', staying projective on E."""
    b, t = ig.b, ig.t
    X, Y, Z = p
    X2 = X.sqr().m(b)
    Z2 = Z.sqr().m(b)
    X3 = (X2 * X).m(b)
    Z3 = (Z2 * Z).m(b)
    X2Z = (X2 * Z).m(b)
    XZ2 = (X * Z2).m(b)
    mono = [X3, X2Z, XZ2, Z3]

    def poly(coeffs):
        acc = None
        for k, m_ in zip(coeffs, mono):
            if k[0] % P == 0 and k[1] % P == 0:
                continue
            term = t.mul_const2(m_, k)
            acc = term if acc is None else acc + term
        return acc.m(b)

    XN, XD, YN, YD = (poly(_ISO3[k]) for k in ("xnum", "xden", "ynum", "yden"))
    YDZ = (YD * Z).m(b)
    Xo = (XN * YDZ).m(b)
    Yo = ((Y * YN).m(b) * XD).m(b)
    Zo = (XD * YDZ).m(b)
    return (Xo, Yo, Zo)


def _hash_to_field_g2(ig: Ingest, buf: int):
    """hash_to_field (index.ts:240-267) tail: four 64-byte big-endian chunks -> u0, u1 in Fp2."""
    b, t = ig.b, ig.t
    two256 = t.fp_const(1 << 256)
    es = []
    for j in range(4):
        hi = Lin.of(b.inp_bytes(buf, 64 * j, 32))
        lo = Lin.of(b.inp_bytes(buf, 64 * j + 32, 32))
        es.append(Lin.of(b.mat(hi * two256 + lo)))
    return E2(es[0], es[1]), E2(es[2], es[3])


def _hash_to_g2_projective(ig: Ingest, buf: int):
    """PointG2.hashToCurve (index.ts:481-490) from the 256 uniform bytes of expand_message_xmd."""
    u0, u1 = _hash_to_field_g2(ig, buf)
    p0 = _swu_g2(ig, u0)
    p1 = _swu_g2(ig, u1)
    s = _add_generic(ig, p0, p1)
    e = _isogeny3_projective(ig, s)
    return ig.g2_clear_cofactor(e)


def _e1_points_in(ig: Ingest, buf: int):
    """The two points of E' per message written by swu_g2_kernel (csrc/swu_g2.cuh): 12 Montgomery residues, 48 bytes
    big-endian each, (X.c0, X.c1, Y.c0, Y.c1, Z.c0, Z.c1) of SWU(u0) then of SWU(u1)."""
    b = ig.b
    f = [Lin.of(b.inp_bytes(buf, 48 * k, 48, montgomery=False)) for k in range(12)]
    return tuple((E2(f[6 * j], f[6 * j + 1]), E2(f[6 * j + 2], f[6 * j + 3]), E2(f[6 * j + 4], f[6 * j + 5])) for j in range(2))


def _hash_tail_projective(ig: Ingest, buf: int):
    """Second half of PointG2.hashToCurve (index.ts:484-489): P0 + P1, the 3-isogeny, clearCofactor."""
    p0, p1 = _e1_points_in(ig, buf)
    s = _add_generic(ig, p0, p1)
    e = _isogeny3_projective(ig, s)
    return ig.g2_clear_cofactor(e)


def build_h2g2_tail(warps=4) -> Builder:
    """buffer 0: n x 576 B (two points of E' per message, from swu_g2_kernel) -> buffer 2: n x 192 B affine H(m)."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    h = _hash_tail_projective(ig, BUF_IN)
    x, y = ig.g2_to_affine(h)
    b.out(x.c0, BUF_OUT, 0)
    b.out(x.c1, BUF_OUT, 1)
    b.out(y.c0, BUF_OUT, 2)
    b.out(y.c1, BUF_OUT, 3)
    return b


PROGRAMS["h2g2_tail"] = build_h2g2_tail


def build_hash_to_g2(warps=8) -> Builder:
    """buffer 0: n x 256 B uniform bytes -> buffer 2: n x 192 B affine H(m) (x.c0, x.c1, y.c0, y.c1)."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    h = _hash_to_g2_projective(ig, BUF_IN)
    x, y = ig.g2_to_affine(h)
    b.out(x.c0, BUF_OUT, 0)
    b.out(x.c1, BUF_OUT, 1)
    b.out(y.c0, BUF_OUT, 2)
    b.out(y.c1, BUF_OUT, 3)
    return b


PROGRAMS["hash_to_g2"] = build_hash_to_g2


# ------------------------------------------------------------------------------------------ hash to G1
# (min-signature deployments hash to G1)
_SWU1_A = 0x144698A3B8E9433D693A02C96D4982B0EA985383EE66A8D8E8981AEFD881AC98936F8DA0E0F97F5CF428082D584C1D  # math.ts:1271-1273
_SWU1_B = 0x12E2908D11688030018B12E8753EEE3B2016C1F0F24F4070A0B9C14FCEF35EF55A23215A316CEAA5D1CC48E98E172BE0  # math.ts:1274-1276
_SWU1_Z = 11


def _swu_g1(ig: Ingest, u: Lin):
    """map_to_curve_simple_swu_3mod4 (math.ts:1270-1313), branch-free, returning the point of E' as (xNum : y * xDen : xDen)."""
    b, t = ig.b, ig.t
    m = lambda q: Lin.of(b.mat(q))
    A, B, Z = t.fp_const(_SWU1_A), t.fp_const(_SWU1_B), t.fp_const(_SWU1_Z)
    c2 = t.fp_const(pow(pow((-_SWU1_Z) % P, 3, P), (P + 1) // 4, P))  # sqrt((-Z)^3)
    assert pow(pow(pow((-_SWU1_Z) % P, 3, P), (P + 1) // 4, P), 2, P) == pow((-_SWU1_Z) % P, 3, P)
    u = m(u)
    tv1 = m(u * u)
    tv3 = m(tv1 * Z)
    xd0 = m(tv3 * tv3 + tv3)
    xn1 = m((xd0 + t.fp_const(1)) * B)
    xn2 = m(tv3 * xn1)
    xd = m((-xd0) * A)
    xd = Lin.of(b.select(b.is_zero(xd), m(t.fp_const(_SWU1_A * _SWU1_Z % P)), xd))
    tv2 = m(xd * xd)
    gxd = m(tv2 * xd)
    atv2 = m(tv2 * A)
    gx1 = m(m(xn1 * xn1 + atv2) * xn1 + gxd * B)
    tv2 = m(gx1 * gxd)
    tv4 = m(m(gxd * gxd) * tv2)
    y1 = m(pow_fixed(lambda x, y: x * y, lambda x: x * x, m, t.fp_const(1), tv4, (P - 3) // 4) * tv2)
    y2 = m(m(m(y1 * c2) * tv1) * u)
    ok = b.is_zero(m(y1 * y1) * gxd - gx1)
    xnum = Lin.of(b.select(ok, xn1, xn2))
    ypos = Lin.of(b.select(ok, y1, y2))
    flip = b.flag_xor(b.parity(u), b.parity(ypos))  # sgn0_m_eq_1(u) != sgn0_m_eq_1(yPos)
    y = Lin.of(b.select(flip, -ypos, ypos))
    return (xnum, m(y * xd), xd)


def _add_generic_fp(ig: Ingest, p, q):
    """add-1998-cmo-2 as coded in math.ts:1008-1024 over Fp (valid on the isogenous curve E', a != 0; the special cases
    P == +-Q cannot be reached by two independent SWU outputs in practice)."""
    b = ig.b
    m = lambda e: Lin.of(b.mat(e))
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    U2 = m(Y1 * Z2)
    V2 = m(X1 * Z2)
    U = m(Y2 * Z1 - U2)
    V = m(X2 * Z1 - V2)
    VV = m(V * V)
    VVV = m(VV * V)
    V2VV = m(V2 * VV)
    W = m(Z1 * Z2)
    Av = m(m(U * U) * W - VVV - V2VV * 2)
    return (m(V * Av), m(U * (V2VV - Av) - VVV * U2), m(VVV * W))


def _isogeny11_projective(ig: Ingest, p):
    """isogenyMapG1 (math.ts:1306-1313, 1327; coefficients math.ts:1612-1790) on a projective point of E', staying projective:
    x' = N_x / (Z D_x), y' = Y N_y / (Z D_y) with the polynomials homogenised -> (N_x D_y : Y N_y D_x : Z D_x D_y)."""
    from . import iso11
    b, t = ig.b, ig.t
    m = lambda e: Lin.of(b.mat(e))
    X, Y, Z = p
    xp, zp = [None, X], [None, Z]
    for k in range(2, 16):
        xp.append(m(xp[k // 2] * xp[k - k // 2]))
        zp.append(m(zp[k // 2] * zp[k - k // 2]))

    def mono(i, deg):  # X^i Z^(deg - i)
        j = deg - i
        if i == 0:
            return zp[j]
        if j == 0:
            return xp[i]
        return m(xp[i] * zp[j])

    def poly(coeffs):  # highest degree first (Horner order of the reference)
        deg = len(coeffs) - 1
        acc = None
        for k, c in enumerate(coeffs):
            if c % P == 0:
                continue
            term = mono(deg - k, deg) * t.fp_const(c)
            acc = term if acc is None else acc + term
        return m(acc)

    NX, DX, NY, DY = poly(iso11.XNUM), poly(iso11.XDEN), poly(iso11.YNUM), poly(iso11.YDEN)
    ZDX = m(Z * DX)
    return (m(NX * DY), m(m(Y * NY) * DX), m(ZDX * DY))


def build_hash_to_g1(warps=4) -> Builder:
    """PointG1.hashToCurve (index.ts:331-339) from the 128 uniform bytes of expand_message_xmd (hash_to_field, m = 1):
    buffer 0: n x 128 B -> buffer 2: n x 96 B affine H(m)."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    two256 = t.fp_const(1 << 256)
    us = []
    for j in range(2):
        hi = Lin.of(b.inp_bytes(BUF_IN, 64 * j, 32))
        lo = Lin.of(b.inp_bytes(BUF_IN, 64 * j + 32, 32))
        us.append(Lin.of(b.mat(hi * two256 + lo)))
    s = _add_generic_fp(ig, _swu_g1(ig, us[0]), _swu_g1(ig, us[1]))
    e = _isogeny11_projective(ig, s)
    G = ig.G1
    h = G.add(G.mul_fixed(e, X_PARAM), e)  # clearCofactor (index.ts:401-405): [|x|]P + P
    x, y = ig.g1_to_affine(h)
    b.out(x, BUF_OUT, 0)
    b.out(y, BUF_OUT, 1)
    return b


PROGRAMS["hash_to_g1"] = build_hash_to_g1


# ------------------------------------------------------------------------------------------ sign / aggregate
BUF_AUX = 1  # second input buffer (scalars for `sign`, status words for the sums)


def _g2_compress_out(ig: Ingest, p, buf_out=BUF_OUT, buf_flags=BUF_STATUS):
    """toSignature (index.ts:586-598): x.c1 || x.c0 as 2 x 48 B plus a flag word (bit0 = sign flag aflag1,
    bit1 = point at infinity); the host ORs the flags into the top bits of byte 0."""
    b, t = ig.b, ig.t
    is_inf = ig.F2.is_zero(p[2])
    x, y = ig.g2_to_affine(p)
    y1_zero = b.is_zero(y.c1)
    aflag = b.select(y1_zero, Lin.of(b.gt_half(y.c0)), Lin.of(b.gt_half(y.c1)))
    word = b.select(is_inf, Lin.of(b.const_raw(2)), Lin.of(aflag))
    b.out(x.c1, buf_out, 0)
    b.out(x.c0, buf_out, 1)
    b.out_word(word, buf_flags)


def _g1_compress_out(ig: Ingest, p, buf_out=BUF_OUT, buf_flags=BUF_STATUS):
    """PointG1.toHex(true) (index.ts:359-371): x as 48 B plus flag word (bit0 = (y*2)/P, bit1 = infinity)."""
    b = ig.b
    is_inf = b.is_zero(p[2])
    x, y = ig.g1_to_affine(p)
    word = b.select(is_inf, Lin.of(b.const_raw(2)), Lin.of(b.gt_half(y)))
    b.out(x, buf_out, 0)
    b.out_word(word, buf_flags)


SIGN_GLS4 = True  # sign: sk given as four base-|x| digits (api.cu: base_z_digits_kernel), joint ladder over psi^i(H(m))


def build_sign(warps=8) -> Builder:
    """sign(message, privateKey) (index.ts:746-752): sk * H(m), compressed.
    buffer 0: n x 256 B uniform bytes of expand_message_xmd; buffer 1: n x 32 B: the four base-|x| digits of the scalar
    (0 < sk < r), a_3 || a_2 || a_1 || a_0, 8 bytes big-endian each (SIGN_GLS4; otherwise the big-endian scalar itself);
    buffer 2: n x 96 B signature body; buffer 5: n x int32 flag words."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    h = _hash_to_g2_projective(ig, BUF_IN)
    if SIGN_GLS4:
        s = ig.g2_mul_secret_gls4(h, lambda i: b.bit(BUF_AUX, 0, 32, i))
    else:
        s = ig.G2.mul_secret(h, R_ORDER.bit_length(), lambda i: b.bit(BUF_AUX, 0, 32, i))
    _g2_compress_out(ig, s)
    return b


def build_sign_tail(warps=4) -> Builder:
    """sign() behind the SWU kernel: buffer 0: n x 576 B (two points of E' per message), buffers 1, 2, 5 as in `sign`."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    h = _hash_tail_projective(ig, BUF_IN)
    if SIGN_GLS4:
        s = ig.g2_mul_secret_gls4(h, lambda i: b.bit(BUF_AUX, 0, 32, i))
    else:
        s = ig.G2.mul_secret(h, R_ORDER.bit_length(), lambda i: b.bit(BUF_AUX, 0, 32, i))
    _g2_compress_out(ig, s)
    return b


def _lane_sum(ig: Ingest, G: Curve, p):
    """Butterfly sum over the 32 lanes with complete additions (every lane ends with the total)."""
    b, F = ig.b, G.F
    for mask in (1, 2, 4, 8, 16):
        other = []
        for c in p:
            cs = [Lin.of(b.xlane(b.mat(x), mask)) for x in F.coeffs(c)]
            other.append(cs[0] if len(cs) == 1 else E2(cs[0], cs[1]))
        p = G.add(p, tuple(other))
    return p


def _pad_identity(ig: Ingest, G: Curve, p, skip_flag=None):
    """Padding lanes (and lanes flagged `skip_flag`) contribute the identity (0 : 1 : 0)."""
    b, F = ig.b, G.F
    one = b.const(1)
    zero = b.const_raw(0)
    ident_consts = [zero, one, zero]
    out = []
    for c, k in zip(p, ident_consts):
        cs = F.coeffs(c)
        sel = []
        for j, x in enumerate(cs):
            const = k if j == 0 else zero
            v = b.mat(x if not (isinstance(x, Lin) and x.is_zero()) else Lin.of(zero))
            if skip_flag is not None:
                v = b.select(skip_flag, Lin.of(const), Lin.of(v))
            sel.append(Lin.of(b.pad_select(v, const)))
        out.append(sel[0] if len(sel) == 1 else E2(sel[0], sel[1]))
    return tuple(out)


def _build_sum(which: str, level: str, warps: int) -> Builder:
    """Per-batch partial sums.  level 'affine': inputs are affine points (buffer 0) + status words (buffer 1, an
    INFINITY status means 'skip'); level 'proj': inputs are projective partial sums (buffer 0).
    Output (buffer 2, one record per 32-item batch): projective X, Y, Z."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    G = ig.G1 if which == "g1" else ig.G2
    nf = 1 if which == "g1" else 2

    def load(k):
        if nf == 1:
            return Lin.of(b.inp(BUF_IN, k))
        return E2(Lin.of(b.inp(BUF_IN, 2 * k)), Lin.of(b.inp(BUF_IN, 2 * k + 1)))

    if level == "affine":
        one = Lin.of(b.mat(t.fp_const(1)))
        z = one if nf == 1 else E2(one, Lin.of(b.mat(Lin.of(b.const_raw(0)))))
        p = (load(0), load(1), z)
        skip = b.bit(BUF_AUX, 0, 1, 0)  # status word INFINITY (1): point at infinity, contributes nothing
        p = _pad_identity(ig, G, p, skip)
    else:
        p = _pad_identity(ig, G, (load(0), load(1), load(2)))
    s = _lane_sum(ig, G, p)
    k = 0
    for c in s:
        for x in G.F.coeffs(c):
            b.out(x, BUF_OUT, k, per_batch=True)
            k += 1
    return b


def _build_compress(which: str, warps: int) -> Builder:
    """projective (buffer 0) -> compressed body (buffer 2) + flag word (buffer 5)."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    if which == "g1":
        p = tuple(Lin.of(b.inp(BUF_IN, k)) for k in range(3))
        _g1_compress_out(ig, p)
    else:
        p = tuple(E2(Lin.of(b.inp(BUF_IN, 2 * k)), Lin.of(b.inp(BUF_IN, 2 * k + 1))) for k in range(3))
        _g2_compress_out(ig, p)
    return b


PROGRAMS.update({
    "sign": build_sign,
    "sign_tail": build_sign_tail,
    "g1_sum_affine": lambda w: _build_sum("g1", "affine", w),
    "g1_sum_proj": lambda w: _build_sum("g1", "proj", w),
    "g2_sum_affine": lambda w: _build_sum("g2", "affine", w),
    "g2_sum_proj": lambda w: _build_sum("g2", "proj", w),
    "g1_compress": lambda w: _build_compress("g1", w),
    "g2_compress": lambda w: _build_compress("g2", w),
})


# ------------------------------------------------------------------------------------------ validity / scalar mul
def _build_validate(which: str, warps: int) -> Builder:
    """assertValidity (index.ts:383-388 / 633-638) of affine points: buffer 0 affine in, buffer 5 status."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    one = Lin.of(b.mat(t.fp_const(1)))
    if which == "g1":
        p = (Lin.of(b.inp(BUF_IN, 0)), Lin.of(b.inp(BUF_IN, 1)), one)
        G, tors = ig.G1, ig.g1_is_torsion_free
    else:
        z = E2(one, Lin.of(b.mat(Lin.of(b.const_raw(0)))))
        p = (E2(Lin.of(b.inp(BUF_IN, 0)), Lin.of(b.inp(BUF_IN, 1))), E2(Lin.of(b.inp(BUF_IN, 2)), Lin.of(b.inp(BUF_IN, 3))), z)
        G, tors = ig.G2, ig.g2_is_torsion_free
    not_on = b.flag_not(G.is_on_curve(p))
    not_sub = b.flag_not(tors(p))
    st = ig.status_chain([(not_on, ST_NOT_ON_CURVE), (not_sub, ST_NOT_IN_SUBGROUP)])
    b.out_word(st, BUF_STATUS)
    return b


def _build_from_uncompressed(which: str, warps: int) -> Builder:
    """PointG1.fromHex for 96-byte / PointG2.fromHex for 192-byte UNCOMPRESSED input (index.ts:315-325, 565-579): the
    coordinates are taken as they are (no flag bits are cleared, `new Fp` reduces them mod p), then assertValidity.
    buffer 0: the wire bytes (G1: x || y; G2: x.c1 || x.c0 || y.c1 || y.c0); buffer 2: canonical affine coordinates in the
    C-ABI order (G2: x.c0, x.c1, y.c0, y.c1); buffer 5: status.  The infinity / encoding flags of byte 0 are handled by the
    caller (api.cu: a byte-level kernel), which overrides the status word."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    one = Lin.of(b.mat(t.fp_const(1)))
    if which == "g1":
        x, y = Lin.of(b.inp(BUF_IN, 0)), Lin.of(b.inp(BUF_IN, 1))
        p = (x, y, one)
        G, tors = ig.G1, ig.g1_is_torsion_free
        outs = [x, y]
    else:
        x1, x0, y1, y0 = (Lin.of(b.inp(BUF_IN, k)) for k in range(4))
        z = E2(one, Lin.of(b.mat(Lin.of(b.const_raw(0)))))
        p = (E2(x0, x1), E2(y0, y1), z)
        G, tors = ig.G2, ig.g2_is_torsion_free
        outs = [x0, x1, y0, y1]
    not_on = b.flag_not(G.is_on_curve(p))
    not_sub = b.flag_not(tors(p))
    st = ig.status_chain([(not_on, ST_NOT_ON_CURVE), (not_sub, ST_NOT_IN_SUBGROUP)])
    for k, v in enumerate(outs):
        b.out(v, BUF_OUT, k)
    b.out_word(st, BUF_STATUS)
    return b


def build_g2_scalar_mul(warps=4) -> Builder:
    """ProjectivePoint#multiply(scalar) on G2 (math.ts:1061-1078): buffer 0 affine point, buffer 1 32-byte scalar
    (0 < k <= r) -> buffer 2 affine result, buffer 5 flag word (bit1 = result is the point at infinity)."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    one = Lin.of(b.mat(t.fp_const(1)))
    z = E2(one, Lin.of(b.mat(Lin.of(b.const_raw(0)))))
    p = (E2(Lin.of(b.inp(BUF_IN, 0)), Lin.of(b.inp(BUF_IN, 1))), E2(Lin.of(b.inp(BUF_IN, 2)), Lin.of(b.inp(BUF_IN, 3))), z)
    s = ig.G2.mul_secret(p, R_ORDER.bit_length(), lambda i: b.bit(BUF_AUX, 0, 32, i))
    is_inf = ig.F2.is_zero(s[2])
    x, y = ig.g2_to_affine(s)
    b.out(x.c0, BUF_OUT, 0)
    b.out(x.c1, BUF_OUT, 1)
    b.out(y.c0, BUF_OUT, 2)
    b.out(y.c1, BUF_OUT, 3)
    b.out_word(b.select(is_inf, Lin.of(b.const_raw(2)), Lin.of(b.const_raw(0))), BUF_STATUS)
    return b


def build_g1_scalar_mul(warps=4) -> Builder:
    """ProjectivePoint#multiply / multiplyUnsafe on G1 (math.ts:1048-1078): buffer 0 affine point, buffer 1 32-byte
    scalar -> buffer 2 affine result, buffer 5 flag word (bit1 = point at infinity)."""
    b = Builder(warps)
    t = Tower(b)
    ig = Ingest(t)
    one = Lin.of(b.mat(t.fp_const(1)))
    p = (Lin.of(b.inp(BUF_IN, 0)), Lin.of(b.inp(BUF_IN, 1)), one)
    s = ig.G1.mul_secret(p, R_ORDER.bit_length(), lambda i: b.bit(BUF_AUX, 0, 32, i))
    is_inf = b.is_zero(s[2])
    x, y = ig.g1_to_affine(s)
    b.out(x, BUF_OUT, 0)
    b.out(y, BUF_OUT, 1)
    b.out_word(b.select(is_inf, Lin.of(b.const_raw(2)), Lin.of(b.const_raw(0))), BUF_STATUS)
    return b


PROGRAMS.update({
    "g1_from_uncompressed": lambda w: _build_from_uncompressed("g1", w),
    "g2_from_uncompressed": lambda w: _build_from_uncompressed("g2", w),
    "g1_scalar_mul": build_g1_scalar_mul,
    "g1_validate": lambda w: _build_validate("g1", w),
    "g2_validate": lambda w: _build_validate("g2", w),
    "g2_scalar_mul": build_g2_scalar_mul,
})
