"""Fp2/Fp6/Fp12 tower, Miller loop and final exponentiation expressed as lazy tower-VM expressions.

Mathematically this follows the reference (math.ts:403-885 tower, :1331-1388 line functions + Miller loop,
:856-874 final exponentiation = f^(3(p^12-1)/r)); all results are canonical residues, so they are
bit-identical to the reference's even though the *evaluation strategy* is different:

  * products are never reduced one by one: every output Fp coefficient is ONE multiply-accumulate
    micro-op over all the Fp products that feed it (schoolbook over the tower, one Montgomery reduction);
  * additions, subtractions, negations, small scalings, conjugations and multiplications by the
    non-residues xi = 1+u and v are free (they are folded into operand modes / term signs);
  * sparse operands (the line 014 element) need no special case: zero coefficients drop out of the sums.
"""
from __future__ import annotations

from . import builder as _builder
from .builder import Builder, Lin, Quad, Val, P

X_PARAM = 0xD201000000010000  # |x| (math.ts:48); the curve parameter is negative


def _zero():
    return Lin()


def _mat(b: Builder, e):
    """Materialise unless the expression is structurally zero (kept symbolic so sparsity survives)."""
    return b.mat_lin(e)


KARATSUBA_FP2 = True  # Fp2 products as three Fp products that feed BOTH coefficients (two-output records)
SHARE = True          # switched off around latency-critical code: a shared product serialises the two coefficients on one warp
SHARE_POLICY = __import__("os").environ.get("BLS381_VM_SHARE", "all")  # all | miller | none  (tuning)
if SHARE_POLICY == "none":
    SHARE = False
    _builder.SHARE_ENABLED = False


class no_sharing:
    """Context manager: Fp2 products built inside are expanded without shared products (one record per coefficient)."""

    def __enter__(self):
        global SHARE
        self.prev, SHARE = SHARE, False
        _builder.SHARE_ENABLED = False

    def __exit__(self, *a):
        global SHARE
        SHARE = self.prev
        _builder.SHARE_ENABLED = self.prev



class E2:
    """Lazy Fp2 element  sum_i (k0_i + k1_i u) X_i Y_i + (l0 + l1 u):  products of Fp expressions with a weight for each
    of the two coefficients, plus a linear part.  A product with k0 != 0 and k1 != 0 is SHARED by the two coefficients: it
    is computed once by a two-output record (Builder.mat2).  `E2(c0, c1)` accepts Lin or Quad coefficients; `.c0` / `.c1`
    give the coefficients back (Lin when the element is materialised / linear, else a Quad)."""

    __slots__ = ("terms", "l0", "l1")

    def __init__(self, c0=None, c1=None, terms=None):
        self.terms = list(terms) if terms else []
        self.l0, self.l1 = Lin(), Lin()
        for idx, c in ((0, c0), (1, c1)):
            if c is None:
                continue
            if isinstance(c, Quad):
                for k, x, y in c.terms:
                    self.terms.append((k, 0, x, y) if idx == 0 else (0, k, x, y))
                lin = c.lin
            else:
                lin = Lin.of(c)
            if idx == 0:
                self.l0 = lin
            else:
                self.l1 = lin

    @staticmethod
    def _raw(terms, l0, l1):
        e = E2()
        e.terms = [t for t in terms if (t[0] or t[1])]
        e.l0, e.l1 = l0, l1
        return e

    @property
    def c0(self):
        t = [(k0, x, y) for k0, _, x, y in self.terms if k0]
        return Quad(t, self.l0) if t else self.l0

    @property
    def c1(self):
        t = [(k1, x, y) for _, k1, x, y in self.terms if k1]
        return Quad(t, self.l1) if t else self.l1

    def is_lin(self):
        return not self.terms

    def __add__(self, o):
        return E2._raw(self.terms + o.terms, self.l0 + o.l0, self.l1 + o.l1)

    def __neg__(self):
        return E2._raw([(-k0, -k1, x, y) for k0, k1, x, y in self.terms], -self.l0, -self.l1)

    def __sub__(self, o):
        return self + (-o)

    def scale(self, k: int):
        return E2._raw([(k0 * k, k1 * k, x, y) for k0, k1, x, y in self.terms], self.l0 * k, self.l1 * k)

    def conj(self):
        return E2._raw([(k0, -k1, x, y) for k0, k1, x, y in self.terms], self.l0, -self.l1)

    def mul_xi(self):  # * (1 + u)   (math.ts:471-475)
        return E2._raw([(k0 - k1, k0 + k1, x, y) for k0, k1, x, y in self.terms], self.l0 - self.l1, self.l0 + self.l1)

    def _need_lin(self):
        if self.terms:
            raise TypeError("cannot multiply an unmaterialised Fp2 product; call .m(builder) first")

    def __mul__(self, o):
        if isinstance(o, int):
            return self.scale(o)
        self._need_lin()
        a0, a1 = self.l0, self.l1
        if isinstance(o, E2):  # (math.ts:451-462)
            o._need_lin()
            b0, b1 = o.l0, o.l1
            if KARATSUBA_FP2 and SHARE and SHARE_POLICY != "none" and _builder.CURRENT_DUAL and not (a0.is_zero() or a1.is_zero() or b0.is_zero() or b1.is_zero()):
                # c0 = a0 b0 - a1 b1 ; c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1: three products, two of them shared
                return E2._raw([(1, -1, a0, b0), (-1, -1, a1, b1), (0, 1, a0 + a1, b0 + b1)], Lin(), Lin())
            t = [(1, 0, a0, b0), (-1, 0, a1, b1), (0, 1, a0, b1), (0, 1, a1, b0)]
            return E2._raw([(k0, k1, x, y) for k0, k1, x, y in t if not x.is_zero() and not y.is_zero()], Lin(), Lin())
        o = Lin.of(o)  # by an Fp expression
        return E2._raw([(1, 0, a0, o), (0, 1, a1, o)], Lin(), Lin())

    def sqr(self):  # (math.ts:477-484)
        self._need_lin()
        a0, a1 = self.l0, self.l1
        return E2._raw([(1, 0, a0 + a1, a0 - a1), (0, 2, a0, a1)], Lin(), Lin())

    def m(self, b: Builder):
        v0, v1 = b.mat2(self.terms, self.l0, self.l1)
        return E2(v0, v1)

    def is_zero(self):
        return not self.terms and self.l0.is_zero() and self.l1.is_zero()


def e2_zero():
    return E2(_zero(), _zero())


class E6:
    """Fp6 = Fp2[v]/(v^3 - xi)  (math.ts:554)."""

    __slots__ = ("c0", "c1", "c2")

    def __init__(self, c0, c1, c2):
        self.c0, self.c1, self.c2 = c0, c1, c2

    def __add__(self, o):
        return E6(self.c0 + o.c0, self.c1 + o.c1, self.c2 + o.c2)

    def __sub__(self, o):
        return E6(self.c0 - o.c0, self.c1 - o.c1, self.c2 - o.c2)

    def __neg__(self):
        return E6(-self.c0, -self.c1, -self.c2)

    def scale(self, k):
        return E6(self.c0.scale(k), self.c1.scale(k), self.c2.scale(k))

    def mul_v(self):  # (math.ts:627-629)
        return E6(self.c2.mul_xi(), self.c0, self.c1)

    def __mul__(self, o):  # schoolbook form of math.ts:601-618; xi is applied to an OPERAND (free)
        a0, a1, a2 = self.c0, self.c1, self.c2
        b0, b1, b2 = o.c0, o.c1, o.c2
        xb1, xb2 = b1.mul_xi(), b2.mul_xi()
        return E6(
            a0 * b0 + a1 * xb2 + a2 * xb1,
            a0 * b1 + a1 * b0 + a2 * xb2,
            a0 * b2 + a1 * b1 + a2 * b0,
        )

    def mul_e2(self, k: E2):
        return E6(self.c0 * k, self.c1 * k, self.c2 * k)

    def sqr(self):  # symmetric schoolbook of math.ts:658-670
        a0, a1, a2 = self.c0, self.c1, self.c2
        return E6(
            a0.sqr() + (a1.scale(2) * a2.mul_xi()),
            (a0.scale(2) * a1) + a2.mul_xi() * a2,
            (a0.scale(2) * a2) + a1.sqr(),
        )

    def m(self, b):
        return E6(self.c0.m(b), self.c1.m(b), self.c2.m(b))


FOLD_T2 = __import__("os").environ.get("BLS381_VM_FOLDT2", "1") != "0"  # 12 xi Rz^2 of the doubling step in one record pair
SCALED_MILLER = __import__("os").environ.get("BLS381_VM_SCALED", "1") != "0"  # 4R doubling step inside `pairing`
EASY_PART_SQR = __import__("os").environ.get("BLS381_VM_EASYSQR", "1") != "0"  # f^(p^6 - 1) as conj(f)^2 / norm
SYMMETRIC_SQR = __import__("os").environ.get("BLS381_VM_SYMSQR", "1") != "0"  # Fp12 squarings of the Miller loop


class E12:
    """Fp12 = Fp6[w]/(w^2 - v)  (math.ts:705)."""

    __slots__ = ("c0", "c1")

    def __init__(self, c0, c1):
        self.c0, self.c1 = c0, c1

    def __mul__(self, o):  # schoolbook form of math.ts:748-759
        return E12(self.c0 * o.c0 + self.c1 * o.c1.mul_v(), self.c0 * o.c1 + self.c1 * o.c0)

    def sqr(self):  # (a0 + a1 w)^2 = a0^2 + v a1^2 + 2 a0 a1 w      (math.ts:783-791)
        a0, a1 = self.c0, self.c1
        if not SYMMETRIC_SQR:
            return E12(a0.sqr() + a1.mul_v() * a1, a0.scale(2) * a1)
        # v a1^2 written as a symmetric square (22 Fp products instead of the 36 of a general Fp6 product):
        # a1^2 = (s0, s1, s2) as in E6.sqr, v (s0, s1, s2) = (xi s2, s0, s1), xi applied to an operand
        b0, b1, b2 = a1.c0, a1.c1, a1.c2
        xb2 = b2.mul_xi()
        va1sq = E6(b0.scale(2) * xb2 + b1 * b1.mul_xi(), b0.sqr() + b1.scale(2) * xb2, b0.scale(2) * b1 + xb2 * b2)
        return E12(a0.sqr() + va1sq, a0.scale(2) * a1)

    def conj(self):  # math.ts:799-801
        return E12(self.c0, -self.c1)

    def m(self, b):
        return E12(self.c0.m(b), self.c1.m(b))

    def coeffs(self):
        return [self.c0.c0, self.c0.c1, self.c0.c2, self.c1.c0, self.c1.c1, self.c1.c2]

    def flat(self):
        return [c for e2 in self.coeffs() for c in (e2.c0, e2.c1)]


# ------------------------------------------------------------------------------------------ constants
def _fp2_pow(a, n):
    r = (1, 0)
    while n:
        if n & 1:
            r = _fp2_mul(r, a)
        a = _fp2_mul(a, a)
        n >>= 1
    return r


def _fp2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


class Tower:
    """Holds a Builder plus the field constants the programs need."""

    def __init__(self, b: Builder):
        self.b = b
        xi = (1, 1)
        # Frobenius coefficients, derived from xi (equal to the literals of math.ts:1454-1543)
        self.frob6_c1 = [_fp2_pow(xi, (P**k - 1) // 3) for k in range(6)]
        self.frob6_c2 = [_fp2_pow(xi, 2 * (P**k - 1) // 3) for k in range(6)]
        self.frob12 = [_fp2_pow(xi, (P**k - 1) // 6) for k in range(12)]

    def fp_const(self, v: int):
        v %= P
        if v == 0:
            return Lin()
        return Lin.of(self.b.const(v))

    def e2_const(self, pair):
        return E2(self.fp_const(pair[0]), self.fp_const(pair[1]))

    def one12(self):
        one = E2(self.fp_const(1), _zero())
        z = e2_zero
        return E12(E6(one, z(), z()), E6(z(), z(), z()))

    # ---- Frobenius (math.ts:529-531, 682-688, 804-809) ------------------------------------------
    def frob2(self, a: E2, power: int):
        return a.conj() if power % 2 else a

    def mul_const2(self, a: E2, pair) -> E2:
        pair = (pair[0] % P, pair[1] % P)
        if pair == (1, 0):
            return a
        if pair == (P - 1, 0):
            return -a
        return a * self.e2_const(pair)

    def frob6(self, a: E6, power: int):
        return E6(
            self.frob2(a.c0, power),
            self.mul_const2(self.frob2(a.c1, power), self.frob6_c1[power % 6]),
            self.mul_const2(self.frob2(a.c2, power), self.frob6_c2[power % 6]),
        )

    def frob12_map(self, a: E12, power: int) -> E12:
        """a^(p^power) for a materialised `a`; power 6 is the conjugation (frob coefficient -1)."""
        if power % 12 == 6:
            return a.conj()
        r0 = self.frob6(a.c0, power)
        t = self.frob6(a.c1, power).m(self.b)
        k = self.frob12[power % 12]
        return E12(r0, E6(self.mul_const2(t.c0, k), self.mul_const2(t.c1, k), self.mul_const2(t.c2, k))).m(self.b)

    def mul12(self, x: E12, y: E12) -> E12:
        """dense Fp12 product of materialised operands, materialised"""
        # (the reference's Karatsuba step over Fp6 on reduced intermediates -- 108 products in 18 records + linear records
        # instead of 144 in 12 -- was costed with the scheduler's model: same total cost, makespan -0.9 %, three times the
        # far slots; not used)
        return (x * y).m(self.b)

    # ---- cyclotomic square (math.ts:811-843) -----------------------------------------------------
    @staticmethod
    def _fp4_square(a: E2, b: E2):
        # first = a^2 + xi b^2 ; second = (a+b)^2 - a^2 - b^2 = 2ab
        return a.sqr() + b.sqr().mul_xi(), a.scale(2) * b

    def cyclotomic_square(self, f: E12) -> E12:
        c0c0, c0c1, c0c2 = f.c0.c0, f.c0.c1, f.c0.c2
        c1c0, c1c1, c1c2 = f.c1.c0, f.c1.c1, f.c1.c2
        t3, t4 = self._fp4_square(c0c0, c1c1)
        t5, t6 = self._fp4_square(c1c0, c0c2)
        t7, t8 = self._fp4_square(c0c1, c1c2)
        t9 = t8.mul_xi()
        return E12(
            E6(t3.scale(3) - c0c0.scale(2), t5.scale(3) - c0c1.scale(2), t7.scale(3) - c0c2.scale(2)),
            E6(t9.scale(3) + c1c0.scale(2), t4.scale(3) + c1c1.scale(2), t6.scale(3) + c1c2.scale(2)),
        )

    def cyclotomic_exp(self, f: E12) -> E12:
        """f^|x| exactly as math.ts:845-852 (the leading squarings of ONE are the identity)."""
        b = self.b
        z = f
        for i in range(X_PARAM.bit_length() - 2, -1, -1):
            z = self.cyclotomic_square(z).m(b)
            if (X_PARAM >> i) & 1:
                z = self.mul12(z, f)
        return z

    # ---- inversion -------------------------------------------------------------------------------
    def fp_inv(self, a: Lin) -> Lin:
        """1/a in one micro-op (fixed-iteration binary GCD, csrc/fp_inv.cuh).  Any algorithm yields the same
        canonical residue as the reference's extended Euclid (math.ts:134-156)."""
        return Lin.of(self.b.inv(a))

    def fp2_inv(self, a: E2) -> E2:  # math.ts:522-526
        b = self.b
        a = a.m(b)
        norm = a.c0 * a.c0 + a.c1 * a.c1
        f = self.fp_inv(Lin.of(b.mat(norm)))
        return E2(f * a.c0, (-f) * a.c1)

    def fp6_inv(self, a: E6) -> E6:  # math.ts:672-680
        b = self.b
        c0, c1, c2 = a.c0, a.c1, a.c2
        t0 = (c0.sqr() - c2.mul_xi() * c1).m(b)
        t1 = (c2.sqr().mul_xi() - c0 * c1).m(b)
        t2 = (c1.sqr() - c0 * c2).m(b)
        d = ((c2 * t1 + c1 * t2).mul_xi() + c0 * t0).m(b)
        t4 = self.fp2_inv(d).m(b)
        return E6(t4 * t0, t4 * t1, t4 * t2)

    def fp12_inv(self, a: E12) -> E12:  # math.ts:793-797
        b = self.b
        t = self.fp6_inv((a.c0.sqr() - a.c1.mul_v() * a.c1).m(b)).m(b)
        return E12(a.c0 * t, -(a.c1 * t))

    # ---- final exponentiation (math.ts:856-874) ---------------------------------------------------
    def final_exponentiate(self, f: E12) -> E12:
        if SHARE_POLICY == "miller" and SHARE:
            with no_sharing():
                return self.final_exponentiate(f)
        b = self.b
        f = f.m(b)
        if EASY_PART_SQR:
            # f^(p^6) / f = conj(f) * conj(f) / (f conj(f)) = conj(f)^2 / N,  N = a0^2 - v a1^2 in Fp6 (math.ts:793-797):
            # one Fp12 squaring and one Fp12-by-Fp6 product (148 Fp products) instead of the inverse and a full Fp12
            # product (216); the squaring does not wait for the inversion
            ninv = self.fp6_inv((f.c0.sqr() - f.c1.mul_v() * f.c1).m(b)).m(b)
            cf2 = f.conj().sqr().m(b)
            t0 = E12(cf2.c0 * ninv, cf2.c1 * ninv).m(b)
        else:
            t0 = (self.frob12_map(f, 6) * self.fp12_inv(f).m(b)).m(b)  # f^(p^6) / f
        t1 = (self.frob12_map(t0, 2) * t0).m(b)
        t2 = self.cyclotomic_exp(t1).conj()
        t3 = self.mul12(self.cyclotomic_square(t1).m(b).conj(), t2)
        t4 = self.cyclotomic_exp(t3).conj()
        t5 = self.cyclotomic_exp(t4).conj()
        t6 = self.mul12(self.cyclotomic_exp(t5).conj(), self.cyclotomic_square(t2).m(b))
        t7 = self.cyclotomic_exp(t6).conj()
        t2_t5_pow_q2 = self.frob12_map(self.mul12(t2, t5), 2)
        t4_t1_pow_q3 = self.frob12_map(self.mul12(t4, t1), 3)
        t6_t1c_pow_q1 = self.frob12_map(self.mul12(t6, t1.conj()), 1)
        t7_t3c_t1 = self.mul12(self.mul12(t7, t3.conj()), t1)
        return self.mul12(self.mul12(self.mul12(t2_t5_pow_q2, t4_t1_pow_q3), t6_t1c_pow_q1), t7_t3c_t1)

    # ---- Miller loop with fused line evaluation (math.ts:1331-1388) --------------------------------
    def miller_loop(self, Px: Lin, Py: Lin, Qx: E2, Qy: E2, scaled=False) -> E12:
        return self.miller_loop_multi([(Px, Py, Qx, Qy)], scaled)

    def miller_loop_multi(self, pairs, scaled=False) -> E12:
        """prod_k millerLoop(P_k, Q_k) for the pairs [(Px, Py, Qx, Qy), ...] held by ONE lane, with the Fp12 squaring of
        every iteration shared by all pairs:  f <- f^2 * l_1 * l_2 * ...  Because (f_a f_b)^2 l_a l_b = (f_a^2 l_a)(f_b^2 l_b),
        the result is exactly the product of the individual Miller loops (math.ts:1373-1388 each, conjugated) -- the
        same canonical bytes as index.ts:815's product, with one squaring per bit instead of one per pair.

        scaled=True (only where the final exponentiation follows in the same program): the doubling step keeps 4R instead
        of R (no multiplications by 1/2, math.ts:1349-1350).  Every formula below is homogeneous in (Rx, Ry, Rz), so each
        line picks up a power of 4: the Miller value changes by a factor in Fp*, which the (p^6 - 1) part of the final
        exponentiation maps to one -- pairing(P, Q) keeps its bytes, pairing(P, Q, false) must not use this form."""
        b = self.b
        inv2 = self.fp_const(pow(2, -1, P))
        c12 = self.fp_const(12)
        R = [(Qx, Qy, E2(self.fp_const(1), _zero())) for (_, _, Qx, Qy) in pairs]
        f = None
        z2 = e2_zero

        def line_mul(f, o0: E2, o1: E2, o4: E2):
            # f * (o0 + o1 v + o4 v w)  == multiplyBy014 (math.ts:768-777); o1, o4 already scaled by Px, Py
            ell = E12(E6(o0, o1, z2()), E6(z2(), o4, z2()))
            if f is None:
                return ell  # ONE * ell
            return (f * ell).m(b)

        for i in range(X_PARAM.bit_length() - 2, -1, -1):
            for k, (Px, Py, Qx, Qy) in enumerate(pairs):
                Rx, Ry, Rz = R[k]
                # ---- doubling step (math.ts:1339-1351)
                t0 = Ry.sqr().m(b)
                if not FOLD_T2:
                    t1 = Rz.sqr().m(b)
                t4 = (Ry.scale(2) * Rz).m(b)  # (Ry+Rz)^2 - t1 - t0
                rx2 = Rx.sqr().m(b)
                rxry = (Rx * Ry).m(b)
                if FOLD_T2:
                    t2 = Rz.sqr().mul_xi().scale(12).m(b)  # 3 * Rz^2 * 4(1+u) as ONE pair of records
                else:
                    t2 = (t1.mul_xi() * c12).m(b)  # 3 * t1 * 4(1+u)
                o0 = t2 - t0
                o1 = (rx2.scale(3) * Px).m(b)
                o4 = ((-t4) * Py).m(b)
                if scaled:
                    g2 = t0 - t2.scale(3)  # 2 (t0 - t3)/2, used as a two-slot operand
                    h2 = (t0 + t2.scale(3)).m(b)  # linear record, no product
                    Rx = (g2 * rxry).scale(2).m(b)
                    Ry = (h2.sqr() - t2.sqr().scale(12)).m(b)
                    Rz = (t0 * t4).scale(4).m(b)
                else:
                    g = ((t0 - t2.scale(3)) * inv2).m(b)  # (t0 - t3)/2
                    h = ((t0 + t2.scale(3)) * inv2).m(b)  # (t0 + t3)/2
                    Rx = (g * rxry).m(b)
                    Ry = (h.sqr() - t2.sqr().scale(3)).m(b)
                    Rz = (t0 * t4).m(b)
                f = line_mul(f, o0, o1, o4)
                if (X_PARAM >> i) & 1:
                    # ---- addition step (math.ts:1354-1367)
                    u0 = (Ry - Qy * Rz).m(b)
                    u1 = (Rx - Qx * Rz).m(b)
                    o0 = (u0 * Qx - u1 * Qy).m(b)
                    o1 = ((-u0) * Px).m(b)
                    o4 = (u1 * Py).m(b)
                    v2 = u1.sqr().m(b)
                    v3 = (v2 * u1).m(b)
                    v4 = (v2 * Rx).m(b)
                    w0 = u0.sqr().m(b)
                    v5 = (v3 - v4.scale(2) + w0 * Rz).m(b)
                    Rx_n = (u1 * v5).m(b)
                    Ry_n = ((v4 - v5) * u0 - v3 * Ry).m(b)
                    Rz = (Rz * v3).m(b)
                    Rx, Ry = Rx_n, Ry_n
                    f = line_mul(f, o0, o1, o4)
                R[k] = (Rx, Ry, Rz)
            if i != 0:
                f = f.sqr().m(b)
        return f.conj()


# ------------------------------------------------------------------------------------------ programs
# Buffer ids of the C ABI (include/bls381_b200.h): 0 = G1 affine (x, y), 1 = G2 affine (x.c0, x.c1, y.c0,
# y.c1), 2 = Fp12 output (12 fields in the flat order of math.ts:709-714), 3 = Fp12 input.
BUF_G1, BUF_G2, BUF_OUT, BUF_F12IN = 0, 1, 2, 3


def _load_g1_g2(b: Builder):
    Px = Lin.of(b.inp(BUF_G1, 0))
    Py = Lin.of(b.inp(BUF_G1, 1))
    Qx = E2(Lin.of(b.inp(BUF_G2, 0)), Lin.of(b.inp(BUF_G2, 1)))
    Qy = E2(Lin.of(b.inp(BUF_G2, 2)), Lin.of(b.inp(BUF_G2, 3)))
    return Px, Py, Qx, Qy


def _store_f12(b: Builder, f: E12, buf=BUF_OUT):
    for k, c in enumerate(f.flat()):
        b.out(c, buf, k)


def _load_f12(b: Builder, buf=BUF_F12IN) -> E12:
    c = [Lin.of(b.inp(buf, k)) for k in range(12)]
    e2 = [E2(c[2 * i], c[2 * i + 1]) for i in range(6)]
    return E12(E6(e2[0], e2[1], e2[2]), E6(e2[3], e2[4], e2[5]))


def build_pairing(warps=6, with_final_exp=True) -> Builder:
    """pairing(P, Q, withFinalExponent) for affine, valid, non-infinity inputs (index.ts:715-722)."""
    b = Builder(warps)
    t = Tower(b)
    Px, Py, Qx, Qy = _load_g1_g2(b)
    f = t.miller_loop(Px, Py, Qx, Qy, scaled=with_final_exp and SCALED_MILLER)
    if with_final_exp:
        f = t.final_exponentiate(f)
    _store_f12(b, f)
    return b


def build_final_exp(warps=6) -> Builder:
    """Fp12#finalExponentiate (math.ts:856-874) on wire-format Fp12 inputs."""
    b = Builder(warps)
    t = Tower(b)
    f = t.final_exponentiate(_load_f12(b))
    _store_f12(b, f)
    return b


def build_fp12_mul_test(warps=6) -> Builder:
    """out = a*b, a from BUF_F12IN, b from BUF_OUT... (test program): a in buffer 3, b in buffer 4."""
    b = Builder(warps)
    a = _load_f12(b, 3)
    c = _load_f12(b, 4)
    _store_f12(b, (a * c), BUF_OUT)
    return b


def _lane_product(b: Builder, t: Tower, f: E12) -> E12:
    """Butterfly product over the 32 lanes of the CTA (every lane ends with the product of all lanes).
    Exact in Fp12, so the result is independent of the combination order (index.ts:815)."""
    for mask in (1, 2, 4, 8, 16):
        f = f.m(b)
        other = E12(*[E6(*[E2(Lin.of(b.xlane(b.mat(e2.c0), mask)), Lin.of(b.xlane(b.mat(e2.c1), mask)))
                           for e2 in (e6.c0, e6.c1, e6.c2)]) for e6 in (f.c0, f.c1)])
        f = (f * other).m(b)
    return f


def _pad_one(b: Builder, t: Tower, f: E12) -> E12:
    """Replace the value of padding lanes (items beyond n_items) by Fp12.ONE."""
    one = b.const(1)
    zero = b.const_raw(0)
    flat = f.m(b).flat()
    sel = [Lin.of(b.pad_select(c if not (isinstance(c, Lin) and c.is_zero()) else Lin.of(zero), one if k == 0 else zero))
           for k, c in enumerate(flat)]
    e2 = [E2(sel[2 * i], sel[2 * i + 1]) for i in range(6)]
    return E12(E6(e2[0], e2[1], e2[2]), E6(e2[3], e2[4], e2[5]))


def build_miller_product(warps=6) -> Builder:
    """One Fp12 per 32-item batch: prod_lanes millerLoop(P_i, Q_i) (padding lanes contribute ONE)."""
    b = Builder(warps)
    t = Tower(b)
    Px, Py, Qx, Qy = _load_g1_g2(b)
    f = t.miller_loop(Px, Py, Qx, Qy)
    f = _lane_product(b, t, _pad_one(b, t, f))
    for k, c in enumerate(f.flat()):
        b.out(c, BUF_OUT, k, per_batch=True)
    return b


def build_miller_product2(warps=6, per_lane=2) -> Builder:
    """One Fp12 per 32-lane batch, `per_lane` pairs per lane: lane l of a batch handles the consecutive items
    per_lane*l .. per_lane*l + per_lane - 1 (the caller passes multiplied strides: per_lane x 96 B of G1 and
    per_lane x 192 B of G2 per lane) and shares the Fp12 squarings between them."""
    b = Builder(warps)
    t = Tower(b)
    pairs = []
    for k in range(per_lane):
        Px = Lin.of(b.inp(BUF_G1, 2 * k))
        Py = Lin.of(b.inp(BUF_G1, 2 * k + 1))
        Qx = E2(Lin.of(b.inp(BUF_G2, 4 * k)), Lin.of(b.inp(BUF_G2, 4 * k + 1)))
        Qy = E2(Lin.of(b.inp(BUF_G2, 4 * k + 2)), Lin.of(b.inp(BUF_G2, 4 * k + 3)))
        pairs.append((Px, Py, Qx, Qy))
    f = t.miller_loop_multi(pairs)
    f = _lane_product(b, t, _pad_one(b, t, f))
    for k, c in enumerate(f.flat()):
        b.out(c, BUF_OUT, k, per_batch=True)
    return b


def build_f12_product(warps=6) -> Builder:
    """One Fp12 per 32-item batch: product of the batch's Fp12 inputs (second level of the tree)."""
    b = Builder(warps)
    t = Tower(b)
    f = _lane_product(b, t, _pad_one(b, t, _load_f12(b)))
    for k, c in enumerate(f.flat()):
        b.out(c, BUF_OUT, k, per_batch=True)
    return b


PROGRAMS = {
    "pairing": lambda w: build_pairing(w, True),
    "miller": lambda w: build_pairing(w, False),
    "final_exp": build_final_exp,
    "miller_product": build_miller_product,
    "miller_product2": build_miller_product2,
    "miller_product3": lambda w: build_miller_product2(w, 3),
    "miller_product4": lambda w: build_miller_product2(w, 4),
    "f12_product": build_f12_product,
    "f12_mul_test": build_fp12_mul_test,
}
