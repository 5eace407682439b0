"""Tower-VM program builder: symbolic Fp algebra -> scheduled, slot-allocated, encoded micro-op streams.

The device executes ONE kind of micro-op (csrc/vm.cuh):

    dst = MontRed( sum_t X_t * Y_t ) + sum_e Z_e        (mod p, canonical)

with every operand a small signed combination cA*A + cB*B of slots/constants.  This module provides

  * `Lin`   - a lazy linear combination of materialised values (free: folded into operand modes),
  * `Quad`  - a lazy sum of products of Lins plus a Lin (free until materialised),
  * `Builder.mat(q)` - materialise a Quad as ONE micro-op (lazy reduction: one Montgomery reduction
     per output, regardless of how many products feed it), splitting when bounds/sizes require,
  * a list scheduler that packs independent micro-ops of a dependency level onto the W warps of a CTA,
  * a step-accurate slot allocator (shared-memory slots, overflow into "far" global-scratch slots),
  * the binary encoder for the 128-byte records.

Nothing here touches the oracle; the tests compare VM results against it.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 1 << 384
R_MOD_P = R % P
R2_MOD_P = R * R % P
P_OVER_R = P / R  # ~0.1016

OP_NOP, OP_MAC, OP_SEL, OP_BIT, OP_INV, OP_MAC2 = 0, 1, 2, 3, 4, 5
F_CONST, F_GLOBAL, F_XLANE, F_SIMPLE, F_FUSE = 1, 2, 4, 8, 0x80
H_SINGLE = 1 << 24
H_BAR, H_DSTG, H_DSTWORD, H_DSTBATCH, H_PADCONST, H_POST_ISZERO, H_POST_GTHALF = 1 << 25, 1 << 26, 1 << 27, 1 << 28, 1 << 29, 1 << 30, 1 << 31
REC_WORDS = 64
WAIT_WORD = 60  # words 60..63: progress requirements
MAX_TERMS = 12
MAX_ENTRIES2 = 24  # product entries of a two-output record (a fused entry pair counts as two)
MAX_KP2 = 63  # largest K of the device table of K*p^2 (csrc/fp_core_gen.cuh: kKP2)
MAX_SUM_BOUND = 80.0  # sum of |x|*|y| bounds in units of p^2 that fits the 768-bit accumulator
COST_MODEL = "v2"  # "v1" = the instruction-count estimate used before the trace calibration
MAX_K = 9.5  # result bound (units of p): must stay below 2^384 / p = 9.84; `correct` handles up to 4 rounds


# ------------------------------------------------------------------------------------------ values
class Val:
    """A materialised Fp value (SSA).  kind: 'op' (result of a micro-op), 'const', 'xlane'."""

    __slots__ = ("id", "kind", "op", "cidx", "name")

    def __init__(self, id, kind, op=None, cidx=None, name=None):
        self.id, self.kind, self.op, self.cidx, self.name = id, kind, op, cidx, name

    def __repr__(self):
        return f"v{self.id}" if self.kind == "op" else f"c{self.cidx}"


class Lin:
    """sum coef[v] * v over Vals (small integer coefficients)."""

    __slots__ = ("t",)

    def __init__(self, t=None):
        self.t = {k: c for k, c in (t or {}).items() if c != 0}

    @staticmethod
    def of(v):
        if isinstance(v, Lin):
            return v
        if isinstance(v, Val):
            return Lin({v: 1})
        raise TypeError(v)

    def __add__(self, o):
        if isinstance(o, Quad):
            return o + self
        o = Lin.of(o)
        t = dict(self.t)
        for k, c in o.t.items():
            t[k] = t.get(k, 0) + c
        return Lin(t)

    def __neg__(self):
        return Lin({k: -c for k, c in self.t.items()})

    def __sub__(self, o):
        if isinstance(o, Quad):
            return (-o) + self
        return self + (-Lin.of(o))

    def scale(self, k: int):
        return Lin({v: c * k for v, c in self.t.items()})

    def __mul__(self, o):
        if isinstance(o, int):
            return self.scale(o)
        if isinstance(o, Quad):
            raise TypeError("cannot multiply by an unmaterialised product")
        return Quad([(1, self, Lin.of(o))], Lin())

    __rmul__ = __mul__

    def is_zero(self):
        return not self.t

    def bound(self):
        return sum(abs(c) for c in self.t.values())


class Quad:
    """sum_i k_i * X_i * Y_i  +  L   (k_i: int, X_i, Y_i, L : Lin)."""

    __slots__ = ("terms", "lin")

    def __init__(self, terms, lin):
        self.terms = [(k, x, y) for k, x, y in terms if k != 0 and not x.is_zero() and not y.is_zero()]
        self.lin = lin

    def __add__(self, o):
        if isinstance(o, Quad):
            return Quad(self.terms + o.terms, self.lin + o.lin)
        return Quad(self.terms, self.lin + Lin.of(o))

    __radd__ = __add__

    def __neg__(self):
        return Quad([(-k, x, y) for k, x, y in self.terms], -self.lin)

    def __sub__(self, o):
        return self + (-o)

    def scale(self, k: int):
        return Quad([(k * kk, x, y) for kk, x, y in self.terms], self.lin.scale(k))

    def __mul__(self, o):
        if isinstance(o, int):
            return self.scale(o)
        raise TypeError("cannot multiply an unmaterialised product; call Builder.mat() first")

    __rmul__ = __mul__


def as_expr(x):
    if isinstance(x, (Lin, Quad)):
        return x
    return Lin.of(x)


# ------------------------------------------------------------------------------------------ ops
@dataclass
class Operand:
    a: Val
    ca: int
    b: Val | None = None
    cb: int = 0
    flags: int = 0  # F_GLOBAL: a=(buf), b=(field) ints
    gl: tuple | None = None

    def vals(self):
        if self.flags & F_GLOBAL:
            return []
        return [v for v in (self.a, self.b) if v is not None and v.kind == "op"]

    def bound(self):
        if self.flags & F_GLOBAL:
            if self.gl[3]:
                return (1 << 256) / P  # 32-byte field
            return (1 << (384 - self.gl[2])) / P  # any (384 - cleared)-bit value
        return abs(self.ca) + abs(self.cb)


@dataclass
class Op:
    id: int
    terms: list  # [(Operand, Operand)]
    epi: list  # [Operand]
    ncorr: int
    out: Val | None = None
    dst_global: tuple | None = None  # (buf, field)
    xmask: int = 0
    kind: str = "mac"  # mac | sel | bit
    post: str = ""  # "" | "iszero" | "gthalf"
    sel: list = field(default_factory=list)  # [flag, a, b] Operands for kind == "sel"
    bit: tuple | None = None  # (buf, off16, nbytes, bitindex) for kind == "bit"
    dst_word: bool = False  # int32 status store
    post_scale: int = 1  # the reduced sum of products is multiplied by this small factor
    per_batch: bool = False  # global store by lane 0 at index item/32
    pad_const: Val | None = None  # padding lanes (item >= n_items) produce this constant instead
    after: list = field(default_factory=list)  # extra ordering deps (Ops)
    # two-output record (kind == "mac2"): out / out2 are the two results; groups M, P, 0, 1 of product entries
    # [(Operand x, Operand y, x2 | None, y2 | None)]: dst0 = ps0*redc(P + M + G0) + epi, dst1 = ps1*redc(P - M + G1) + epi2
    out2: Val | None = None
    groups: dict = field(default_factory=dict)
    epi2: list = field(default_factory=list)
    ncorr2: int = 0
    post_scale2: int = 1
    kp2: int = 0
    step: int = -1
    warp: int = -1
    sidx: int = -1
    tag: str = ""

    def cost(self):
        """Estimated duration in units of one 12x12 product.  Calibrated on the per-record clock trace of the
        pairing program on a B200 (profiles/r1_notes.md): cycles ~ 2500 + 1700*T + 1100*E + 125*ncorr for a
        multiply-accumulate record, ~1600 for an epilogue-only record, ~193 k for the inversion."""
        if COST_MODEL == "v1":
            if self.kind == "inv":
                return 235.0
            if self.kind != "mac":
                return 0.5
            t = len(self.terms)
            return (t + 1.7 if t else 0.6) + 0.1 * len(self.epi)
        if self.kind == "inv":
            return 114.0
        if self.kind == "mac2":
            ne = sum(len(g) for g in self.groups.values())
            nf = sum(1 for g in self.groups.values() for e in g if e[2] is not None or e[3] is not None)
            shared = 1 if (self.groups.get("M") or self.groups.get("P")) else 0
            return (ne + 0.15 * nf + 2 * 1.1 + 0.45 + 0.3 * shared + (0.15 if self.groups.get("M") else 0.0)
                    + 0.65 * (len(self.epi) + len(self.epi2)) + 0.07 * (self.ncorr + self.ncorr2))
        if self.kind != "mac":
            return 0.4
        t = len(self.terms)
        if not t:
            return 0.6 + 0.35 * len(self.epi) + 0.07 * self.ncorr
        return t + 1.5 + 0.65 * len(self.epi) + 0.07 * self.ncorr

    def src_vals(self):
        vs = []
        for x, y in self.terms:
            vs += x.vals() + y.vals()
        for e in self.epi:
            vs += e.vals()
        for o in self.sel:
            vs += o.vals()
        for g in self.groups.values():
            for x, y, x2, y2 in g:
                for o in (x, y, x2, y2):
                    if o is not None:
                        vs += o.vals()
        for e in self.epi2:
            vs += e.vals()
        return vs

    def outs(self):
        return [v for v in (self.out, self.out2) if v is not None]

    def n_products(self):
        if self.kind == "mac2":
            return sum(len(g) for g in self.groups.values())
        return len(self.terms) if self.kind == "mac" else 0

    def n_reductions(self):
        if self.kind == "mac2":
            return 2
        return 1 if (self.kind == "mac" and self.terms) else 0


# Two-output records (products shared by the two coefficients of an Fp2 value: Karatsuba without an extra reduction) are
# an EXPERIMENT that lost on the B200 (profiles/r2_notes.md: 0.84 M pairings/s against 1.29 M for one-output records; the
# multiplier is not the bottleneck, per-warp dependent-instruction latency is, and the 768-bit recombination adds more of
# it than three saved products remove).  BLS381_VM_DUAL=1 builds the programs in that format for A/B runs.
DUAL_DEFAULT = __import__("os").environ.get("BLS381_VM_DUAL", "0") == "1"
SHARE_ENABLED = True  # False: coefficient pairs are materialised as independent records even if they share products (tuning)
CURRENT_DUAL = True   # policy of the Builder created last (read by tower.E2 when it expands an Fp2 product)


class Builder:
    def __init__(self, warps=6, dual=None):
        global CURRENT_DUAL
        self.warps = warps
        self.dual = DUAL_DEFAULT if dual is None else dual
        CURRENT_DUAL = self.dual
        self.ops: list[Op] = []
        self.vals: list[Val] = []
        self.consts: list[int] = []  # integer values exactly as stored in the device table
        self._cidx: dict[int, Val] = {}
        self.tag = ""
        self._after: list[Op] = []
        self.R2 = self.const_raw(R2_MOD_P)
        self.ONE_PLAIN = self.const_raw(1)

    # ---- constants --------------------------------------------------------------------------
    def const_raw(self, stored: int) -> Val:
        stored %= P
        if stored in self._cidx:
            return self._cidx[stored]
        v = Val(len(self.vals), "const", cidx=len(self.consts))
        self.vals.append(v)
        self.consts.append(stored)
        self._cidx[stored] = v
        return v

    def const(self, value: int) -> Val:
        """Field constant `value` (stored in Montgomery form)."""
        return self.const_raw(value % P * R_MOD_P % P)

    # ---- op creation ------------------------------------------------------------------------
    def _new_op(self, terms, epi, ncorr, dst_global=None, xmask=0):
        op = Op(len(self.ops), terms, epi, ncorr, dst_global=dst_global, xmask=xmask, tag=self.tag)
        op.after = list(self._after)
        if dst_global is None:
            v = Val(len(self.vals), "op", op=op)
            self.vals.append(v)
            op.out = v
        self.ops.append(op)
        return op

    def set_after(self, ops):
        """Every op created from now on is ordered after `ops` (software-pipelining fence)."""
        self._after = list(ops)

    def inp(self, buf: int, fld: int) -> Val:
        """Wire-format 48-byte field number `fld` of input buffer `buf` -> Montgomery-form value."""
        return self.inp_bytes(buf, fld * 48)

    def inp_bytes(self, buf: int, byte_off: int, nbytes: int = 48, clear_top: int = 0, montgomery: bool = True) -> Val:
        """Big-endian field of 48 or 32 bytes at `byte_off` (multiple of 16); the top `clear_top` bits are
        masked off (compression flags).  Returns the value mod p in Montgomery form (or the plain integer
        mod p when montgomery=False)."""
        assert byte_off % 16 == 0 and nbytes in (48, 32) and 0 <= clear_top <= 15
        x = Operand(None, 1, flags=F_GLOBAL, gl=(buf, byte_off // 16, clear_top, 1 if nbytes == 32 else 0))
        y = Operand(self.R2 if montgomery else self.const_raw(R_MOD_P), 1, flags=F_CONST)
        # x < 2^384, y < p: x*y/R < 2^384*p/R + p = 2p
        return self._new_op([(x, y)], [], 1).out

    def out(self, e, buf: int, fld: int, per_batch=False):
        """Store canonical, out-of-Montgomery value to wire-format output."""
        if isinstance(e, Lin) and e.is_zero():
            x = Operand(self.const_raw(0), 1, flags=F_CONST)
        else:
            x = self._operand(Lin.of(self.mat(e)))
        y = Operand(self.ONE_PLAIN, 1, flags=F_CONST)
        op = self._new_op([(x, y)], [], 1, dst_global=(buf, fld * 3))
        op.per_batch = per_batch
        return op

    def out_word(self, flag_val, buf: int, byte_off: int = 0):
        """Store a plain small integer (flag / status code) as int32 at `byte_off` of the item's record."""
        assert byte_off % 16 == 0
        op = self._new_op([], [self._operand(Lin.of(self.mat(flag_val)))], 0, dst_global=(buf, byte_off // 16))
        op.dst_word = True
        return op

    # ---- predicates / selection (flags are plain integers 0 / 1 held in slots) -----------------------
    def is_zero(self, e) -> Val:
        """1 if the expression is 0 mod p else 0."""
        v = self.mat_expr_op(e)
        v.op.post = "iszero"
        return v

    def gt_half(self, e) -> Val:
        """1 if the canonical NON-Montgomery value of `e` is > (p-1)/2  (the reference's `(y*2)/P` flag)."""
        x = self._operand(Lin.of(self.mat(e)))
        y = Operand(self.ONE_PLAIN, 1, flags=F_CONST)
        op = self._new_op([(x, y)], [], 1)
        op.post = "gthalf"
        return op.out

    def parity(self, e) -> Val:
        """Least significant bit of the canonical NON-Montgomery value of `e` (sgn0, math.ts:1179-1189)."""
        x = self._operand(Lin.of(self.mat(e)))
        y = Operand(self.ONE_PLAIN, 1, flags=F_CONST)
        op = self._new_op([(x, y)], [], 1)
        op.post = "parity"
        return op.out

    def flag_xor(self, f1: Val, f2: Val) -> Val:
        return self.select(f1, Lin.of(self.flag_not(f2)), Lin.of(f2))

    def mat_expr_op(self, e) -> Val:
        """Materialise `e` with a FRESH micro-op (so that post-processing flags can be attached)."""
        e = as_expr(e)
        if isinstance(e, Lin):
            e = Quad([], e)
        n_before = len(self.ops)
        v = self.mat(e)
        if v.kind != "op" or v.op.id < n_before or v.op.post:
            v = self._new_op([], [self._operand(Lin.of(v))], 0).out  # pre-existing value: test a copy
        return v

    def select(self, flag: Val, a, b) -> Val:
        """flag ? a : b  (per lane)."""
        oa = self._operand(self.lin_operand(as_expr(a) if isinstance(as_expr(a), Lin) else Lin.of(self.mat(a))))
        ob = self._operand(self.lin_operand(as_expr(b) if isinstance(as_expr(b), Lin) else Lin.of(self.mat(b))))
        of = self._operand(Lin.of(flag))
        k = max(oa.bound(), ob.bound())
        ncorr = 0
        while (1 << ncorr) <= k + 1e-9:
            ncorr += 1
        if k <= 1 and oa.ca >= 0 and ob.ca >= 0 and oa.cb >= 0 and ob.cb >= 0:
            ncorr = 0
        op = self._new_op([], [], ncorr)
        op.kind = "sel"
        op.sel = [of, oa, ob]
        return op.out

    def inv(self, e) -> Val:
        """1/e mod p in ONE micro-op (binary GCD in registers, csrc/fp_inv.cuh); inv(0) = 0."""
        v = self.mat(e)
        op = self._new_op([], [], 0)
        op.kind = "inv"
        op.sel = [self._operand(Lin.of(v))]
        return op.out

    def bit(self, buf: int, byte_off: int, nbytes: int, bitindex: int) -> Val:
        """Bit `bitindex` (0 = least significant) of the big-endian `nbytes`-byte field at `byte_off`."""
        assert byte_off % 16 == 0
        op = self._new_op([], [], 0)
        op.kind = "bit"
        op.bit = (buf, byte_off // 16, nbytes, bitindex)
        return op.out

    def flag_and(self, f1: Val, f2: Val) -> Val:
        return self.select(f1, Lin.of(f2), Lin.of(self.const_raw(0)))

    def flag_or(self, f1: Val, f2: Val) -> Val:
        return self.select(f1, Lin.of(self.const_raw(1)), Lin.of(f2))

    def flag_not(self, f: Val) -> Val:
        return self.select(f, Lin.of(self.const_raw(0)), Lin.of(self.const_raw(1)))

    def pad_select(self, e, const_val: Val) -> Val:
        """Copy of `e` in which padding lanes (items beyond n_items) hold the constant instead."""
        v = self.mat(e)
        op = self._new_op([], [Operand(v, 1)], 0)
        op.pad_const = const_val
        return op.out

    def xlane(self, v: Val, mask: int) -> Val:
        """Value of `v` in lane (lane ^ mask): a copy op with the XLANE operand flag."""
        o = Operand(v, 1, flags=F_XLANE)
        return self._new_op([], [o], 0, xmask=mask).out

    # ---- materialisation --------------------------------------------------------------------
    def _operand(self, lin: Lin) -> Operand:
        items = list(lin.t.items())
        assert 1 <= len(items) <= 2, "operand must have 1..2 terms"
        kinds = {v.kind for v, _ in items}
        assert len(kinds) == 1, "operand terms must be all slots or all constants"
        flags = F_CONST if "const" in kinds else 0
        (a, ca) = items[0]
        (b, cb) = items[1] if len(items) == 2 else (None, 0)
        assert -4 <= ca <= 4 and -4 <= cb <= 4
        return Operand(a, ca, b, cb, flags)

    def _fits_operand(self, lin: Lin) -> bool:
        items = list(lin.t.items())
        if not (1 <= len(items) <= 2):
            return False
        if len({v.kind for v, _ in items}) != 1:
            return False
        return all(abs(c) <= 4 for _, c in items)

    def lin_operand(self, lin: Lin) -> Lin:
        """Returns a Lin usable as a single operand, materialising it if it is too wide."""
        if self._fits_operand(lin):
            return lin
        return Lin.of(self.mat(lin))

    def mat(self, e) -> Val:
        """Materialise an expression as (usually) one micro-op; returns the resulting Val."""
        if isinstance(e, Val):
            return e
        e = as_expr(e)
        if isinstance(e, Lin):
            if len(e.t) == 1:
                (v, c), = e.t.items()
                if c == 1 and v.kind == "op":
                    return v
            q = Quad([], e)
        else:
            q = e
        if not q.terms and q.lin.is_zero():
            q = Quad([], Lin.of(self.const_raw(0)))
        # product terms: make each side a valid operand
        # normalise every product to k * x' * y' with primitive operands, then pull the common integer factor g of
        # all products out of the sum: it is applied ONCE to the reduced result (post-scale) instead of to an
        # operand of every term
        import math
        norm = []
        for k, x, y in q.terms:
            x, y = self.lin_operand(x), self.lin_operand(y)
            gx = math.gcd(*[abs(c) for c in x.t.values()])
            gy = math.gcd(*[abs(c) for c in y.t.values()])
            x = Lin({v: c // gx for v, c in x.t.items()})
            y = Lin({v: c // gy for v, c in y.t.items()})
            norm.append((k * gx * gy, x, y))
        g = 0
        for k, _, _ in norm:
            g = math.gcd(g, abs(k))
        epi_chunks = self._chunk_lin(q.lin)

        def terms_for(post):
            terms = []
            for k, x, y in norm:
                k //= post
                if k < 0:
                    x, k = -x, -k
                while k > 0:
                    # fold as much of the remaining integer factor as the 4-bit operand coefficients allow
                    fx = 4 // max(abs(c) for c in x.t.values())
                    fy = 4 // max(abs(c) for c in y.t.values())
                    best = None
                    for dx in range(1, fx + 1):
                        for dy in range(1, fy + 1):
                            if dx * dy <= k and (best is None or dx * dy > best[0] * best[1]):
                                best = (dx, dy)
                    dx, dy = best
                    terms.append((x.scale(dx), y.scale(dy)))
                    k -= dx * dy
            return terms

        def pack(terms, post):
            """greedy packing into micro-ops under T<=12, E<=2, bound limits: [(terms, epilogue chunks, bound)]"""
            pending_terms = list(terms)
            pending_epi = list(epi_chunks)
            plan = []
            while True:
                cur_t, cur_e = [], []
                sumb, kb = 0.0, 0.0
                while pending_terms and len(cur_t) < MAX_TERMS:
                    x, y = pending_terms[0]
                    b = x.bound() * y.bound()
                    if sumb + b > MAX_SUM_BOUND or post * ((sumb + b) * P_OVER_R + 1) + kb > MAX_K - 0.01:
                        break
                    cur_t.append(pending_terms.pop(0))
                    sumb += b
                base = post * (sumb * P_OVER_R + 1) if cur_t else 0.0
                while pending_epi and len(cur_e) < 2:
                    b = pending_epi[0].bound()
                    if base + kb + b > MAX_K - 0.01:
                        break
                    cur_e.append(pending_epi.pop(0))
                    kb += b
                assert cur_t or cur_e, "cannot make progress materialising expression"
                plan.append((cur_t, cur_e, base + kb))
                if not pending_terms and not pending_epi:
                    return plan

        # the common factor is applied as a post-scale of 4, 3 or 2 (first that divides g) -- unless the scaled result
        # bound then forces the sum to be split over several records and a smaller post-scale (more of the factor folded
        # into the operands) does it in fewer
        cands = [c for c in (4, 3, 2) if g and g % c == 0][:1] + [1]
        post, plan = None, None
        for cand in cands:
            pl = pack(terms_for(cand), cand)
            if plan is None or len(pl) < len(plan):
                post, plan = cand, pl
        if len(plan) == 1:
            cur_t, cur_e, kb = plan[0]
            return self._emit(cur_t, cur_e, kb, post)
        partial = [self._emit(cur_t, cur_e, kb, post) for cur_t, cur_e, kb in plan]
        # sum the partial results
        return self.mat(sum((Lin.of(v) for v in partial[1:]), Lin.of(partial[0])))

    # ---- two-output records ---------------------------------------------------------------------------
    def mat_lin(self, e) -> Lin:
        """Materialise unless the expression is structurally zero (kept symbolic so sparsity survives)."""
        if isinstance(e, Lin) and e.is_zero():
            return e
        if isinstance(e, Quad) and not e.terms and e.lin.is_zero():
            return Lin()
        return Lin.of(self.mat(e))

    def mat2(self, terms, l0: Lin, l1: Lin):
        """Materialise the coefficient pair  (sum k0_i X_i Y_i + l0,  sum k1_i X_i Y_i + l1)  given as
        terms = [(k0, k1, X: Lin, Y: Lin)].  Products with k0 != 0 and k1 != 0 are computed ONCE (two-output record,
        csrc/vm.cuh OP_MAC2); without shared products the two coefficients become independent records."""
        terms = [(k0, k1, x, y) for k0, k1, x, y in terms if (k0 or k1) and not x.is_zero() and not y.is_zero()]

        def split():
            t0 = [(k0, x, y) for k0, k1, x, y in terms if k0]
            t1 = [(k1, x, y) for k0, k1, x, y in terms if k1]
            return self.mat_lin(Quad(t0, l0) if t0 else l0), self.mat_lin(Quad(t1, l1) if t1 else l1)

        if not (self.dual and SHARE_ENABLED and any(k0 and k1 for k0, k1, _, _ in terms)):
            return split()
        r = self._try_emit2(terms, l0, l1)
        return r if r is not None else split()

    def _halves(self, lin: Lin):
        """Operand(s) for a Lin of up to four terms: [Operand] or [Operand, Operand] (summed on load, F_FUSE)."""
        if self._fits_operand(lin):
            return [self._operand(lin)]
        items = list(lin.t.items())
        if len(items) <= 4 and all(abs(c) <= 4 for _, c in items):
            slots = [(v, c) for v, c in items if v.kind != "const"]
            consts = [(v, c) for v, c in items if v.kind == "const"]
            halves = []
            for grp in (slots, consts):
                for i in range(0, len(grp), 2):
                    halves.append(Lin(dict(grp[i:i + 2])))
            if len(halves) <= 2:
                return [self._operand(h) for h in halves]
        return [self._operand(Lin.of(self.mat(lin)))]

    def _try_emit2(self, terms, l0: Lin, l1: Lin):
        import math

        def post_of(ks):
            g = 0
            for k in ks:
                g = math.gcd(g, abs(k))
            for cand in (4, 3, 2):
                if g and g % cand == 0:
                    return cand
            return 1

        ps0 = post_of([k0 for k0, _, _, _ in terms if k0])
        ps1 = post_of([k1 for _, k1, _, _ in terms if k1])
        groups = {"M": [], "P": [], "0": [], "1": []}

        def add(cls, k, x, y):
            gx = math.gcd(*[abs(c) for c in x.t.values()])
            gy = math.gcd(*[abs(c) for c in y.t.values()])
            x = Lin({v: c // gx for v, c in x.t.items()})
            y = Lin({v: c // gy for v, c in y.t.items()})
            k *= gx * gy
            if k < 0:
                x, k = -x, -k
            while k > 0:
                fx = 4 // max(abs(c) for c in x.t.values())
                fy = 4 // max(abs(c) for c in y.t.values())
                best = None
                for dx in range(1, fx + 1):
                    for dy in range(1, fy + 1):
                        if dx * dy <= k:
                            key = (dx * dy, -max(dx, dy))
                            if best is None or key > best[0]:
                                best = (key, dx, dy)
                _, dx, dy = best
                groups[cls].append((x.scale(dx), y.scale(dy)))
                k -= dx * dy

        for k0, k1, x, y in terms:
            assert k0 % ps0 == 0 and k1 % ps1 == 0
            a, c = k0 // ps0, k1 // ps1
            if a and c:
                sgn = 1 if a > 0 else -1
                m = min(abs(a), abs(c)) * sgn
                if (a > 0) == (c > 0):
                    add("P", m, x, y)
                    a, c = a - m, c - m
                else:
                    add("M", m, x, y)
                    a, c = a - m, c + m
            if a:
                add("0", a, x, y)
            if c:
                add("1", c, x, y)
        # operands (up to four slots each, summed on load)
        enc = {}
        nent = 0
        bsum = {}
        for cls, lst in groups.items():
            out = []
            b_ = 0.0
            for x, y in lst:
                hx, hy = self._halves(x), self._halves(y)
                x2 = hx[1] if len(hx) > 1 else None
                y2 = hy[1] if len(hy) > 1 else None
                out.append((hx[0], hy[0], x2, y2))
                nent += 2 if (x2 is not None or y2 is not None) else 1
                bx = hx[0].bound() + (x2.bound() if x2 is not None else 0)
                by = hy[0].bound() + (y2.bound() if y2 is not None else 0)
                if bx > 9 or by > 9:
                    return None
                b_ += bx * by
            enc[cls] = out
            bsum[cls] = b_
        if nent > MAX_ENTRIES2:
            return None
        K = int(math.ceil(bsum["M"] - 1e-9)) if enc["M"] else 0
        if K > MAX_KP2:
            return None
        acc0 = bsum["P"] + bsum["M"] + bsum["0"]
        acc1 = bsum["P"] + K + bsum["1"]
        if acc0 > MAX_SUM_BOUND or acc1 > MAX_SUM_BOUND:
            return None
        res = []
        for acc, ps, lin in ((acc0, ps0, l0), (acc1, ps1, l1)):
            chunks = self._chunk_lin(lin)
            if len(chunks) > 2:
                chunks = self._chunk_lin(Lin.of(self.mat(lin)))
            kb = ps * (acc * P_OVER_R + 1) + sum(ch.bound() for ch in chunks)
            if kb > MAX_K - 0.01:
                if not chunks:
                    return None
                # keep the epilogue out of the record: added by a later (lazy) consumer instead
                return None
            ncorr = 0
            while (1 << ncorr) <= kb + 1e-9:
                ncorr += 1
            assert ncorr <= 4, kb
            res.append(([self._operand(z) for z in chunks], ncorr))
        op = self._new_op([], res[0][0], res[0][1])
        op.kind = "mac2"
        op.groups = enc
        op.epi2, op.ncorr2 = res[1]
        op.post_scale, op.post_scale2, op.kp2 = ps0, ps1, K
        v2 = Val(len(self.vals), "op", op=op)
        self.vals.append(v2)
        op.out2 = v2
        return Lin.of(op.out), Lin.of(v2)

    def _chunk_lin(self, lin: Lin):
        chunks = []
        slots = [(v, c) for v, c in lin.t.items() if v.kind != "const"]
        consts = [(v, c) for v, c in lin.t.items() if v.kind == "const"]
        for group in (slots, consts):
            # split big coefficients
            flat = []
            for v, c in group:
                while abs(c) > 4:
                    flat.append((v, 4 if c > 0 else -4))
                    c -= 4 if c > 0 else -4
                flat.append((v, c))
            i = 0
            while i < len(flat):
                if i + 1 < len(flat) and flat[i][0] is not flat[i + 1][0]:
                    chunks.append(Lin({flat[i][0]: flat[i][1], flat[i + 1][0]: flat[i + 1][1]}))
                    i += 2
                else:
                    chunks.append(Lin({flat[i][0]: flat[i][1]}))
                    i += 1
        return chunks

    def _emit(self, lin_terms, lin_epi, k_bound, post_scale=1) -> Val:
        terms = [(self._operand(x), self._operand(y)) for x, y in lin_terms]
        epi = [self._operand(z) for z in lin_epi]
        # operands with negative coefficients are bounded INCLUSIVELY (p - 0 = p), so require 2^ncorr > k
        k = max(k_bound, 1.0)
        ncorr = 0
        while (1 << ncorr) <= k + 1e-9:
            ncorr += 1
        if not lin_terms and all(c > 0 for z in lin_epi for c in z.t.values()) and abs(k - round(k)) < 1e-9:
            # purely positive sum of canonical values: strictly below k*p
            ncorr = 0
            while (1 << ncorr) < k - 1e-9:
                ncorr += 1
        assert ncorr <= 4, k
        op = self._new_op(terms, epi, ncorr)
        if terms:
            op.post_scale = post_scale
        return op.out

    # ------------------------------------------------------------------------------------------
    # reference evaluation of the SSA program on plain integers (no scheduling/slots): used by tests
    def eval_ssa(self, inputs, lanes=1):
        """inputs: {(buf, fld): [int per lane]} ; returns ({val_id: [ints]}, {(buf,fld): [ints]})"""
        rinv = pow(R, -1, P)
        env = {}
        outs = {}

        def ev_operand(o: Operand, lane, xmask):
            if o.flags & F_GLOBAL:
                return inputs[(o.gl[0], o.gl[1])][lane] & ((1 << (256 if o.gl[3] else 384 - o.gl[2])) - 1)
            ln = lane ^ xmask if (o.flags & F_XLANE) else lane

            def one(v):
                return self.consts[v.cidx] if v.kind == "const" else env[v.id][ln]

            r = (one(o.a) if o.ca >= 0 else P - one(o.a)) * abs(o.ca)
            if o.b is not None and o.cb != 0:
                r += (one(o.b) if o.cb >= 0 else P - one(o.b)) * abs(o.cb)
            return r

        for op in self.ops:
            res = []
            if op.kind == "mac2":
                res2 = []
                for lane in range(lanes):
                    def prod(e):
                        x, y, x2, y2 = e
                        xv = ev_operand(x, lane, 0) + (ev_operand(x2, lane, 0) if x2 is not None else 0)
                        yv = ev_operand(y, lane, 0) + (ev_operand(y2, lane, 0) if y2 is not None else 0)
                        assert xv < R and yv < R
                        return xv * yv
                    S = {k: sum(prod(e) for e in g) for k, g in op.groups.items()}
                    assert S["M"] <= op.kp2 * P * P
                    a0 = S["P"] + S["M"] + S["0"]
                    a1 = S["P"] - S["M"] + op.kp2 * P * P + S["1"]
                    assert 0 <= a0 < R * R and 0 <= a1 < R * R
                    r0 = a0 * rinv % P * op.post_scale + sum(ev_operand(e, lane, 0) for e in op.epi)
                    r1 = a1 * rinv % P * op.post_scale2 + sum(ev_operand(e, lane, 0) for e in op.epi2)
                    res.append(r0 % P)
                    res2.append(r1 % P)
                env[op.out.id] = res
                env[op.out2.id] = res2
                continue
            for lane in range(lanes):
                if op.kind == "sel":
                    f, a_, b_ = (ev_operand(o, lane, op.xmask) for o in op.sel)
                    r = a_ if f else b_
                elif op.kind == "inv":
                    x_ = ev_operand(op.sel[0], lane, op.xmask) % P  # Montgomery form a*R
                    r = pow(x_ * rinv % P, -1, P) * R % P if x_ else 0
                elif op.kind == "bit":
                    buf, off16, nbytes, bitindex = op.bit
                    r = (inputs[(buf, off16)][lane] >> bitindex) & 1
                else:
                    acc = 0
                    for x, y in op.terms:
                        acc += ev_operand(x, lane, op.xmask) * ev_operand(y, lane, op.xmask)
                    r = acc * rinv % P * op.post_scale if op.terms else 0
                    for e in op.epi:
                        r += ev_operand(e, lane, op.xmask)
                r %= P
                if op.post == "iszero":
                    r = 1 if r == 0 else 0
                elif op.post == "gthalf":
                    r = 1 if r > (P - 1) // 2 else 0
                elif op.post == "parity":
                    r = r & 1
                res.append(r)
            if op.dst_global is not None:
                outs[op.dst_global] = res
            else:
                env[op.out.id] = res
        return env, outs

    # ------------------------------------------------------------------------------------------
    def schedule(self):
        """Static dataflow list scheduling onto the W warps of a CTA.

        There are no CTA-wide barriers: every warp executes its own stream in order and, before a record,
        waits until the other warps' progress counters (number of completed records) reach the values the
        record carries.  Here: (1) priorities = longest path to a sink, (2) event-driven list scheduling with
        an estimated duration per micro-op, (3) the resulting start times define one linear order L; each
        warp's stream is its ops in L order."""
        W = self.warps
        n = len(self.ops)
        deps = [set() for _ in range(n)]
        users = [[] for _ in range(n)]
        for op in self.ops:
            for v in op.src_vals():
                deps[op.id].add(v.op.id)
            for a in op.after:
                deps[op.id].add(a.id)
        for i in range(n):
            for d in deps[i]:
                users[d].append(i)
        cost = [op.cost() for op in self.ops]
        prio = [0.0] * n
        for i in range(n - 1, -1, -1):
            prio[i] = cost[i] + max((prio[u] for u in users[i]), default=0.0)
        remaining = [len(d) for d in deps]
        est = [0.0] * n  # earliest start (all producers finished)
        finish = [0.0] * n
        start = [0.0] * n
        ready = [i for i in range(n) if remaining[i] == 0]
        warp_free = [0.0] * W
        scheduled = 0
        while scheduled < n:
            assert ready, "dependency cycle"
            w = min(range(W), key=lambda k: warp_free[k])
            t = warp_free[w]
            avail = [i for i in ready if est[i] <= t + 1e-9]
            if avail:
                i = max(avail, key=lambda k: prio[k])
            else:
                i = min(ready, key=lambda k: (est[k], -prio[k]))
                t = est[i]
            ready.remove(i)
            start[i] = t
            finish[i] = t + cost[i]
            warp_free[w] = finish[i]
            self.ops[i].warp = w
            scheduled += 1
            for u in users[i]:
                remaining[u] -= 1
                est[u] = max(est[u], finish[i])
                if remaining[u] == 0:
                    ready.append(u)
        order = sorted(range(n), key=lambda i: (start[i], i))
        self.order = order
        self.pos = {i: k for k, i in enumerate(order)}  # position in the linear order L
        self.streams = [[] for _ in range(W)]
        for i in order:
            op = self.ops[i]
            op.step = self.pos[i]
            op.sidx = len(self.streams[op.warp])  # index in its warp's stream
            self.streams[op.warp].append(i)
        makespan = max(finish) if n else 0.0
        total = sum(cost)
        self.sched_stats = {
            "ops": n,
            "steps": n,
            "total_cost": total,
            "critical_cost": makespan,
            "efficiency": total / (W * makespan) if makespan else 1.0,
        }
        return self.streams

    # ------------------------------------------------------------------------------------------
    def allocate(self, nslots: int, nfar_max: int = 180):
        """Slot allocation along the linear order L.  A value occupies its slot from its definition to its
        last reader (positions in L); a slot is re-written only by an op later in L than every reader of the
        previous value, and that op WAITS for those readers (WAR) -- see `_compute_waits`."""
        last_use = {}
        readers = {}
        for op in self.ops:
            for v in op.src_vals():
                last_use[v.id] = max(last_use.get(v.id, -1), op.step)
                readers.setdefault(v.id, []).append(op.id)
        self.readers = readers
        defs = [(op, v) for op in self.ops for v in op.outs()]
        for op, v in defs:
            last_use.setdefault(v.id, op.step)
        npos = len(self.ops)
        far = set()
        uses = {vid: len(r) for vid, r in readers.items()}

        def pressure():
            delta = [0] * (npos + 2)
            for op, v in defs:
                if v.id in far:
                    continue
                delta[op.step] += 1
                delta[last_use[v.id] + 1] -= 1
            cur, peak, at = 0, 0, 0
            for s_ in range(npos + 1):
                cur += delta[s_]
                if cur > peak:
                    peak, at = cur, s_
            return peak, at

        peak, at = pressure()
        while peak > nslots:
            cands = [(op, v) for op, v in defs if v.id not in far and op.step <= at <= last_use[v.id]]
            cands.sort(key=lambda ov: (last_use[ov[1].id] - ov[0].step) / (1 + uses.get(ov[1].id, 0)), reverse=True)
            far.add(cands[0][1].id)
            peak, at = pressure()
        self.peak_slots = peak
        nslots = max(peak, 1)  # never reserve more shared memory than the program needs
        slot_of = {}
        import heapq
        free_near = [(0, k) for k in range(nslots)]  # (position at which it became free, slot): LRU first
        heapq.heapify(free_near)
        free_far = []
        nfar = 0
        busy = []  # heap of (last_use, slot, is_far)
        self.prev_in_slot = {}  # op id -> value ids previously held by the slots it writes
        cur_val = {}
        for op, v in sorted(defs, key=lambda ov: ov[0].step):
            while busy and busy[0][0] < op.step:
                lu, k, isfar = heapq.heappop(busy)
                heapq.heappush(free_far if isfar else free_near, (lu, k))
            vid = v.id
            if vid in far:
                if free_far:
                    _, k = heapq.heappop(free_far)
                else:
                    k = nslots + nfar
                    nfar += 1
                heapq.heappush(busy, (last_use[vid], k, True))
            else:
                assert free_near, "slot allocation failed despite pressure check"
                _, k = heapq.heappop(free_near)
                heapq.heappush(busy, (last_use[vid], k, False))
            slot_of[vid] = k
            if k in cur_val:
                self.prev_in_slot.setdefault(op.id, []).append(cur_val[k])
            cur_val[k] = vid
        self.nslots = nslots
        self.nfar = nfar
        assert self.nfar <= nfar_max and nslots + self.nfar <= 255, (self.nfar, nslots)
        self.slot_of = slot_of
        self._compute_waits()
        return slot_of

    def _compute_waits(self):
        """Per record: the progress each other warp must have reached (RAW on operands, WAR/WAW on the
        destination slot).  Requirements already implied by earlier records of the same warp are dropped."""
        W = self.warps
        val_op = {v.id: op for op in self.ops for v in op.outs()}
        known = [[0] * W for _ in range(W)]
        self.waits = {}
        self.full_reqs = {}
        for i in self.order:
            op = self.ops[i]
            req = [0] * W
            for v in op.src_vals():
                y = v.op
                req[y.warp] = max(req[y.warp], y.sidx + 1)
            for a in op.after:
                req[a.warp] = max(req[a.warp], a.sidx + 1)
            for pv in self.prev_in_slot.get(i, []):
                y = val_op[pv]
                req[y.warp] = max(req[y.warp], y.sidx + 1)
                for zid in self.readers.get(pv, []):
                    z = self.ops[zid]
                    req[z.warp] = max(req[z.warp], z.sidx + 1)
            req[op.warp] = 0
            self.full_reqs[i] = list(req)
            k = known[op.warp]
            emit = [r if r > k[w2] else 0 for w2, r in enumerate(req)]
            for w2, r in enumerate(req):
                k[w2] = max(k[w2], r)
            self.waits[i] = emit

    def check_hazards(self):
        """Vector-clock proof that the emitted waits order every RAW, WAR and WAW pair."""
        W = self.warps
        vc = {}  # op id -> vector clock after completion
        last_vc = [[0] * W for _ in range(W)]
        for i in self.order:
            op = self.ops[i]
            c = list(last_vc[op.warp])
            for w2, need in enumerate(self.waits[i]):
                if need:
                    other = self.streams[w2][need - 1]
                    assert other in vc, "wait on a record later in the linear order (deadlock)"
                    c = [max(a, b) for a, b in zip(c, vc[other])]
            # every requirement must be implied by the clock BEFORE executing
            for w2, need in enumerate(self.full_reqs[i]):
                assert c[w2] >= need, f"op {i}: missing ordering on warp {w2}"
            c[op.warp] = op.sidx + 1
            vc[i] = c
            last_vc[op.warp] = c
        # slot exclusivity along L
        holder = {}
        for i in self.order:
            op = self.ops[i]
            for v in op.src_vals():
                assert holder.get(self.slot_of[v.id]) == v.id, "operand slot overwritten before use"
            for v in op.outs():
                holder[self.slot_of[v.id]] = v.id
            if op.kind == "mac2":  # the destination slots are scratch while the record runs: no operand may live there
                for v in op.src_vals():
                    assert self.slot_of[v.id] not in (self.slot_of[op.out.id], self.slot_of[op.out2.id])

    def _enc_operand(self, o: Operand) -> int:
        if o.flags & F_GLOBAL:
            buf, off16, clear_top, short32 = o.gl
            assert off16 < 256
            return buf | (off16 << 8) | (clear_top << 16) | (short32 << 20) | (F_GLOBAL << 24)

        def idx(v):
            return v.cidx if v.kind == "const" else self.slot_of[v.id]

        a = idx(o.a)
        b = idx(o.b) if o.b is not None else 0
        flags = o.flags
        cb = o.cb if o.b is not None else 0
        near = a < self.nslots and (cb == 0 or b < self.nslots)
        if flags == 0 and near:
            if (o.ca, cb) == (1, 0):
                flags |= F_SIMPLE
            elif (o.ca, cb) == (-1, 0):
                flags |= 1 << 4
            elif (o.ca, cb) == (1, 1):
                flags |= 2 << 4
            elif (o.ca, cb) == (1, -1):
                flags |= 3 << 4
            elif (o.ca, cb) == (-1, 1):
                a, b = b, a
                flags |= 3 << 4
            elif (o.ca, cb) == (-1, -1):
                flags |= 4 << 4
            elif (o.ca, cb) == (2, 0):
                flags |= 5 << 4
        return a | (b << 8) | ((o.ca & 0xF) << 16) | ((o.cb & 0xF) << 20) | (flags << 24)

    def _enc_waits(self, words, wt):
        """Words 60..63: eight 16-bit progress requirements (warps 0..7), or ten 12-bit / twelve 10-bit fields."""
        W = self.warps
        if any(wt):
            words[0] |= H_BAR  # "has waits"
        if W <= 8:  # eight 16-bit fields
            wt = wt + [0] * (8 - W)
            assert max(wt) < 65536
            for j in range(4):
                words[WAIT_WORD + j] = wt[2 * j] | (wt[2 * j + 1] << 16)
        else:  # W fields of 12 bits (W <= 10) or 10 bits (W = 12), little-endian over the four words
            fw = 12 if W <= 10 else 10
            assert max(wt) < (1 << fw)
            big = 0
            for k, v_ in enumerate(wt):
                big |= v_ << (fw * k)
            for j in range(4):
                words[WAIT_WORD + j] = (big >> (32 * j)) & 0xFFFFFFFF

    def encode(self):
        """Returns (prog: bytes [W][nrec][64 words], nrec)."""
        W = self.warps
        assert W <= 12
        streams = []
        for w in range(W):
            recs = []
            for i in self.streams[w]:
                op = self.ops[i]
                words = [0] * REC_WORDS
                dst = 0 if op.dst_global is not None else self.slot_of[op.out.id]
                opcode = {"mac": OP_MAC, "sel": OP_SEL, "bit": OP_BIT, "inv": OP_INV, "mac2": OP_MAC2}[op.kind]
                if op.kind == "mac2" or (self.dual and op.kind == "mac" and op.terms):
                    # two-output format; a one-output multiply-accumulate is the H_SINGLE form (group 0 only)
                    single = op.kind == "mac"
                    groups = {"0": [(x, y, None, None) for x, y in op.terms]} if single else op.groups
                    epi2, ncorr2, ps2, kp2 = ([], 0, 1, 0) if single else (op.epi2, op.ncorr2, op.post_scale2, op.kp2)
                    words[0] = OP_MAC2 | (dst << 8) | ((0 if single else self.slot_of[op.out2.id]) << 16)
                    if single:
                        words[0] |= H_SINGLE
                        if op.post == "iszero":
                            words[0] |= H_POST_ISZERO
                        elif op.post == "gthalf":
                            words[0] |= H_POST_GTHALF
                        elif op.post == "parity":
                            words[0] |= H_POST_ISZERO | H_POST_GTHALF
                        if op.dst_word:
                            words[0] |= H_DSTWORD
                        assert op.xmask == 0  # cross-lane operands only occur in epilogue-only records
                        gaux = 0
                        if op.dst_global is not None:
                            words[0] |= H_DSTG
                            gaux |= op.dst_global[0] | (op.dst_global[1] << 8)
                            if op.per_batch:
                                words[0] |= H_DSTBATCH
                        if op.pad_const is not None:
                            words[0] |= H_PADCONST
                            gaux |= op.pad_const.cidx << 24
                        words[3] = gaux
                    else:
                        assert not op.post and op.dst_global is None and op.pad_const is None and op.xmask == 0
                    cnt = {}
                    wi = 4
                    for cls in ("M", "P", "0", "1"):
                        n = 0
                        for x, y, x2, y2 in groups.get(cls, []):
                            fused = x2 is not None or y2 is not None
                            words[wi] = self._enc_operand(x) | ((F_FUSE << 24) if fused else 0)
                            words[wi + 1] = self._enc_operand(y)
                            wi += 2
                            n += 1
                            if fused:
                                words[wi] = self._enc_operand(x2) if x2 is not None else 0
                                words[wi + 1] = self._enc_operand(y2) if y2 is not None else 0
                                assert (x2 is None or words[wi] != 0) and (y2 is None or words[wi + 1] != 0)
                                wi += 2
                                n += 1
                        cnt[cls] = n
                    assert wi <= 4 + 2 * MAX_ENTRIES2 and max(cnt.values()) < 32
                    assert len(op.epi) <= 2 and len(epi2) <= 2 and kp2 <= MAX_KP2
                    words[1] = (cnt["M"] | (cnt["P"] << 5) | (cnt["0"] << 10) | (cnt["1"] << 15) | (len(op.epi) << 20)
                                | (len(epi2) << 22) | (op.ncorr << 24) | (ncorr2 << 27))
                    words[2] = kp2 | (op.post_scale << 8) | (ps2 << 12)
                    for e, z in enumerate(op.epi):
                        words[52 + e] = self._enc_operand(z)
                    for e, z in enumerate(epi2):
                        words[54 + e] = self._enc_operand(z)
                    self._enc_waits(words, self.waits[i])
                    recs.append(words)
                    continue
                hdr = opcode | (dst << 8) | (len(op.terms) << 16) | (len(op.epi) << 20) | (op.ncorr << 22)
                if op.post == "iszero":
                    hdr |= H_POST_ISZERO
                elif op.post == "gthalf":
                    hdr |= H_POST_GTHALF
                elif op.post == "parity":
                    hdr |= H_POST_ISZERO | H_POST_GTHALF
                if op.dst_word:
                    hdr |= H_DSTWORD
                aux = op.xmask << 16
                if op.dst_global is not None:
                    hdr |= H_DSTG
                    aux |= op.dst_global[0] | (op.dst_global[1] << 8)
                    if op.per_batch:
                        hdr |= H_DSTBATCH
                if op.pad_const is not None:
                    assert op.post_scale == 1
                    hdr |= H_PADCONST
                    aux |= op.pad_const.cidx << 24
                elif op.post_scale > 1:
                    aux |= op.post_scale << 24
                words[0], words[1] = hdr, aux
                for t, (x, y) in enumerate(op.terms):
                    words[2 + 2 * t] = self._enc_operand(x)
                    words[3 + 2 * t] = self._enc_operand(y)
                for e, z in enumerate(op.epi):
                    words[26 + 2 * e] = self._enc_operand(z)
                if op.kind == "sel":
                    words[2], words[3], words[4] = (self._enc_operand(o) for o in op.sel)
                elif op.kind == "inv":
                    words[2] = self._enc_operand(op.sel[0])
                elif op.kind == "bit":
                    buf, off16, nbytes, bitindex = op.bit
                    words[2], words[3], words[4] = buf | (off16 << 8), bitindex, nbytes
                self._enc_waits(words, self.waits[i])
                recs.append(words)
            streams.append(recs)
        nrec = max(len(s_) for s_ in streams)
        blob = bytearray()
        for w in range(W):
            recs = streams[w] + [[OP_NOP] + [0] * (REC_WORDS - 1)] * (nrec - len(streams[w]))
            for words in recs:
                blob += struct.pack("<%dI" % REC_WORDS, *words)
        return bytes(blob), nrec

    def const_table(self) -> bytes:
        out = bytearray()
        for c in self.consts:
            out += struct.pack("<12I", *[(c >> (32 * i)) & 0xFFFFFFFF for i in range(12)])
        return bytes(out)

