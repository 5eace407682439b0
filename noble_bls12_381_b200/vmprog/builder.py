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

OP_NOP, OP_MAC = 0, 1
F_CONST, F_GLOBAL, F_XLANE = 1, 2, 4
H_BAR, H_DSTG, H_DSTBATCH, H_PADCONST = 1 << 25, 1 << 26, 1 << 28, 1 << 29
REC_WORDS = 32
MAX_TERMS = 12
MAX_SUM_BOUND = 80.0  # sum of |x|*|y| bounds in units of p^2 that fits the 768-bit accumulator
MAX_K = 8.0  # result bound (units of p) that `correct` can bring back to [0,p)


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
            return R / P  # any 384-bit value
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
    per_batch: bool = False  # global store by lane 0 at index item/32
    pad_const: Val | None = None  # padding lanes (item >= n_items) produce this constant instead
    after: list = field(default_factory=list)  # extra ordering deps (Ops)
    step: int = -1
    warp: int = -1
    tag: str = ""

    def cost(self):
        t = len(self.terms)
        return (t + 1.1 if t else 0.0) + 0.35 + 0.1 * len(self.epi)

    def src_vals(self):
        vs = []
        for x, y in self.terms:
            vs += x.vals() + y.vals()
        for e in self.epi:
            vs += e.vals()
        return vs


class Builder:
    def __init__(self, warps=6):
        self.warps = warps
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
        """Wire-format field `fld` of input buffer `buf` -> Montgomery-form value."""
        x = Operand(None, 1, flags=F_GLOBAL, gl=(buf, fld))
        y = Operand(self.R2, 1, flags=F_CONST)
        # x < 2^384, R2 < p: x*R2/R < 2^384*p/R + p = 2p
        return self._new_op([(x, y)], [], 1).out

    def out(self, e, buf: int, fld: int, per_batch=False):
        """Store canonical, out-of-Montgomery value to wire-format output."""
        if isinstance(e, Lin) and e.is_zero():
            x = Operand(self.const_raw(0), 1, flags=F_CONST)
        else:
            x = self._operand(Lin.of(self.mat(e)))
        y = Operand(self.ONE_PLAIN, 1, flags=F_CONST)
        op = self._new_op([(x, y)], [], 1, dst_global=(buf, fld))
        op.per_batch = per_batch
        return op

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
        terms = []
        for k, x, y in q.terms:
            x, y = self.lin_operand(x), self.lin_operand(y)
            if k < 0:
                x, k = -x, -k
            while k > 0:
                # fold as much of the integer factor as the 4-bit operand coefficients allow
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
        # merge identical (x, y) operand pairs?  (rare; skip)
        # epilogue: split the Lin into chunks of <= 2 same-kind terms with bound <= 4
        epi_chunks = self._chunk_lin(q.lin)
        # greedy packing into micro-ops under T<=12, E<=2, bound limits
        pending_terms = list(terms)
        pending_epi = list(epi_chunks)
        partial: list[Val] = []
        while True:
            cur_t, cur_e = [], []
            sumb, kb = 0.0, 0.0
            while pending_terms and len(cur_t) < MAX_TERMS:
                x, y = pending_terms[0]
                b = x.bound() * y.bound()
                if sumb + b > MAX_SUM_BOUND or (sumb + b) * P_OVER_R + 1 + kb > MAX_K - 0.01:
                    break
                cur_t.append(pending_terms.pop(0))
                sumb += b
            base = (sumb * P_OVER_R + 1) if cur_t else 0.0
            while pending_epi and len(cur_e) < 2:
                b = pending_epi[0].bound()
                if base + kb + b > MAX_K - 0.01:
                    break
                cur_e.append(pending_epi.pop(0))
                kb += b
            assert cur_t or cur_e, "cannot make progress materialising expression"
            done = not pending_terms and not pending_epi
            if done and not partial:
                return self._emit(cur_t, cur_e, base + kb)
            partial.append(self._emit(cur_t, cur_e, base + kb))
            if done:
                break
            # fold partial results into the epilogue queue
        # sum the partial results
        return self.mat(sum((Lin.of(v) for v in partial[1:]), Lin.of(partial[0])))

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

    def _emit(self, lin_terms, lin_epi, k_bound) -> Val:
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
        assert ncorr <= 3, k
        return self._new_op(terms, epi, ncorr).out

    # ------------------------------------------------------------------------------------------
    # reference evaluation of the SSA program on plain integers (no scheduling/slots): used by tests
    def eval_ssa(self, inputs, lanes=1):
        """inputs: {(buf, fld): [int per lane]} ; returns ({val_id: [ints]}, {(buf,fld): [ints]})"""
        rinv = pow(R, -1, P)
        env = {}
        outs = {}

        def ev_operand(o: Operand, lane, xmask):
            if o.flags & F_GLOBAL:
                return inputs[o.gl][lane]
            ln = lane ^ xmask if (o.flags & F_XLANE) else lane

            def one(v):
                return self.consts[v.cidx] if v.kind == "const" else env[v.id][ln]

            r = (one(o.a) if o.ca >= 0 else P - one(o.a)) * abs(o.ca)
            if o.b is not None and o.cb != 0:
                r += (one(o.b) if o.cb >= 0 else P - one(o.b)) * abs(o.cb)
            return r

        for op in self.ops:
            res = []
            for lane in range(lanes):
                acc = 0
                for x, y in op.terms:
                    acc += ev_operand(x, lane, op.xmask) * ev_operand(y, lane, op.xmask)
                r = acc * rinv % P if op.terms else 0
                for e in op.epi:
                    r += ev_operand(e, lane, op.xmask)
                res.append(r % P)
            if op.dst_global is not None:
                outs[op.dst_global] = res
            else:
                env[op.out.id] = res
        return env, outs

    # ------------------------------------------------------------------------------------------
    def schedule(self):
        """Level-based list scheduling: a step holds ops whose producers finished in earlier steps;
        ops of a step are packed onto the warps longest-first.  Returns list of steps (lists per warp)."""
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
        # priority = longest path to a sink
        prio = [0.0] * n
        for i in range(n - 1, -1, -1):
            c = self.ops[i].cost()
            prio[i] = c + max((prio[u] for u in users[i]), default=0.0)
        remaining = [len(d) for d in deps]
        ready = [i for i in range(n) if remaining[i] == 0]
        steps = []
        scheduled = 0
        while scheduled < n:
            assert ready, "dependency cycle"
            ready.sort(key=lambda i: -prio[i])
            maxprio = prio[ready[0]]
            # urgent ops define the step length; the rest only fill idle capacity
            bins = [0.0] * W
            assign = [[] for _ in range(W)]
            chosen = []
            urgent = [i for i in ready if prio[i] >= maxprio - 1e-9 or True]
            total = sum(self.ops[i].cost() for i in urgent)
            cap = max(total / W, max(self.ops[i].cost() for i in urgent))
            for i in sorted(ready, key=lambda i: -self.ops[i].cost()):
                b = min(range(W), key=lambda k: bins[k])
                c = self.ops[i].cost()
                if bins[b] > 0 and bins[b] + c > cap * 1.05:
                    continue
                bins[b] += c
                assign[b].append(i)
                chosen.append(i)
            step_idx = len(steps)
            for w in range(W):
                for i in assign[w]:
                    self.ops[i].step = step_idx
                    self.ops[i].warp = w
            steps.append(assign)
            scheduled += len(chosen)
            cs = set(chosen)
            ready = [i for i in ready if i not in cs]
            for i in chosen:
                for u in users[i]:
                    remaining[u] -= 1
                    if remaining[u] == 0:
                        ready.append(u)
        self.steps = steps
        loads = [max(sum(self.ops[i].cost() for i in wl) for wl in st) for st in steps]
        total = sum(op.cost() for op in self.ops)
        self.sched_stats = {
            "ops": n,
            "steps": len(steps),
            "total_cost": total,
            "critical_cost": sum(loads),
            "efficiency": total / (W * sum(loads)) if loads else 1.0,
        }
        return steps

    # ------------------------------------------------------------------------------------------
    def allocate(self, nslots: int, nfar_max: int = 160):
        """Step-accurate slot allocation.  A value occupies its slot from its defining step to its last
        reading step; a slot may be re-written only in a step strictly after the last read."""
        last_use = {}
        for op in self.ops:
            for v in op.src_vals():
                last_use[v.id] = max(last_use.get(v.id, -1), op.step)
        defs = [op for op in self.ops if op.out is not None]
        for op in defs:
            last_use.setdefault(op.out.id, op.step)  # dead value: free right after
        # choose far values: repeatedly move the longest-lived values out until pressure fits
        nsteps = len(self.steps)
        far = set()

        def pressure():
            delta = [0] * (nsteps + 2)
            for op in defs:
                if op.out.id in far:
                    continue
                delta[op.step] += 1
                delta[last_use[op.out.id] + 1] -= 1
            cur, peak, prof = 0, 0, []
            for s in range(nsteps + 1):
                cur += delta[s]
                prof.append(cur)
                peak = max(peak, cur)
            return peak, prof

        peak, prof = pressure()
        while peak > nslots:
            # among values live at the peak step, evict the one with the fewest uses per lifetime
            s_peak = prof.index(peak)
            uses = {}
            for op in self.ops:
                for v in op.src_vals():
                    uses[v.id] = uses.get(v.id, 0) + 1
            cands = [op for op in defs if op.out.id not in far and op.step <= s_peak <= last_use[op.out.id]]
            cands.sort(key=lambda op: (last_use[op.out.id] - op.step) / (1 + uses.get(op.out.id, 0)), reverse=True)
            far.add(cands[0].out.id)
            peak, prof = pressure()
        self.peak_slots = peak
        slot_of = {}
        free_at = [0] * nslots  # first step in which the slot may be written again
        far_free_at = []
        by_step = sorted(defs, key=lambda op: (op.step, op.id))
        for op in by_step:
            vid = op.out.id
            lu = last_use[vid]
            if vid in far:
                for k in range(len(far_free_at)):
                    if far_free_at[k] <= op.step:
                        far_free_at[k] = lu + 1
                        slot_of[vid] = nslots + k
                        break
                else:
                    far_free_at.append(lu + 1)
                    slot_of[vid] = nslots + len(far_free_at) - 1
            else:
                for k in range(nslots):
                    if free_at[k] <= op.step:
                        free_at[k] = lu + 1
                        slot_of[vid] = k
                        break
                else:
                    raise AssertionError("slot allocation failed despite pressure check")
        self.nslots = nslots
        self.nfar = len(far_free_at)
        assert self.nfar <= nfar_max and nslots + self.nfar <= 255, (self.nfar, nslots)
        self.slot_of = slot_of
        return slot_of

    # ------------------------------------------------------------------------------------------
    def _enc_operand(self, o: Operand) -> int:
        if o.flags & F_GLOBAL:
            buf, fld = o.gl
            return buf | (fld << 8) | (1 << 16) | (F_GLOBAL << 24)

        def idx(v):
            return v.cidx if v.kind == "const" else self.slot_of[v.id]

        a = idx(o.a)
        b = idx(o.b) if o.b is not None else 0
        return a | (b << 8) | ((o.ca & 0xF) << 16) | ((o.cb & 0xF) << 20) | (o.flags << 24)

    def encode(self):
        """Returns (prog: bytes [W][nrec][32 words], nrec)."""
        W = self.warps
        streams = [[] for _ in range(W)]
        for s, st in enumerate(self.steps):
            for w in range(W):
                recs = []
                for i in st[w]:
                    op = self.ops[i]
                    words = [0] * REC_WORDS
                    dst = 0 if op.dst_global is not None else self.slot_of[op.out.id]
                    hdr = OP_MAC | (dst << 8) | (len(op.terms) << 16) | (len(op.epi) << 20) | (op.ncorr << 22)
                    aux = op.xmask << 16
                    if op.dst_global is not None:
                        hdr |= H_DSTG
                        aux |= op.dst_global[0] | (op.dst_global[1] << 8)
                        if op.per_batch:
                            hdr |= H_DSTBATCH
                    if op.pad_const is not None:
                        hdr |= H_PADCONST
                        aux |= op.pad_const.cidx << 24
                    words[0], words[1] = hdr, aux
                    for t, (x, y) in enumerate(op.terms):
                        words[2 + 2 * t] = self._enc_operand(x)
                        words[3 + 2 * t] = self._enc_operand(y)
                    for e, z in enumerate(op.epi):
                        words[26 + 2 * e] = self._enc_operand(z)
                    recs.append(words)
                if not recs:
                    recs.append([OP_NOP] + [0] * (REC_WORDS - 1))
                if s > 0:
                    recs[0][0] |= H_BAR
                streams[w] += recs
        nrec = max(len(s) for s in streams)
        # pad with NOPs WITHOUT barrier flags (all barriers already matched per step)
        blob = bytearray()
        for w in range(W):
            recs = streams[w] + [[OP_NOP] + [0] * (REC_WORDS - 1)] * (nrec - len(streams[w]))
            for words in recs:
                blob += struct.pack("<32I", *words)
        return bytes(blob), nrec

    def const_table(self) -> bytes:
        out = bytearray()
        for c in self.consts:
            out += struct.pack("<12I", *[(c >> (32 * i)) & 0xFFFFFFFF for i in range(12)])
        return bytes(out)

    def check_hazards(self):
        """No slot is read and written (or written twice) inside one step."""
        for s, st in enumerate(self.steps):
            reads, writes = set(), set()
            for wl in st:
                for i in wl:
                    op = self.ops[i]
                    for v in op.src_vals():
                        reads.add(self.slot_of[v.id])
                    if op.out is not None:
                        sl = self.slot_of[op.out.id]
                        assert sl not in writes, f"WAW hazard in step {s}"
                        writes.add(sl)
            assert not (reads & writes), f"RAW/WAR hazard in step {s}: {reads & writes}"
