"""Synthetic circuits for the generic constraint-program engine (`eval_check`, `accumulate`, witgen).

The real rv32im circuit (taps, `poly_ext` program, witgen) is a generated artifact of the un-vendored crate
risc0-circuit-rv32im 4.0.2 (/root/reference Cargo.lock:5320-5323) and is not available in this environment
(SURVEY.md §0.4).  The prover is therefore generic over a *circuit blob* whose constraint program uses exactly
the step vocabulary of risc0-zkp 3.0.2 `adapter.rs` (`PolyExtStep`: Const, ConstExt, Get, GetGlobal, Add, Sub,
Mul, True, AndEqz, AndCond).  This module builds satisfiable circuits of configurable
width in that vocabulary, shaped like the real one (three register groups accum/code/data, selector-gated
constraints of degree <= 5, bit constraints, back-taps, LogUp-style ext-field accumulators fed by Fiat-Shamir
mix values, globals pinned on the first row), together with the witness-generation program that satisfies them.

Blob layout (u32 words; both the oracle `oracle/circuit.c` and the product `zeth_b200/csrc/circuit.cpp` parse it):
  [0] magic 'B2ZC' [1] version=1 [2..4] group sizes (accum, code, data) [5] globals [6] mix size [7] n_taps
  [8] n_steps [9] ret mix-var [10..13] 16-byte circuit info [14] n_wit_steps [15] n_wit_stores [16] n_accum
  [17] n_pins [18..23] reserved; then taps (offset | back<<16 | group<<24), steps (5 words each: op a b c d),
  witgen steps (5 words each), witgen stores (data col, var), accumulator source columns, pins (data col, global),
  data column kinds (0 free, 1 free bit, 2 derived).
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field

import numpy as np

P = 15 * (1 << 27) + 1
MAGIC = 0x435A3242
G_ACCUM, G_CODE, G_DATA = 0, 1, 2
OP_CONST, OP_CONST_EXT, OP_GET, OP_GET_GLOBAL, OP_ADD, OP_SUB, OP_MUL, OP_TRUE, OP_AND_EQZ, OP_AND_COND = range(10)


@dataclass
class _Prog:
    """Append-only step list with hash-consing of Const/Get/GetGlobal nodes."""

    steps: list = field(default_factory=list)
    n_fp: int = 0
    n_mix: int = 0
    memo: dict = field(default_factory=dict)

    def _fp(self, op, a=0, b=0, c=0, d=0, memo_key=None):
        if memo_key is not None and memo_key in self.memo:
            return self.memo[memo_key]
        self.steps.append((op, a, b, c, d))
        v = self.n_fp
        self.n_fp += 1
        if memo_key is not None:
            self.memo[memo_key] = v
        return v

    def const(self, v):
        return self._fp(OP_CONST, v % P, memo_key=("c", v % P))

    def const_ext(self, a, b, c, d):
        return self._fp(OP_CONST_EXT, a, b, c, d, memo_key=("ce", a, b, c, d))

    def get(self, tap_index):
        return self._fp(OP_GET, tap_index, memo_key=("g", tap_index))

    def get_global(self, base, off):
        return self._fp(OP_GET_GLOBAL, base, off, memo_key=("gg", base, off))

    def add(self, a, b):
        return self._fp(OP_ADD, a, b)

    def sub(self, a, b):
        return self._fp(OP_SUB, a, b)

    def mul(self, a, b):
        return self._fp(OP_MUL, a, b)

    def _mix(self, op, a=0, b=0, c=0):
        self.steps.append((op, a, b, c, 0))
        v = self.n_mix
        self.n_mix += 1
        return v

    def true(self):
        return self._mix(OP_TRUE)

    def and_eqz(self, chain, val):
        return self._mix(OP_AND_EQZ, chain, val)

    def and_cond(self, chain, cond, inner):
        return self._mix(OP_AND_COND, chain, cond, inner)


@dataclass
class SynthCircuit:
    blob: np.ndarray
    group_sizes: tuple
    globals_size: int
    mix_size: int
    n_taps: int
    n_steps: int
    info: bytes
    name: str

    @property
    def w_accum(self):
        return self.group_sizes[G_ACCUM]

    @property
    def w_code(self):
        return self.group_sizes[G_CODE]

    @property
    def w_data(self):
        return self.group_sizes[G_DATA]


def make_synth_circuit(w_code=16, w_data=192, w_accum=48, n_globals=32, seed=1, free_frac=0.4, name=None) -> SynthCircuit:
    """Build a satisfiable synthetic circuit; expression trees are (kind, ...) tuples over tap references."""
    assert w_code >= 3 and w_accum % 4 == 0 and w_data >= 8
    rnd = random.Random(seed)
    n_accum = w_accum // 4
    n_free = max(6, int(w_data * free_frac))
    n_free = min(n_free, w_data - 1)
    kinds = [2] * w_data
    free_cols = sorted(rnd.sample(range(w_data), n_free))
    bit_cols = set(rnd.sample(free_cols, max(1, n_free // 4)))
    for c in free_cols:
        kinds[c] = 1 if c in bit_cols else 0
    derived_cols = [c for c in range(w_data) if kinds[c] == 2]
    n_pins = min(8, n_globals, n_free)
    pins = [(free_cols[i], i) for i in range(n_pins)]

    # ---- expression trees for derived columns ----
    def f():  # free data tap
        return ("tap", G_DATA, rnd.choice(free_cols), rnd.choice([0, 0, 0, 1, 1, 2]))

    def k():  # code constant column tap
        return ("tap", G_CODE, rnd.randrange(2, w_code), 0)

    def cst():
        return ("const", rnd.randrange(1, P))

    def glb():
        return ("global", 0, rnd.randrange(n_globals))

    def mk_expr():
        t = rnd.random()
        if t < 0.30:
            return ("add", ("mul", f(), f()), k())
        if t < 0.50:
            return ("add", ("mul", ("mul", f(), f()), f()), f())
        if t < 0.60:
            return ("mul", ("mul", ("add", f(), k()), ("sub", f(), f())), ("mul", f(), f()))
        if t < 0.78:
            return ("sub", ("add", f(), ("mul", cst(), f())), k())
        if t < 0.90:
            a = f()
            return ("add", ("mul", a, a), ("mul", glb(), f()))
        a, b = f(), f()
        d = ("sub", a, b)
        return ("mul", d, ("sub", d, cst()))

    exprs = {c: mk_expr() for c in derived_cols}
    accum_src = [rnd.choice(free_cols + derived_cols) for _ in range(n_accum)]

    # ---- tap set: every column at back 0, plus the backs the expressions use, accum at {0,1} ----
    taps = set()
    for g, w in ((G_ACCUM, w_accum), (G_CODE, w_code), (G_DATA, w_data)):
        for c in range(w):
            taps.add((g, c, 0))
    for c in range(w_accum):
        taps.add((G_ACCUM, c, 1))

    def collect(e):
        if e[0] == "tap":
            taps.add((e[1], e[2], e[3]))
            # a register's back-list is contiguous from 0 in the real circuit; keep that shape
            for b in range(e[3]):
                taps.add((e[1], e[2], b))
        elif e[0] in ("add", "sub", "mul"):
            collect(e[1])
            collect(e[2])

    for e in exprs.values():
        collect(e)
    taps = sorted(taps)
    tap_index = {t: i for i, t in enumerate(taps)}

    def emit(p: _Prog, e):
        if e[0] == "tap":
            return p.get(tap_index[(e[1], e[2], e[3])])
        if e[0] == "const":
            return p.const(e[1])
        if e[0] == "global":
            return p.get_global(e[1], e[2])
        a, b = emit(p, e[1]), emit(p, e[2])
        return {"add": p.add, "sub": p.sub, "mul": p.mul}[e[0]](a, b)

    # ---- constraint program ----
    p = _Prog()
    active = p.get(tap_index[(G_CODE, 0, 0)])
    first = p.get(tap_index[(G_CODE, 1, 0)])
    one = p.const(1)
    inner = p.true()
    for c in derived_cols:
        lhs = p.get(tap_index[(G_DATA, c, 0)])
        inner = p.and_eqz(inner, p.sub(lhs, emit(p, exprs[c])))
    for c in sorted(bit_cols):
        b = p.get(tap_index[(G_DATA, c, 0)])
        inner = p.and_eqz(inner, p.mul(b, p.sub(one, b)))
    # accumulators: (A - (1 - first) * Aprev) * (M + D) - 1 == 0 over the extension field
    xs = [p.const_ext(1, 0, 0, 0), p.const_ext(0, 1, 0, 0), p.const_ext(0, 0, 1, 0), p.const_ext(0, 0, 0, 1)]
    not_first = p.sub(one, first)

    def ext_from(vals):
        acc = p.mul(xs[0], vals[0])
        for j in range(1, 4):
            acc = p.add(acc, p.mul(xs[j], vals[j]))
        return acc

    for kacc in range(n_accum):
        a_now = ext_from([p.get(tap_index[(G_ACCUM, 4 * kacc + j, 0)]) for j in range(4)])
        a_prev = ext_from([p.get(tap_index[(G_ACCUM, 4 * kacc + j, 1)]) for j in range(4)])
        m = ext_from([p.get_global(1, 4 * kacc + j) for j in range(4)])
        d = p.get(tap_index[(G_DATA, accum_src[kacc], 0)])
        val = p.sub(p.mul(p.sub(a_now, p.mul(a_prev, not_first)), p.add(m, d)), one)
        inner = p.and_eqz(inner, val)
    pin_chain = p.true()
    for col, g in pins:
        pin_chain = p.and_eqz(pin_chain, p.sub(p.get(tap_index[(G_DATA, col, 0)]), p.get_global(0, g)))
    outer = p.true()
    outer = p.and_cond(outer, active, inner)
    outer = p.and_cond(outer, first, pin_chain)
    ret = outer

    # ---- witgen program (base field only) ----
    wp = _Prog()
    stores = []
    for c in derived_cols:
        stores.append((c, emit(wp, exprs[c])))

    info = (name or f"synth-{w_code}-{w_data}-{w_accum}-{seed}").encode()[:16].ljust(16, b"_")
    words = [MAGIC, 1, w_accum, w_code, w_data, n_globals, 4 * n_accum, len(taps), len(p.steps), ret]
    words += list(np.frombuffer(info, dtype="<u4"))
    words += [len(wp.steps), len(stores), n_accum, n_pins, 0, 0, 0, 0, 0, 0]
    assert len(words) == 24
    for g, c, b in taps:
        words.append(c | (b << 16) | (g << 24))
    for s in p.steps:
        words.extend(s)
    for s in wp.steps:
        words.extend(s)
    for c, v in stores:
        words.extend((c, v))
    words.extend(accum_src)
    for col, g in pins:
        words.extend((col, g))
    words.extend(kinds)
    blob = np.array(words, dtype=np.uint32)
    return SynthCircuit(blob=blob, group_sizes=(w_accum, w_code, w_data), globals_size=n_globals, mix_size=4 * n_accum,
                        n_taps=len(taps), n_steps=len(p.steps), info=info, name=info.decode())


def make_mux_circuit(w_code=16, w_data=192, w_accum=48, n_globals=32, n_modes=16, n_free=90, seed=2, max_backs=(5, 6, 7), name=None) -> SynthCircuit:
    """A circuit of the SIZE class of rv32im 4.0.2's constraint program (10^4 - 10^5 steps, ~10^3 taps) with
    its SHAPE: every row is in exactly one of `n_modes` modes (one-hot selector columns, like the major / minor selects of
    the real circuit) and every derived column has its own degree <= 3 definition per mode, gated AndCond(chain, sel_m,
    block_m) -- total degree 5 with the `active` selector; free columns are tapped at backs 0 .. 5-7 (registers of up to 8
    taps, 6 combos), accumulators at backs {0, 1, 2}.  The witness generator evaluates the mux (sum_m sel_m * expr_m)."""
    assert w_code >= 3 and w_accum % 4 == 0 and n_modes + n_free < w_data and n_modes <= 255
    rnd = random.Random(seed)
    n_accum = w_accum // 4
    kinds = [2] * w_data
    sel_cols = list(range(n_modes))
    for i, c in enumerate(sel_cols):
        kinds[c] = 3 | (i << 8) | (n_modes << 16)
    free_cols = sorted(rnd.sample(range(n_modes, w_data), n_free))
    bit_cols = set(rnd.sample(free_cols, max(1, n_free // 5)))
    for c in free_cols:
        kinds[c] = 1 if c in bit_cols else 0
    derived_cols = [c for c in range(w_data) if kinds[c] == 2]
    tappable = sel_cols + free_cols
    back_of = {c: rnd.choice(max_backs) for c in tappable}
    n_pins = min(8, n_globals, n_free)
    pins = [(free_cols[i], i) for i in range(n_pins)]

    def f():
        c = rnd.choice(tappable)
        return ("tap", G_DATA, c, rnd.randint(0, back_of[c]))

    def k():
        return ("tap", G_CODE, rnd.randrange(2, w_code), rnd.choice([0, 0, 1]))

    def cst():
        return ("const", rnd.randrange(1, P))

    def glb():
        return ("global", 0, rnd.randrange(n_globals))

    def term():  # degree <= 3
        t = rnd.random()
        if t < 0.30:
            return ("add", ("mul", f(), f()), k())
        if t < 0.55:
            return ("add", ("mul", ("mul", f(), f()), f()), f())
        if t < 0.70:
            return ("sub", ("add", f(), ("mul", cst(), f())), k())
        if t < 0.85:
            a = f()
            return ("add", ("mul", a, a), ("mul", glb(), f()))
        a, b = f(), f()
        d = ("sub", a, b)
        return ("mul", d, ("sub", d, cst()))

    def mk_expr():
        return ("add", term(), ("mul", cst(), term())) if rnd.random() < 0.8 else term()

    exprs = {m: {c: mk_expr() for c in derived_cols} for m in range(n_modes)}
    accum_src = [rnd.choice(free_cols + derived_cols) for _ in range(n_accum)]

    taps = set()
    for g, w in ((G_ACCUM, w_accum), (G_CODE, w_code), (G_DATA, w_data)):
        for c in range(w):
            taps.add((g, c, 0))
    for c in range(w_accum):
        taps.update({(G_ACCUM, c, 1), (G_ACCUM, c, 2)})
    for c in range(2, w_code):
        taps.add((G_CODE, c, 1))
    for c in tappable:
        for b in range(back_of[c] + 1):
            taps.add((G_DATA, c, b))
    taps = sorted(taps)
    tap_index = {t: i for i, t in enumerate(taps)}

    def emit(p: _Prog, e):
        if e[0] == "tap":
            return p.get(tap_index[(e[1], e[2], e[3])])
        if e[0] == "const":
            return p.const(e[1])
        if e[0] == "global":
            return p.get_global(e[1], e[2])
        a, b = emit(p, e[1]), emit(p, e[2])
        return {"add": p.add, "sub": p.sub, "mul": p.mul}[e[0]](a, b)

    p = _Prog()
    active = p.get(tap_index[(G_CODE, 0, 0)])
    first = p.get(tap_index[(G_CODE, 1, 0)])
    one = p.const(1)
    inner = p.true()
    sels = [p.get(tap_index[(G_DATA, c, 0)]) for c in sel_cols]
    tot = sels[0]
    for sv in sels:
        inner = p.and_eqz(inner, p.mul(sv, p.sub(one, sv)))
    for sv in sels[1:]:
        tot = p.add(tot, sv)
    inner = p.and_eqz(inner, p.sub(tot, one))
    for m in range(n_modes):
        block = p.true()
        for c in derived_cols:
            block = p.and_eqz(block, p.sub(p.get(tap_index[(G_DATA, c, 0)]), emit(p, exprs[m][c])))
        inner = p.and_cond(inner, sels[m], block)
    for c in sorted(bit_cols):
        b = p.get(tap_index[(G_DATA, c, 0)])
        inner = p.and_eqz(inner, p.mul(b, p.sub(one, b)))
    xs = [p.const_ext(1, 0, 0, 0), p.const_ext(0, 1, 0, 0), p.const_ext(0, 0, 1, 0), p.const_ext(0, 0, 0, 1)]
    not_first = p.sub(one, first)

    def ext_from(vals):
        acc = p.mul(xs[0], vals[0])
        for j in range(1, 4):
            acc = p.add(acc, p.mul(xs[j], vals[j]))
        return acc

    for kacc in range(n_accum):
        a_now = ext_from([p.get(tap_index[(G_ACCUM, 4 * kacc + j, 0)]) for j in range(4)])
        a_prev = ext_from([p.get(tap_index[(G_ACCUM, 4 * kacc + j, 1)]) for j in range(4)])
        mm = ext_from([p.get_global(1, 4 * kacc + j) for j in range(4)])
        d = p.get(tap_index[(G_DATA, accum_src[kacc], 0)])
        inner = p.and_eqz(inner, p.sub(p.mul(p.sub(a_now, p.mul(a_prev, not_first)), p.add(mm, d)), one))
    pin_chain = p.true()
    for col, g in pins:
        pin_chain = p.and_eqz(pin_chain, p.sub(p.get(tap_index[(G_DATA, col, 0)]), p.get_global(0, g)))
    outer = p.and_cond(p.true(), active, inner)
    outer = p.and_cond(outer, first, pin_chain)
    ret = outer

    wp = _Prog()
    wsels = [wp.get(tap_index[(G_DATA, c, 0)]) for c in sel_cols]
    stores = []
    for c in derived_cols:
        acc = None
        for m in range(n_modes):
            t = wp.mul(wsels[m], emit(wp, exprs[m][c]))
            acc = t if acc is None else wp.add(acc, t)
        stores.append((c, acc))

    info = (name or f"mux-{w_data}-{n_modes}-{seed}").encode()[:16].ljust(16, b"_")
    words = [MAGIC, 1, w_accum, w_code, w_data, n_globals, 4 * n_accum, len(taps), len(p.steps), ret]
    words += list(np.frombuffer(info, dtype="<u4"))
    words += [len(wp.steps), len(stores), n_accum, n_pins, 0, 0, 0, 0, 0, 0]
    assert len(words) == 24
    for g, c, b in taps:
        words.append(c | (b << 16) | (g << 24))
    for st in p.steps:
        words.extend(st)
    for st in wp.steps:
        words.extend(st)
    for c, v in stores:
        words.extend((c, v))
    words.extend(accum_src)
    for col, g in pins:
        words.extend((col, g))
    words.extend(kinds)
    blob = np.array(words, dtype=np.uint32)
    return SynthCircuit(blob=blob, group_sizes=(w_accum, w_code, w_data), globals_size=n_globals, mix_size=4 * n_accum,
                        n_taps=len(taps), n_steps=len(p.steps), info=info, name=info.decode())


# Named workloads (BASELINE.md §4 planning widths: code 16, data 192, accum 48 -> W = 256)
def circuit_planning() -> SynthCircuit:
    return make_synth_circuit(16, 192, 48, 32, seed=0xB200, name="b2z-plan-256")


def circuit_small() -> SynthCircuit:
    return make_synth_circuit(4, 24, 8, 8, seed=7, name="b2z-small-36")


def circuit_tiny() -> SynthCircuit:
    return make_synth_circuit(3, 8, 4, 4, seed=3, name="b2z-tiny-15")


def circuit_rv32im_scale() -> SynthCircuit:
    """W = 256 like the planning circuit, constraint program of the rv32im size class (>= 3 * 10^4 steps, ~10^3 taps)."""
    return make_mux_circuit(16, 192, 48, 32, n_modes=32, n_free=80, seed=0xB2004, name="b2z-rv32im-scale")


def circuit_mux_small() -> SynthCircuit:
    """The same mux shape at test size (one-hot modes, AndCond-gated blocks, deep back taps)."""
    return make_mux_circuit(4, 28, 8, 8, n_modes=4, n_free=10, seed=11, max_backs=(2, 3, 5), name="b2z-mux-small")
