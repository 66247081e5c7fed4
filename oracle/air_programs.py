"""TEST INFRASTRUCTURE (oracle side): the AIRs of oracle/air.py recorded as aero_air_programs
(include/aero_b200.h) -- what a Rust caller produces once per AIR by running Air::evaluate_transition /
evaluate_aux_transition over a symbolic element type -- with each constraint's degree adjustment and each
assertion's divisor column taken from the restated constraint groups.  Used by the parity tests and by the in-run
parity check of bench.py --gpus N; never by the product.  `to_abi_int` maps a canonical value to the context's ABI
form."""
from __future__ import annotations

from .air import MulFib2Air, P


def fib2_program(air, to_abi_int):
    """The fib2 / mulfib2 AIR as an aero_air_program: evaluate_transition (fib2/air.rs:41-58,
    mulfib2/air.rs:47-63) recorded node by node, the transition group's degree adjustment, and the assertions
    in the reference's coefficient order with their divisor columns (oracle/air.py boundary_groups restates
    the grouping)."""
    from aero_b200 import AirProgramBuilder

    b = AirProgramBuilder()
    c0, c1, n0, n1 = b.cur(0), b.cur(1), b.next(0), b.next(1)
    if isinstance(air, MulFib2Air):
        t0 = b.sub(n0, b.mul(c0, c1))   # next[0] - cur[0] * cur[1]
        t1 = b.sub(n1, b.mul(c1, n0))   # next[1] - cur[1] * next[0]
    else:
        t0 = b.sub(n0, b.add(c0, c1))   # next[0] - (cur[0] + cur[1])
        t1 = b.sub(n1, b.add(c1, n0))   # next[1] - (cur[1] + next[0])
    pairs = [(0, 0)] * (air.num_constraint_coefficients() // 2)
    (adj_t, members), = air.transition_groups(pairs[:2])
    assert [m[0] for m in members] == [0, 1]
    b.transition(t0, adj_t)
    b.transition(t1, adj_t)
    # coefficient pairs follow the sorted assertions; each lands in its group's divisor column
    assertions = sorted(air.get_assertions(), key=lambda a: (0, a.step, a.column))
    groups = air.boundary_groups(pairs[2:])
    for a in assertions:
        for j, (div, adj, mem) in enumerate(groups):
            if any(col == a.column and val == a.value for col, val, _ in mem) and div.b == (pow(air.g, a.step, P) if a.step else 1):
                b.assertion(a.column, to_abi_int(a.value), adj, 1 + j)
                break
        else:
            raise AssertionError("assertion without a group")
    return b.finish()


def masked_chain_program(air, to_abi_int):
    """MaskedChainAir.evaluate_transition recorded node by node, the periodic columns handed over by their cycle
    values, each constraint with its group's degree adjustment (degrees with cycles)."""
    from aero_b200 import AirProgramBuilder

    b = AirProgramBuilder()
    mask, ark0, ark1 = (b.periodic(b.periodic_column([to_abi_int(v) for v in col])) for col in air.periodic_columns)
    c0, c1, c2, n0, n1, n2 = b.cur(0), b.cur(1), b.cur(2), b.next(0), b.next(1), b.next(2)
    one = b.const(to_abi_int(1))
    not_mask = b.sub(one, mask)
    round0 = b.sub(n0, b.add(b.add(b.mul(b.mul(c0, c0), c0), c1), ark0))
    round1 = b.sub(n1, b.add(b.mul(c0, c1), ark1))
    t = [b.add(b.mul(mask, round0), b.mul(not_mask, b.sub(n0, c1))),
         b.add(b.mul(mask, round1), b.mul(not_mask, b.sub(n1, c0))),
         b.mul(ark1, b.sub(b.sub(n2, c2), one))]
    nt = len(t)
    pairs = [(0, 0)] * (air.num_constraint_coefficients() // 2)
    adj_of = {idx: adj for adj, members in air.transition_groups(pairs[:nt]) for idx, _ in members}
    for i in range(nt):
        b.transition(t[i], adj_of[i])
    assertions = sorted(air.get_assertions(), key=lambda a: (0, a.step, a.column))
    groups = air.boundary_groups(pairs[nt:])
    for a in assertions:
        (j, adj), = [(j, adj) for j, (div, adj, mem) in enumerate(groups)
                     if div.b == (pow(air.g, a.step, P) if a.step else 1)]
        b.assertion(a.column, to_abi_int(a.value), adj, 1 + j)
    return b.finish()


def bitwise_program(air, to_abi_int):
    """BitwiseChipletAir.evaluate_transition -- i.e. chiplets::enforce_selectors (s0, s1) + bitwise::enforce_constraints
    -- recorded node by node for the device evaluator."""
    from aero_b200 import AirProgramBuilder

    A = air
    b = AirProgramBuilder()
    k0, k1 = (b.periodic(b.periodic_column([to_abi_int(v) for v in col])) for col in A.periodic_columns)
    cur = [b.cur(c) for c in range(A.trace_width)]
    nxt = [b.next(c) for c in range(A.trace_width)]
    one, two, sixteen = (b.const(to_abi_int(v)) for v in (1, 2, 16))
    pow2 = [one, two, b.const(to_abi_int(4)), b.const(to_abi_int(8))]
    is_binary = lambda v: b.sub(b.mul(v, v), v)

    def agg(r, start):
        acc = b.mul(pow2[0], r[start])
        for i in range(1, 4):
            acc = b.add(acc, b.mul(pow2[i], r[start + i]))
        return acc

    s0, s1, sel = cur[A.S0], cur[A.S1], cur[A.SEL]
    t = [is_binary(s0), b.mul(s0, is_binary(s1)), b.mul(s0, b.sub(s0, nxt[A.S0])),
         b.mul(b.mul(s0, s1), b.sub(s1, nxt[A.S1]))]
    flag = b.mul(s0, b.sub(one, nxt[A.S1]))
    t.append(b.mul(flag, is_binary(sel)))
    t.append(b.mul(b.mul(flag, k1), b.sub(sel, nxt[A.SEL])))
    t += [b.mul(flag, is_binary(cur[A.A_BITS + i])) for i in range(4)]
    t += [b.mul(flag, is_binary(cur[A.B_BITS + i])) for i in range(4)]
    first, trans = b.mul(flag, k0), b.mul(flag, k1)
    t.append(b.mul(first, b.sub(cur[A.A], agg(cur, A.A_BITS))))
    t.append(b.mul(first, b.sub(cur[A.B], agg(cur, A.B_BITS))))
    t.append(b.mul(trans, b.sub(nxt[A.A], b.add(b.mul(sixteen, cur[A.A]), agg(nxt, A.A_BITS)))))
    t.append(b.mul(trans, b.sub(nxt[A.B], b.add(b.mul(sixteen, cur[A.B]), agg(nxt, A.B_BITS)))))
    t.append(b.mul(b.mul(k0, flag), cur[A.OUT_PREV]))
    t.append(b.mul(b.mul(k1, flag), b.sub(nxt[A.OUT_PREV], cur[A.OUT])))
    shifted = b.mul(cur[A.OUT_PREV], sixteen)
    and_acc = xor_acc = None
    for i in range(4):
        x, y = cur[A.A_BITS + i], cur[A.B_BITS + i]
        xy = b.mul(x, y)
        a_term = b.mul(pow2[i], xy)
        x_term = b.mul(pow2[i], b.sub(b.add(x, y), b.mul(two, xy)))
        and_acc = a_term if and_acc is None else b.add(and_acc, a_term)
        xor_acc = x_term if xor_acc is None else b.add(xor_acc, x_term)
    and_flag, xor_flag = b.mul(flag, b.sub(one, sel)), b.mul(flag, sel)
    t.append(b.add(b.mul(and_flag, b.sub(cur[A.OUT], b.add(shifted, and_acc))),
                   b.mul(xor_flag, b.sub(cur[A.OUT], b.add(shifted, xor_acc)))))
    nt = len(t)
    assert nt == len(A.transition_degrees)
    pairs = [(0, 0)] * (A.num_constraint_coefficients() // 2)
    adj_of = {idx: adj for adj, members in A.transition_groups(pairs[:nt]) for idx, _ in members}
    for i in range(nt):
        b.transition(t[i], adj_of[i])
    groups = A.boundary_groups(pairs[nt:])
    for a in sorted(A.get_assertions(), key=lambda a: (0, a.step, a.column)):
        (j, adj), = [(j, adj) for j, (div, adj, mem) in enumerate(groups) if div.b == (pow(A.g, a.step, P) if a.step else 1)]
        b.assertion(a.column, to_abi_int(a.value), adj, 1 + j)
    return b.finish()


def permutation_program(air, to_abi_int):
    """PermutationAir as a device program: main and auxiliary transition constraints over the concatenated frame
    (auxiliary column 0 = trace column main_width + 0); consts[0], consts[1] = the segment's random elements
    (alpha, beta), filled in by the aux_builder callback."""
    from aero_b200 import AirProgramBuilder

    n = air.n
    b = AirProgramBuilder()
    x, y, x_next = b.cur(0), b.cur(1), b.next(0)
    p_cur, p_next = b.cur(2), b.next(2)              # auxiliary column 0 = trace column main_width + 0
    alpha, beta, step = b.const(0), b.const(0), b.const(to_abi_int(air.STEP))
    v = lambda t: b.add(alpha, b.mul(beta, t))
    t_main = b.sub(b.sub(x_next, x), step)
    t_aux = b.sub(b.mul(p_next, v(y)), b.mul(p_cur, v(x)))
    pairs = [(0, 0)] * (air.num_constraint_coefficients() // 2)
    adj_of = {idx: adj for adj, members in air.transition_groups(pairs[:2]) for idx, _ in members}
    b.transition(t_main, adj_of[0])
    b.transition(t_aux, adj_of[1])
    # assertions in coefficient order: main (sorted), then auxiliary (sorted); divisor column = group index + 1
    groups = air.boundary_groups(pairs[2:])
    col_of = {(col, div.b): (1 + j, adj) for j, (div, adj, mem) in enumerate(groups) for col, _, _ in mem}
    first, last = 1, pow(air.g, n - 1, P)
    for col, value, div_b in ((0, air.X0, first), (2, 1, first), (2, 1, last)):
        j, adj = col_of[(col, div_b)]
        b.assertion(col, to_abi_int(value), adj, j)
    prog, keep = b.finish()
    return prog, keep


# ---- a program interpreter and a symbolic recorder (CPU side) -----------------------------------------------
def interpret_program(builder_or_nodes, consts, cur, nxt, periodic=()):
    """Node values of a transition program over one evaluation frame, in canonical big-int arithmetic: what
    air_evaluate_kernel computes per step (aero_b200/csrc/poly.cu), restated for CPU-side checks of the recorded
    programs.  `builder_or_nodes`: AirProgramBuilder.nodes (op, a, b triples); consts / frame rows / periodic values
    canonical."""
    val = []
    for op, a, b in builder_or_nodes:
        if op == 0:
            val.append(cur[a] % P)
        elif op == 1:
            val.append(nxt[a] % P)
        elif op == 2:
            val.append(consts[a] % P)
        elif op == 6:
            val.append(periodic[a] % P)
        elif op == 3:
            val.append((val[a] + val[b]) % P)
        elif op == 4:
            val.append((val[a] - val[b]) % P)
        elif op == 5:
            val.append(val[a] * val[b] % P)
        else:
            raise ValueError("unknown node op %r" % op)
    return val


class _Sym:
    """A field element that records instead of computing: the Python twin of the symbolic FieldElement a Rust
    caller runs through Air::evaluate_transition.  Integers met in the arithmetic become
    constants of the program; `% P` is the identity (the device reduces)."""

    __slots__ = ("rec", "node")

    def __init__(self, rec, node):
        self.rec, self.node = rec, node

    def _lift(self, other):
        return other if isinstance(other, _Sym) else self.rec.const(int(other))

    def __add__(self, o):
        return _Sym(self.rec, self.rec.b.add(self.node, self._lift(o).node))

    def __radd__(self, o):
        return _Sym(self.rec, self.rec.b.add(self._lift(o).node, self.node))

    def __sub__(self, o):
        return _Sym(self.rec, self.rec.b.sub(self.node, self._lift(o).node))

    def __rsub__(self, o):
        return _Sym(self.rec, self.rec.b.sub(self._lift(o).node, self.node))

    def __mul__(self, o):
        return _Sym(self.rec, self.rec.b.mul(self.node, self._lift(o).node))

    def __rmul__(self, o):
        return _Sym(self.rec, self.rec.b.mul(self._lift(o).node, self.node))

    def __mod__(self, m):
        assert m == P
        return self


class _Recorder:
    def __init__(self, builder, to_abi_int):
        self.b, self.to_abi_int, self.const_nodes = builder, to_abi_int, {}

    def const(self, v: int) -> _Sym:
        v %= P
        if v not in self.const_nodes:          # one constant node per distinct value
            self.const_nodes[v] = self.b.const(self.to_abi_int(v))
        return _Sym(self, self.const_nodes[v])


def record_program(air, to_abi_int, aux_rand_const_slots: int = 0):
    """Runs air.evaluate_transition (and evaluate_aux_transition) ONCE over symbolic frame rows and returns the
    recorded aero_air_program with the degree adjustments and assertions of the restated constraint groups --
    no hand-written node list.  The auxiliary segment's random elements become the first `aux_rand_const_slots`
    constants (value 0 until the aux_builder callback fills them in).  Returns (program, keep-alive, builder)."""
    from aero_b200 import AirProgramBuilder

    b = AirProgramBuilder()
    rec = _Recorder(b, to_abi_int)
    rand = [_Sym(rec, b.const(0)) for _ in range(aux_rand_const_slots)]   # consts[0 .. k): the random elements
    periodic = [_Sym(rec, b.periodic(b.periodic_column([to_abi_int(v) for v in col]))) for col in air.periodic_columns]
    w, wa = air.trace_width, air.aux_width
    cur = [_Sym(rec, b.cur(c)) for c in range(w + wa)]
    nxt = [_Sym(rec, b.next(c)) for c in range(w + wa)]
    outs = list(air.evaluate_transition(cur[:w], nxt[:w], periodic))
    if wa:
        outs += list(air.evaluate_aux_transition(cur[:w], nxt[:w], cur[w:], nxt[w:], periodic, rand))
    nt = len(outs)
    assert nt == len(air._all_degrees())
    pairs = [(0, 0)] * (air.num_constraint_coefficients() // 2)
    adj_of = {idx: adj for adj, members in air.transition_groups(pairs[:nt]) for idx, _ in members}
    for i, o in enumerate(outs):
        b.transition(rec.const(o).node if not isinstance(o, _Sym) else o.node, adj_of[i])
    # assertions in coefficient order (main sorted, then auxiliary sorted), each with its group's divisor column
    groups = air.boundary_groups(pairs[nt:])
    main_as = sorted(air.get_assertions(), key=lambda a: (0, a.step, a.column))
    aux_as = sorted(air.get_aux_assertions([0] * air.num_aux_rands), key=lambda a: (0, a.step, a.column)) if wa else []
    for a, off in [(a, 0) for a in main_as] + [(a, w) for a in aux_as]:
        div_b = pow(air.g, a.step, P) if a.step else 1
        (j, adj), = [(j, adj) for j, (div, adj, mem) in enumerate(groups) if div.b == div_b]
        b.assertion(a.column + off, to_abi_int(a.value), adj, 1 + j)
    prog, keep = b.finish()
    return prog, keep, b
