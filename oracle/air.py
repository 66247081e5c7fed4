"""TEST INFRASTRUCTURE (oracle): a small AIR with REAL constraint evaluations, restated from the
reference, so that the oracle's verifier can run the reference's out-of-domain consistency check --
the first reference-derived constraint on ConstraintEvaluationTable::into_poly
(winterfell/prover/src/constraints/evaluation_table.rs:166-190,330-419: divisor and exemption
handling) and on CompositionPoly's column transposition (constraints/composition_poly.rs:111-128).
Synthetic constraint columns (random data) can never fail that check's prover side silently: here a
wrong divisor convention, a missing exemption or a swapped composition column makes `verify` fail.

Restated, each with the file:line it follows (paths relative to winterfell/):
  * the Fibonacci AIR of examples/src/fibonacci/fib2 (air.rs:15-71, prover.rs:22-40) and the multiplicative
    Fibonacci AIR of examples/src/fibonacci/mulfib2 (air.rs:15-80, prover.rs:20-40: degree-2 constraints);
  * AirContext (air/src/air/context.rs:87-193): ce_blowup_factor, composition degree;
  * TransitionConstraints / TransitionConstraintGroup (air/src/air/transition/mod.rs:36-290,
    degree.rs:102-131): grouping by evaluation degree, degree adjustment, merge_evaluations;
  * BoundaryConstraints / BoundaryConstraintGroup / BoundaryConstraint (air/src/air/boundary/
    mod.rs:44-190, constraint_group.rs:44-110, constraint.rs:46-113) for single-value assertions;
  * ConstraintDivisor (air/src/air/divisor.rs:36-108);
  * the prover's ConstraintEvaluator (prover/src/constraints/evaluator.rs:54-230, boundary.rs:59-100,
    255-275, domain.rs:99-117): one column per divisor over the constraint evaluation domain;
  * the verifier's evaluate_constraints (verifier/src/evaluator.rs:14-107) and the comparison in
    verifier/src/lib.rs:248-290.
  * periodic columns: Air::get_periodic_column_polys (air/src/air/mod.rs:310-344), the prover's
    PeriodicValueTable (prover/src/constraints/periodic_table.rs:25-90), the verifier's evaluation of the
    column polynomials at x^(n / cycle) (verifier/src/evaluator.rs:27-36) and
    TransitionConstraintDegree::with_cycles (air/src/air/transition/degree.rs:57-131) -- what Miden's
    ProcessorAir needs beyond the Fibonacci examples (miden/air/src/lib.rs:117-120: hasher and bitwise
    chiplet columns); exercised by MaskedChainAir, built like examples/src/rescue/air.rs:60-116 (a cycle
    mask switching between a round function with periodic round constants and a plain step).
  * one auxiliary trace segment: Air::evaluate_aux_transition / get_aux_assertions with the segment's random
    elements (air/src/air/mod.rs:262-306), the coefficient order main transition, auxiliary transition, main
    assertions, auxiliary assertions (mod.rs:511-533), auxiliary boundary groups merged into the main group
    with the same divisor or appended (prover/src/constraints/boundary.rs:58-72) and the full evaluation
    frame of the prover (evaluator.rs:210-270) and the verifier (verifier/src/evaluator.rs:38-57,77-100);
    exercised by PermutationAir, a running-product argument of the kind Miden's auxiliary columns are
    (miden/processor/src/trace/utils.rs:153-199; examples/src/rescue_raps/air.rs:162-240 has the same shape).
  * Miden's BITWISE CHIPLET (miden/air/src/chiplets/bitwise/mod.rs:25-60,78-240 with the chiplet selector
    flags and constraints of miden/air/src/chiplets/mod.rs:100-130,178-185 and the trace rows of
    miden/processor/src/chiplets/bitwise/mod.rs:87-150,210-224) as a stand-alone AIR, BitwiseChipletAir: a
    piece of the ProcessorAir that the path's real caller evaluates, with its two periodic columns.
Only what these use is covered: single-value assertions, at most one auxiliary segment.
Pure Python big-int arithmetic: small traces only."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

P = 0xFFFFFFFF00000001
GENERATOR = 7
TWO_ADIC_ROOT = 1753635133440165772


def root_of_unity(k: int) -> int:
    return pow(TWO_ADIC_ROOT, 1 << (32 - k), P)


def inv(x: int) -> int:
    return pow(x % P, P - 2, P)


def log2(n: int) -> int:
    assert n > 0 and n & (n - 1) == 0
    return n.bit_length() - 1


@dataclass
class ConstraintDivisor:  # air/src/air/divisor.rs: (x^a - b) / prod (x - e)
    a: int
    b: int
    exemptions: List[int] = field(default_factory=list)

    def degree(self) -> int:  # divisor.rs:69-76
        return self.a - len(self.exemptions)

    def evaluate_at(self, x: int) -> int:  # divisor.rs:79-96
        num = (pow(x, self.a, P) - self.b) % P
        den = 1
        for e in self.exemptions:
            den = den * ((x - e) % P) % P
        return num * inv(den) % P


@dataclass
class Assertion:  # single-value assertions only (air/src/air/assertions/mod.rs:79-96)
    column: int
    step: int
    value: int


class SimpleAir:
    """The AIR-generic machinery (AirContext, constraint groups, prover- and verifier-side evaluation);
    subclasses supply trace_width, transition_degrees, build_trace, evaluate_transition, get_assertions."""

    trace_width = 0
    # TransitionConstraintDegree::new(d) as the int d, ::with_cycles(d, cycles) as the pair (d, [cycles])
    transition_degrees: list = []
    periodic_columns: List[List[int]] = []  # Air::get_periodic_column_values (air/src/air/mod.rs:246-248)
    num_transition_exemptions = 1  # context.rs:159
    # auxiliary segment (TraceLayout: aux_segment_widths / aux_segment_rands, air/src/air/trace_info.rs)
    aux_width = 0
    num_aux_rands = 0
    aux_transition_degrees: list = []
    aux_rand_elements: Sequence[int] = ()  # AuxTraceRandElements: set by whoever drew them (prover / verifier)

    def __init__(self, trace_length: int, result: int, blowup: int = 8):
        self.n, self.result, self.blowup = trace_length, result, blowup
        # context.rs:124-137 with degree.rs:113-131: max over constraints of
        # max(next_power_of_two(base + cycles - 1), MIN_BLOWUP_FACTOR = 2)
        self.ce_blowup = max(max(_next_pow2(b + len(cyc) - 1), 2)
                             for b, cyc in map(_degree, list(self.transition_degrees) + list(self.aux_transition_degrees)))
        assert blowup >= self.ce_blowup
        self.g = root_of_unity(log2(trace_length))
        for col in self.periodic_columns:  # air/src/air/mod.rs:319-335
            assert len(col) >= 2 and len(col) & (len(col) - 1) == 0 and len(col) <= trace_length

    # context.rs:183-193
    def ce_domain_size(self) -> int:
        return self.n * self.ce_blowup

    def composition_degree(self) -> int:
        return self.ce_domain_size() - 1

    def trace_poly_degree(self) -> int:
        return self.n - 1

    def evaluate_transition(self, cur: Sequence[int], nxt: Sequence[int], periodic: Sequence[int] = ()) -> List[int]:
        raise NotImplementedError

    # ---- periodic columns ---------------------------------------------------------------------
    def periodic_column_polys(self) -> List[List[int]]:
        """Air::get_periodic_column_polys (air/src/air/mod.rs:310-344): each column's values interpolated
        over the subgroup of order cycle_length (fft::interpolate_poly: natural-order values in,
        coefficients out)."""
        polys = []
        for col in self.periodic_columns:
            c = len(col)
            w_inv, c_inv = inv(root_of_unity(log2(c))), inv(c)
            polys.append([sum(int(v) * pow(w_inv, j * k, P) for k, v in enumerate(col)) % P * c_inv % P for j in range(c)])
        return polys

    def periodic_value_table(self) -> List[List[int]]:
        """PeriodicValueTable::new (prover/src/constraints/periodic_table.rs:25-75): every column
        polynomial evaluated over the coset offset^(n / cycle) * <w_(cycle * ce_blowup)>
        (fft::evaluate_poly_with_offset, natural order); column k of the table's row for step s is
        table[k][s % len(table[k])] (get_row, :84-90, with the per-column wrap of :66)."""
        table = []
        for poly in self.periodic_column_polys():
            c = len(poly)
            offset = pow(GENERATOR, self.n // c, P)
            w = root_of_unity(log2(c * self.ce_blowup))
            table.append([_eval_poly(poly, offset * pow(w, i, P) % P) for i in range(c * self.ce_blowup)])
        return table

    def periodic_values_at(self, x: int) -> List[int]:
        """verifier/src/evaluator.rs:27-36: poly(x^(n / cycle))."""
        return [_eval_poly(poly, pow(x, self.n // len(poly), P)) for poly in self.periodic_column_polys()]

    def get_assertions(self) -> List[Assertion]:
        raise NotImplementedError

    # auxiliary segment (air/src/air/mod.rs:262-306); column indices are relative to the segment
    def evaluate_aux_transition(self, main_cur, main_nxt, aux_cur, aux_nxt, periodic, rand) -> List[int]:
        return []

    def get_aux_assertions(self, rand) -> List[Assertion]:
        return []

    def _all_degrees(self) -> list:
        return list(self.transition_degrees) + list(self.aux_transition_degrees)

    def _num_aux_assertions(self) -> int:  # the count does not depend on the random elements (context.rs:139-157)
        return len(self.get_aux_assertions([0] * self.num_aux_rands)) if self.aux_width else 0

    def _transition_all(self, cur: Sequence[int], nxt: Sequence[int], periodic: Sequence[int]) -> List[int]:
        """Main constraints, then auxiliary ones (the order of their coefficient pairs), over a frame whose rows
        hold the main columns followed by the auxiliary ones."""
        w = self.trace_width
        t = self.evaluate_transition(cur[:w], nxt[:w], periodic)
        if self.aux_width:
            t = list(t) + list(self.evaluate_aux_transition(cur[:w], nxt[:w], cur[w:], nxt[w:], periodic, self.aux_rand_elements))
        return t

    def num_constraint_coefficients(self) -> int:
        """Field elements drawn by get_constraint_composition_coefficients (air/src/air/mod.rs:511-533):
        a pair per transition constraint (main, then auxiliary), then a pair per assertion (likewise)."""
        return 2 * (len(self._all_degrees()) + len(self.get_assertions()) + self._num_aux_assertions())

    # ---- constraint structure shared by prover and verifier ---------------------------------
    def transition_divisor(self) -> ConstraintDivisor:  # divisor.rs:36-45
        ex = [pow(self.g, s, P) for s in range(self.n - self.num_transition_exemptions, self.n)]
        return ConstraintDivisor(self.n, 1, ex)

    def transition_groups(self, coeffs: Sequence[Tuple[int, int]]):
        """transition/mod.rs:305-343: constraints grouped by evaluation degree (BTreeMap order);
        each group: (degree_adjustment, [(constraint index, (c0, c1))]).  Auxiliary constraints follow the
        main ones in index and coefficient order; the reference keeps them in a second list of groups
        (transition/mod.rs:62-79) whose merged evaluations are added to the first (:165-178), which is the
        same sum as grouping them together."""
        div_deg = self.transition_divisor().degree()
        groups: Dict[int, Tuple[int, list]] = {}
        for i, d in enumerate(self._all_degrees()):
            base, cycles = _degree(d)
            ev_deg = base * (self.n - 1) + sum((self.n // c) * (c - 1) for c in cycles)  # degree.rs:102-108
            if ev_deg not in groups:
                target = self.composition_degree() + div_deg  # transition/mod.rs:224-226
                groups[ev_deg] = (target - ev_deg, [])
            groups[ev_deg][1].append((i, coeffs[i]))
        return [groups[k] for k in sorted(groups)]

    def boundary_groups(self, coeffs: Sequence[Tuple[int, int]]):
        """boundary/mod.rs:120-190: assertions sorted by (stride, first step, column) -- the BTreeSet
        order of assertions/mod.rs:310-322 -- zipped with the coefficient pairs, grouped by
        (stride, first step), groups sorted by degree adjustment (stable).  Each group:
        (divisor, degree_adjustment, [(column, value, (c0, c1))]).
        Assertions against the auxiliary segment take the coefficient pairs after the main ones
        (boundary/mod.rs:104-106), are grouped the same way on their own, and each of their groups joins the
        main group with the same divisor or is appended (prover/src/constraints/boundary.rs:58-72); their
        column index is returned offset by the main width (a row = main columns, then auxiliary)."""
        def group(assertions, ccs, col_offset):
            assertions = sorted(assertions, key=lambda a: (0, a.step, a.column))
            groups: Dict[Tuple[int, int], Tuple[ConstraintDivisor, int, list]] = {}
            for a, cc in zip(assertions, ccs):
                key = (0, a.step)
                if key not in groups:
                    # divisor.rs:47-61 (num_steps = 1): x - g^step
                    div = ConstraintDivisor(1, pow(self.g, a.step, P) if a.step else 1, [])
                    adj = self.composition_degree() + div.degree() - self.trace_poly_degree()  # constraint_group.rs:28-30
                    groups[key] = (div, adj, [])
                groups[key][2].append((a.column + col_offset, a.value, cc))
            out = [groups[k] for k in sorted(groups)]
            out.sort(key=lambda gr: gr[1])
            return out

        main_assertions = self.get_assertions()
        out = group(main_assertions, coeffs[:len(main_assertions)], 0)
        if self.aux_width:
            rand = self.aux_rand_elements if len(self.aux_rand_elements) else [0] * self.num_aux_rands
            for g_aux in group(self.get_aux_assertions(rand), coeffs[len(main_assertions):], self.trace_width):
                same = [g_main for g_main in out if (g_main[0].a, g_main[0].b) == (g_aux[0].a, g_aux[0].b)]
                if same:
                    same[0][2].extend(g_aux[2])
                else:
                    out.append(g_aux)
        return out

    def divisors(self) -> List[ConstraintDivisor]:
        """Columns of the prover's evaluation table (evaluator.rs:66-67): transition first."""
        pairs = [(0, 0)] * (self.num_constraint_coefficients() // 2)
        nt = len(self._all_degrees())
        return [self.transition_divisor()] + [g[0] for g in self.boundary_groups(pairs[nt:])]

    @staticmethod
    def split_coefficients(flat: Sequence[int], n_transition: int):
        pairs = [(int(flat[2 * i]), int(flat[2 * i + 1])) for i in range(len(flat) // 2)]
        return pairs[:n_transition], pairs[n_transition:]

    # ---- prover side: ConstraintEvaluator::evaluate -----------------------------------------
    def evaluate_constraints_over_ce_domain(self, trace_lde: Sequence[Sequence[int]], coeffs: Sequence[int]) -> np.ndarray:
        """trace_lde: natural-order LDE columns (N = blowup * n values each); coeffs: the drawn
        composition coefficients, flat.  Returns (1 + boundary groups, ce_domain_size) merged
        evaluations, one column per divisor (evaluator.rs:121-160, 207-224; boundary.rs:59-72)."""
        t_cc, b_cc = self.split_coefficients(coeffs, len(self._all_degrees()))
        tg, bg = self.transition_groups(t_cc), self.boundary_groups(b_cc)
        ce = self.ce_domain_size()
        N = self.blowup * self.n
        lde_shift = log2(self.blowup // self.ce_blowup)
        g_ce = root_of_unity(log2(ce))
        out = np.zeros((1 + len(bg), ce), np.uint64)
        periodic = self.periodic_value_table()  # evaluator.rs:104-105
        x = GENERATOR  # domain.rs:99-101: ce_domain[step] * offset
        for step in range(ce):
            row = step << lde_shift
            cur = [int(c[row]) for c in trace_lde]
            nxt = [int(c[(row + self.blowup) % N]) for c in trace_lde]  # trace_lde.rs: next = + blowup, wrapping
            t = self._transition_all(cur, nxt, [col[step % len(col)] for col in periodic])  # evaluator.rs:183-186, 230-246
            acc = 0
            for adj, members in tg:
                xp = pow(x, adj, P)  # domain.rs:109-117: ce_domain[step*power mod ce] * offset^power = x^power
                for idx, (c0, c1) in members:
                    acc = (acc + (c0 + c1 * xp) * t[idx]) % P  # transition/mod.rs:272-283
            out[0, step] = acc
            for j, (_, adj, members) in enumerate(bg):
                xp = pow(x, adj, P)
                acc = 0
                for col, value, (c0, c1) in members:
                    acc = (acc + (c0 + c1 * xp) * ((cur[col] - value) % P)) % P  # boundary.rs:260-263
                out[1 + j, step] = acc
            x = x * g_ce % P
        return out

    # ---- verifier side: evaluate_constraints at the out-of-domain point ---------------------
    def evaluate_constraints_at(self, coeffs: Sequence[int], ood_cur: Sequence[int], ood_next: Sequence[int], z: int) -> int:
        """verifier/src/evaluator.rs:14-107."""
        t_cc, b_cc = self.split_coefficients(coeffs, len(self._all_degrees()))
        t = self._transition_all(list(ood_cur), list(ood_next), self.periodic_values_at(z))
        result = 0
        for adj, members in self.transition_groups(t_cc):  # transition/mod.rs:165-185
            xp = pow(z, adj, P)
            for idx, (c0, c1) in members:
                result = (result + (c0 + c1 * xp) * t[idx]) % P
        result = result * inv(self.transition_divisor().evaluate_at(z)) % P
        for div, adj, members in self.boundary_groups(b_cc):  # constraint_group.rs:82-108
            xp = pow(z, adj, P)
            num = 0
            for col, value, (c0, c1) in members:
                num = (num + ((ood_cur[col] - value) % P) * (c0 + c1 * xp)) % P
            result = (result + num * inv(div.evaluate_at(z))) % P
        return result


class Fib2Air(SimpleAir):
    """examples/src/fibonacci/fib2/air.rs: two registers, two terms of the sequence per row."""

    trace_width = 2
    transition_degrees = [1, 1]  # TransitionConstraintDegree::new(1) twice (air.rs:27-30)

    @staticmethod
    def build_trace(n: int) -> np.ndarray:  # prover.rs:22-40
        s0, s1 = 1, 1
        cols = np.empty((2, n), np.uint64)
        for i in range(n):
            cols[0, i], cols[1, i] = s0, s1
            s0 = (s0 + s1) % P
            s1 = (s1 + s0) % P
        return cols

    def evaluate_transition(self, cur: Sequence[int], nxt: Sequence[int], periodic: Sequence[int] = ()) -> List[int]:  # air.rs:41-58
        return [(nxt[0] - (cur[0] + cur[1])) % P, (nxt[1] - (cur[1] + nxt[0])) % P]

    def get_assertions(self) -> List[Assertion]:  # air.rs:60-70
        return [Assertion(0, 0, 1), Assertion(1, 0, 1), Assertion(1, self.n - 1, self.result)]


class MulFib2Air(SimpleAir):
    """examples/src/fibonacci/mulfib2/air.rs: the multiplicative Fibonacci sequence, two registers --
    transition constraints of degree 2 (a different evaluation degree, hence a different degree adjustment
    and a composition polynomial that really needs both of its columns)."""

    trace_width = 2
    transition_degrees = [2, 2]  # air.rs:29-32

    @staticmethod
    def build_trace(n: int) -> np.ndarray:  # prover.rs:24-37 (the trace has `length / 2` rows there; n rows here)
        s0, s1 = 1, 2
        cols = np.empty((2, n), np.uint64)
        for i in range(n):
            cols[0, i], cols[1, i] = s0, s1
            s0 = s0 * s1 % P
            s1 = s1 * s0 % P
        return cols

    def evaluate_transition(self, cur: Sequence[int], nxt: Sequence[int], periodic: Sequence[int] = ()) -> List[int]:  # air.rs:47-63
        return [(nxt[0] - cur[0] * cur[1]) % P, (nxt[1] - cur[1] * nxt[0]) % P]

    def get_assertions(self) -> List[Assertion]:  # air.rs:65-75: starts with 1, 2; register 0 ends with the result
        return [Assertion(0, 0, 1), Assertion(1, 0, 2), Assertion(0, self.n - 1, self.result)]


class MaskedChainAir(SimpleAir):
    """An AIR with periodic columns, shaped like the reference's Rescue hash-chain example
    (examples/src/rescue/air.rs:60-116, prover.rs:30-58): a periodic CYCLE MASK selects, row by row, between a
    round function that adds periodic ROUND CONSTANTS to a non-linear map of the state and a plain step that
    carries the state over.  Three registers: on rows where mask = 1
        s0' = s0^3 + s1 + ark0,   s1' = s0 * s1 + ark1      (the cube is Rescue's S-box, rescue.rs:121-135)
    and on the last row of every 4-row cycle (mask = 0) the two registers are swapped; s2 counts rows
    (enforced through a periodic multiplier that is nowhere zero, so its constraint has a periodic degree
    component but a different evaluation degree than the first two: three transition groups in all).  Cycle lengths 4
    and 8, so the table lookups wrap at different periods."""

    trace_width = 3
    CYCLE_MASK = [1, 1, 1, 0]
    ARK0 = [3, 1 << 40, 0xFFFFFFFF00000000, 17, 0x123456789ABCDEF, 5, 0, 0xDEADBEEF]
    ARK1 = [7, 0xFFFFFFFF, 2, 0xFEDCBA9876543210 % P, 11, 1 << 63, 13, 0xC0FFEE]
    periodic_columns = [CYCLE_MASK, ARK0, ARK1]
    # mask * (cubic / quadratic in the trace): with_cycles(3, [4]), with_cycles(2, [4]); ark1 * (linear):
    # with_cycles(1, [8])  (rescue/air.rs:41-46 declares its degrees the same way) -> ce_blowup 4, three groups
    transition_degrees = [(3, [4]), (2, [4]), (1, [8])]
    SEED = (42, 43)

    @classmethod
    def build_trace(cls, n: int) -> np.ndarray:
        s0, s1 = cls.SEED
        cols = np.empty((3, n), np.uint64)
        for i in range(n):
            cols[0, i], cols[1, i], cols[2, i] = s0, s1, i
            if cls.CYCLE_MASK[i % 4]:
                s0, s1 = (s0 * s0 * s0 + s1 + cls.ARK0[i % 8]) % P, (s0 * s1 + cls.ARK1[i % 8]) % P
            else:
                s0, s1 = s1, s0
        return cols

    def evaluate_transition(self, cur: Sequence[int], nxt: Sequence[int], periodic: Sequence[int] = ()) -> List[int]:
        mask, ark0, ark1 = periodic
        not_mask = (1 - mask) % P
        round0 = (nxt[0] - (cur[0] * cur[0] * cur[0] + cur[1] + ark0)) % P
        round1 = (nxt[1] - (cur[0] * cur[1] + ark1)) % P
        return [(mask * round0 + not_mask * (nxt[0] - cur[1])) % P,
                (mask * round1 + not_mask * (nxt[1] - cur[0])) % P,
                ark1 * (nxt[2] - cur[2] - 1) % P]

    def get_assertions(self) -> List[Assertion]:
        return [Assertion(0, 0, self.SEED[0]), Assertion(1, 0, self.SEED[1]), Assertion(2, 0, 0),
                Assertion(0, self.n - 1, self.result)]


class PermutationAir(SimpleAir):
    """A randomized AIR with one auxiliary column, the shape of Miden's running-product columns
    (miden/processor/src/trace/utils.rs:153-199: col[i + 1] = col[i] * multiplicand[i]) and of the
    reference's rescue_raps example (examples/src/rescue_raps/air.rs:162-240).  Main segment: x counts up in
    steps of 3 from 5; y holds the first n - 1 values of x rotated by 7 rows.  Auxiliary column p proves that
    the two multisets agree: with v(t) = alpha + beta * t for the segment's random elements (alpha, beta),
        p[0] = 1,   p[i + 1] * v(y[i]) = p[i] * v(x[i]),   p[n - 1] = 1."""

    trace_width = 2
    transition_degrees = [1]
    aux_width = 1
    num_aux_rands = 2
    aux_transition_degrees = [2]
    X0, STEP, ROT = 5, 3, 7

    @classmethod
    def build_trace(cls, n: int) -> np.ndarray:
        cols = np.zeros((2, n), np.uint64)
        for i in range(n):
            cols[0, i] = cls.X0 + cls.STEP * i
        for i in range(n - 1):
            cols[1, i] = cols[0, (i + cls.ROT) % (n - 1)]
        return cols

    @staticmethod
    def multiplicands(main: np.ndarray, rand: Sequence[int]) -> Tuple[List[int], List[int]]:
        """Row values (alpha + beta * x[i], alpha + beta * y[i]): numerators and denominators of the updates."""
        a, b = int(rand[0]), int(rand[1])
        return [(a + b * int(v)) % P for v in main[0]], [(a + b * int(v)) % P for v in main[1]]

    @classmethod
    def build_aux(cls, main: np.ndarray, rand: Sequence[int]) -> np.ndarray:  # Trace::build_aux_segment
        num, den = cls.multiplicands(main, rand)
        n = main.shape[1]
        col = np.empty((1, n), np.uint64)
        acc = 1
        for i in range(n):
            col[0, i] = acc
            acc = acc * num[i] % P * inv(den[i]) % P
        return col

    def evaluate_transition(self, cur, nxt, periodic=()):
        return [(nxt[0] - cur[0] - self.STEP) % P]

    def evaluate_aux_transition(self, main_cur, main_nxt, aux_cur, aux_nxt, periodic, rand):
        a, b = rand[0], rand[1]
        return [(aux_nxt[0] * (a + b * main_cur[1]) - aux_cur[0] * (a + b * main_cur[0])) % P]

    def get_assertions(self):
        return [Assertion(0, 0, self.X0)]

    def get_aux_assertions(self, rand):
        return [Assertion(0, 0, 1), Assertion(0, self.n - 1, 1)]


class BitwiseChipletAir(SimpleAir):
    """Miden's bitwise chiplet on its own: the 17 constraints of miden/air/src/chiplets/bitwise/mod.rs:78-240
    under the chiplet flag s0 * (1 - s1') (chiplets/mod.rs:183-185), preceded by the four chiplet-selector
    constraints that involve s0 and s1 (chiplets/mod.rs:100-130, results 0, 1, 3, 4), over the periodic columns
    k0 = 1,0,0,0,0,0,0,0 and k1 = 1,1,1,1,1,1,1,0 (bitwise/mod.rs:395-417).  Columns (CHIPLETS_OFFSET = 0,
    core/src/chiplets/mod.rs:17-86, core/src/chiplets/bitwise.rs:7-56): s0, s1, selector, a, b, a0..a3, b0..b3,
    output_prev, output.  Every row belongs to the chiplet (s0 = 1, s1 = 0); each operation fills an 8-row
    cycle, most significant 4-bit limb first (processor/src/chiplets/bitwise/mod.rs:87-150, 210-224)."""

    trace_width = 15
    S0, S1, SEL, A, B, A_BITS, B_BITS, OUT_PREV, OUT = 0, 1, 2, 3, 4, 5, 9, 13, 14
    K0 = [1, 0, 0, 0, 0, 0, 0, 0]
    K1 = [1, 1, 1, 1, 1, 1, 1, 0]
    periodic_columns = [K0, K1]
    # chiplets/mod.rs:20-23 (entries 0, 1, 3, 4), then bitwise/mod.rs:37-60
    transition_degrees = [2, 3, 2, 3,
                          4, (3, [8])] + [4] * 8 + [(3, [8])] * 4 + [(3, [8]), (3, [8]), 5]

    @staticmethod
    def operations(n: int) -> List[Tuple[int, int, int]]:
        """(selector, a, b) per 8-row cycle: AND = 0, XOR = 1 (core/src/chiplets/bitwise.rs:18-24)."""
        ops, x = [], 0x9E3779B97F4A7C15
        for k in range(n // 8):
            x = (x * 6364136223846793005 + 1442695040888963407) % (1 << 64)
            a = (x >> 32) & 0xFFFFFFFF
            x = (x * 6364136223846793005 + 1442695040888963407) % (1 << 64)
            b = (x >> 32) & 0xFFFFFFFF
            ops.append((k % 3 != 0, a, b))
        return [(int(sel), a, b) for sel, a, b in ops]

    @classmethod
    def build_trace(cls, n: int) -> np.ndarray:
        assert n >= 8
        cols = np.zeros((cls.trace_width, n), np.uint64)
        cols[cls.S0, :] = 1
        row = 0
        for sel, a, b in cls.operations(n):
            result = 0
            for bit_offset in range(28, -1, -4):  # u32and / u32xor, processor/.../bitwise/mod.rs:94-112
                cols[cls.OUT_PREV, row] = result
                av, bv = a >> bit_offset, b >> bit_offset
                cols[cls.SEL, row], cols[cls.A, row], cols[cls.B, row] = sel, av, bv  # add_bitwise_trace_row :210-224
                for i in range(4):
                    cols[cls.A_BITS + i, row] = (av >> i) & 1
                    cols[cls.B_BITS + i, row] = (bv >> i) & 1
                result = (result << 4) | (((av ^ bv) if sel else (av & bv)) & 0xF)
                cols[cls.OUT, row] = result
                row += 1
        return cols

    def evaluate_transition(self, cur, nxt, periodic=()):
        k0, k1 = periodic
        is_binary = lambda v: (v * v - v) % P                      # air/src/utils.rs:14-16
        agg = lambda r, start: sum((1 << i) * r[start + i] for i in range(4)) % P   # bitwise/mod.rs:381-391
        s0, s1 = cur[self.S0], cur[self.S1]
        out = [is_binary(s0), s0 * is_binary(s1) % P,              # chiplets/mod.rs:107-110
               s0 * (s0 - nxt[self.S0]) % P, s0 * s1 * (s1 - nxt[self.S1]) % P]   # :118-121
        flag = s0 * (1 - nxt[self.S1]) % P                         # bitwise_flag, chiplets/mod.rs:183-185
        sel = cur[self.SEL]
        out.append(flag * is_binary(sel) % P)                      # bitwise/mod.rs:100
        out.append(flag * k1 * (sel - nxt[self.SEL]) % P)          # :106
        out += [flag * is_binary(cur[self.A_BITS + i]) % P for i in range(4)]   # :134-136
        out += [flag * is_binary(cur[self.B_BITS + i]) % P for i in range(4)]   # :140-146
        first = flag * k0 % P
        out.append(first * (cur[self.A] - agg(cur, self.A_BITS)) % P)           # :151-152
        out.append(first * (cur[self.B] - agg(cur, self.B_BITS)) % P)           # :157
        trans = flag * k1 % P
        out.append(trans * (nxt[self.A] - (16 * cur[self.A] + agg(nxt, self.A_BITS))) % P)   # :162-164
        out.append(trans * (nxt[self.B] - (16 * cur[self.B] + agg(nxt, self.B_BITS))) % P)   # :169-170
        out.append(k0 * flag * cur[self.OUT_PREV] % P)             # :200
        out.append(k1 * flag * (nxt[self.OUT_PREV] - cur[self.OUT]) % P)        # :205-206
        shifted = cur[self.OUT_PREV] * 16
        a_b = [(cur[self.A_BITS + i], cur[self.B_BITS + i]) for i in range(4)]
        b_and = sum((1 << i) * a * b for i, (a, b) in enumerate(a_b))            # :231-240
        b_xor = sum((1 << i) * (a + b - 2 * a * b) for i, (a, b) in enumerate(a_b))   # :244-253
        and_flag, xor_flag = flag * (1 - sel) % P, flag * sel % P                # :195-196, 363-369
        out.append((and_flag * (cur[self.OUT] - (shifted + b_and)) + xor_flag * (cur[self.OUT] - (shifted + b_xor))) % P)  # :211-221
        return out

    def get_assertions(self):
        return [Assertion(self.S0, 0, 1), Assertion(self.OUT_PREV, 0, 0), Assertion(self.OUT, self.n - 1, self.result)]


def _degree(d) -> Tuple[int, List[int]]:
    """(base, cycles) of a TransitionConstraintDegree given as an int or a (base, cycles) pair."""
    return (d, []) if isinstance(d, int) else (d[0], list(d[1]))


def _eval_poly(poly: Sequence[int], x: int) -> int:  # polynom::eval (math/src/polynom/mod.rs:44-55), Horner
    acc = 0
    for c in reversed(poly):
        acc = (acc * x + c) % P
    return acc


def _next_pow2(x: int) -> int:
    """usize::next_power_of_two: 0 and 1 map to 1."""
    return 1 if x <= 1 else 1 << (x - 1).bit_length()


def ood_consistency_check(air: SimpleAir, coeffs: Sequence[int], ood_cur, ood_next, ood_comp: Sequence[int], z: int) -> None:
    """verifier/src/lib.rs:248-290: constraints evaluated over the OOD frame must equal
    sum_i z^i * H_i(z^m) reduced from the composition-column evaluations the prover sent."""
    lhs = air.evaluate_constraints_at(coeffs, ood_cur, ood_next, z)
    rhs = 0
    for i, v in enumerate(ood_comp):
        rhs = (rhs + pow(z, i, P) * v) % P
    if lhs != rhs:
        raise AssertionError("InconsistentOodConstraintEvaluations")
