"""Development aid: a pure-Python mirror of the index arithmetic of aero_b200/csrc/ntt.cu and the
table construction in abi.cu (get_plan), checked against the CPU oracle.  Not part of the product;
lets the four-step / DIT-round bookkeeping be validated on a machine without a GPU."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle import stark_oracle as so

P = so.P
mul = lambda a, b: a * b % P


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


MAX_ROUND_LOG = 4  # ntt.cuh NTT_MAX_ROUND_LOG


def rounds_of(logM):
    """NttRounds: fewest rounds of <= 2^MAX_ROUND_LOG points, as even as possible, larger first."""
    count = max(1, (logM + MAX_ROUND_LOG - 1) // MAX_ROUND_LOG)
    base, extra = logM // count, logM % count
    return [base + (1 if i < extra else 0) for i in range(count)]


def fill_stage_table(logM, sigma, wM):
    M = 1 << logM
    tw = [0] * M
    s0 = 0
    for R in rounds_of(logM):
        m = 1 << (s0 + R)
        sg = pow(sigma, M // m, P)
        wm = pow(wM, M // m, P)
        x = sg
        for low in range(1 << s0):
            pw = x
            for rho in range(1, 1 << R):
                e = bitrev(rho, R)
                tw[(e << s0) + low] = pw
                pw = mul(pw, x)
            x = mul(x, wm)
        s0 += R
    return tw


def unit_exp(j, inv):
    u = (39 << (6 - j)) % 192
    return (192 - u) % 192 if inv else u


def need_canon(R, q, e):
    if q >= R:
        return False
    bit = 1 << q
    if e & bit:
        return (e & (bit - 1)) == 0
    return need_canon(R, q + 1, e) or need_canon(R, q + 1, e | bit)


def dit_round(a, tw, s0, logM, R, T, RS, nthreads, plain0, inv):
    """Values are tracked with a known-canonical flag to check the kernel's canonical-form
    bookkeeping: add_cc / unit-twiddle v operands must be canonical."""
    ngroups = (1 << logM) >> R
    items = ngroups * T
    for tid in range(nthreads):
        it = tid
        while it < items:
            t, g = it % T, it // T
            low = g & ((1 << s0) - 1)
            base = ((g >> s0) << (s0 + R)) | low
            x, canon = [], []
            for e in range(1 << R):
                v = a[(base + (e << s0)) * RS + t]
                if e > 0 and not plain0:
                    v = mul(v, tw[(e << s0) + low])
                    canon.append(True)
                elif need_canon(R, 0, e):
                    canon.append(True)   # canon_any
                else:
                    canon.append(False)
                x.append(v)
            for q in range(R):
                for e in range(1 << R):
                    if e & (1 << q):
                        continue
                    f = e | (1 << q)
                    ex = (unit_exp(q + 1, inv) * (e & ((1 << q) - 1))) % 192
                    canon_sum = need_canon(R, q + 1, e)
                    u = x[e]
                    if ex == 0:
                        assert canon[f], "unit twiddle needs a canonical v"
                        v = x[f]
                    else:
                        assert (ex % 96) % 32 != 0
                        v = mul(x[f], pow(2, ex % 96, P))
                    if canon_sum:
                        assert canon[e], "canonical sum needs a canonical u"
                    if ex < 96:
                        x[e], x[f] = (u + v) % P, (u - v) % P
                        canon[e], canon[f] = canon_sum, canon[e]
                    else:
                        assert not canon_sum
                        x[e], x[f] = (u - v) % P, (u + v) % P
                        canon[e], canon[f] = canon[e], False
            for e in range(1 << R):
                a[(base + (e << s0)) * RS + t] = x[e]
            it += nthreads


def dit_tile(a, tw, logM, T, RS, nthreads, plain=False, inv=False):
    s0 = 0
    for i, R in enumerate(rounds_of(logM)):
        dit_round(a, tw, s0, logM, R, T, RS, nthreads, plain and i == 0, inv)
        s0 += R


def pow_table(base, total_bits, lo_bits, hi_scale):
    lo = [pow(base, i, P) for i in range(1 << lo_bits)]
    step = pow(base, 1 << lo_bits, P)
    hi = [mul(hi_scale, pow(step, i, P)) for i in range(1 << max(0, total_bits - lo_bits))]
    return lo, hi


def root_pow(lo, hi, lo_bits, e):
    return mul(lo[e & ((1 << lo_bits) - 1)], hi[e >> lo_bits])


def out_index(i, logn, deint):
    if deint == 0:
        return i
    return (i & ((1 << deint) - 1)) * (1 << (logn - deint)) + (i >> deint)


def dft(src, logn, inverse, shifts, scale_c, post_base, deint=0, single_max=11, T=8, nthreads=64):
    """Returns list per coset of outputs; mirrors get_plan + dft_run."""
    n = 1 << logn
    w = so.root_of_unity(logn)
    if inverse:
        w = so.inv(w)
    outs = []
    plain = all(sh == 1 for sh in shifts)
    if logn <= single_max:
        for r, sh in enumerate(shifts):
            tw = fill_stage_table(logn, sh, w)
            a = [0] * n
            for i in range(n):
                a[bitrev(i, logn)] = int(src[i])
            dit_tile(a, tw, logn, 1, 1, nthreads, plain, inverse)
            o = [0] * n
            for i in range(n):
                v = a[i]
                if post_base:
                    v = mul(v, mul(scale_c, pow(post_base, i, P)))
                elif scale_c != 1:
                    v = mul(v, scale_c)
                o[out_index(i, logn, deint)] = v
            outs.append(o)
        return outs
    log2_ = logn // 2
    log1 = logn - log2_
    n1, n2 = 1 << log1, 1 << log2_
    RS = T + 1
    w1, w2 = pow(w, n2, P), pow(w, n1, P)
    lo_bits = (logn + 1) // 2
    wlo, whi = pow_table(w, logn, lo_bits, 1)
    st2 = fill_stage_table(log2_, 1, w2)
    for r, sh in enumerate(shifts):
        st1 = fill_stage_table(log1, pow(sh, n2, P), w1)
        ib = [mul(scale_c, pow(sh, j2, P)) for j2 in range(n2)]
        tmp = [0] * n
        for bx in range(n2 // T):  # pass 1 blocks
            j2_0 = bx * T
            a = [0] * (n1 * RS)
            for it in range(n1 * T):
                t, j1 = it % T, it // T
                a[bitrev(j1, log1) * RS + t] = int(src[(j1 << log2_) + j2_0 + t])
            dit_tile(a, st1, log1, T, RS, nthreads, plain, inverse)
            bd = max(T * T, (nthreads // (T * T)) * (T * T))
            for tid in range(bd):
                ii, t = tid % T, (tid // T) % T
                j2 = j2_0 + t
                cstep = bd // (T * T)
                step_i1 = cstep * T
                c = tid // (T * T)
                i1 = c * T + ii
                f = mul(ib[j2], root_pow(wlo, whi, lo_bits, i1 * j2))
                d = root_pow(wlo, whi, lo_bits, step_i1 * j2)
                while c < n1 // T:
                    tmp[(c << log2_) * T + j2 * T + ii] = mul(a[i1 * RS + t], f)
                    f = mul(f, d)
                    c += cstep
                    i1 += step_i1
        o = [0] * n
        pu = [pow(post_base, i, P) for i in range(n1)] if post_base else None
        pv = [pow(pow(post_base, n1, P), i, P) for i in range(n2)] if post_base else None
        for bx in range(n1 // T):  # pass 2 blocks
            i1_0 = bx * T
            a = [0] * (n2 * RS)
            base = (bx << log2_) * T
            for it in range(n2 * T):
                t, j2 = it % T, it // T
                a[bitrev(j2, log2_) * RS + t] = tmp[base + it]
            dit_tile(a, st2, log2_, T, RS, nthreads, True, inverse)
            for it in range(n2 * T):
                t, i2 = it % T, it // T
                v = a[i2 * RS + t]
                i1 = i1_0 + t
                if post_base:
                    v = mul(v, mul(pu[i1], pv[i2]))
                i = i1 + (i2 << log1)
                o[out_index(i, logn, deint)] = v
        outs.append(o)
    return outs


def check(logn, single_max):
    n = 1 << logn
    col = so.synthetic_trace(1, n, 0x1234 + logn)
    # iNTT
    polys = so.interpolate_columns(col)
    got = dft(col[0], logn, True, [1], so.inv(n % P), 0, single_max=single_max)[0]
    assert got == [int(v) for v in polys[0]], ("intt", logn)
    # LDE blowup 8
    lde = so.evaluate_columns_over(polys, 8)[0]
    gN = so.root_of_unity(logn + 3)
    shifts = [7 * pow(gN, r, P) % P for r in range(8)]
    outs = dft(polys[0], logn, False, shifts, 1, 0, single_max=single_max)
    for r in range(8):
        assert outs[r] == [int(lde[8 * i + r]) for i in range(n)], ("lde", logn, r)
    if logn < 3:
        print('ok logn=%d (no deinterleave case)' % logn)
        return
    # coset iNTT with deinterleave 3
    ev = so.synthetic_trace(1, n, 0x777)[0]
    itw = np.empty(n // 2, np.uint64)
    so.lib().aero_or_get_inv_twiddles(so.u64(n), so._a64(itw))
    c = ev.copy()
    so.lib().aero_or_interpolate_poly_with_offset(so._a64(c), so.u64(n), so._a64(itw), so.u64(7))
    got = dft(ev, logn, True, [1], so.inv(n % P), so.inv(7), deint=3, single_max=single_max)[0]
    want = [0] * n
    for i in range(n):
        want[(i % 8) * (n // 8) + i // 8] = int(c[i])
    assert got == want, ("cintt", logn)
    print("ok logn=%d single_max=%d" % (logn, single_max))


if __name__ == "__main__":
    for logn in (1, 2, 3, 4, 5, 6, 7, 8, 10, 11):
        check(logn, 11)
    for logn in (6, 7, 9, 12):  # force the two-pass path at small sizes (T=8 needs n1,n2 >= 8)
        check(logn, 2)
