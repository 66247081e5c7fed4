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


def fill_stage_table(logM, sigma, wM):
    M = 1 << logM
    tw = [0] * M
    for s in range(logM):
        m = 2 << s
        sg = pow(sigma, M // m, P)
        wm = pow(wM, M // m, P)
        x = sg
        for k in range(m // 2):
            tw[m // 2 + k] = x
            x = mul(x, wm)
    return tw


def dit_round(a, tw, s0, logM, R, T, RS, nthreads):
    ngroups = (1 << logM) >> R
    items = ngroups * T
    for tid in range(nthreads):
        it = tid
        while it < items:
            t, g = it % T, it // T
            low = g & ((1 << s0) - 1)
            base = ((g >> s0) << (s0 + R)) | low
            x = [a[(base + (e << s0)) * RS + t] for e in range(1 << R)]
            for q in range(R):
                for e in range(1 << R):
                    if e & (1 << q):
                        continue
                    k = low + ((e & ((1 << q) - 1)) << s0)
                    w = tw[(1 << (s0 + q)) + k]
                    u = x[e]
                    v = mul(x[e | (1 << q)], w)
                    x[e] = (u + v) % P
                    x[e | (1 << q)] = (u - v) % P
            for e in range(1 << R):
                a[(base + (e << s0)) * RS + t] = x[e]
            it += nthreads


def dit_tile(a, tw, logM, T, RS, nthreads):
    s0 = 0
    while s0 < logM:
        left = logM - s0
        if left >= 3 and left != 4:
            R = 3
        elif left >= 2:
            R = 2
        else:
            R = 1
        dit_round(a, tw, s0, logM, R, T, RS, nthreads)
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
    if logn <= single_max:
        for r, sh in enumerate(shifts):
            tw = fill_stage_table(logn, sh, w)
            a = [0] * n
            for i in range(n):
                a[bitrev(i, logn)] = int(src[i])
            dit_tile(a, tw, logn, 1, 1, nthreads)
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
            dit_tile(a, st1, log1, T, RS, nthreads)
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
            dit_tile(a, st2, log2_, T, RS, nthreads)
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
    for logn in (3, 4, 5, 7, 10):
        check(logn, 11)
    for logn in (6, 7, 9):  # force the two-pass path at small sizes (T=8 needs n1,n2 >= 8)
        check(logn, 2)
