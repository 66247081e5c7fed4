/*
 * aero_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference algorithms on the Winterfell/Miden hot path that
 * starkoracles/Aero drives (Goldilocks LDE, blake2s row commitment, DEEP composition, FRI).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library -- and only as the checker / the timed CPU baseline, never as a fallback for
 * the CUDA path.
 *
 * Parity status: PINNED.  oracle/stark_oracle.py uses these routines to (a) re-verify the
 * reference's golden proof proofs/fib.bin (committed as tests/golden/fib.bin): leaf layout,
 * Merkle batch proofs, Fiat-Shamir chain, minimum grinding nonce, query positions, DEEP values at
 * all 27 queries, both FRI folds and the remainder commitment; and (b) reproduce the coin KAT in
 * tests/integration/test_verifier.cairo:104,108.  The forward prover intermediates
 * (trace -> polys -> LDE) of fib.bin are not reproducible without the Rust VM; they are pinned
 * by the NTT == naive-evaluation property tests the reference itself uses
 * (winterfell/math/src/fft/tests.rs:17-58) and by prover->verifier self-consistency.
 *
 * The reference (Rust, nightly-2023-02-17 + un-vendored crates.io deps) cannot be built in this
 * image, so there is no oracle/_ref.  BLAKE2s lives in the third-party crate blake2-rfc = "0.2"
 * (winterfell/crypto/Cargo.toml:35, version un-pinned: no Cargo.lock); it is restated here from
 * RFC 7693 and cross-checked against Python's hashlib.blake2s in tests/.
 *
 * All field elements here are CANONICAL u64 in [0, p); the reference keeps Montgomery form in
 * memory (winterfell/math/src/field/f64/mod.rs:59-61) which yields identical canonical values
 * because every operation is exact arithmetic mod p.
 *
 * Threading: OpenMP at the reference's rayon cut points so the same
 * library is the multi-threaded "port" CPU baseline.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */

/* ---------------------------------------------------------------------------------------------
 * Field: winterfell/math/src/field/f64/mod.rs:273 (add), :293 (sub), :311 (mul), :120 (inv),
 * :103 (exp via square-and-multiply), :43,:222 (2^32-th root), :218 (GENERATOR = 7).
 * ------------------------------------------------------------------------------------------- */
static inline u64 gl_add(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a || s >= GL_P) s -= GL_P;
    return s;
}
static inline u64 gl_sub(u64 a, u64 b) { return a >= b ? a - b : a + (GL_P - b); }
static inline u64 gl_reduce128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    /* x = lo + 2^64*hi_lo + 2^96*hi_hi ; 2^64 = 2^32-1, 2^96 = -1 (mod p) */
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS; /* borrow: add p  (== subtract 2^32-1 mod 2^64) */
    u64 t1 = hi_lo * GL_EPS;
    u64 r = t0 + t1;
    if (r < t0) r += GL_EPS; /* carry: 2^64 = 2^32-1 */
    if (r >= GL_P) r -= GL_P;
    return r;
}
static inline u64 gl_mul(u64 a, u64 b) { return gl_reduce128((u128)a * b); }
u64 aero_or_gl_add(u64 a, u64 b) { return gl_add(a, b); }
u64 aero_or_gl_sub(u64 a, u64 b) { return gl_sub(a, b); }
u64 aero_or_gl_mul(u64 a, u64 b) { return gl_mul(a, b); }
u64 aero_or_gl_exp(u64 b, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_mul(b, b);
        e >>= 1;
    }
    return r;
}
u64 aero_or_gl_inv(u64 a) { return aero_or_gl_exp(a, GL_P - 2); }
/* StarkField::get_root_of_unity, winterfell/math/src/field/traits.rs:224-233 */
u64 aero_or_gl_root_of_unity(u32 k) {
    return aero_or_gl_exp(1753635133440165772ULL, 1ULL << (32 - k));
}
/* Montgomery <-> canonical at the ABI boundary: f64/mod.rs:59-61 (new), :234 (as_int), :578. */
u64 aero_or_mont_to_canon(u64 x) { /* x * 2^-64 mod p */
    static u64 rinv = 0;
    if (!rinv) rinv = aero_or_gl_inv(GL_EPS);
    return gl_mul(x % GL_P, rinv);
}
u64 aero_or_canon_to_mont(u64 x) { return gl_mul(x, GL_EPS); }
void aero_or_mont_to_canon_vec(const u64 *in, u64 *out, u64 n) {
    for (u64 i = 0; i < n; i++) out[i] = aero_or_mont_to_canon(in[i]);
}
void aero_or_canon_to_mont_vec(const u64 *in, u64 *out, u64 n) {
    for (u64 i = 0; i < n; i++) out[i] = aero_or_canon_to_mont(in[i]);
}

/* ---------------------------------------------------------------------------------------------
 * FFT: winterfell/math/src/fft/mod.rs:464 (get_twiddles), :504 (get_inv_twiddles), :598
 * (permute_index); fft/serial.rs:108 (permute), :124 (fft_in_place), :32, :70, :86.
 * ------------------------------------------------------------------------------------------- */
static inline u64 permute_index(u64 size, u64 index) {
    if (size == 1) return 0;
    int bits = __builtin_ctzll(size);
    u64 r = 0;
    for (int i = 0; i < bits; i++) r |= ((index >> i) & 1ULL) << (bits - 1 - i);
    return r;
}
void aero_or_permute(u64 *v, u64 n) {
    for (u64 i = 0; i < n; i++) {
        u64 j = permute_index(n, i);
        if (j > i) {
            u64 t = v[i];
            v[i] = v[j];
            v[j] = t;
        }
    }
}
static void power_series(u64 b, u64 s, u64 *out, u64 n) { /* math/src/utils/mod.rs:36,72 */
    u64 x = s;
    for (u64 i = 0; i < n; i++) {
        out[i] = x;
        x = gl_mul(x, b);
    }
}
void aero_or_get_twiddles(u64 n, u64 *out) {
    u64 root = aero_or_gl_root_of_unity((u32)__builtin_ctzll(n));
    power_series(root, 1, out, n / 2);
    aero_or_permute(out, n / 2);
}
void aero_or_get_inv_twiddles(u64 n, u64 *out) {
    u64 root = aero_or_gl_root_of_unity((u32)__builtin_ctzll(n));
    u64 inv_root = aero_or_gl_exp(root, n - 1);
    power_series(inv_root, 1, out, n / 2);
    aero_or_permute(out, n / 2);
}
#define MAX_LOOP 256 /* fft/serial.rs:14 */
static void fft_in_place(u64 *values, u64 len, const u64 *tw, u64 count, u64 stride, u64 offset) {
    u64 size = len / stride;
    if (size > 2) {
        if (stride == count && count < MAX_LOOP) {
            fft_in_place(values, len, tw, 2 * count, 2 * stride, offset);
        } else {
            fft_in_place(values, len, tw, count, 2 * stride, offset);
            fft_in_place(values, len, tw, count, 2 * stride, offset + stride);
        }
    }
    for (u64 o = offset; o < offset + count; o++) {
        u64 i = o, j = o + stride, t = values[i];
        values[i] = gl_add(t, values[j]);
        values[j] = gl_sub(t, values[j]);
    }
    u64 last = offset + size * stride;
    u64 idx = 1;
    for (u64 o = offset + 2 * stride; o < last; o += 2 * stride, idx++) {
        u64 w = tw[idx];
        for (u64 j0 = o; j0 < o + count; j0++) {
            u64 i = j0, j = j0 + stride, t = values[i];
            u64 vj = gl_mul(values[j], w);
            values[i] = gl_add(t, vj);
            values[j] = gl_sub(t, vj);
        }
    }
}
void aero_or_fft_in_place(u64 *values, u64 n, const u64 *tw) { fft_in_place(values, n, tw, 1, 1, 0); }
/* fft/serial.rs:21-29 evaluate_poly */
void aero_or_evaluate_poly(u64 *p, u64 n, const u64 *tw) {
    fft_in_place(p, n, tw, 1, 1, 0);
    aero_or_permute(p, n);
}
/* fft/serial.rs:70-81 interpolate_poly */
void aero_or_interpolate_poly(u64 *ev, u64 n, const u64 *inv_tw) {
    fft_in_place(ev, n, inv_tw, 1, 1, 0);
    u64 inv_len = aero_or_gl_inv(n % GL_P);
    for (u64 i = 0; i < n; i++) ev[i] = gl_mul(ev[i], inv_len);
    aero_or_permute(ev, n);
}
/* fft/serial.rs:86-103 interpolate_poly_with_offset */
void aero_or_interpolate_poly_with_offset(u64 *ev, u64 n, const u64 *inv_tw, u64 domain_offset) {
    fft_in_place(ev, n, inv_tw, 1, 1, 0);
    aero_or_permute(ev, n);
    u64 inv_off = aero_or_gl_inv(domain_offset);
    u64 off = aero_or_gl_inv(n % GL_P);
    for (u64 i = 0; i < n; i++) {
        ev[i] = gl_mul(ev[i], off);
        off = gl_mul(off, inv_off);
    }
}
/* fft/serial.rs:32-63 evaluate_poly_with_offset; out has n*blowup entries */
void aero_or_evaluate_poly_with_offset(const u64 *p, u64 n, const u64 *tw, u64 domain_offset,
                                       u64 blowup, u64 *out) {
    u64 domain_size = n * blowup;
    u64 g = aero_or_gl_root_of_unity((u32)__builtin_ctzll(domain_size));
    for (u64 i = 0; i < blowup; i++) {
        u64 *chunk = out + i * n;
        u64 idx = permute_index(blowup, i);
        u64 offset = gl_mul(aero_or_gl_exp(g, idx), domain_offset);
        u64 factor = 1;
        for (u64 j = 0; j < n; j++) {
            chunk[j] = gl_mul(p[j], factor);
            factor = gl_mul(factor, offset);
        }
        fft_in_place(chunk, n, tw, 1, 1, 0);
    }
    aero_or_permute(out, domain_size);
}

/* ---------------------------------------------------------------------------------------------
 * Matrix: winterfell/prover/src/matrix.rs:151 (interpolate_columns), :189
 * (evaluate_columns_over).  Column-major flat storage: column c at [c*rows, (c+1)*rows).
 * Parallel over columns like iter!(self.columns) under `concurrent` (matrix.rs:153,190).
 * ------------------------------------------------------------------------------------------- */
void aero_or_interpolate_columns(const u64 *trace, u64 w, u64 n, u64 *polys) {
    u64 *itw = (u64 *)malloc(sizeof(u64) * (n / 2 + 1));
    aero_or_get_inv_twiddles(n, itw);
#pragma omp parallel for schedule(dynamic, 1)
    for (long c = 0; c < (long)w; c++) {
        memcpy(polys + c * n, trace + c * n, n * sizeof(u64));
        aero_or_interpolate_poly(polys + c * n, n, itw);
    }
    free(itw);
}
void aero_or_evaluate_columns_over(const u64 *polys, u64 w, u64 n, u64 blowup, u64 offset, u64 *lde) {
    u64 *tw = (u64 *)malloc(sizeof(u64) * (n / 2 + 1));
    aero_or_get_twiddles(n, tw);
#pragma omp parallel for schedule(dynamic, 1)
    for (long c = 0; c < (long)w; c++)
        aero_or_evaluate_poly_with_offset(polys + c * n, n, tw, offset, blowup, lde + c * n * blowup);
    free(tw);
}

/* ---------------------------------------------------------------------------------------------
 * BLAKE2s-256 (RFC 7693), unkeyed, digest length 32 -- the function blake2-rfc's
 * blake2s(32, &[], bytes) computes (winterfell/crypto/src/hash/blake2s/mod.rs:15-20,140-156).
 * ------------------------------------------------------------------------------------------- */
static const u32 B2S_IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A,
                              0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
static const u8 B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
    {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4},
    {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13},
    {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11},
    {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5},
    {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
static inline u32 rotr32(u32 x, int n) { return (x >> n) | (x << (32 - n)); }
#define B2S_G(a, b, c, d, x, y)   \
    do {                          \
        a = a + b + (x);          \
        d = rotr32(d ^ a, 16);    \
        c = c + d;                \
        b = rotr32(b ^ c, 12);    \
        a = a + b + (y);          \
        d = rotr32(d ^ a, 8);     \
        c = c + d;                \
        b = rotr32(b ^ c, 7);     \
    } while (0)
static void b2s_compress(u32 h[8], const u32 m[16], u64 t, int last) {
    u32 v[16];
    for (int i = 0; i < 8; i++) {
        v[i] = h[i];
        v[i + 8] = B2S_IV[i];
    }
    v[12] ^= (u32)t;
    v[13] ^= (u32)(t >> 32);
    if (last) v[14] = ~v[14];
    for (int r = 0; r < 10; r++) {
        const u8 *s = B2S_SIGMA[r];
        B2S_G(v[0], v[4], v[8], v[12], m[s[0]], m[s[1]]);
        B2S_G(v[1], v[5], v[9], v[13], m[s[2]], m[s[3]]);
        B2S_G(v[2], v[6], v[10], v[14], m[s[4]], m[s[5]]);
        B2S_G(v[3], v[7], v[11], v[15], m[s[6]], m[s[7]]);
        B2S_G(v[0], v[5], v[10], v[15], m[s[8]], m[s[9]]);
        B2S_G(v[1], v[6], v[11], v[12], m[s[10]], m[s[11]]);
        B2S_G(v[2], v[7], v[8], v[13], m[s[12]], m[s[13]]);
        B2S_G(v[3], v[4], v[9], v[14], m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
static void b2s_init(u32 h[8]) {
    for (int i = 0; i < 8; i++) h[i] = B2S_IV[i];
    h[0] ^= 0x01010020u; /* digest_length 32, key 0, fanout 1, depth 1 */
}
void aero_or_blake2s(const u8 *msg, u64 len, u8 out[32]) {
    u32 h[8], m[16];
    b2s_init(h);
    u64 off = 0;
    while (len - off > 64) {
        memcpy(m, msg + off, 64);
        off += 64;
        b2s_compress(h, m, off, 0);
    }
    u8 last[64];
    memset(last, 0, 64);
    memcpy(last, msg + off, len - off);
    memcpy(m, last, 64);
    b2s_compress(h, m, len, 1);
    memcpy(out, h, 32);
}
/* hash_elements: crypto/src/hash/blake2s/mod.rs:52-77 -- every element is written as its 8
 * canonical LE bytes followed by 24 zero bytes (fork-specific padding, :64-69). */
void aero_or_hash_elements(const u64 *e, u64 count, u8 out[32]) {
    u32 h[8], m[16];
    b2s_init(h);
    u64 nblocks = (count + 1) / 2;
    for (u64 b = 0; b < nblocks; b++) {
        memset(m, 0, sizeof m);
        m[0] = (u32)e[2 * b];
        m[1] = (u32)(e[2 * b] >> 32);
        if (2 * b + 1 < count) {
            m[8] = (u32)e[2 * b + 1];
            m[9] = (u32)(e[2 * b + 1] >> 32);
        }
        int last = (b + 1 == nblocks);
        u64 t = last ? 32 * count : 64 * (b + 1);
        b2s_compress(h, m, t, last);
    }
    memcpy(out, h, 32);
}
/* merge: blake2s/mod.rs:37-39 ; merge_with_int: :41-46 */
void aero_or_merge(const u8 a[32], const u8 b[32], u8 out[32]) {
    u32 h[8], m[16];
    b2s_init(h);
    memcpy(m, a, 32);
    memcpy(m + 8, b, 32);
    b2s_compress(h, m, 64, 1);
    memcpy(out, h, 32);
}
void aero_or_merge_with_int(const u8 seed[32], u64 v, u8 out[32]) {
    u32 h[8], m[16];
    b2s_init(h);
    memset(m, 0, sizeof m);
    memcpy(m, seed, 32);
    m[8] = (u32)v;
    m[9] = (u32)(v >> 32);
    b2s_compress(h, m, 40, 1);
    memcpy(out, h, 32);
}

/* ---------------------------------------------------------------------------------------------
 * Row commitment: prover/src/matrix.rs:222-246 (commit_to_rows: hash every row, batches of >=128
 * rows per task) and crypto/src/merkle/mod.rs:316-340 (build_merkle_nodes; rayon variant in
 * merkle/concurrent.rs:21-70 gives identical nodes).
 * ------------------------------------------------------------------------------------------- */
void aero_or_hash_rows(const u64 *m, u64 w, u64 rows, u8 *leaves /* rows*32 */) {
#pragma omp parallel
    {
        u64 *row = (u64 *)malloc(sizeof(u64) * w);
#pragma omp for schedule(static)
        for (long r = 0; r < (long)rows; r++) {
            for (u64 c = 0; c < w; c++) row[c] = m[c * rows + r]; /* read_row_into, matrix.rs:109 */
            aero_or_hash_elements(row, w, leaves + 32 * r);
        }
        free(row);
    }
}
void aero_or_build_merkle_nodes(const u8 *leaves, u64 num_leaves, u8 *nodes /* num_leaves*32 */) {
    u64 n = num_leaves / 2;
    memset(nodes, 0, 32);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++)
        aero_or_merge(leaves + 64 * i, leaves + 64 * i + 32, nodes + 32 * (n + i));
    for (u64 lvl = n / 2; lvl >= 1; lvl /= 2) { /* nodes[lvl .. 2*lvl) from nodes[2*lvl .. 4*lvl) */
#pragma omp parallel for schedule(static) if (lvl >= 1024)
        for (long i = (long)lvl; i < (long)(2 * lvl); i++)
            aero_or_merge(nodes + 64 * i, nodes + 64 * i + 32, nodes + 32 * i);
        if (lvl == 1) break;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Polynomial helpers: math/src/polynom/mod.rs:53 (eval, Horner), :524-542 (syn_div_in_place,
 * a == 1 branch); math/src/utils/mod.rs:143 (mul_acc), :177 (batch_inversion).
 * ------------------------------------------------------------------------------------------- */
u64 aero_or_polynom_eval(const u64 *p, u64 n, u64 x) {
    u64 acc = 0;
    for (u64 i = n; i-- > 0;) acc = gl_add(gl_mul(acc, x), p[i]);
    return acc;
}
void aero_or_syn_div_in_place(u64 *p, u64 n, u64 b) { /* divide by (x - b) */
    u64 c = 0;
    for (u64 i = n; i-- > 0;) {
        u64 v = gl_add(p[i], gl_mul(b, c));
        p[i] = c;
        c = v;
    }
}
void aero_or_mul_acc(u64 *a, const u64 *b, u64 n, u64 c) {
#pragma omp parallel for schedule(static) if (n >= 65536)
    for (long i = 0; i < (long)n; i++) a[i] = gl_add(a[i], gl_mul(b[i], c));
}
void aero_or_batch_inversion(const u64 *v, u64 n, u64 *out) {
    /* serial_batch_inversion semantic: zero maps to zero (math/src/utils/mod.rs:192-218) */
    u64 last = 1;
    for (u64 i = 0; i < n; i++) {
        out[i] = last;
        if (v[i] != 0) last = gl_mul(last, v[i]);
    }
    last = aero_or_gl_inv(last);
    for (u64 i = n; i-- > 0;) {
        if (v[i] == 0) {
            out[i] = 0;
        } else {
            out[i] = gl_mul(last, out[i]);
            last = gl_mul(last, v[i]);
        }
    }
}

/* Running-product auxiliary column: miden/processor/src/trace/utils.rs:153-199 (build_aux_column) with the
 * table hints flattened to one multiplicand per row (1 where no update happens): result[0] = init,
 * result[clk + 1] = result[clk] * multiplicand(clk); rows between updates repeat the last value. */
void aero_or_build_aux_column(const u64 *multiplicands, u64 n, u64 init, u64 *out) {
    out[0] = init;
    for (u64 i = 0; i + 1 < n; i++) out[i + 1] = gl_mul(out[i], multiplicands[i]);
}

/* OOD frame: prover/src/trace/poly_table.rs:59-72 (evaluate_at / get_ood_frame) and
 * constraints/composition_poly.rs:93-96 (evaluate_at z^m): Horner per column. */
void aero_or_eval_columns_at(const u64 *polys, u64 w, u64 n, u64 x, u64 *out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (long c = 0; c < (long)w; c++) out[c] = aero_or_polynom_eval(polys + c * n, n, x);
}

/* ---------------------------------------------------------------------------------------------
 * Constraint composition: prover/src/constraints/evaluation_table.rs:166-190 (into_poly),
 * :330-380 (acc_column), :383-419 (get_inv_evaluation); composition_poly.rs:111-128 (transpose).
 * Divisor i is (x^a_i - b_i) / prod_k (x - ex_i[k])  (air/src/air/divisor.rs:14-17).
 * eval_cols: ndiv columns of N = ce_domain_size entries.  out_cols: (N/trace_len) columns of
 * trace_len coefficients (column j holds C[ncols*i + j]).
 * ------------------------------------------------------------------------------------------- */
void aero_or_constraints_into_poly(const u64 *eval_cols, u64 ndiv, const u64 *div_a, const u64 *div_b,
                                   const u64 *div_nex, const u64 *div_ex /* ndiv x 8 */, u64 N,
                                   u64 trace_len, u64 domain_offset, u64 *out_cols) {
    u64 *combined = (u64 *)calloc(N, sizeof(u64));
    u64 g = aero_or_gl_root_of_unity((u32)__builtin_ctzll(N));
    u64 *ce = (u64 *)malloc(sizeof(u64) * N); /* domain.rs:38: ce_domain = power series of g */
    power_series(g, 1, ce, N);
    for (u64 d = 0; d < ndiv; d++) {
        u64 a = div_a[d], b = div_b[d];
        u64 zn = N / a;
        u64 off_exp = aero_or_gl_exp(domain_offset, a);
        u64 *ev = (u64 *)malloc(sizeof(u64) * zn), *z = (u64 *)malloc(sizeof(u64) * zn);
        for (u64 i = 0; i < zn; i++) /* get_ce_x_power_at, domain.rs:108-116 */
            ev[i] = gl_sub(gl_mul(ce[(i * a) & (N - 1)], off_exp), b);
        aero_or_batch_inversion(ev, zn, z);
        const u64 *col = eval_cols + d * N;
        u64 nex = div_nex[d];
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)N; i++) {
            u64 zi = z[(u64)i % zn];
            if (nex) {
                u64 x = gl_mul(ce[i], domain_offset); /* get_ce_x_at */
                u64 e = 1;
                for (u64 k = 0; k < nex; k++) e = gl_mul(e, gl_sub(x, div_ex[d * 8 + k]));
                zi = gl_mul(zi, e);
            }
            combined[i] = gl_add(combined[i], gl_mul(col[i], zi));
        }
        free(ev);
        free(z);
    }
    u64 *itw = (u64 *)malloc(sizeof(u64) * (N / 2 + 1));
    aero_or_get_inv_twiddles(N, itw);
    aero_or_interpolate_poly_with_offset(combined, N, itw, domain_offset);
    u64 ncols = N / trace_len;
    for (u64 i = 0; i < N; i++) out_cols[(i % ncols) * trace_len + (i / ncols)] = combined[i];
    free(itw);
    free(ce);
    free(combined);
}

/* ---------------------------------------------------------------------------------------------
 * DEEP composition in coefficient form: prover/src/composer/mod.rs:71-181 (add_trace_polys),
 * :184-214 (add_composition_poly), :222-238 (adjust_degree), :258-276
 * (merge_trace_compositions), :280-288 (acc_trace_poly).  No extension field (Miden:
 * FieldExtension::None) so the third coefficient of every triple is drawn but unused.
 * trace_polys: w columns of n coefficients (main then aux); comp_polys: m columns of n.
 * cc_trace: w triples; cc_comp: m; cc_deg: 2.
 * ------------------------------------------------------------------------------------------- */
void aero_or_deep_compose(const u64 *trace_polys, u64 w, const u64 *comp_polys, u64 m, u64 n, u64 z,
                          const u64 *ood_z, const u64 *ood_zg, const u64 *ood_comp, const u64 *cc_trace,
                          const u64 *cc_comp, const u64 *cc_deg, u64 *out) {
    u64 g = aero_or_gl_root_of_unity((u32)__builtin_ctzll(n));
    u64 next_z = gl_mul(z, g);
    u64 *t1 = (u64 *)calloc(n, sizeof(u64)), *t2 = (u64 *)calloc(n, sizeof(u64));
    for (u64 i = 0; i < w; i++) {
        const u64 *poly = trace_polys + i * n;
        aero_or_mul_acc(t1, poly, n, cc_trace[3 * i]);
        t1[0] = gl_sub(t1[0], gl_mul(ood_z[i], cc_trace[3 * i]));
        aero_or_mul_acc(t2, poly, n, cc_trace[3 * i + 1]);
        t2[0] = gl_sub(t2[0], gl_mul(ood_zg[i], cc_trace[3 * i + 1]));
    }
    aero_or_syn_div_in_place(t1, n, z);
    aero_or_syn_div_in_place(t2, n, next_z);
    for (u64 i = 0; i < n; i++) t1[i] = gl_add(t1[i], t2[i]);
    u64 z_m = aero_or_gl_exp(z, m);
    u64 *col = (u64 *)malloc(sizeof(u64) * n);
    for (u64 j = 0; j < m; j++) {
        memcpy(col, comp_polys + j * n, n * sizeof(u64));
        col[0] = gl_sub(col[0], ood_comp[j]);
        aero_or_syn_div_in_place(col, n, z_m);
        aero_or_mul_acc(t1, col, n, cc_comp[j]);
    }
    memset(out, 0, n * sizeof(u64));
    aero_or_mul_acc(out, t1, n, cc_deg[0]);
    aero_or_mul_acc(out + 1, t1, n - 1, cc_deg[1]);
    free(col);
    free(t1);
    free(t2);
}

/* ---------------------------------------------------------------------------------------------
 * FRI layer: fri/src/prover/mod.rs:197-218 (build_layer), utils/core/src/lib.rs:557-582
 * (transpose_slice), fri/src/utils.rs:41-51 (hash_values), fri/src/folding/mod.rs:86-118
 * (apply_drp), :181-190 (get_inv_offsets).  Folding factor FF (2,4,8,16); Miden uses 8.
 * ------------------------------------------------------------------------------------------- */
void aero_or_fri_transpose(const u64 *src, u64 len, u64 ff, u64 *dst /* (len/ff) rows x ff */) {
    u64 rows = len / ff;
    for (u64 i = 0; i < rows; i++)
        for (u64 j = 0; j < ff; j++) dst[i * ff + j] = src[i + j * rows];
}
void aero_or_fri_hash_values(const u64 *transposed, u64 rows, u64 ff, u8 *leaves) {
#pragma omp parallel for schedule(static) if (rows >= 1024)
    for (long i = 0; i < (long)rows; i++) aero_or_hash_elements(transposed + i * ff, ff, leaves + 32 * i);
}
void aero_or_fri_apply_drp(const u64 *transposed, u64 rows, u64 ff, u64 domain_offset, u64 alpha,
                           u64 *out /* rows */) {
    u64 n = rows * ff;
    u64 g = aero_or_gl_root_of_unity((u32)__builtin_ctzll(n));
    u64 *offs = (u64 *)malloc(sizeof(u64) * rows), *inv_offs = (u64 *)malloc(sizeof(u64) * rows);
    power_series(g, domain_offset, offs, rows);
    aero_or_batch_inversion(offs, rows, inv_offs);
    u64 itw[16];
    aero_or_get_inv_twiddles(ff, itw);
    u64 len_offset = aero_or_gl_inv(ff);
#pragma omp parallel for schedule(static) if (rows >= 1024)
    for (long i = 0; i < (long)rows; i++) {
        u64 poly[16];
        memcpy(poly, transposed + i * ff, ff * sizeof(u64));
        fft_in_place(poly, ff, itw, 1, 1, 0); /* fft::serial_fft = fft_in_place + permute */
        aero_or_permute(poly, ff);
        u64 off = len_offset;
        for (u64 k = 0; k < ff; k++) {
            poly[k] = gl_mul(poly[k], off);
            off = gl_mul(off, inv_offs[i]);
        }
        out[i] = aero_or_polynom_eval(poly, ff, alpha);
    }
    free(offs);
    free(inv_offs);
}

/* ---------------------------------------------------------------------------------------------
 * Grinding: prover/src/channel.rs:151-167 (serial build: smallest nonce >= 1) with
 * crypto/src/random/mod.rs:112-117 (check_leading_zeros = trailing_zeros of the LE u64 head).
 * ------------------------------------------------------------------------------------------- */
u64 aero_or_grind_min_nonce(const u8 seed[32], u32 grinding_factor) {
    for (u64 nonce = 1;; nonce++) {
        u8 d[32];
        aero_or_merge_with_int(seed, nonce, d);
        u64 head;
        memcpy(&head, d, 8);
        u32 tz = head ? (u32)__builtin_ctzll(head) : 64;
        if (tz >= grinding_factor) return nonce;
    }
}

int aero_or_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void aero_or_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
