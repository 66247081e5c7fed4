// Polynomial kernels: boundary conversions, OOD evaluation (K7), DEEP composition in coefficient
// form (K6), constraint combination (front half of K5), query gathers (K9).
#include "kernels.cuh"

namespace aero {

// ---------------------------------------------------------------------------------------------
// Montgomery <-> canonical at the ABI boundary (winterfell/math/src/field/f64/mod.rs:59-61, :234)
// ---------------------------------------------------------------------------------------------
__global__ void convert_form_kernel(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst, size_t count,
                                    int to_mont) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = to_mont ? gl::canon_to_mont(src[i]) : gl::mont_to_canon(src[i]);
}
void convert_form(const uint64_t *src, uint64_t *dst, size_t count, int to_montgomery, cudaStream_t s) {
    if (!count) return;
    size_t blocks = (count + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    AERO_COUNT_LAUNCH(1);
    convert_form_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, dst, count, to_montgomery);
}

// Field-arithmetic self test: arbitrary u64 operands (reduced first where the operation requires
// canonical inputs).  out: [mul, add, sub, mul of the raw (possibly non-canonical) operands,
// raw a * 2^S for S = 12, 24, ..., 84 (the shifts inside the NTT rounds), canonical-sum add]
__global__ void field_ops_kernel(const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, size_t n,
                                 uint64_t *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t x = gl::canon(a[i]), y = gl::canon(b[i]);
    out[i] = gl::mul(x, y);
    out[n + i] = gl::add(x, y);
    out[2 * n + i] = gl::sub(x, y);
    out[3 * n + i] = gl::canon_any(gl::mul_any(a[i], b[i]));
    out[4 * n + i] = gl::mul_pow2<12>(a[i]);
    out[5 * n + i] = gl::mul_pow2<24>(a[i]);
    out[6 * n + i] = gl::mul_pow2<36>(a[i]);
    out[7 * n + i] = gl::mul_pow2<48>(a[i]);
    out[8 * n + i] = gl::mul_pow2<60>(a[i]);
    out[9 * n + i] = gl::mul_pow2<72>(a[i]);
    out[10 * n + i] = gl::mul_pow2<84>(a[i]);
    out[11 * n + i] = gl::add_cc(x, y);
}
void field_ops(const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out, cudaStream_t s) {
    AERO_COUNT_LAUNCH(1);
    field_ops_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, b, n, out);
}

// coset-major LDE column (B cosets of n) -> natural order out[B*i + r] = lde[r*n + i]
// (the order Matrix::evaluate_columns_over returns, matrix.rs:189-201)
__global__ void lde_to_natural_kernel(const uint64_t *__restrict__ lde, uint64_t *__restrict__ out, int logn,
                                      int log_blowup, int to_mont) {
    const size_t N = (size_t)1 << (logn + log_blowup);
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (size_t)gridDim.x * blockDim.x) {
        const size_t r = k & (((size_t)1 << log_blowup) - 1), i = k >> log_blowup;
        const uint64_t v = lde[(r << logn) + i];
        out[k] = to_mont ? gl::canon_to_mont(v) : v;
    }
}
void lde_to_natural(const uint64_t *lde_cm, uint64_t *out, int logn, int log_blowup, int to_montgomery,
                    cudaStream_t s) {
    const size_t N = (size_t)1 << (logn + log_blowup);
    size_t blocks = (N + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    AERO_COUNT_LAUNCH(1);
    lde_to_natural_kernel<<<(unsigned)blocks, 256, 0, s>>>(lde_cm, out, logn, log_blowup, to_montgomery);
}

// natural order -> coset-major: out[r*n + i] = in[B*i + r]
__global__ void natural_to_coset_major_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, int logn,
                                              int log_blowup) {
    const size_t N = (size_t)1 << (logn + log_blowup);
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (size_t)gridDim.x * blockDim.x) {
        const size_t r = k & (((size_t)1 << log_blowup) - 1), i = k >> log_blowup;
        out[(r << logn) + i] = in[k];
    }
}
void natural_to_coset_major(const uint64_t *in, uint64_t *out, int logn, int log_blowup, cudaStream_t s) {
    const size_t N = (size_t)1 << (logn + log_blowup);
    size_t blocks = (N + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    AERO_COUNT_LAUNCH(1);
    natural_to_coset_major_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, out, logn, log_blowup);
}

// Transforms beyond the two-pass NTT (n = B*m > 2^24, B = 2 or 4; the 2^25 / 2^26-row end of the NTT sweep):
// one outer decimation-in-time step.  With x_rho[j'] = x[B j' + rho] and Y_rho the size-m transform of x_rho on
// the coset shift s^B,
//     X[i0 + k m] = sum_rho (s w_n^i0)^rho  w_B^(k rho)  Y_rho[i0]          (i0 < m, k < B)
// i.e. a twist by powers of t = s w_n^i0 and a B-point DFT across the B sub-transforms.  The inverse transform
// is the same with the inverse roots, no shift, and 1/B folded into `scale`.
__global__ void __launch_bounds__(256) large_deinterleave_kernel(const uint64_t *__restrict__ x, uint64_t *__restrict__ out,
                                                                 int logm, int logB) {
    const size_t n = (size_t)1 << (logm + logB);
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
        out[((j & (((size_t)1 << logB) - 1)) << logm) + (j >> logB)] = x[j];
}
void large_deinterleave(const uint64_t *x, uint64_t *out, int logm, int logB, cudaStream_t s) {
    AERO_COUNT_LAUNCH(1);
    large_deinterleave_kernel<<<148 * 16, 256, 0, s>>>(x, out, logm, logB);
}
// Y: [B][ncosets][m] ; dst: coset q at dst + q * n ; wn: powers of the n-th root (inverse root for INV) ;
// shifts: per-coset s (nullptr = 1) ; w4: primitive 4th root for the direction (B = 4)
template <int LOGB>
__global__ void __launch_bounds__(256) large_combine_kernel(const uint64_t *__restrict__ Y, uint64_t *__restrict__ dst,
                                                            PowTable wn, const uint64_t *__restrict__ shifts, uint64_t scale,
                                                            uint64_t w4, int logm, int ncosets) {
    constexpr int B = 1 << LOGB;
    const size_t m = (size_t)1 << logm, n = m << LOGB;
    const int q = blockIdx.y;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= m) return;
    uint64_t t = pow_lookup(wn, (uint32_t)i0);
    if (shifts) t = gl::mul(t, shifts[q]);
    uint64_t z[B], pw = scale;
#pragma unroll
    for (int r = 0; r < B; r++) {
        z[r] = gl::mul(Y[((size_t)r * ncosets + q) * m + i0], pw);
        pw = gl::mul(pw, t);
    }
    uint64_t *o = dst + (size_t)q * n + i0;
    if (B == 2) {
        o[0] = gl::add(z[0], z[1]);
        o[m] = gl::sub(z[0], z[1]);
    } else {  // out_k = sum_r z_r w4^(k r)
        const uint64_t a0 = gl::add(z[0], z[2]), a1 = gl::sub(z[0], z[2]);
        const uint64_t b0 = gl::add(z[1], z[3]), b1 = gl::mul(gl::sub(z[1], z[3]), w4);
        o[0] = gl::add(a0, b0);
        o[m] = gl::add(a1, b1);
        o[2 * m] = gl::sub(a0, b0);
        o[3 * m] = gl::sub(a1, b1);
    }
}
void large_combine(const uint64_t *Y, uint64_t *dst, PowTable wn, const uint64_t *d_shifts, uint64_t scale, uint64_t w4, int logm,
                   int logB, int ncosets, cudaStream_t s) {
    const size_t m = (size_t)1 << logm;
    dim3 g((unsigned)((m + 255) / 256), ncosets);
    AERO_COUNT_LAUNCH(1);
    if (logB == 1) large_combine_kernel<1><<<g, 256, 0, s>>>(Y, dst, wn, d_shifts, scale, w4, logm, ncosets);
    else large_combine_kernel<2><<<g, 256, 0, s>>>(Y, dst, wn, d_shifts, scale, w4, logm, ncosets);
}

// Last step of a size-N = B*n coset interpolation done as B size-n interpolations (the route taken
// when N exceeds the two-pass NTT, i.e. for traces above 2^21 rows).  With a_r = plain inverse DFT
// (scale 1/n) of the evaluations on coset r (x = s_r w_n^i, s_r = offset g_N^r), the residue of f
// modulo x^n - s_r^n has coefficients s_r^-j a_r[j]; writing f = sum_k F_k x^(kn) gives
//   s_r^-j a_r[j] = sum_k F_k[j] (offset^n)^k w_B^(rk)
// so F_k[j] = offset^(-nk)/B * sum_r w_B^(-rk) s_r^-j a_r[j]: a B-point inverse DFT across cosets.
// M[k*B + r] = offset^(-nk)/B * w_B^(-rk).  Coefficient c = k*n + j goes to composition column
// c mod B, row c div B (CompositionPoly::transpose, composition_poly.rs:111-128).
template <int LOGB>
__global__ void __launch_bounds__(256) coset_interp_combine_kernel(const uint64_t *__restrict__ a, PowTable ginv,
                                                                   PowTable oinv, const uint64_t *__restrict__ M,
                                                                   int logn, uint64_t *__restrict__ polys) {
    constexpr int B = 1 << LOGB;
    const size_t n = (size_t)1 << logn;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t step = pow_lookup(ginv, (uint32_t)j);  // g_N^-j
    uint64_t t = pow_lookup(oinv, (uint32_t)j);           // offset^-j
    uint64_t y[B];
#pragma unroll
    for (int r = 0; r < B; r++) {
        y[r] = gl::mul(a[((size_t)r << logn) + j], t);
        t = gl::mul(t, step);
    }
#pragma unroll
    for (int k = 0; k < B; k++) {
        uint64_t acc = 0;
#pragma unroll
        for (int r = 0; r < B; r++) acc = gl::add(acc, gl::mul(y[r], __ldg(M + k * B + r)));
        const size_t c = ((size_t)k << logn) + j;
        polys[((c & (B - 1)) << logn) + (c >> LOGB)] = acc;
    }
}
bool coset_interp_combine(const uint64_t *a, PowTable ginv, PowTable oinv, const uint64_t *M, int logn, int log_b,
                          uint64_t *polys, cudaStream_t s) {
    const size_t n = (size_t)1 << logn;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    AERO_COUNT_LAUNCH(1);
    switch (log_b) {
        case 1: coset_interp_combine_kernel<1><<<blocks, 256, 0, s>>>(a, ginv, oinv, M, logn, polys); return true;
        case 2: coset_interp_combine_kernel<2><<<blocks, 256, 0, s>>>(a, ginv, oinv, M, logn, polys); return true;
        case 3: coset_interp_combine_kernel<3><<<blocks, 256, 0, s>>>(a, ginv, oinv, M, logn, polys); return true;
        case 4: coset_interp_combine_kernel<4><<<blocks, 256, 0, s>>>(a, ginv, oinv, M, logn, polys); return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// Block-wide helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t block_sum(uint64_t v, uint64_t *sm /* >= 32 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = gl::add(v, __shfl_down_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0ULL;
    if (wid == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = gl::add(v, __shfl_down_sync(0xffffffffu, v, o));
    }
    return v;  // valid in thread 0
}

// ---------------------------------------------------------------------------------------------
// K7: OOD evaluation.  TracePolyTable::get_ood_frame (prover/src/trace/poly_table.rs:59-72) and
// CompositionPoly::evaluate_at (constraints/composition_poly.rs:93-96) are Horner evaluations
// (math/src/polynom/mod.rs:53-62); here sum_j p[j] x^j is split into chunks of CH coefficients:
// thread t takes j = j0 + t + 256*l (coalesced), Horner in x^256, times x^t, block-reduced, times
// x^j0; a second tiny kernel adds the chunk partials.
// d_tab layout per point p (stride tab_stride): [0..255] x^t ; [256] x^256 ; [257 + chunk] x^(chunk*CH) ;
// [257 + nchunks + l] x^(256 l), l < 16
// ---------------------------------------------------------------------------------------------
constexpr int OOD_THREADS = 256;
constexpr int OOD_L = 16;
__global__ void __launch_bounds__(OOD_THREADS) ood_partial_kernel(const uint64_t *__restrict__ polys, size_t col_stride,
                                                                  int logn, const uint64_t *__restrict__ d_tab,
                                                                  int tab_stride, int npoints, int chunk_len,
                                                                  uint64_t *__restrict__ partial) {
    __shared__ uint64_t sm[32];
    const int chunk = blockIdx.x, col = blockIdx.y, nchunks = gridDim.x;
    const uint64_t *p = polys + (size_t)col * col_stride + (size_t)chunk * chunk_len;
    const int t = threadIdx.x;
    const int L = chunk_len / OOD_THREADS > 0 ? chunk_len / OOD_THREADS : 1;
    uint64_t c[OOD_L];
#pragma unroll
    for (int l = 0; l < OOD_L; l++) c[l] = (l < L && t + OOD_THREADS * l < chunk_len) ? __ldg(p + t + OOD_THREADS * l) : 0ULL;
    for (int pt = 0; pt < npoints; pt++) {
        const uint64_t *tab = d_tab + (size_t)pt * tab_stride;
        const uint64_t *xp = tab + 257 + nchunks;  // x^(256 l), l < OOD_L
        gl::Acc160 dot;                            // sum_l c[l] x^(256 l), reduced once
#pragma unroll
        for (int l = 0; l < OOD_L; l++) dot.mac(c[l], __ldg(xp + l));
        uint64_t acc = gl::mul(dot.reduce(), tab[t]);
        acc = block_sum(acc, sm);
        if (t == 0) partial[((size_t)col * npoints + pt) * nchunks + chunk] = gl::mul(acc, tab[257 + chunk]);
    }
}
__global__ void ood_final_kernel(const uint64_t *__restrict__ partial, int nchunks, int total,
                                 uint64_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint64_t acc = 0;
    for (int c = 0; c < nchunks; c++) acc = gl::add(acc, partial[(size_t)i * nchunks + c]);
    out[i] = acc;
}
static int ood_chunk_len(int logn) {
    const int n = 1 << logn;
    return n < OOD_THREADS * OOD_L ? n : OOD_THREADS * OOD_L;
}
size_t ood_scratch_elems(int ncols, int logn, int npoints) {
    const int nchunks = (1 << logn) / ood_chunk_len(logn);
    return (size_t)ncols * npoints * nchunks;
}
// d_points here is the device table described above (built by the host driver).
void ood_eval(const uint64_t *polys, size_t col_stride, int ncols, int logn, const uint64_t *d_tab, int npoints,
              uint64_t *d_out, uint64_t *d_scratch, cudaStream_t s) {
    const int chunk_len = ood_chunk_len(logn);
    const int nchunks = (1 << logn) / chunk_len;
    const int tab_stride = 257 + nchunks + OOD_L;
    dim3 g(nchunks, ncols);
    AERO_COUNT_LAUNCH(2);
    ood_partial_kernel<<<g, OOD_THREADS, 0, s>>>(polys, col_stride, logn, d_tab, tab_stride, npoints, chunk_len,
                                                 d_scratch);
    const int total = ncols * npoints;
    ood_final_kernel<<<(total + 127) / 128, 128, 0, s>>>(d_scratch, nchunks, total, d_out);
}

// ---------------------------------------------------------------------------------------------
// K6: DEEP composition, coefficient form (prover/src/composer/mod.rs:71-238).
//   t1 = sum_i cc_i.0 * T_i  - [x^0] sum_i cc_i.0 * T_i(z)         (acc_trace_poly, :280-288)
//   t2 = sum_i cc_i.1 * T_i  - [x^0] sum_i cc_i.1 * T_i(z g)
//   h  = sum_j cc'_j  * H_j  - [x^0] sum_j cc'_j  * H_j(z^m)       (add_composition_poly, :184-214;
//        division is linear, so the m divisions by (x - z^m) collapse into one)
// then t1/(x-z) + t2/(x-zg) + h/(x-z^m) (syn_div_in_place, math/src/polynom/mod.rs:535-542), and
// out[j] = d0*c[j] + d1*c[j-1] (adjust_degree, :222-238).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) deep_accumulate_kernel(DeepSegs segs, const uint64_t *__restrict__ cp, int m,
                                                              uint32_t n, uint32_t j_begin, uint32_t j_count,
                                                              const uint64_t *__restrict__ cc,
                                                              const uint64_t *__restrict__ consts,
                                                              uint64_t *__restrict__ t1, uint64_t *__restrict__ t2,
                                                              uint64_t *__restrict__ h) {
    const uint32_t jl = blockIdx.x * blockDim.x + threadIdx.x;
    if (jl >= j_count) return;
    const uint32_t j = j_begin + jl;
    // three dot products over the columns, reduced once each (gl::Acc160)
    gl::Acc160 s1, s2, sh;
    int i = 0;  // running trace-column index across segments (composer/mod.rs:96-99)
    for (int s = 0; s < segs.nseg; s++) {
        const uint64_t *tp = segs.p[s] + j;
#pragma unroll 4
        for (int c = 0; c < segs.ncols[s]; c++, i++) {
            const uint64_t v = __ldg(tp + (size_t)c * n);
            s1.mac(v, __ldg(cc + 2 * i));
            s2.mac(v, __ldg(cc + 2 * i + 1));
        }
    }
    for (int c = 0; c < m; c++) sh.mac(__ldg(cp + (size_t)c * n + j), __ldg(cc + 2 * i + c));
    uint64_t a1 = s1.reduce(), a2 = s2.reduce(), ah = sh.reduce();
    if (j == 0) {
        a1 = gl::sub(a1, consts[0]);
        a2 = gl::sub(a2, consts[1]);
        ah = gl::sub(ah, consts[2]);
    }
    t1[j] = a1;
    t2[j] = a2;
    h[j] = ah;
}
void deep_accumulate(const DeepSegs &segs, const uint64_t *comp_polys, int m, int logn, uint32_t j_begin, uint32_t j_count,
                     const uint64_t *d_cc, const uint64_t *d_consts, uint64_t *t1, uint64_t *t2, uint64_t *h,
                     cudaStream_t s) {
    const uint32_t n = 1u << logn;
    if (!j_count) return;
    AERO_COUNT_LAUNCH(1);
    deep_accumulate_kernel<<<(j_count + 255) / 256, 256, 0, s>>>(segs, comp_polys, m, n, j_begin, j_count, d_cc, d_consts,
                                                               t1, t2, h);
}

// Synthetic division by (x - b) as a suffix recurrence c <- p[i] + b*c, q[i] = previous c.
// Blocks own chunks of SYN_T*SYN_L coefficients; a thread owns SYN_L consecutive ones.
//   pass A: chunk totals Tot[ch] = sum_{k in chunk} p[k] b^(k-lo)
//   pass B: CarryIn[ch] = sum_{ch' > ch} Tot[ch'] b^(C (ch'-ch-1))          (one thread per polynomial)
//   pass C: thread totals + block suffix scan (carry appended) -> rerun the recurrence, write q.
constexpr int SYN_T = 256;
constexpr int SYN_L = 16;
struct SynParams {
    uint64_t *p[3];
    uint64_t b[3];
    uint64_t bl_pow[3][9];  // b^(L * 2^s), s = 0..8
    uint64_t bc[3];         // b^(chunk_len)
    uint32_t n, chunk_len, L;
};
// inclusive suffix scan over SYN_T+1 values with multiplier b^L: S[t] = v[t] + b^L * S[t+1]
__device__ __forceinline__ void block_suffix_scan(uint64_t *sv /* SYN_T+1 */, const uint64_t *blp) {
    const int t = threadIdx.x;
    int s = 0;
    for (int d = 1; d <= SYN_T; d <<= 1, s++) {
        __syncthreads();
        uint64_t add0 = 0;
        if (t + d <= SYN_T) add0 = gl::mul(sv[t + d], blp[s]);
        __syncthreads();
        if (t + d <= SYN_T) sv[t] = gl::add(sv[t], add0);
    }
    __syncthreads();
}
template <bool APPLY>
__global__ void __launch_bounds__(SYN_T) syn_div_kernel(SynParams sp, uint64_t *__restrict__ tot /* [3][nchunks] */,
                                                        const uint64_t *__restrict__ carry_in /* [3][nchunks] */) {
    __shared__ uint64_t sv[SYN_T + 1];
    const int which = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    uint64_t *p = sp.p[which] + (size_t)chunk * sp.chunk_len;
    const uint64_t b = sp.b[which];
    const int t = threadIdx.x;
    const uint32_t L = sp.L;
    const uint32_t start = t * L;
    uint64_t vals[SYN_L];
    uint64_t c = 0;
    const bool active = start < sp.chunk_len;
#pragma unroll
    for (int l = SYN_L - 1; l >= 0; l--) {
        vals[l] = (active && (uint32_t)l < L) ? p[start + l] : 0ULL;
        if (active && (uint32_t)l < L) c = gl::add(vals[l], gl::mul(b, c));
    }
    sv[t] = active ? c : 0ULL;
    if (t == 0) sv[SYN_T] = APPLY ? carry_in[(size_t)which * nchunks + chunk] : 0ULL;
    // inactive threads (chunk shorter than SYN_T*L) hold zero totals but still shift the carry by
    // b^L per slot; chunk_len is always SYN_T*L or (n < SYN_T*SYN_L) n with L = 1 and n >= SYN_T?
    // -> the launcher guarantees chunk_len == SYN_T * L.
    block_suffix_scan(sv, sp.bl_pow[which]);
    if (!APPLY) {
        if (t == 0) tot[(size_t)which * nchunks + chunk] = sv[0];
        return;
    }
    c = sv[t + 1];
#pragma unroll
    for (int l = SYN_L - 1; l >= 0; l--) {
        if (active && (uint32_t)l < L) {
            const uint64_t v = gl::add(vals[l], gl::mul(b, c));
            p[start + l] = c;
            c = v;
        }
    }
}
__global__ void syn_carry_kernel(const uint64_t *__restrict__ tot, uint64_t *__restrict__ carry_in, int nchunks,
                                 SynParams sp) {
    const int which = threadIdx.x;
    if (which >= 3) return;
    uint64_t c = 0;
    for (int ch = nchunks - 1; ch >= 0; ch--) {
        carry_in[(size_t)which * nchunks + ch] = c;
        c = gl::add(tot[(size_t)which * nchunks + ch], gl::mul(sp.bc[which], c));
    }
}
void syn_div3(uint64_t *t1, uint64_t *t2, uint64_t *h, int logn, const uint64_t b[3], uint64_t *d_carry,
              cudaStream_t s) {
    SynParams sp;
    sp.n = 1u << logn;
    // chunk_len == SYN_T * L always; small inputs shrink L and, below SYN_T, pad with a single
    // chunk handled by L = 1 and inactive threads contributing zero (their slots still carry
    // powers of b^L, which is harmless because everything to their right is zero).
    if (sp.n >= (uint32_t)SYN_T * SYN_L) {
        sp.L = SYN_L;
    } else if (sp.n >= (uint32_t)SYN_T) {
        sp.L = sp.n / SYN_T;
    } else {
        sp.L = 1;
    }
    sp.chunk_len = sp.n >= (uint32_t)SYN_T ? SYN_T * sp.L : sp.n;
    const int nchunks = sp.n / sp.chunk_len;
    uint64_t *ps[3] = {t1, t2, h};
    for (int i = 0; i < 3; i++) {
        sp.p[i] = ps[i];
        sp.b[i] = b[i];
        uint64_t bl = gl::pow(b[i], sp.L);
        for (int k = 0; k < 9; k++) {
            sp.bl_pow[i][k] = bl;
            bl = gl::mul(bl, bl);
        }
        // the carry crosses SYN_T thread slots of L coefficients each, even when n < SYN_T
        sp.bc[i] = gl::pow(b[i], (uint64_t)SYN_T * sp.L);
    }
    uint64_t *tot = d_carry, *carry = d_carry + 3 * (size_t)nchunks;
    dim3 g(nchunks, 3);
    AERO_COUNT_LAUNCH(3);
    syn_div_kernel<false><<<g, SYN_T, 0, s>>>(sp, tot, carry);
    syn_carry_kernel<<<1, 32, 0, s>>>(tot, carry, nchunks, sp);
    syn_div_kernel<true><<<g, SYN_T, 0, s>>>(sp, tot, carry);
}

__global__ void deep_finish_kernel(const uint64_t *__restrict__ q1, const uint64_t *__restrict__ q2,
                                   const uint64_t *__restrict__ q3, uint32_t n, uint64_t d0, uint64_t d1,
                                   uint64_t *__restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t c = gl::add(gl::add(q1[j], q2[j]), q3[j]);
    uint64_t r = gl::mul(c, d0);
    if (j > 0) {
        const uint64_t cm = gl::add(gl::add(q1[j - 1], q2[j - 1]), q3[j - 1]);
        r = gl::add(r, gl::mul(cm, d1));
    }
    out[j] = r;
}
void deep_finish(const uint64_t *t1, const uint64_t *t2, const uint64_t *h, int logn, uint64_t d0, uint64_t d1,
                 uint64_t *out, cudaStream_t s) {
    const uint32_t n = 1u << logn;
    AERO_COUNT_LAUNCH(1);
    deep_finish_kernel<<<(n + 255) / 256, 256, 0, s>>>(t1, t2, h, n, d0, d1, out);
}

// ---------------------------------------------------------------------------------------------
// Running-product auxiliary columns.  Every auxiliary column of the Miden trace is
//   col[0] = init,  col[i + 1] = col[i] * m[i]
// with m[i] the multiplicand of the table update at row i, 1 where nothing happens
// (miden/processor/src/trace/utils.rs:153-199 build_aux_column; decoder / stack / range / hasher / chiplets
// columns all go through it, processor/src/trace/mod.rs:188-249): an exclusive prefix product, done here as
// a blocked scan -- chunk products, a carry per chunk, then the prefix inside each chunk.
// ---------------------------------------------------------------------------------------------
constexpr int RP_T = 256, RP_L = 16;  // threads per block, rows per thread: chunks of 4096 rows
// exclusive scan (product) of one value per thread; returns this thread's prefix, *total = block product
__device__ __forceinline__ uint64_t block_exclusive_product(uint64_t v, uint64_t *sm /* RP_T */, uint64_t *total) {
    const int t = threadIdx.x;
    sm[t] = v;
    __syncthreads();
    for (int d = 1; d < RP_T; d <<= 1) {
        const uint64_t x = t >= d ? gl::mul(sm[t], sm[t - d]) : sm[t];
        __syncthreads();
        sm[t] = x;
        __syncthreads();
    }
    const uint64_t excl = t ? sm[t - 1] : 1ULL;
    *total = sm[RP_T - 1];
    __syncthreads();
    return excl;
}
template <bool APPLY>
__global__ void __launch_bounds__(RP_T) running_product_kernel(const uint64_t *__restrict__ m, size_t m_stride, uint64_t n,
                                                               int mont, uint64_t *__restrict__ tot,
                                                               const uint64_t *__restrict__ carry,
                                                               uint64_t *__restrict__ out, size_t out_stride) {
    __shared__ uint64_t sm[RP_T];
    const int col = blockIdx.y, nchunks = gridDim.x;
    const uint64_t i0 = ((uint64_t)blockIdx.x * RP_T + threadIdx.x) * RP_L;
    const uint64_t *mc = m + (size_t)col * m_stride;
    uint64_t p[RP_L];  // p[l] = product of this thread's multiplicands before row i0 + l
    uint64_t acc = 1;
#pragma unroll
    for (int l = 0; l < RP_L; l++) {
        p[l] = acc;
        const uint64_t i = i0 + l;
        if (i + 1 < n) {  // the multiplicand of the last row is never used
            uint64_t v = mc[i];
            v = mont ? gl::mont_to_canon(v) : gl::canon(v);
            acc = gl::mul(acc, v);
        }
    }
    uint64_t total;
    const uint64_t excl = block_exclusive_product(acc, sm, &total);
    if (!APPLY) {
        if (threadIdx.x == 0) tot[(size_t)col * nchunks + blockIdx.x] = total;
        return;
    }
    const uint64_t base = gl::mul(carry[(size_t)col * nchunks + blockIdx.x], excl);
    uint64_t *oc = out + (size_t)col * out_stride;
#pragma unroll
    for (int l = 0; l < RP_L; l++) {
        const uint64_t i = i0 + l;
        if (i < n) {
            const uint64_t v = gl::mul(base, p[l]);
            oc[i] = mont ? gl::canon_to_mont(v) : v;
        }
    }
}
// carry[col][chunk] = init[col] * prod_{ch' < chunk} tot[col][ch']  (one thread per column)
__global__ void running_product_carry_kernel(const uint64_t *__restrict__ tot, const uint64_t *__restrict__ init, int mont,
                                             int ncols, int nchunks, uint64_t *__restrict__ carry) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    uint64_t c = mont ? gl::mont_to_canon(init[col]) : gl::canon(init[col]);
    for (int ch = 0; ch < nchunks; ch++) {
        carry[(size_t)col * nchunks + ch] = c;
        c = gl::mul(c, tot[(size_t)col * nchunks + ch]);
    }
}
size_t running_product_scratch_elems(int ncols, uint64_t n) {
    const uint64_t chunk = (uint64_t)RP_T * RP_L;
    return (size_t)2 * ncols * ((n + chunk - 1) / chunk);
}
void running_product(const uint64_t *m, size_t m_stride, const uint64_t *d_init, int ncols, uint64_t n, int montgomery,
                     uint64_t *out, size_t out_stride, uint64_t *scratch, cudaStream_t s) {
    const uint64_t chunk = (uint64_t)RP_T * RP_L;
    const int nchunks = (int)((n + chunk - 1) / chunk);
    uint64_t *tot = scratch, *carry = scratch + (size_t)ncols * nchunks;
    dim3 g(nchunks, ncols);
    AERO_COUNT_LAUNCH(3);
    running_product_kernel<false><<<g, RP_T, 0, s>>>(m, m_stride, n, montgomery, tot, carry, out, out_stride);
    running_product_carry_kernel<<<(ncols + 31) / 32, 32, 0, s>>>(tot, d_init, montgomery, ncols, nchunks, carry);
    running_product_kernel<true><<<g, RP_T, 0, s>>>(m, m_stride, n, montgomery, tot, carry, out, out_stride);
}
// out[i] = 1 / v[i], zero -> zero (math::batch_inversion, winterfell/math/src/utils/mod.rs:218-238; the lookup-table
// row inverses of processor/src/trace/utils.rs:build_lookup_table_row_values are this over non-zero values)
__global__ void __launch_bounds__(256) batch_inverse_kernel(const uint64_t *__restrict__ v_in, uint64_t count, int mont,
                                                            uint64_t *__restrict__ out) {
    constexpr int NB = 32;
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * NB;
    if (i0 >= count) return;
    uint64_t v[NB], pre[NB];
    uint64_t acc = 1;
#pragma unroll
    for (int k = 0; k < NB; k++) {
        const uint64_t i = i0 + k;
        v[k] = 0;
        if (i < count) v[k] = mont ? gl::mont_to_canon(v_in[i]) : gl::canon(v_in[i]);
        pre[k] = acc;
        if (v[k]) acc = gl::mul(acc, v[k]);
    }
    acc = gl::inv(acc);
#pragma unroll
    for (int k = NB - 1; k >= 0; k--) {
        const uint64_t i = i0 + k;
        uint64_t r = 0;
        if (v[k]) {
            r = gl::mul(acc, pre[k]);
            acc = gl::mul(acc, v[k]);
        }
        if (i < count) out[i] = mont ? gl::canon_to_mont(r) : r;
    }
}
void batch_inverse(const uint64_t *v, uint64_t count, int montgomery, uint64_t *out, cudaStream_t s) {
    if (!count) return;
    const uint64_t threads = (count + 31) / 32;
    AERO_COUNT_LAUNCH(1);
    batch_inverse_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(v, count, montgomery, out);
}

// ---------------------------------------------------------------------------------------------
// K9: query gathers (prover/src/trace/commitment.rs:115-140, constraints/commitment.rs:54-70,
// fri/src/prover/mod.rs:282-302, crypto/src/merkle/mod.rs:188-250 for the node list)
// ---------------------------------------------------------------------------------------------
// Every opened row / node has one owning rank, which writes it into the result buffer of every rank.
__global__ void gather_rows_kernel(const uint64_t *__restrict__ lde, size_t col_stride, int ncols, int logn,
                                   int log_blowup, int coset_begin, int coset_count, int G,
                                   const uint32_t *__restrict__ pos, int npos, RankPtrs out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npos * ncols) return;
    const int p = i / ncols, c = i - p * ncols;
    const uint32_t k = pos[p];
    const int q = (int)(k & ((1u << log_blowup) - 1)) - coset_begin;
    if (q < 0 || q >= coset_count) return;  // another rank stores this coset
    const size_t rho = ((size_t)q << logn) + (k >> log_blowup);
    const uint64_t v = lde[(size_t)c * col_stride + rho];
    for (int r = 0; r < G; r++) reinterpret_cast<uint64_t *>(out.p[r])[i] = v;
}
void gather_rows(const uint64_t *lde_cm, size_t col_stride, int ncols, int logn, int log_blowup, int coset_begin,
                 int coset_count, int G, const uint32_t *d_positions, int npos, const RankPtrs &out, cudaStream_t s) {
    const int total = npos * ncols;
    AERO_COUNT_LAUNCH(1);
    gather_rows_kernel<<<(total + 127) / 128, 128, 0, s>>>(lde_cm, col_stride, ncols, logn, log_blowup, coset_begin,
                                                           coset_count, G, d_positions, npos, out);
}
// idx: heap indices of the WHOLE tree (leaf k = N + k, internal node j, root 1)
__global__ void gather_tree_digests_kernel(SegTreeView t, const uint32_t *__restrict__ idx, int count, RankPtrs out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count * 8) return;
    const uint32_t g = idx[i >> 3], w = i & 7;
    const int logN = t.logn + t.log_blowup, log_nb = t.logn - t.logG;
    const int G = 1 << t.logG;
    const uint32_t *src;
    int owner;
    if (g >> logN) {  // leaf
        const uint32_t k = g - (1u << logN);
        const uint32_t r = k >> t.log_blowup, c = k & ((1u << t.log_blowup) - 1);
        owner = (int)(r >> log_nb);
        src = t.stage + ((((size_t)c << log_nb) + (r & ((1u << log_nb) - 1))) << 3);
    } else if (g >= (uint32_t)G) {  // node inside a rank's subtree
        const int lv = 31 - __clz(g);          // level of 2^lv nodes
        const uint32_t j = g - (1u << lv);
        const int ll = lv - t.logG;            // 2^ll of them per rank
        owner = (int)(j >> ll);
        src = t.heap + ((((size_t)1 << ll) + (j & ((1u << ll) - 1))) << 3);
    } else {  // top levels: on every rank
        reinterpret_cast<uint32_t *>(out.p[t.rank])[i] = t.top[(size_t)g * 8 + w];
        return;
    }
    if (owner != t.rank) return;
    const uint32_t v = src[w];
    for (int r = 0; r < G; r++) reinterpret_cast<uint32_t *>(out.p[r])[i] = v;
}
void gather_tree_digests(const SegTreeView &t, const uint32_t *d_idx, int count, const RankPtrs &out, cudaStream_t s) {
    if (!count) return;
    AERO_COUNT_LAUNCH(1);
    gather_tree_digests_kernel<<<(count * 8 + 127) / 128, 128, 0, s>>>(t, d_idx, count, out);
}
__global__ void gather_fri_rows_kernel(const uint64_t *__restrict__ f, uint32_t rows, int log_cosets,
                                       const uint32_t *__restrict__ pos, int npos, uint64_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npos * 8) return;
    const int p = i >> 3, k = i & 7;
    const size_t q = (size_t)pos[p] + (size_t)k * rows;  // natural index into the layer
    size_t addr = q;
    if (log_cosets) {
        const size_t coset_len = ((size_t)rows * 8) >> log_cosets;
        addr = (q & (((size_t)1 << log_cosets) - 1)) * coset_len + (q >> log_cosets);
    }
    out[i] = f[addr];
}
void gather_fri_rows(const uint64_t *f, uint32_t rows, int log_cosets, const uint32_t *d_positions, int npos,
                     uint64_t *d_out, cudaStream_t s) {
    AERO_COUNT_LAUNCH(1);
    gather_fri_rows_kernel<<<(npos * 8 + 127) / 128, 128, 0, s>>>(f, rows, log_cosets, d_positions, npos, d_out);
}
__global__ void gather_digests_kernel(const uint32_t *__restrict__ full, const uint32_t *__restrict__ idx, int count,
                                      uint32_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count * 8) return;
    out[i] = full[(size_t)idx[i >> 3] * 8 + (i & 7)];
}
void gather_digests(const uint32_t *full, const uint32_t *d_idx, int count, uint32_t *d_out, cudaStream_t s) {
    if (!count) return;
    AERO_COUNT_LAUNCH(1);
    gather_digests_kernel<<<(count * 8 + 127) / 128, 128, 0, s>>>(full, d_idx, count, d_out);
}

// ---------------------------------------------------------------------------------------------
// K5 front half: ConstraintEvaluationTable::into_poly's accumulation
// (prover/src/constraints/evaluation_table.rs:166-190, acc_column :330-380,
// get_inv_evaluation :383-419).  x_i = offset * g_N^i ; divisor d = (x^a - b) / prod (x - ex_k).
// ---------------------------------------------------------------------------------------------
constexpr int INV_BATCH = 32;  // one field inversion (~95 multiplications) per 32 elements; 8 cost 17 multiplications per element, 32 cost 8
__global__ void __launch_bounds__(256) divisor_inverses_kernel(DivisorDev d, uint64_t *__restrict__ zinv, PowTable gN) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i0 = t * INV_BATCH;
    if (i0 >= d.zn) return;
    uint64_t v[INV_BATCH], pre[INV_BATCH];
    uint64_t acc = 1;
#pragma unroll
    for (int k = 0; k < INV_BATCH; k++) {
        const uint32_t i = i0 + k;
        v[k] = 0;
        if (i < d.zn) v[k] = gl::sub(gl::mul(pow_lookup(gN, (uint32_t)((uint64_t)i * d.a)), d.off_pow_a), d.b);
        pre[k] = acc;
        if (v[k]) acc = gl::mul(acc, v[k]);
    }
    acc = gl::inv(acc);
#pragma unroll
    for (int k = INV_BATCH - 1; k >= 0; k--) {
        const uint32_t i = i0 + k;
        uint64_t r = 0;
        if (v[k]) {  // zero maps to zero: serial_batch_inversion, math/src/utils/mod.rs:218-238
            r = gl::mul(acc, pre[k]);
            acc = gl::mul(acc, v[k]);
        }
        if (i < d.zn) zinv[i] = r;
    }
}
void divisor_inverses(const DivisorDev &d, uint64_t *zinv_out, int logN, PowTable gN, cudaStream_t s) {
    (void)logN;
    const uint32_t threads = (d.zn + INV_BATCH - 1) / INV_BATCH;
    AERO_COUNT_LAUNCH(1);
    divisor_inverses_kernel<<<(threads + 255) / 256, 256, 0, s>>>(d, zinv_out, gN);
}

struct DivisorSet {
    DivisorDev d[8];
    int n;
};
__global__ void __launch_bounds__(256) constraint_combine_kernel(const uint64_t *__restrict__ cols, size_t col_stride,
                                                                 DivisorSet ds, uint32_t i_begin, uint32_t i_count,
                                                                 uint64_t offset, PowTable gN, int cm_logn, int cm_logb,
                                                                 uint64_t *__restrict__ combined) {
    const uint32_t il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= i_count) return;
    // cm_logn >= 0: the range [i_begin, i_begin + i_count) and the output are COSET-MAJOR (t = r * n + k for the
    // natural row i = k * B + r, B = 2^cm_logb) -- the rows a coset-sharded rank evaluated itself
    const uint32_t t = i_begin + il;
    const uint32_t i = cm_logn < 0 ? t : (((t & ((1u << cm_logn) - 1)) << cm_logb) | (t >> cm_logn));
    const uint64_t x = gl::mul(pow_lookup(gN, i), offset);  // domain.get_ce_x_at, domain.rs:101-103
    uint64_t acc = 0;
    for (int k = 0; k < ds.n; k++) {
        const DivisorDev &d = ds.d[k];
        uint64_t z = __ldg(d.zinv + (i % d.zn));
        for (uint32_t e = 0; e < d.nex; e++) z = gl::mul(z, gl::sub(x, d.ex[e]));
        acc = gl::add(acc, gl::mul(__ldg(cols + (size_t)k * col_stride + i), z));
    }
    combined[t] = acc;
}
void constraint_combine(const uint64_t *cols, size_t col_stride, const DivisorDev *divs, int ndiv, int logN,
                        uint64_t offset, PowTable gN, uint32_t i_begin, uint32_t i_count, uint64_t *combined,
                        cudaStream_t s, int cm_logn, int cm_logb) {
    (void)logN;
    DivisorSet ds;
    ds.n = ndiv;
    for (int i = 0; i < ndiv; i++) ds.d[i] = divs[i];
    if (!i_count) return;
    AERO_COUNT_LAUNCH(1);
    constraint_combine_kernel<<<(i_count + 255) / 256, 256, 0, s>>>(cols, col_stride, ds, i_begin, i_count, offset, gN, cm_logn, cm_logb, combined);
}

// ---- AIR constraint evaluation -----------------------------------------------------
// ConstraintEvaluator::evaluate (prover/src/constraints/evaluator.rs:121-230) for a transition program:
// one thread per step of the constraint evaluation domain.  Threads are numbered coset-major like the LDE
// (thread = rc * n + i for step s = i * ce_blowup + rc), so frame loads are coalesced; the next row of the
// frame is the same LDE coset one position on (trace_lde.rs: + blowup, wrapping).  Node values live in
// per-thread SLOTS (local memory: coalesced across the warp for equal slot indices); the host assigns slots
// by liveness (a slot is reused once the last consumer of its value has run), so the per-thread footprint is
// the widest cut of the expression DAG, not its size.
template <int MAXN>
__global__ void __launch_bounds__(128) air_evaluate_kernel(AirSegs segs, AirProgramDev p, int logn, int log_blowup, int log_ce,
                                                          PowTable g_ce, int to_montgomery, uint64_t *__restrict__ out,
                                                          size_t out_stride, uint32_t tau0, uint32_t tau_count) {
    // [tau0, tau0 + tau_count): the cosets of the evaluation domain this rank holds (all of them on one GPU).
    // One thread per step by default (the loop runs once).  The grid-stride form exists for the experiment
    // recorded in DESIGN.md: fewer resident threads whose slots stay in L2 remove the slots' DRAM traffic (ncu: 6.2 GB
    // for 0.6 GB of frame and result) but lose more to the lower occupancy than they gain (2.25 vs 1.83 ms).
    const uint32_t n = 1u << logn;
    uint64_t val[MAXN];
    for (uint32_t t_ = blockIdx.x * blockDim.x + threadIdx.x; t_ < tau_count; t_ += gridDim.x * blockDim.x) {
    const uint32_t tau = tau0 + t_;
    const uint32_t i = tau & (n - 1), rc = tau >> logn;
    const uint32_t step = (i << log_ce) | rc;                     // natural index in the evaluation domain
    const size_t cur = ((size_t)(rc << (log_blowup - log_ce)) << logn) + i;   // (LDE coset, i)
    const size_t nxt = cur - i + ((i + 1) & (n - 1));
    auto load = [&](uint32_t col, size_t pos) -> uint64_t {
        for (int sg = 0; sg < segs.nseg; sg++) {
            if (col < (uint32_t)segs.ncols[sg]) return segs.lde[sg][(size_t)col * segs.stride[sg] + pos];
            col -= segs.ncols[sg];
        }
        return 0;
    };
    // x^adjustment for x = offset * g_ce^step, the way StarkDomain::get_ce_x_power_at does it (domain.rs:109-117):
    // g_ce^(step * adj mod CE) from the domain's table, times offset^adj from the host -- two products instead of
    // a square-and-multiply chain per adjustment
    uint64_t xp[AIR_MAX_ADJ];
    const uint32_t ce_mask = (n << log_ce) - 1;
    for (int a = 0; a < p.n_adj; a++)
        xp[a] = gl::mul(pow_lookup(g_ce, (uint32_t)(((uint64_t)step * __ldg(p.adj + a)) & ce_mask)), __ldg(p.adj_off + a));
    for (int k = 0; k < p.n_nodes; k++) {
        const uint4 nd = __ldg(reinterpret_cast<const uint4 *>(p.nodes) + k);   // op, a, b, destination slot
        uint64_t v;
        switch (nd.x) {
        case 0: v = load(nd.y, cur); break;
        case 1: v = load(nd.y, nxt); break;
        case 2: v = __ldg(p.consts + nd.y); break;
        case 6: v = __ldg(p.periodic + nd.y + (step & nd.z)); break;  // periodic_table.rs:84-90 (get_row: step % length)
        case 3: v = gl::add(val[nd.y], val[nd.z]); break;
        case 4: v = gl::sub(val[nd.y], val[nd.z]); break;
        default: v = gl::mul(val[nd.y], val[nd.z]); break;
        }
        val[nd.w] = v;
    }
    uint64_t acc[8];
    for (int d = 0; d < 8; d++) acc[d] = 0;
    // transition/mod.rs:272-283: sum (c0 + c1 * x^adj) * evaluation over the constraints of each group
    for (int t = 0; t < p.nt; t++) {
        const uint64_t w = gl::add(__ldg(p.coeffs + 2 * t), gl::mul(__ldg(p.coeffs + 2 * t + 1), xp[__ldg(p.t_adj + t)]));
        acc[0] = gl::add(acc[0], gl::mul(w, val[__ldg(p.t_out + t)]));
    }
    // boundary.rs:255-275: (trace value - asserted value) * (c0 + c1 * x^adj), one column per divisor
    for (int j = 0; j < p.nb; j++) {
        const uint64_t *cc = p.coeffs + 2 * (p.nt + j);
        const uint64_t w = gl::add(__ldg(cc), gl::mul(__ldg(cc + 1), xp[__ldg(p.b_adj + j)]));
        const uint64_t v = gl::sub(load(__ldg(p.b_col + j), cur), __ldg(p.b_val + j));
        const uint32_t d = __ldg(p.b_div + j);
        acc[d] = gl::add(acc[d], gl::mul(w, v));
    }
    for (int d = 0; d < p.n_div; d++) out[(size_t)d * out_stride + step] = to_montgomery ? gl::canon_to_mont(acc[d]) : acc[d];
    }
}
void air_evaluate(const AirSegs &segs, const AirProgramDev &p, int logn, int log_blowup, int log_ce, PowTable g_ce,
                  int to_montgomery, uint64_t *out, size_t out_stride, cudaStream_t s, uint32_t tau0, uint32_t tau_count,
                  int num_sms, int blocks_per_sm) {
    if (!tau_count) return;
    // blocks_per_sm <= 0: one thread per step; n > 0: n resident blocks per SM, each thread looping over steps
    const uint64_t need = ((uint64_t)tau_count + 127) / 128;
    const uint64_t cap = blocks_per_sm > 0 ? (uint64_t)num_sms * blocks_per_sm : need;
    const unsigned grid = (unsigned)(need < cap ? need : cap);
    AERO_COUNT_LAUNCH(1);
    if (p.n_slots <= 32) air_evaluate_kernel<32><<<grid, 128, 0, s>>>(segs, p, logn, log_blowup, log_ce, g_ce, to_montgomery, out, out_stride, tau0, tau_count);
    else if (p.n_slots <= 128) air_evaluate_kernel<128><<<grid, 128, 0, s>>>(segs, p, logn, log_blowup, log_ce, g_ce, to_montgomery, out, out_stride, tau0, tau_count);
    else air_evaluate_kernel<1024><<<grid, 128, 0, s>>>(segs, p, logn, log_blowup, log_ce, g_ce, to_montgomery, out, out_stride, tau0, tau_count);
}

}  // namespace aero
