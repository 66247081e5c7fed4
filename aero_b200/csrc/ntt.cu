// Goldilocks NTT kernels for sm_100a: shared-memory DIT transforms with register radix-8 rounds,
// four-step two-pass decomposition for n > 2^11.  See ntt.cuh for the mapping to the reference.
#include "gl.cuh"
#include "ntt.cuh"
#include "kernels.cuh"

namespace aero {

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return __brev(x) >> (32 - bits); }

// One register-blocked round of R DIT stages (s0 .. s0+R-1) over an M x T tile in shared memory.
// a[idx*RS + t]; tw[2^s + k] is the stage-s twiddle for pair offset k (k < 2^s).
// PLAIN0: this is the first round (s0 == 0) of a transform without coset shift, where the
// twiddle of pair offset k == 0 is 1 -- known at compile time, so 7 of the 12 multiplications of a
// radix-8 round disappear.
template <int R, int T, int RS, bool PLAIN0>
__device__ __forceinline__ void dit_round(uint64_t *a, const uint64_t *tw, int s0, int logM) {
    if (PLAIN0) s0 = 0;
    const int ngroups = (1 << logM) >> R;
    const int items = ngroups * T;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int t = it % T, g = it / T;
        const int low = g & ((1 << s0) - 1);
        const int base = ((g >> s0) << (s0 + R)) | low;
        uint64_t x[1 << R];
#pragma unroll
        for (int e = 0; e < (1 << R); e++) x[e] = a[(base + (e << s0)) * RS + t];
#pragma unroll
        for (int q = 0; q < R; q++) {
#pragma unroll
            for (int e = 0; e < (1 << R); e++) {
                if (e & (1 << q)) continue;
                const int k = low + ((e & ((1 << q) - 1)) << s0);
                // values in the tile stay "any" (congruent mod p, < 2^64); only v is canonicalised
                const uint64_t u = x[e];
                uint64_t v;
                if (PLAIN0 && (e & ((1 << q) - 1)) == 0) {
                    v = gl::canon_any(x[e | (1 << q)]);
                } else {
                    v = gl::mul_canon(x[e | (1 << q)], tw[(1 << (s0 + q)) + k]);
                }
                x[e] = gl::add_ac(u, v);
                x[e | (1 << q)] = gl::sub_ac(u, v);
            }
        }
#pragma unroll
        for (int e = 0; e < (1 << R); e++) a[(base + (e << s0)) * RS + t] = x[e];
    }
}

// Full M-point DIT over the tile: input in bit-reversed row order, output in natural row order.
template <int T, int RS, bool PLAIN>
__device__ __forceinline__ void dit_tile(uint64_t *a, const uint64_t *tw, int logM) {
    int s0 = 0;
    // first round: s0 == 0 is a compile-time fact, which lets PLAIN transforms drop unit twiddles
    if (logM >= 3 && logM != 4) {
        dit_round<3, T, RS, PLAIN>(a, tw, 0, logM);
        s0 = 3;
    } else if (logM >= 2) {
        dit_round<2, T, RS, PLAIN>(a, tw, 0, logM);
        s0 = 2;
    } else {
        dit_round<1, T, RS, PLAIN>(a, tw, 0, logM);
        s0 = 1;
    }
    __syncthreads();
    while (s0 < logM) {
        const int left = logM - s0;
        // prefer 3-stage rounds; avoid a trailing 1-stage round when 4 stages remain (2+2).  A 16-point
        // register round for 4 remaining stages was measured slower (64 registers, LDE 14.1 -> 14.3 ms)
        if (left >= 3 && left != 4) {
            dit_round<3, T, RS, false>(a, tw, s0, logM);
            s0 += 3;
        } else if (left >= 2) {
            dit_round<2, T, RS, false>(a, tw, s0, logM);
            s0 += 2;
        } else {
            dit_round<1, T, RS, false>(a, tw, s0, logM);
            s0 += 1;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ uint64_t root_pow(const uint64_t *__restrict__ wlo, const uint64_t *__restrict__ whi,
                                             int lo_bits, uint32_t e) {
    return gl::mul(__ldg(wlo + (e & ((1u << lo_bits) - 1))), __ldg(whi + (e >> lo_bits)));
}

__device__ __forceinline__ size_t out_index(uint32_t i, int logn, int deint) {
    if (deint == 0) return i;
    return (size_t)(i & ((1u << deint) - 1)) * ((size_t)1 << (logn - deint)) + (i >> deint);
}

// ---- pass 1 -------------------------------------------------------------------------------
// grid: (n2/T, ncosets, ncols).  tmp layout per (col, coset): [i1/T][j2][i1%T].
template <int T, bool PLAIN>
__global__ void __launch_bounds__(1024) dft_pass1_kernel(const uint64_t *__restrict__ src, uint64_t *__restrict__ tmp,
                                                        const uint64_t *__restrict__ stage1,
                                                        const uint64_t *__restrict__ inter_b,
                                                        const uint64_t *__restrict__ wlo,
                                                        const uint64_t *__restrict__ whi, int lo_bits, int log1,
                                                        int log2, size_t src_col_stride, int ncosets) {
    constexpr int RS = T + 1;
    extern __shared__ uint64_t smem[];
    const int n1 = 1 << log1;
    const size_t n = (size_t)1 << (log1 + log2);
    uint64_t *a = smem;            // n1 * RS
    uint64_t *tw = smem + n1 * RS; // n1
    const int coset = blockIdx.y, col = blockIdx.z;
    const uint32_t j2_0 = blockIdx.x * T;
    const uint64_t *s = src + (size_t)col * src_col_stride;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) tw[i] = stage1[(size_t)coset * n1 + i];
    for (int it = threadIdx.x; it < n1 * T; it += blockDim.x) {
        const int t = it % T, j1 = it / T;
        a[bitrev(j1, log1) * RS + t] = s[((size_t)j1 << log2) + j2_0 + t];
    }
    __syncthreads();
    dit_tile<T, RS, PLAIN>(a, tw, log1);
    // inter-pass factor F(i1, j2) = b[j2] * w_n^(i1*j2), advanced by a running product per thread
    uint64_t *o = tmp + ((size_t)col * ncosets + coset) * n;
    const int it0 = threadIdx.x;
    const int ii = it0 % T, t = (it0 / T) % T;
    const uint32_t j2 = j2_0 + t;
    const int cstep = blockDim.x / (T * T);   // blockDim.x is a multiple of T*T
    const uint32_t step_i1 = cstep * T;
    int c = it0 / (T * T);
    uint32_t i1 = c * T + ii;
    uint64_t f = gl::mul(__ldg(inter_b + (size_t)coset * ((size_t)1 << log2) + j2), root_pow(wlo, whi, lo_bits, i1 * j2));
    const uint64_t d = root_pow(wlo, whi, lo_bits, step_i1 * j2);
    const int nchunks = n1 / T;
    for (; c < nchunks; c += cstep, i1 += step_i1) {
        const uint64_t v = gl::mul_canon(a[i1 * RS + t], f);
        o[((size_t)c << log2) * T + (size_t)j2 * T + ii] = v;
        f = gl::mul_any(f, d);
    }
}

// ---- pass 2 -------------------------------------------------------------------------------
// grid: (n1/T, ncosets, ncols)
template <int T>
__global__ void __launch_bounds__(1024) dft_pass2_kernel(const uint64_t *__restrict__ tmp, uint64_t *__restrict__ dst,
                                                        const uint64_t *__restrict__ stage2,
                                                        const uint64_t *__restrict__ post_u,
                                                        const uint64_t *__restrict__ post_v, int log1, int log2,
                                                        size_t dst_col_stride, int ncosets, int deint) {
    constexpr int RS = T + 1;
    extern __shared__ uint64_t smem[];
    const int n2 = 1 << log2;
    const int logn = log1 + log2;
    const size_t n = (size_t)1 << logn;
    uint64_t *a = smem;
    uint64_t *tw = smem + n2 * RS;
    const int coset = blockIdx.y, col = blockIdx.z;
    const uint32_t i1_0 = blockIdx.x * T;
    const uint64_t *s = tmp + ((size_t)col * ncosets + coset) * n + ((size_t)blockIdx.x << log2) * T;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) tw[i] = stage2[i];
    for (int it = threadIdx.x; it < n2 * T; it += blockDim.x) {
        const int t = it % T, j2 = it / T;
        a[bitrev(j2, log2) * RS + t] = s[it];
    }
    __syncthreads();
    dit_tile<T, RS, true>(a, tw, log2);
    uint64_t *o = dst + (size_t)col * dst_col_stride + (size_t)coset * n;
    for (int it = threadIdx.x; it < n2 * T; it += blockDim.x) {
        const int t = it % T, i2 = it / T;
        uint64_t v = a[i2 * RS + t];
        const uint32_t i1 = i1_0 + t;
        if (post_u) v = gl::mul(v, gl::mul(__ldg(post_u + i1), __ldg(post_v + i2)));
        else v = gl::canon_any(v);
        const uint32_t i = i1 + ((uint32_t)i2 << log1);
        o[out_index(i, logn, deint)] = v;
    }
}

// ---- single pass (n <= 2^11) --------------------------------------------------------------
// grid: (1, ncosets, ncols)
template <bool PLAIN>
__global__ void __launch_bounds__(256) dft_single_kernel(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst,
                                                         const uint64_t *__restrict__ stage,
                                                         const uint64_t *__restrict__ post_u, uint64_t scale, int logn,
                                                         size_t src_col_stride, size_t dst_col_stride, int deint) {
    extern __shared__ uint64_t smem[];
    const int n = 1 << logn;
    uint64_t *a = smem;
    uint64_t *tw = smem + n;
    const int coset = blockIdx.y, col = blockIdx.z;
    const uint64_t *s = src + (size_t)col * src_col_stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        tw[i] = stage[(size_t)coset * n + i];
        a[bitrev(i, logn)] = s[i];
    }
    __syncthreads();
    dit_tile<1, 1, PLAIN>(a, tw, logn);
    uint64_t *o = dst + (size_t)col * dst_col_stride + (size_t)coset * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        uint64_t v = a[i];
        if (post_u) v = gl::mul(v, __ldg(post_u + i));
        else if (scale != 1) v = gl::mul(v, scale);
        else v = gl::canon_any(v);
        o[out_index(i, logn, deint)] = v;
    }
}

template <int T>
static void launch_two_pass(const DftTables &t, const DftLaunch &l, cudaStream_t s) {
    const int n1 = 1 << t.log1, n2 = 1 << t.log2;
    const int nc = l.coset_count ? l.coset_count : t.ncosets;
    const uint64_t *stage1 = t.stage1 + (size_t)l.coset_begin * n1;
    const uint64_t *inter_b = t.inter_b + (size_t)l.coset_begin * n2;
    const size_t smem1 = (size_t)n1 * (T + 1) * 8 + (size_t)n1 * 8;
    const size_t smem2 = (size_t)n2 * (T + 1) * 8 + (size_t)n2 * 8;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(dft_pass1_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dft_pass1_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dft_pass2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dft_pass1_kernel<T, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(dft_pass1_kernel<T, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(dft_pass2_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    // 512 threads while two blocks fit the 227 KB of shared memory of an SM; a 2^12-point pass
    // (192 KB, one block per SM) runs 1024 threads instead so the SM still holds 32 warps
    auto threads_for = [](int m, size_t smem) {
        const int cap = smem > 113 * 1024 ? 1024 : 512;
        int items = (m / 8) * T;  // radix-8 work items per round
        int th = items < 64 ? 64 : (items > cap ? cap : items);
        th = (th / (T * T)) * (T * T);
        return th < T * T ? T * T : th;
    };
    dim3 g1(n2 / T, nc, l.ncols), g2(n1 / T, nc, l.ncols);
    AERO_COUNT_LAUNCH(2);
    if (t.plain)
        dft_pass1_kernel<T, true><<<g1, threads_for(n1, smem1), smem1, s>>>(l.src, l.tmp, stage1, inter_b, t.wlo, t.whi,
                                                                     t.lo_bits, t.log1, t.log2, l.src_col_stride, nc);
    else
        dft_pass1_kernel<T, false><<<g1, threads_for(n1, smem1), smem1, s>>>(l.src, l.tmp, stage1, inter_b, t.wlo, t.whi,
                                                                      t.lo_bits, t.log1, t.log2, l.src_col_stride, nc);
    dft_pass2_kernel<T><<<g2, threads_for(n2, smem2), smem2, s>>>(l.tmp, l.dst, t.stage2, t.post_u, t.post_v, t.log1, t.log2,
                                                           l.dst_col_stride, nc, l.deinterleave_log);
}

void dft_run(const DftTables &t, const DftLaunch &l, cudaStream_t s) {
    if (t.log1 == 0) {
        const int n = 1 << t.logn;
        int th = n / 8 < 32 ? 32 : (n / 8 > 256 ? 256 : n / 8);
        const int nc = l.coset_count ? l.coset_count : t.ncosets;
        dim3 g(1, nc, l.ncols);
        AERO_COUNT_LAUNCH(1);
        if (t.plain)
            dft_single_kernel<true><<<g, th, (size_t)n * 16, s>>>(l.src, l.dst, t.stage2 + (size_t)l.coset_begin * n,
                                                                 t.post_u, t.single_scale, t.logn, l.src_col_stride,
                                                                 l.dst_col_stride, l.deinterleave_log);
        else
            dft_single_kernel<false><<<g, th, (size_t)n * 16, s>>>(l.src, l.dst, t.stage2 + (size_t)l.coset_begin * n,
                                                                  t.post_u, t.single_scale, t.logn, l.src_col_stride,
                                                                  l.dst_col_stride, l.deinterleave_log);
        return;
    }
    const int big = t.log1 > t.log2 ? t.log1 : t.log2;
    // tile width T: 2^10-point passes with T = 8 and 2^11-point passes with T = 4 both leave room for
    // two blocks per SM (82 / 96 KB each); measured 4.2e11 butterflies/s either way, against 3.0e11
    // with one 512-thread block per SM
    if (big <= 10) launch_two_pass<8>(t, l, s);
    else launch_two_pass<4>(t, l, s);
}

}  // namespace aero
