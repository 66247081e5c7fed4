// FRI degree-respecting projection, folding factor 8 (K8 fold half).
//
// Replaces fri::folding::apply_drp (winterfell/fri/src/folding/mod.rs:86-118) on the transposed
// layer (winterfell/utils/core/src/lib.rs:574-581): next[j] = P_j(alpha), P_j the degree-<8
// interpolant through (x_j * w8^k, f[j + k*rows]), x_j = offset * g_M^j (get_inv_offsets, :181-190;
// the offset is the constant GENERATOR at every layer, fri/src/options.rs:61-63).
// The reference interpolates (8-point inverse DFT, coefficients scaled by (1/8) x_j^-k) and then
// evaluates at alpha by Horner; sum_k (1/8) X_k x_j^-k alpha^k = (1/8) * Horner_k(X; alpha/x_j) is
// the same field element.
#include "blake2s.cuh"
#include "kernels.cuh"

namespace aero {

struct FoldConsts {
    uint64_t w[4];  // w8^-k, k = 0..3
    uint64_t inv8;
};

// next[j] from the 8 evaluations v of leaf j (see the header comment): 8-point inverse DFT + Horner at alpha / x_j
__device__ __forceinline__ uint64_t fri_fold8(const uint64_t v[8], uint64_t alpha, uint64_t xinv_j, const FoldConsts &fc) {
    // 8-point DFT with root w = w8^-1 (decimation in frequency, natural-order outputs X0..X7)
    uint64_t a[4], b[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
        a[m] = gl::add(v[m], v[m + 4]);
        b[m] = gl::sub(v[m], v[m + 4]);
    }
    b[1] = gl::mul(b[1], fc.w[1]);
    b[2] = gl::mul(b[2], fc.w[2]);
    b[3] = gl::mul(b[3], fc.w[3]);
    uint64_t X[8];
    {
        const uint64_t c0 = gl::add(a[0], a[2]), c1 = gl::add(a[1], a[3]);
        const uint64_t d0 = gl::sub(a[0], a[2]), d1 = gl::mul(gl::sub(a[1], a[3]), fc.w[2]);
        X[0] = gl::add(c0, c1); X[4] = gl::sub(c0, c1);
        X[2] = gl::add(d0, d1); X[6] = gl::sub(d0, d1);
    }
    {
        const uint64_t c0 = gl::add(b[0], b[2]), c1 = gl::add(b[1], b[3]);
        const uint64_t d0 = gl::sub(b[0], b[2]), d1 = gl::mul(gl::sub(b[1], b[3]), fc.w[2]);
        X[1] = gl::add(c0, c1); X[5] = gl::sub(c0, c1);
        X[3] = gl::add(d0, d1); X[7] = gl::sub(d0, d1);
    }
    const uint64_t beta = gl::mul(alpha, xinv_j);  // alpha / x_j
    uint64_t acc = X[7];
#pragma unroll
    for (int k = 6; k >= 0; k--) acc = gl::add(gl::mul(acc, beta), X[k]);
    return gl::mul(acc, fc.inv8);
}

__global__ void __launch_bounds__(256) fri_fold_kernel(const uint64_t *__restrict__ f, uint32_t rows, int log_cosets,
                                                       uint64_t alpha, const uint64_t *__restrict__ alpha_dev,
                                                       PowTable xinv, FoldConsts fc, uint64_t *__restrict__ out) {
    const uint32_t tau = blockIdx.x * blockDim.x + threadIdx.x;
    if (tau >= rows) return;
    if (alpha_dev) alpha = *alpha_dev;  // drawn by the device coin earlier on this stream (fri_coin)
    uint64_t v[8];
    uint32_t j;
    fri_gather8(f, rows, log_cosets, tau, j, v);
    out[j] = fri_fold8(v, alpha, pow_lookup(xinv, j), fc);
}

// Fused fold-and-hash: folds the layer AND hashes the leaves of the next one.  Thread = leaf j' of the next
// layer (rows' = rows / 8 leaves): it folds the eight leaves j' + k' * rows' of this layer -- exactly the
// eight values of its own leaf (transpose_slice of the folded layer) -- stores them to next[] and writes
// hash_elements of them to leaves[j'], so the next layer is neither re-read nor hashed by a second kernel.
// The thread-to-leaf map keeps the loads coalesced for a coset-major layer (see fri_gather8); rows' must
// then be a multiple of the number of cosets.
__global__ void __launch_bounds__(128) fri_fold_hash_kernel(const uint64_t *__restrict__ f, uint32_t rows, int log_cosets,
                                                            const uint64_t *__restrict__ alpha_dev, PowTable xinv, FoldConsts fc,
                                                            uint64_t *__restrict__ next, uint32_t *__restrict__ leaves) {
    const uint32_t rows2 = rows >> 3;
    const uint32_t tau = blockIdx.x * blockDim.x + threadIdx.x;
    if (tau >= rows2) return;
    const uint64_t alpha = *alpha_dev;
    uint32_t j2;
    const uint64_t *p;
    size_t kstride, k2stride;       // between the k-th values of a leaf; between the leaves j' + k' * rows'
    if (log_cosets == 0) {
        j2 = tau;
        p = f + j2;
        kstride = rows;
        k2stride = rows2;
    } else {
        const uint32_t per2 = rows2 >> log_cosets;
        const uint32_t r = tau / per2, a = tau - r * per2;
        j2 = (a << log_cosets) | r;
        p = f + (size_t)r * (((size_t)rows * 8) >> log_cosets) + a;
        kstride = rows >> log_cosets;
        k2stride = per2;
    }
    uint64_t w[8];
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) {
        uint64_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __ldg(p + (size_t)k2 * k2stride + (size_t)k * kstride);
        const uint32_t j = j2 + (uint32_t)k2 * rows2;
        w[k2] = fri_fold8(v, alpha, pow_lookup(xinv, j), fc);
        next[j] = w[k2];
    }
    uint32_t h[8];
    b2s::init(h);
    b2s::compress_pair(h, w[0], w[1], 64u, false);
    b2s::compress_pair(h, w[2], w[3], 128u, false);
    b2s::compress_pair(h, w[4], w[5], 192u, false);
    b2s::compress_pair(h, w[6], w[7], 256u, true);
    store_digest(leaves + (size_t)j2 * 8, h);
}

void fri_fold(const uint64_t *f, uint32_t rows, int log_cosets, uint64_t alpha, const uint64_t *alpha_dev, PowTable xinv,
              const uint64_t w8inv[4], uint64_t inv8, uint64_t *out, cudaStream_t s) {
    FoldConsts fc;
    for (int i = 0; i < 4; i++) fc.w[i] = w8inv[i];
    fc.inv8 = inv8;
    AERO_COUNT_LAUNCH(1);
    fri_fold_kernel<<<(rows + 255) / 256, 256, 0, s>>>(f, rows, log_cosets, alpha, alpha_dev, xinv, fc, out);
}

void fri_fold_hash(const uint64_t *f, uint32_t rows, int log_cosets, const uint64_t *alpha_dev, PowTable xinv,
                   const uint64_t w8inv[4], uint64_t inv8, uint64_t *next, uint32_t *next_leaves, cudaStream_t s) {
    FoldConsts fc;
    for (int i = 0; i < 4; i++) fc.w[i] = w8inv[i];
    fc.inv8 = inv8;
    AERO_COUNT_LAUNCH(1);
    const uint32_t rows2 = rows >> 3;
    fri_fold_hash_kernel<<<(rows2 + 127) / 128, 128, 0, s>>>(f, rows, log_cosets, alpha_dev, xinv, fc, next, next_leaves);
}

}  // namespace aero
