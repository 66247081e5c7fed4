// FRI degree-respecting projection, folding factor 8 (K8 fold half).
//
// Replaces fri::folding::apply_drp (winterfell/fri/src/folding/mod.rs:86-118) on the transposed
// layer (winterfell/utils/core/src/lib.rs:574-581): next[j] = P_j(alpha), P_j the degree-<8
// interpolant through (x_j * w8^k, f[j + k*rows]), x_j = offset * g_M^j (get_inv_offsets, :181-190;
// the offset is the constant GENERATOR at every layer, fri/src/options.rs:61-63).
// The reference interpolates (8-point inverse DFT, coefficients scaled by (1/8) x_j^-k) and then
// evaluates at alpha by Horner; sum_k (1/8) X_k x_j^-k alpha^k = (1/8) * Horner_k(X; alpha/x_j) is
// the same field element.
#include "kernels.cuh"

namespace aero {

struct FoldConsts {
    uint64_t w[4];  // w8^-k, k = 0..3
    uint64_t inv8;
};

__global__ void __launch_bounds__(256) fri_fold_kernel(const uint64_t *__restrict__ f, uint32_t rows, int log_cosets,
                                                       uint64_t alpha, const uint64_t *__restrict__ alpha_dev,
                                                       PowTable xinv, FoldConsts fc, uint64_t *__restrict__ out) {
    const uint32_t tau = blockIdx.x * blockDim.x + threadIdx.x;
    if (tau >= rows) return;
    if (alpha_dev) alpha = *alpha_dev;  // drawn by the device coin earlier on this stream (fri_coin)
    uint64_t v[8];
    uint32_t j;
    fri_gather8(f, rows, log_cosets, tau, j, v);
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
    const uint64_t beta = gl::mul(alpha, pow_lookup(xinv, j));  // alpha / x_j
    uint64_t acc = X[7];
#pragma unroll
    for (int k = 6; k >= 0; k--) acc = gl::add(gl::mul(acc, beta), X[k]);
    out[j] = gl::mul(acc, fc.inv8);
}

void fri_fold(const uint64_t *f, uint32_t rows, int log_cosets, uint64_t alpha, const uint64_t *alpha_dev, PowTable xinv,
              const uint64_t w8inv[4], uint64_t inv8, uint64_t *out, cudaStream_t s) {
    FoldConsts fc;
    for (int i = 0; i < 4; i++) fc.w[i] = w8inv[i];
    fc.inv8 = inv8;
    AERO_COUNT_LAUNCH(1);
    fri_fold_kernel<<<(rows + 255) / 256, 256, 0, s>>>(f, rows, log_cosets, alpha, alpha_dev, xinv, fc, out);
}

}  // namespace aero
