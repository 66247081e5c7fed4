// Goldilocks NTT kernels (K1 iNTT, K2 coset LDE, K5 coset iNTT) -- declarations.
//
// Replaces winterfell/math/src/fft/serial.rs:32-103 (evaluate_poly_with_offset, interpolate_poly,
// interpolate_poly_with_offset) as driven by winterfell/prover/src/matrix.rs:151-201.  The
// reference's in-place DIF + bit-reversal is mathematically the natural-order DFT
// X[i] = sum_j x[j] w^(ij); any exact algorithm yields the same canonical values, so the device
// uses a four-step (n = n1*n2) decomposition of shared-memory DIT transforms instead:
//   pass 1: n1-point DFTs over j1 (stride n2), coset shift absorbed into the stage twiddles,
//           times the inter-pass factor s^j2 * w_n^(i1*j2), stored in T x T tiles;
//   pass 2: n2-point DFTs over j2, natural-order store.
// Sizes up to 2^11 run in a single shared-memory pass.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace aero {

constexpr int NTT_SINGLE_MAX_LOG = 11;  // largest single-pass transform
constexpr int NTT_MAX_LOG = 24;         // two passes of <= 2^12 points

// Device-resident tables of one transform "plan" (built on the host once per shape, cached).
struct DftTables {
    int logn = 0, log1 = 0, log2 = 0;  // n = n1 * n2 ; log1 == 0 means single pass (n2 = n)
    int ncosets = 1;
    bool plain = false;               // no coset shift (all shifts == 1): unit twiddles are skipped
    int lo_bits = 0;                  // two-level w_n^e table split
    uint64_t *stage1 = nullptr;       // [ncosets][n1]  pass-1 stage twiddles tw[m/2+k] = sigma^(n1/m) w_m^k
    uint64_t *stage2 = nullptr;       // [n2]           pass-2 stage twiddles (plain); single pass: [ncosets][n]
    uint64_t *inter_b = nullptr;      // [ncosets][n2]  c * s_r^j2
    uint64_t *wlo = nullptr, *whi = nullptr;
    uint64_t *post_u = nullptr;       // optional output scale: out[i] *= post_u[i1] * post_v[i2]
    uint64_t *post_v = nullptr;       //   (single pass: post_u[i], post_v unused)
    uint64_t single_scale = 1;        // single pass: constant output scale when post_u == nullptr
};

struct DftLaunch {
    const uint64_t *src;   // column c at src + c*src_col_stride, n entries
    uint64_t *dst;         // (column c, coset r) at dst + c*dst_col_stride + r*n
    uint64_t *tmp;         // >= ncols*ncosets*n entries (two-pass only)
    size_t src_col_stride, dst_col_stride;
    int ncols;
    int deinterleave_log;  // >0: natural index i is stored at (i & (2^d-1))*(n>>d) + (i>>d)
    int coset_begin = 0;   // transform only cosets [coset_begin, coset_begin + coset_count) of the plan;
    int coset_count = 0;   //   0 = all.  Coset (coset_begin + q) is stored at local slot q of dst/tmp.
};

void dft_run(const DftTables &t, const DftLaunch &l, cudaStream_t s);

}  // namespace aero
