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
// Above 2^20 points a third factor n0 <= 16 is split off (n = n1*n2*n0, n1 = n2 = 2^10) so that both
// shared-memory passes stay 2^10-point transforms (two 32-point rounds each, the fastest shape):
//   pass 1: as above with n2*n0 in the place of n2 (tile columns q = j2*n0 + j0);
//   pass 2: for every j0, n2-point DFTs over j2 (rows n0 tile columns apart), times w_(n2 n0)^(i2 j0),
//           stored at i1 + n1*i2 + n1*n2*j0 of the destination;
//   pass 3: in place, n0-point DFTs over j0 (stride n1*n2) whose twiddles are powers of two.
//
// Each shared-memory transform is a mixed-radix DIT of register-resident rounds of 2^R points
// (R <= NTT_MAX_ROUND_LOG).  Only the inputs of a round are multiplied by general twiddles; the
// twiddles inside a round are 2^R-th roots of unity, which are powers of two in this field
// (w_64 = 2^39), i.e. shifts (gl::mul_pow2).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace aero {

constexpr int NTT_SINGLE_MAX_LOG = 11;  // largest single-pass transform
constexpr int NTT_MAX_LOG = 24;         // two passes of <= 2^12 points, or 2^10 x 2^10 x (<= 2^4)
constexpr int NTT_OUTER_MAX_LOG = 4;    // largest third factor n0 of a three-pass plan
constexpr int NTT_LARGE_MAX_LOG = 26;   // with one outer radix-2 / radix-4 step over 2^24-point transforms (abi.cu)
constexpr int NTT_MAX_ROUND_LOG = 5;    // largest register-resident round: 32 points

// Round schedule of a 2^logM-point shared-memory transform, shared by the host (stage tables) and the
// device (dit_tile): the fewest rounds of at most 2^NTT_MAX_ROUND_LOG points, sizes as even as
// possible, larger rounds first.  Round i covers DIT stages [start, start + log).
struct NttRounds {
    int count, base, extra;
    __host__ __device__ explicit NttRounds(int logM)
        : count((logM + NTT_MAX_ROUND_LOG - 1) / NTT_MAX_ROUND_LOG), base(0), extra(0) {
        if (count < 1) count = 1;
        base = logM / count;
        extra = logM % count;
    }
    __host__ __device__ int log(int i) const { return base + (i < extra ? 1 : 0); }
};

// Device-resident tables of one transform "plan" (built on the host once per shape, cached).
struct DftTables {
    int logn = 0, log1 = 0, log2 = 0;  // n = n1 * n2 * n0; log1 == 0 means single pass (n2 = n)
    int log0 = 0;                     // > 0: three-pass plan (pass 1 sees n2 * n0 tile columns)
    int ncosets = 1;
    bool plain = false;               // no coset shift (all shifts == 1): unit twiddles are skipped
    bool inverse = false;             // transform root is w_n^-1 (selects the shifts inside a round)
    int lo_bits = 0;                  // two-level w_n^e table split
    // stage tables: for the round covering stages [s0, s0+R), sub-transform e (1 <= e < 2^R) and
    // offset low < 2^s0:  tw[(e << s0) + low] = (sigma^(M/m) w_m^low)^bitrev_R(e),  m = 2^(s0+R)
    uint64_t *stage1 = nullptr;       // [ncosets][n1]  pass 1 (coset shift sigma = s_r^n2 absorbed)
    uint64_t *stage2 = nullptr;       // [n2]           pass 2 (plain); single pass: [ncosets][n]
    uint64_t *inter_b = nullptr;      // [ncosets][n2]  c * s_r^j2
    uint64_t *inter_full = nullptr;   // optional [ncosets][n]: the whole inter-pass factor c s_r^j2 w_n^(i1 j2) in
                                      //   the tile order pass 1 stores in (one multiplication instead of two)
    uint64_t *wlo = nullptr, *whi = nullptr;
    uint64_t *post_u = nullptr;       // optional output scale: out[i] *= post_u[i1] * post_v[i2]
    uint64_t *post_v = nullptr;       //   (single pass: post_u[i], post_v unused)
    uint64_t *post_j = nullptr;       // three-pass plans: [n0][n2]  w_(n2 n0)^(i2 j0), applied by pass 2
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
// two-pass plans: writes the [ncosets][n] table DftTables::inter_full points at (t.inter_full itself
// is ignored; the other tables of t must be resident)
void dft_fill_inter_table(const DftTables &t, uint64_t *out, cudaStream_t s);
int dft_tile_width(int log1, int log2);

}  // namespace aero
