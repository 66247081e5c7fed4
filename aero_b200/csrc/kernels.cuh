// Shared declarations for the aero_b200 CUDA kernels (host-callable launchers + device helpers).
#pragma once
#include <atomic>
#include <mutex>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "gl.cuh"

namespace aero {

extern std::atomic<unsigned long long> g_launch_count;  // kernels launched by this library (aero_launch_count)
#define AERO_COUNT_LAUNCH(n) (aero::g_launch_count += (n))

// cudaFuncSetAttribute is per device: run(f) calls f once per device the calling thread has current
// (a process may hold contexts on several GPUs, one host thread each).
struct DeviceOnce {
    std::mutex mu;
    unsigned long long mask = 0;
    template <class F>
    void run(F &&f) {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ULL << (dev & 63);
        std::lock_guard<std::mutex> lock(mu);
        if (mask & bit) return;
        f();
        mask |= bit;
    }
};

// Two-level table for powers of a fixed base: base^e = lo[e & (2^lo_bits-1)] * hi[e >> lo_bits].
// `hi` may carry an extra constant factor (folded scale).
struct PowTable {
    const uint64_t *lo = nullptr;
    const uint64_t *hi = nullptr;
    int lo_bits = 0;
};
#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t pow_lookup(const PowTable &t, uint32_t e) {
    return gl::mul(__ldg(t.lo + (e & ((1u << t.lo_bits) - 1))), __ldg(t.hi + (e >> t.lo_bits)));
}

__device__ __forceinline__ void store_digest(uint32_t *dst, const uint32_t h[8]) {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    d[0] = make_uint4(h[0], h[1], h[2], h[3]);
    d[1] = make_uint4(h[4], h[5], h[6], h[7]);
}
__device__ __forceinline__ void load_digest(const uint32_t *src, uint32_t h[8]) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 a = s[0], b = s[1];
    h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w;
    h[4] = b.x; h[5] = b.y; h[6] = b.z; h[7] = b.w;
}

// Gather the 8 evaluations of FRI leaf j = {f[j + k*rows]} (fri transpose_slice,
// utils/core/src/lib.rs:574-581).  Thread index tau maps to leaf j so that loads are coalesced:
// natural layout: j = tau; coset-major layout (B = 2^log_cosets cosets of M/B entries, natural q at
// (q & (B-1))*(M/B) + (q >> b)): j = B*a + r with tau = r*(rows/B) + a (rows % B == 0).
__device__ __forceinline__ void fri_gather8(const uint64_t *__restrict__ f, uint32_t rows, int log_cosets,
                                            uint32_t tau, uint32_t &j, uint64_t v[8]) {
    if (log_cosets == 0) {
        j = tau;
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __ldg(f + (size_t)j + (size_t)k * rows);
    } else {
        const uint32_t per = rows >> log_cosets;     // rows / B
        const uint32_t r = tau / per, a = tau - r * per;
        j = (a << log_cosets) | r;
        const size_t coset_len = ((size_t)rows * 8) >> log_cosets;  // M / B
        const uint64_t *p = f + (size_t)r * coset_len + a;
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __ldg(p + (size_t)k * per);
    }
}
#endif

// Multi-GPU: every rank of a sharded proof owns an exchange window; a buffer
// bump-allocated inside it sits at the same offset on every rank, so a kernel can address all G copies:
// p[r] = the copy on rank r (p[own rank] = the local buffer).  Kernels store results straight into the
// copy of the rank that consumes them (NVLink peer stores), so the exchange rides inside the producer.
constexpr int AERO_MAX_RANKS = 8;
struct RankPtrs {
    void *p[AERO_MAX_RANKS];
};

// hash.cu
// Row hashing of the LDE (K3).  nrows = (locally stored cosets) * n; local coset q holds natural rows
// B*i + coset_begin + q.  The digest of (coset c, row i) goes to the LEAF STAGE of the rank that owns
// the leaf block of i: rank i >> log_nb, slot (c << log_nb) + (i & (nb - 1)) -- coset-major inside the
// block, so consecutive threads store consecutive 32-byte digests, locally or over NVLink.  With one
// rank log_nb = logn and the stage is simply [B][n].  Columns [c0, c0 + ncols) of a total_cols-wide
// row (see hash_rows_kernel); c0 must be even and, when > 0, all destinations local.
// max_blocks > 0 caps the grid (the kernel strides over the rows).
void hash_rows_lde(const uint64_t *lde, size_t col_stride, int c0, int ncols, int total_cols, int logn,
                   uint32_t coset_begin, uint32_t nrows, int log_nb, const RankPtrs &stage, int max_blocks,
                   cudaStream_t s);
// Tree over one rank's leaf block (K4, bottom part): stage = [B][nb] digests, heap = B*nb digests in
// heap layout for the block's subtree (heap[1] = sub-root; the level of L nodes at [L, 2L)).  Leaf
// B*il + c of the block is stage[c*nb + il].  Builds every level up to heap[1].
void merkle_build_block(const uint32_t *stage, uint32_t *heap, uint32_t nb, int log_blowup, cudaStream_t s);
// heap[1] of every rank -> top[G + rank] on all ranks (top = nodes 1 .. 2G-1 of the whole tree)
void merkle_push_subroot(const uint32_t *heap, const RankPtrs &top, int G, int rank, cudaStream_t s);
// top[j] = merge(top[2j], top[2j+1]) for j = G-1 .. 1
void merkle_top(uint32_t *top, int G, cudaStream_t s);
// copies [off, off + bytes) of the local buffer to the same range of every other rank's copy (16-byte units)
void peer_push(const RankPtrs &buf, int G, int rank, size_t off, size_t bytes, cudaStream_t s);
// same for [first, first + count) 8-byte words (small, unaligned ranges)
void peer_push_words(const RankPtrs &buf, int G, int rank, size_t first, size_t count, cudaStream_t s);
// [off, off + bytes) of the local buffer -> the copy on rank `dest` only; then (a second, one-thread kernel:
// all stores of the first have completed) *dest_flag = epoch, the arrival flag the receiver polls
void peer_send(const RankPtrs &buf, int rank, int dest, size_t off, size_t bytes, unsigned long long *dest_flag,
               unsigned long long epoch, cudaStream_t s);
// spins until flags[0 .. nflags) >= epoch (nflags <= 32; ~4 s time-out counted in *d_timeout)
void peer_wait(const unsigned long long *flags, int nflags, unsigned long long epoch, unsigned int *d_timeout, cudaStream_t s);
// loads the exchange kernels (see hash.cu)
void preload_exchange_kernels();
// all ranks arrive (epoch) before any leaves: flags[r] of rank q's window is written by rank r
void peer_barrier(const RankPtrs &flags, int G, int rank, unsigned long long epoch, unsigned int *d_timeout,
                  cudaStream_t s);
// natural-order leaf digests of a single-rank segment (tests / aero_segment_download_leaves)
void leaves_to_natural(const uint32_t *stage, uint32_t *out, int logn, int log_blowup, cudaStream_t s);
void hash_rows_natural(const uint64_t *m, size_t col_stride, int ncols, uint32_t nrows, uint32_t *leaves,
                       cudaStream_t s);
void merkle_build(uint32_t *full, uint64_t num_leaves, cudaStream_t s);
void fri_leaf_hash(const uint64_t *f, uint32_t rows, int log_cosets, uint32_t *leaves, cudaStream_t s);
// Fiat-Shamir for one FRI layer on the device: seed <- H(seed || root) (RandomCoin::reseed,
// crypto/src/random/mod.rs:105-108), alpha <- first canonical draw (:179-196); the root is also
// copied to root_out so that all layer roots leave in one download.  alpha = 2^64-1 if 1000 draws fail.
void fri_coin(uint32_t *seed, const uint32_t *root, uint64_t *alpha_out, uint32_t *root_out, cudaStream_t s);
void pow_search(const uint32_t *seed, uint64_t base, uint32_t count, uint32_t bits, unsigned long long *best,
                cudaStream_t s);

// poly.cu
void field_ops(const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out, cudaStream_t s);
void convert_form(const uint64_t *src, uint64_t *dst, size_t count, int to_montgomery, cudaStream_t s);
void lde_to_natural(const uint64_t *lde_cm, uint64_t *out, int logn, int log_blowup, int to_montgomery,
                    cudaStream_t s);
void natural_to_coset_major(const uint64_t *in, uint64_t *out, int logn, int log_blowup, cudaStream_t s);
// outer radix-B step of transforms above 2^24 points (poly.cu)
void large_deinterleave(const uint64_t *x, uint64_t *out, int logm, int logB, cudaStream_t s);
void large_combine(const uint64_t *Y, uint64_t *dst, PowTable wn, const uint64_t *d_shifts, uint64_t scale, uint64_t w4, int logm,
                   int logB, int ncosets, cudaStream_t s);
bool coset_interp_combine(const uint64_t *a, PowTable ginv, PowTable oinv, const uint64_t *M, int logn, int log_b,
                          uint64_t *polys, cudaStream_t s);
// out[p*2 + 0/1] partial sums; see poly.cu
void ood_eval(const uint64_t *polys, size_t col_stride, int ncols, int logn, const uint64_t *d_points, int npoints,
              uint64_t *d_out /* ncols*npoints */, uint64_t *d_scratch, cudaStream_t s);
size_t ood_scratch_elems(int ncols, int logn, int npoints);
struct DeepSegs {  // trace segments in order (main, aux...): column c of segment s at p[s] + c*n
    const uint64_t *p[4];
    int ncols[4];
    int nseg;
};
// coefficient indices [j_begin, j_begin + j_count) only (a rank's share of a sharded proof)
void deep_accumulate(const DeepSegs &segs, const uint64_t *comp_polys, int m, int logn, uint32_t j_begin,
                     uint32_t j_count, const uint64_t *d_cc /* W*2 + m */,
                     const uint64_t *d_consts /* 3: subtract from coefficient 0 of t1,t2,h */, uint64_t *t1,
                     uint64_t *t2, uint64_t *h, cudaStream_t s);
void syn_div3(uint64_t *t1, uint64_t *t2, uint64_t *h, int logn, const uint64_t b[3], uint64_t *d_carry,
              cudaStream_t s);
void deep_finish(const uint64_t *t1, const uint64_t *t2, const uint64_t *h, int logn, uint64_t d0, uint64_t d1,
                 uint64_t *out, cudaStream_t s);
// Openings of a (possibly sharded) segment: every entry has ONE owning rank, which writes it into the
// result buffer of every rank (out.p[r] + same offset); after a barrier all ranks hold all results.
// Rows belong to the rank that stores their LDE coset; tree nodes to the rank whose leaf block they
// cover; the top log2(G) levels exist on every rank and are written locally only.
struct SegTreeView {
    const uint32_t *stage, *heap, *top;
    int logn, log_blowup, logG, rank;
};
void gather_rows(const uint64_t *lde_cm, size_t col_stride, int ncols, int logn, int log_blowup, int coset_begin,
                 int coset_count, int G, const uint32_t *d_positions, int npos, const RankPtrs &out, cudaStream_t s);
void gather_tree_digests(const SegTreeView &t, const uint32_t *d_idx, int count, const RankPtrs &out, cudaStream_t s);
void gather_fri_rows(const uint64_t *f, uint32_t rows, int log_cosets, const uint32_t *d_positions, int npos,
                     uint64_t *d_out, cudaStream_t s);
void gather_digests(const uint32_t *full, const uint32_t *d_idx, int count, uint32_t *d_out, cudaStream_t s);

// running-product columns and batch inversion (poly.cu)
size_t running_product_scratch_elems(int ncols, uint64_t n);
void running_product(const uint64_t *m, size_t m_stride, const uint64_t *d_init, int ncols, uint64_t n, int montgomery,
                     uint64_t *out, size_t out_stride, uint64_t *scratch, cudaStream_t s);
void batch_inverse(const uint64_t *v, uint64_t count, int montgomery, uint64_t *out, cudaStream_t s);

struct DivisorDev {
    uint64_t a;            // numerator degree: (x^a - b)
    uint64_t b;
    uint64_t off_pow_a;    // offset^a
    uint32_t nex;
    uint64_t ex[8];
    const uint64_t *zinv;  // device: 1/(x^a - b) over the N/a distinct values
    uint32_t zn;           // N / a
};
void divisor_inverses(const DivisorDev &d, uint64_t *zinv_out, int logN, PowTable gN, cudaStream_t s);
// rows [i_begin, i_begin + i_count) of the constraint evaluation domain; with cm_logn >= 0 the range and the output
// are coset-major (t = r * n + k for the natural row k * B + r, n = 2^cm_logn, B = 2^cm_logb), the evaluation
// columns stay in natural order
void constraint_combine(const uint64_t *cols, size_t col_stride, const DivisorDev *divs, int ndiv, int logN,
                        uint64_t offset, PowTable gN, uint32_t i_begin, uint32_t i_count, uint64_t *combined,
                        cudaStream_t s, int cm_logn = -1, int cm_logb = 0);

// AIR constraint evaluation over the constraint evaluation domain (include/aero_b200.h, aero_air_program)
struct AirSegs {  // trace segments in order: column c of segment s at lde[s] + c*stride[s], coset-major
    const uint64_t *lde[4];
    size_t stride[4];
    int ncols[4];
    int nseg;
};
constexpr int AIR_MAX_ADJ = 32;  // distinct degree adjustments of one AIR (Miden's ProcessorAir has about twenty)
struct AirProgramDev {       // device copies; field elements canonical
    const uint32_t *nodes;   // 4 words per node (16-byte aligned): op, a, b, destination slot; operands of
                             //   add / sub / mul and t_out are SLOTS (assigned by liveness on the host); a
                             //   periodic node carries its column's offset into `periodic` and length - 1
    const uint64_t *consts;
    const uint64_t *periodic;  // PeriodicValueTable, column after column: cycle * ce_blowup values each
    const uint32_t *t_out, *t_adj;           // per transition constraint: node, index into adj
    const uint32_t *b_col, *b_adj, *b_div;   // per boundary constraint
    const uint64_t *b_val;
    const uint64_t *coeffs;  // pairs: transition constraints, then boundary constraints
    const uint64_t *adj;     // distinct degree adjustments
    const uint64_t *adj_off; // offset^adj for each of them (canonical)
    int n_nodes, n_slots, nt, nb, n_adj, n_div;
};
// threads [tau0, tau0 + tau_count) of the coset-major numbering (tau = rc * n + i): a coset-sharded rank passes
// the cosets it holds, with segs.lde rebased so that (LDE coset, i) indexes its compact storage
void air_evaluate(const AirSegs &segs, const AirProgramDev &p, int logn, int log_blowup, int log_ce, PowTable g_ce /* g_ce^s */,
                  int to_montgomery, uint64_t *out, size_t out_stride, cudaStream_t s, uint32_t tau0, uint32_t tau_count,
                  int num_sms, int blocks_per_sm /* 0 = one thread per step */);

// peak.cu
double measure_alu_peak(int num_sms, uint32_t *scratch, cudaStream_t s);

// fri.cu
// alpha_dev != nullptr: the folding challenge is read from device memory (written by fri_coin)
void fri_fold(const uint64_t *f, uint32_t rows, int log_cosets, uint64_t alpha, const uint64_t *alpha_dev,
              PowTable xinv /* (7 g_M^j)^-1 */, const uint64_t w8inv[4], uint64_t inv8, uint64_t *out, cudaStream_t s);
// fused fold-and-hash: the fold above plus the leaf digests of the folded layer (rows / 8 leaves of 8 values:
// leaf j' = {next[j' + k * rows / 8]}), written to next_leaves[j'].  rows % 8 == 0; for a coset-major layer
// rows / 8 must be a multiple of the number of cosets.
void fri_fold_hash(const uint64_t *f, uint32_t rows, int log_cosets, const uint64_t *alpha_dev, PowTable xinv,
                   const uint64_t w8inv[4], uint64_t inv8, uint64_t *next, uint32_t *next_leaves, cudaStream_t s);

}  // namespace aero
