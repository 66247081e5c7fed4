// blake2s row hashing (K3), Merkle levels (K4), FRI leaf hashing (part of K8), grinding (K10).
//
// K3 replaces Matrix::commit_to_rows' row loop (winterfell/prover/src/matrix.rs:222-242) +
// Blake2s_256::hash_elements (winterfell/crypto/src/hash/blake2s/mod.rs:52-77).
// K4 replaces build_merkle_nodes (winterfell/crypto/src/merkle/mod.rs:316-340).
// K10 replaces ProverChannel::grind_query_seed (winterfell/prover/src/channel.rs:151-167).
#include "blake2s.cuh"
#include "kernels.cuh"

namespace aero {

// One thread per LDE row.  The LDE is stored coset-major: storage row rho = q*n + i (local coset q)
// holds natural row k = B*i + coset_begin + q, so thread rho reads column c at lde[c*col_stride + rho]:
// a fully coalesced 8-byte-per-lane stream.  The digest goes to the leaf stage of the rank that owns
// the leaf block of i (kernels.cuh): coset-major inside the block, so consecutive threads store
// consecutive 32-byte digests -- into local memory or, for foreign blocks, over NVLink.  A sharded proof
// thereby sends each digest to exactly ONE peer (an all-to-all of N*32*(G-1)/G^2 bytes per rank)
// instead of all-gathering the leaves.
//
// BLAKE2s absorbs the row left to right, so the chaining value after the first 2b columns depends
// on those columns only: the kernel hashes the column range [c0, c0 + ncols) of a `total_cols`-wide
// row (c0 even), reading the chaining value left by the previous range from the leaf slot (c0 > 0) and
// leaving its own there.  A whole row in one launch is the case c0 = 0, ncols = total_cols.  This
// lets the row hash of one column batch run on a second stream while the next batch is extended.
__global__ void __launch_bounds__(256) hash_rows_kernel(const uint64_t *__restrict__ lde, size_t col_stride, int c0,
                                                        int ncols, int total_cols, uint32_t nrows, int logn,
                                                        uint32_t coset_begin, int log_nb,
                                                        const __grid_constant__ RankPtrs stage) {
    const uint32_t n_mask = (1u << logn) - 1, nb_mask = (1u << log_nb) - 1;
    const int nblocks = (ncols + 1) >> 1;
    const uint32_t blocks_before = (uint32_t)c0 >> 1;
    const bool tail = (c0 + ncols == total_cols);
    // grid-stride over local storage rows rho (coset q = rho >> logn): the launcher bounds the grid so
    // that an overlapped launch leaves room on every SM for the NTT blocks of the next column batch
    for (uint32_t rho = blockIdx.x * blockDim.x + threadIdx.x; rho < nrows; rho += gridDim.x * blockDim.x) {
        const uint32_t i = rho & n_mask, coset = (rho >> logn) + coset_begin;
        uint32_t *slot = reinterpret_cast<uint32_t *>(stage.p[i >> log_nb]) + ((((size_t)coset << log_nb) + (i & nb_mask)) << 3);
        uint32_t h[8];
        if (c0 == 0) b2s::init(h);
        else load_digest(slot, h);
        const uint64_t *p = lde + (size_t)c0 * col_stride + rho;
        uint64_t e0 = __ldg(p), e1 = ncols > 1 ? __ldg(p + col_stride) : 0ULL;
        for (int b = 0; b < nblocks; b++) {
            uint64_t n0 = 0, n1 = 0;
            if (b + 1 < nblocks) {  // prefetch the next pair while this block is compressed
                n0 = __ldg(p + (size_t)(2 * b + 2) * col_stride);
                if (2 * b + 3 < ncols) n1 = __ldg(p + (size_t)(2 * b + 3) * col_stride);
            }
            const bool last = tail && (b + 1 == nblocks);
            const uint32_t t = last ? 32u * (uint32_t)total_cols : 64u * (blocks_before + (uint32_t)b + 1);
            b2s::compress_pair(h, e0, e1, t, last);
            e0 = n0;
            e1 = n1;
        }
        store_digest(slot, h);
    }
}

// ---- peer exchange helpers (multi-GPU) ----------------------------------------------------------
__global__ void __launch_bounds__(256) peer_push_kernel(RankPtrs buf, int G, int rank, size_t first, size_t count) {
    const uint4 *local = reinterpret_cast<const uint4 *>(buf.p[rank]);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = local[first + i];
        for (int q = 0; q < G; q++)
            if (q != rank) reinterpret_cast<uint4 *>(buf.p[q])[first + i] = v;
    }
}
void peer_push(const RankPtrs &buf, int G, int rank, size_t off, size_t bytes, cudaStream_t s) {
    if (G <= 1 || bytes == 0) return;
    const size_t count = bytes / 16;
    size_t blocks = (count + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    AERO_COUNT_LAUNCH(1);
    peer_push_kernel<<<(unsigned)blocks, 256, 0, s>>>(buf, G, rank, off / 16, count);
}
__global__ void __launch_bounds__(256) peer_push_words_kernel(RankPtrs buf, int G, int rank, size_t first, size_t count) {
    const uint64_t *local = reinterpret_cast<const uint64_t *>(buf.p[rank]);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t v = local[first + i];
        for (int q = 0; q < G; q++)
            if (q != rank) reinterpret_cast<uint64_t *>(buf.p[q])[first + i] = v;
    }
}
void peer_push_words(const RankPtrs &buf, int G, int rank, size_t first, size_t count, cudaStream_t s) {
    if (G <= 1 || count == 0) return;
    size_t blocks = (count + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    AERO_COUNT_LAUNCH(1);
    peer_push_words_kernel<<<(unsigned)blocks, 256, 0, s>>>(buf, G, rank, first, count);
}
__global__ void peer_flag_kernel(volatile unsigned long long *flag, unsigned long long epoch) {
    __threadfence_system();
    *flag = epoch;
    __threadfence_system();
}
void peer_send(const RankPtrs &buf, int rank, int dest, size_t off, size_t bytes, unsigned long long *dest_flag,
               unsigned long long epoch, cudaStream_t s) {
    // One destination at a time, through a copy engine: every SM stays with the LDE running beside it.  The
    // flag kernel starts when the copy has completed.
    AERO_COUNT_LAUNCH(1);
    if (bytes)
        cudaMemcpyAsync((uint8_t *)buf.p[dest] + off, (const uint8_t *)buf.p[rank] + off, bytes, cudaMemcpyDeviceToDevice, s);
    peer_flag_kernel<<<1, 1, 0, s>>>(dest_flag, epoch);
}
__global__ void peer_wait_kernel(const volatile unsigned long long *flag, int nflags, unsigned long long epoch,
                                 unsigned int *d_timeout) {
    if ((int)threadIdx.x >= nflags) return;
    flag += threadIdx.x;
    const long long t0 = clock64();
    while (*flag < epoch) {
        if (clock64() - t0 > 8000000000LL) {
            atomicAdd(d_timeout, 1u);
            break;
        }
    }
    __threadfence_system();
}
void peer_wait(const unsigned long long *flags, int nflags, unsigned long long epoch, unsigned int *d_timeout, cudaStream_t s) {
    AERO_COUNT_LAUNCH(1);
    peer_wait_kernel<<<1, 32, 0, s>>>(flags, nflags, epoch, d_timeout);
}
// The kernels of the device-side exchange are first LAUNCHED by the second proof of a shape, when the
// peers may already be spinning in a barrier -- and with lazy module loading (the CUDA 12 default) a first
// launch loads the kernel, which can synchronise the device.  Load them while nothing spins.
void preload_exchange_kernels();
// Stream-ordered barrier between the ranks of a proof: thread q publishes `epoch` in rank q's flag
// slot for this rank (after a system-scope fence, so every peer store of earlier kernels on this
// stream is visible first), then waits for rank q's flag in the local window.  A peer that never
// arrives (it failed) is given ~4 s; the time-out is reported through *d_timeout instead of hanging.
__global__ void peer_barrier_kernel(RankPtrs flags, int G, int rank, unsigned long long epoch, unsigned int *d_timeout) {
    const int q = threadIdx.x;
    if (q >= G || q == rank) return;
    volatile unsigned long long *mine = reinterpret_cast<volatile unsigned long long *>(flags.p[rank]);
    __threadfence_system();
    reinterpret_cast<volatile unsigned long long *>(flags.p[q])[rank] = epoch;
    __threadfence_system();
    const long long t0 = clock64();
    while (mine[q] < epoch) {
        if (clock64() - t0 > 8000000000LL) {
            atomicAdd(d_timeout, 1u);
            break;
        }
    }
    __threadfence_system();
}
void preload_exchange_kernels() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, peer_barrier_kernel);
    cudaFuncGetAttributes(&a, peer_flag_kernel);
    cudaFuncGetAttributes(&a, peer_wait_kernel);
    cudaFuncGetAttributes(&a, peer_push_kernel);
    cudaFuncGetAttributes(&a, peer_push_words_kernel);
}
void peer_barrier(const RankPtrs &flags, int G, int rank, unsigned long long epoch, unsigned int *d_timeout,
                  cudaStream_t s) {
    if (G <= 1) return;
    AERO_COUNT_LAUNCH(1);
    peer_barrier_kernel<<<1, 32, 0, s>>>(flags, G, rank, epoch, d_timeout);
}

// stage[c*n + i] -> out[B*i + c] (32 bytes each; single-rank segments only)
__global__ void leaves_to_natural_kernel(const uint4 *__restrict__ stage, uint4 *__restrict__ out, int logn, int log_blowup) {
    const size_t total = (size_t)2 << (logn + log_blowup);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t k = t >> 1;
        const size_t c = k & (((size_t)1 << log_blowup) - 1), i = k >> log_blowup;
        out[t] = stage[(((c << logn) + i) << 1) | (t & 1)];
    }
}
void leaves_to_natural(const uint32_t *stage, uint32_t *out, int logn, int log_blowup, cudaStream_t s) {
    const size_t total = (size_t)2 << (logn + log_blowup);
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    AERO_COUNT_LAUNCH(1);
    leaves_to_natural_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const uint4 *>(stage),
                                                             reinterpret_cast<uint4 *>(out), logn, log_blowup);
}

// Generic variant: rows of a plain column-major matrix in natural order (used for small inputs /
// tests): leaf k = hash(m[c][k]).
__global__ void __launch_bounds__(256) hash_rows_natural_kernel(const uint64_t *__restrict__ m, size_t col_stride,
                                                                int ncols, uint32_t nrows,
                                                                uint32_t *__restrict__ leaves) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrows) return;
    uint32_t h[8];
    b2s::init(h);
    const int nblocks = (ncols + 1) >> 1;
    for (int b = 0; b < nblocks; b++) {
        const uint64_t e0 = m[(size_t)(2 * b) * col_stride + k];
        const uint64_t e1 = (2 * b + 1 < ncols) ? m[(size_t)(2 * b + 1) * col_stride + k] : 0ULL;
        const bool last = (b + 1 == nblocks);
        const uint32_t t = last ? 32u * (uint32_t)ncols : 64u * (uint32_t)(b + 1);
        b2s::compress_pair(h, e0, e1, t, last);
    }
    store_digest(leaves + (size_t)k * 8, h);
}

// Merkle tree over `full` = 2*N digests: full[N + k] = leaf k, full[i] = merge(full[2i], full[2i+1])
// for 1 <= i < N (heap layout of merkle/mod.rs:316-340; full[1] is the root, full[0] unused/zero).
// A block takes `width` consecutive nodes of the node level that starts at heap index `level_size`
// (children from global memory) and climbs `levels` levels through shared memory; every level is
// written to `full`.  Wide levels climb only MERKLE_LEVELS_PER_LAUNCH levels per launch so that
// every warp stays full (256, 128, 64, 32 nodes); the last launch (<= 256 nodes) runs to the root.
constexpr int MERKLE_LEVELS_PER_LAUNCH = 4;
// stage != nullptr: the children of the bottom level are the leaf digests of a segment's coset-major
// leaf stage ([B][nb]: leaf B*il + c at stage[c*nb + il], kernels.cuh) instead of heap nodes.
// Digests travel between levels through shared memory WORD-MAJOR (sbuf[word][node]): a warp's stores
// of one word are 32 consecutive banks and its loads of the (2t, 2t+1) child pair one 8-byte access per
// lane, both conflict-free (the node-major layout of round 1 had 8-way conflicts on every access).
__global__ void __launch_bounds__(256) merkle_subtree_kernel(uint32_t *__restrict__ full, uint32_t level_size, int width,
                                                             int levels, const uint32_t *__restrict__ stage, uint32_t nb,
                                                             int log_blowup) {
    __shared__ __align__(16) uint32_t sbuf[2][8][256];
    int cur = 0;
    // bottom level: children from global memory
    {
        const uint32_t first = level_size + blockIdx.x * width;
        for (int t = threadIdx.x; t < width; t += blockDim.x) {
            const uint32_t node = first + t;
            uint32_t a[8], b[8], o[8];
            if (stage) {
                const uint32_t j = node - level_size;                   // pair index: leaves 2j, 2j + 1
                const uint32_t il = j >> (log_blowup - 1), c = (j & ((1u << (log_blowup - 1)) - 1)) << 1;
                load_digest(stage + ((size_t)c * nb + il) * 8, a);
                load_digest(stage + ((size_t)(c + 1) * nb + il) * 8, b);
            } else {
                load_digest(full + (size_t)(2 * node) * 8, a);
                load_digest(full + (size_t)(2 * node + 1) * 8, b);
            }
            b2s::merge(a, b, o);
            store_digest(full + (size_t)node * 8, o);
#pragma unroll
            for (int i = 0; i < 8; i++) sbuf[cur][i][t] = o[i];
        }
    }
    for (int l = 1; l < levels; l++) {
        __syncthreads();
        width >>= 1;
        const uint32_t first = (level_size >> l) + blockIdx.x * width;
        for (int t = threadIdx.x; t < width; t += blockDim.x) {
            uint32_t a[8], b[8], o[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint2 ab = *reinterpret_cast<const uint2 *>(&sbuf[cur][i][2 * t]);
                a[i] = ab.x;
                b[i] = ab.y;
            }
            b2s::merge(a, b, o);
            store_digest(full + (size_t)(first + t) * 8, o);
#pragma unroll
            for (int i = 0; i < 8; i++) sbuf[cur ^ 1][i][t] = o[i];
        }
        cur ^= 1;
    }
}

// levels from the node level of size `level` (children: heap nodes, or the leaf stage) up to the root
static void merkle_levels(uint32_t *full, uint64_t level, const uint32_t *stage, uint32_t nb, int log_blowup, cudaStream_t s) {
    while (level >= 1) {
        AERO_COUNT_LAUNCH(1);
        if (level > 256) {
            merkle_subtree_kernel<<<(unsigned)(level / 256), 256, 0, s>>>(full, (uint32_t)level, 256, MERKLE_LEVELS_PER_LAUNCH,
                                                                        stage, nb, log_blowup);
            level >>= MERKLE_LEVELS_PER_LAUNCH;
        } else {
            int levels = 0;
            for (uint64_t l = level; l >= 1; l >>= 1) levels++;
            merkle_subtree_kernel<<<1, 256, 0, s>>>(full, (uint32_t)level, (int)level, levels, stage, nb, log_blowup);
            break;
        }
        stage = nullptr;
    }
}
void merkle_build(uint32_t *full, uint64_t num_leaves, cudaStream_t s) {
    // node levels have sizes num_leaves/2, ..., 1 ; the level of size L occupies heap indices [L, 2L)
    merkle_levels(full, num_leaves / 2, nullptr, 0, 0, s);
}
void merkle_build_block(const uint32_t *stage, uint32_t *heap, uint32_t nb, int log_blowup, cudaStream_t s) {
    merkle_levels(heap, ((uint64_t)nb << log_blowup) / 2, stage, nb, log_blowup, s);
}
__global__ void merkle_push_subroot_kernel(const uint32_t *__restrict__ heap, RankPtrs top, int G, int rank) {
    const int r = threadIdx.x >> 3, w = threadIdx.x & 7;
    if (r < G) reinterpret_cast<uint32_t *>(top.p[r])[(size_t)(G + rank) * 8 + w] = heap[8 + w];
}
void merkle_push_subroot(const uint32_t *heap, const RankPtrs &top, int G, int rank, cudaStream_t s) {
    AERO_COUNT_LAUNCH(1);
    merkle_push_subroot_kernel<<<1, 8 * AERO_MAX_RANKS, 0, s>>>(heap, top, G, rank);
}
__global__ void merkle_top_kernel(uint32_t *__restrict__ top, int G) {
    if (threadIdx.x != 0) return;
    for (int j = G - 1; j >= 1; j--) {
        uint32_t a[8], b[8], o[8];
        load_digest(top + (size_t)(2 * j) * 8, a);
        load_digest(top + (size_t)(2 * j + 1) * 8, b);
        b2s::merge(a, b, o);
        store_digest(top + (size_t)j * 8, o);
    }
}
void merkle_top(uint32_t *top, int G, cudaStream_t s) {
    if (G <= 1) return;
    AERO_COUNT_LAUNCH(1);
    merkle_top_kernel<<<1, 32, 0, s>>>(top, G);
}

void hash_rows_lde(const uint64_t *lde, size_t col_stride, int c0, int ncols, int total_cols, int logn,
                   uint32_t coset_begin, uint32_t nrows, int log_nb, const RankPtrs &stage, int max_blocks,
                   cudaStream_t s) {
    if (nrows == 0 || ncols == 0) return;
    uint32_t grid = (nrows + 255) / 256;
    if (max_blocks > 0 && grid > (uint32_t)max_blocks) grid = (uint32_t)max_blocks;
    static DeviceOnce once;
    // same shared-memory carve-out as the NTT passes, so both can be resident on one SM
    once.run([] { cudaFuncSetAttribute(hash_rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); });
    AERO_COUNT_LAUNCH(1);
    hash_rows_kernel<<<grid, 256, 0, s>>>(lde, col_stride, c0, ncols, total_cols, nrows, logn, coset_begin, log_nb, stage);
}
void hash_rows_natural(const uint64_t *m, size_t col_stride, int ncols, uint32_t nrows, uint32_t *leaves,
                       cudaStream_t s) {
    AERO_COUNT_LAUNCH(1);
    hash_rows_natural_kernel<<<(nrows + 255) / 256, 256, 0, s>>>(m, col_stride, ncols, nrows, leaves);
}

// FRI leaf j (j < rows = M/8) = hash_elements(f[j + k*rows], k = 0..7): 256-byte message, 4 blocks
// (fri/src/prover/mod.rs:202-203: transpose_slice + hash_values).  `log_cosets` > 0 means f is
// stored coset-major (natural q at (q & (B-1))*(M/B) + (q >> b)), as the DEEP LDE is.
__global__ void __launch_bounds__(256) fri_leaf_hash_kernel(const uint64_t *__restrict__ f, uint32_t rows,
                                                            int log_cosets, uint32_t *__restrict__ leaves) {
    const uint32_t tau = blockIdx.x * blockDim.x + threadIdx.x;
    if (tau >= rows) return;
    uint64_t v[8];
    uint32_t j;
    fri_gather8(f, rows, log_cosets, tau, j, v);
    uint32_t h[8];
    b2s::init(h);
    b2s::compress_pair(h, v[0], v[1], 64u, false);
    b2s::compress_pair(h, v[2], v[3], 128u, false);
    b2s::compress_pair(h, v[4], v[5], 192u, false);
    b2s::compress_pair(h, v[6], v[7], 256u, true);
    store_digest(leaves + (size_t)j * 8, h);
}
void fri_leaf_hash(const uint64_t *f, uint32_t rows, int log_cosets, uint32_t *leaves, cudaStream_t s) {
    AERO_COUNT_LAUNCH(1);
    fri_leaf_hash_kernel<<<(rows + 255) / 256, 256, 0, s>>>(f, rows, log_cosets, leaves);
}

__global__ void fri_coin_kernel(uint32_t *__restrict__ seed, const uint32_t *__restrict__ root,
                                uint64_t *__restrict__ alpha_out, uint32_t *__restrict__ root_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t s[8], r[8], t[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s[i] = seed[i];
        r[i] = root[i];
    }
    b2s::merge(s, r, t);
    uint64_t alpha = ~0ULL;
    for (uint64_t counter = 1; counter <= 1000; counter++) {
        b2s::merge_with_int(t, counter, o);
        const uint64_t v = ((uint64_t)o[1] << 32) | o[0];
        if (v < gl::P) {  // from_random_bytes rejects non-canonical values (f64/mod.rs:438-454)
            alpha = v;
            break;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        seed[i] = t[i];
        root_out[i] = r[i];
    }
    *alpha_out = alpha;
}
void fri_coin(uint32_t *seed, const uint32_t *root, uint64_t *alpha_out, uint32_t *root_out, cudaStream_t s) {
    AERO_COUNT_LAUNCH(1);
    fri_coin_kernel<<<1, 32, 0, s>>>(seed, root, alpha_out, root_out);
}

// Grinding: smallest nonce >= 1 whose digest head has >= `bits` trailing zero bits.  Batches are
// scanned in ascending order; inside a batch atomicMin keeps the smallest hit, so the result is
// the global minimum, matching the reference's serial `find` (channel.rs:154-157).
__global__ void __launch_bounds__(256) pow_search_kernel(const uint32_t *__restrict__ seed, uint64_t base,
                                                         uint32_t bits, unsigned long long *__restrict__ best) {
    const uint64_t nonce = base + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = seed[i];
    b2s::merge_with_int(s, nonce, o);
    const uint64_t head = ((uint64_t)o[1] << 32) | o[0];
    const uint32_t tz = head ? (uint32_t)(__ffsll((long long)head) - 1) : 64u;
    if (tz >= bits) atomicMin(best, (unsigned long long)nonce);
}
void pow_search(const uint32_t *seed, uint64_t base, uint32_t count, uint32_t bits, unsigned long long *best,
                cudaStream_t s) {
    AERO_COUNT_LAUNCH(1);
    pow_search_kernel<<<count / 256, 256, 0, s>>>(seed, base, bits, best);
}

}  // namespace aero
