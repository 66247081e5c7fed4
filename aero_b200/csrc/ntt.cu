// Goldilocks NTT kernels for sm_100a: shared-memory DIT transforms with register radix-8 rounds,
// four-step two-pass decomposition for n > 2^11.  See ntt.cuh for the mapping to the reference.
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include <cuda.h>  // CUtensorMap (types only: the encoder is resolved through cudaGetDriverEntryPoint)

#include "gl.cuh"
#include "ntt.cuh"
#include "kernels.cuh"

namespace aero {

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return __brev(x) >> (32 - bits); }

// 8-byte asynchronous global -> shared copy (LDGSTS): a tile is fetched with every load of a thread in
// flight at once and no register staging; a plain load/store loop exposes one DRAM latency per
// iteration (ncu: 40 % of the stall samples of the first version sat on that loop's STS).
__device__ __forceinline__ void cp_async8(uint64_t *smem_dst, const uint64_t *gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8_saddr(unsigned sa, const uint64_t *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- TMA (bulk asynchronous copies completing on an mbarrier) --------------------------------------
// The 2^9 / 2^10-point passes (BULK instantiations) fetch their tile and twiddles with a handful of TMA
// operations issued by one thread -- a 1-D cp.async.bulk where the tile is contiguous (pass 2, twiddles),
// a tiled cp.async.bulk.tensor through a CUtensorMap where it is M rows of T words at a row stride
// (pass 1) -- instead of 32 LDGSTS + BREV + 64-bit add per thread.  The tile lands DENSE and in natural
// row order; the first DIT round reads its bit-reversed inputs from there (dit_round<DENSE>).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// Spins until the phase with the given parity has completed.  A copy that can never complete (which would
// be a bug in the byte count) traps after ~1 s instead of hanging the device.
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spins = 0; !done; spins++) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
        if (spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(unsigned smem_dst, const void *gmem_src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void tensor4d_g2s(unsigned smem_dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_dst),
                 "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tensor3d_g2s(unsigned smem_dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_dst),
                 "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// Shared-memory layout of an M x T tile: row `row` starts at word row*RS + T*(row >> r1), RS = T + 1 for
// T > 1 (so that the transposed accesses of the store stages spread over the banks), plus T words of
// padding after every 2^r1 rows, r1 = log2 of the first round's size.  Without the group padding the
// 32/T row groups a warp touches in the first round (rows 2^r1 apart: 2^r1 * RS words, a multiple of the
// 32 banks) all fell on the same banks -- 4-way conflicts for 8-wide tiles, 8-way for 4-wide ones (ncu:
// 46-51 % of the shared wavefronts of a pass were conflicts); with it they tile the banks exactly.  In
// every round the rows of one butterfly group stay an arithmetic progression, so the padding costs no
// instruction there.
// Measured: 8-wide tiles (2^10-point passes) do not gain from it (LDE of 72 columns x 2^20: 9.99 ms without,
// 10.2 ms with -- those passes are bound by ALU issue and two of the four wavefronts were already hidden),
// so they keep the plain layout; the 4-wide tiles of 2^11 / 2^12-point passes (traces of 2^22 rows and
// more), where every first-round access was an 8-way conflict, and single-pass transforms use it.
template <int T>
__host__ __device__ constexpr int tile_pad() { return T == 8 ? 0 : T; }
template <int T, int RS>
__host__ __device__ __forceinline__ int tile_off(int row, int r1) {
    return row * RS + tile_pad<T>() * (row >> r1);
}
template <int T, int RS>
__host__ __device__ __forceinline__ int tile_words(int logM, int r1) {
    return (RS << logM) + tile_pad<T>() * ((1 << logM) >> r1);
}

// Fetches an M x T tile (M = 2^logM rows of T consecutive words, row j at g + (j << gsh_row)) into
// shared memory with row j stored at row bitrev_logM(j).  blockDim.x = J*T threads, J a power of two
// <= M: thread (j0 = tid / T, t = tid % T) owns rows j0 + k*J, whose bit-reversed positions are the
// 2^m consecutive rows starting at bitrev_logJ(j0) << m  (m = logM - logJ).  Walking those in order
// leaves one BREV and one 64-bit add per element; the plain loop over `it` spent ~20 instructions per
// element on index arithmetic (ncu: a fifth of all instructions of a pass).
template <int T, int RS>
__device__ __forceinline__ void load_tile_bitrev(uint64_t *a, const uint64_t *g, int logM, int gsh_row, int r1) {
    const int logJ = 31 - __clz((int)blockDim.x / T);
    const int m = logM - logJ;
    const int t = threadIdx.x % T, j0 = threadIdx.x / T;
    const uint32_t row0 = (logJ ? (__brev((uint32_t)j0) >> (32 - logJ)) : 0u) << m;
    const uint64_t *gp = g + ((size_t)j0 << gsh_row) + t;
    const int gsh = logJ + gsh_row;
    if (m == 0) {
        cp_async8(a + tile_off<T, RS>((int)row0, r1) + t, gp);
        return;
    }
    const int rsh = 32 - m;
    if (m <= r1) {  // the 2^m rows share one padding group
        const unsigned sa = (unsigned)__cvta_generic_to_shared(a + tile_off<T, RS>((int)row0, r1) + t);
#pragma unroll 8
        for (int r = 0; r < (1 << m); r++)
            cp_async8_saddr(sa + r * (RS * 8), gp + ((size_t)(__brev((uint32_t)r) >> rsh) << gsh));
    } else {
        for (int r = 0; r < (1 << m); r++)
            cp_async8(a + tile_off<T, RS>((int)row0 + r, r1) + t, gp + ((size_t)(__brev((uint32_t)r) >> rsh) << gsh));
    }
}

// ---- compile-time structure of a 2^R-point round ------------------------------------------------
template <int N, class F, int I = 0>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<N, F, I + 1>(static_cast<F &&>(f));
    }
}
// w_{2^j} = 2^dft_unit_exp(j): TWO_ADIC_ROOT^(2^(32-j)) is 2^(39 * 2^(6-j)) for j <= 6 (2^192 = 1)
__host__ __device__ constexpr int dft_unit_exp(int j, bool inv) {
    const int u = (39 << (6 - j)) % 192;
    return inv ? (192 - u) % 192 : u;
}
// twiddle w_{2^(q+1)}^k of inner stage q as an exponent of two in [0, 192)
__host__ __device__ constexpr int dft_tw_exp(int q, int k, bool inv) { return (dft_unit_exp(q + 1, inv) * k) % 192; }
// Must x[e] be canonical when it enters inner stage q?  Yes where a unit twiddle hands it to a
// butterfly as v unchanged, and for the u of a butterfly one of whose outputs must be canonical
// (add_cc / sub_ac give canonical sums only from canonical u).
__host__ __device__ constexpr bool dft_need_canon(int R, int q, int e) {
    if (q >= R) return false;
    const int bit = 1 << q;
    if (e & bit) return (e & (bit - 1)) == 0;
    return dft_need_canon(R, q + 1, e) || dft_need_canon(R, q + 1, e | bit);
}

// One register-resident round of a mixed-radix DIT over an M x T tile in shared memory
// (a[idx*RS + t]): the 2^R sub-transforms of size 2^s0 that merge into one of size m = 2^(s0+R).
// Input e (a bit-reversed residue) is first multiplied by the general twiddle
// tw[(e << s0) + low] = (sigma_m w_m^low)^bitrev(e); what remains is a plain 2^R-point DFT whose
// twiddles are powers of two (gl::mul_pow2), negated where the exponent is >= 96.
// PLAIN0: first round (s0 == 0) of a transform without coset shift: all general twiddles are 1.
// IN_CANON: the tile holds canonical values already (pass 2 reads what pass 1 stored with
// mul_canon), so a PLAIN0 round has nothing to canonicalise on the way in.
// The 2^R-point DFT network of a round on registers: x[e] holds the input of bit-reversed residue e (any
// u64 congruent to the value; canonical where dft_need_canon(R, 0, e)), afterwards x[k] is output k
// ("any").  All twiddles are powers of two (gl::mul_pow2), negated where the exponent is >= 96.
template <int R, bool INV>
__device__ __forceinline__ void dft_network(uint64_t (&x)[1 << R]) {
    static_for<R>([&](auto Q) {
        constexpr int q = decltype(Q)::value;
        static_for<(1 << R)>([&](auto E) {
            constexpr int e = decltype(E)::value;
            if constexpr ((e & (1 << q)) == 0) {
                constexpr int f = e | (1 << q);
                constexpr int ex = dft_tw_exp(q, e & ((1 << q) - 1), INV);
                constexpr bool canon_sum = dft_need_canon(R, q + 1, e);
                const uint64_t u = x[e];
                if constexpr (ex == 0) {
                    const uint64_t v = x[f];
                    x[e] = canon_sum ? gl::add_cc(u, v) : gl::add_ac(u, v);
                    x[f] = gl::sub_ac(u, v);
                } else if constexpr (ex < 96) {
                    const uint64_t v = gl::mul_pow2<ex>(x[f]);
                    x[e] = canon_sum ? gl::add_cc(u, v) : gl::add_ac(u, v);
                    x[f] = gl::sub_ac(u, v);
                } else {
                    const uint64_t v = gl::mul_pow2<ex - 96>(x[f]);  // twiddle = -2^(ex-96)
                    static_assert(!canon_sum, "only unit-twiddle butterflies feed canonical sums");
                    x[e] = gl::sub_ac(u, v);
                    x[f] = gl::add_ac(u, v);
                }
            }
        });
    });
}

__host__ __device__ constexpr int dft_bitrev(int e, int bits) {
    int r = 0;
    for (int b = 0; b < bits; b++) r |= ((e >> b) & 1) << (bits - 1 - b);
    return r;
}
// DENSE (first round of a BULK pass): the tile was delivered by TMA, dense (row j at a[j*T]) and in NATURAL
// row order.  Group p of the round (bit-reversed positions p*2^R + e) is the natural rows
// bitrev_R(e)*G + q, G = M >> R, q = bitrev(p): thread (q, t) reads those -- a warp reads 32 consecutive
// words per e -- and, after a barrier (every thread owns exactly one item, so all reads of the dense tile
// precede all writes), stores its outputs to rows (p << R) + e of the padded layout, in place.
template <int R, int T, int RS, bool PLAIN0, bool INV, bool IN_CANON = false, bool FIRST = PLAIN0, bool DENSE = false>
__device__ __forceinline__ void dit_round(uint64_t *a, const uint64_t *tw, int s0, int logM, int r1) {
    static_assert(!DENSE || FIRST, "only the first round reads the dense tile");
    if (FIRST) s0 = 0;
    const int ngroups = (1 << logM) >> R;
    const int items = ngroups * T;
    // rows base + (e << s0) of a group: words astep apart.  First round (s0 = 0, R = r1): the group is
    // one padding group, RS apart -- compile-time offsets; later rounds have s0 >= r1 (see tile_off).
    const int astep = FIRST ? RS : (RS << s0) + (tile_pad<T>() << (s0 - r1));
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int t = it % T, g = it / T;
        const int low = g & ((1 << s0) - 1);
        int base = ((g >> s0) << (s0 + R)) | low;
        uint64_t x[1 << R];
        if constexpr (DENSE) {
            const int lg = logM - R;
            base = (lg ? (int)(__brev((uint32_t)g) >> (32 - lg)) : 0) << R;
            const uint64_t *dg = a + g * T + t;
            const int dstep = T << lg;
            static_for<(1 << R)>([&](auto E) {
                constexpr int e = decltype(E)::value;
                x[e] = dg[dft_bitrev(e, R) * dstep];
            });
            __syncthreads();
        }
        uint64_t *ag = a + tile_off<T, RS>(base, r1) + t;
        // tile values are "any" (congruent mod p, < 2^64); products are canonical
        static_for<(1 << R)>([&](auto E) {
            constexpr int e = decltype(E)::value;
            uint64_t v;
            if constexpr (DENSE) v = x[e];
            else v = ag[e * astep];
            if constexpr (e > 0 && !PLAIN0) v = gl::mul_canon(v, tw[(e << s0) + low]);
            else if constexpr (!IN_CANON && dft_need_canon(R, 0, e)) v = gl::canon_any(v);
            x[e] = v;
        });
        dft_network<R, INV>(x);
        static_for<(1 << R)>([&](auto E) {
            constexpr int e = decltype(E)::value;
            ag[e * astep] = x[e];
        });
    }
}

template <int T, int RS, bool PLAIN0, bool INV, int RMAX, bool IN_CANON = false, bool FIRST = PLAIN0, bool DENSE = false>
__device__ __forceinline__ void dit_round_dispatch(int R, uint64_t *a, const uint64_t *tw, int s0, int logM, int r1) {
    if constexpr (RMAX >= 5) {
        if (DENSE || R == 5) {  // a BULK pass starts with a 32-point round by construction (see launch_two_pass)
            dit_round<5, T, RS, PLAIN0, INV, IN_CANON, FIRST, DENSE>(a, tw, s0, logM, r1);
            return;
        }
    }
    static_assert(!DENSE || RMAX >= 5, "dense tiles are read by 32-point first rounds only");
    switch (R) {
    case 4: dit_round<4, T, RS, PLAIN0, INV, IN_CANON, FIRST>(a, tw, s0, logM, r1); break;
    case 3: dit_round<3, T, RS, PLAIN0, INV, IN_CANON, FIRST>(a, tw, s0, logM, r1); break;
    case 2: dit_round<2, T, RS, PLAIN0, INV, IN_CANON, FIRST>(a, tw, s0, logM, r1); break;
    default: dit_round<1, T, RS, PLAIN0, INV, IN_CANON, FIRST>(a, tw, s0, logM, r1); break;
    }
}

// Full M-point DIT over the tile: input in bit-reversed row order, output in natural row order.
template <int T, int RS, bool PLAIN, bool INV, int RMAX, bool IN_CANON = false, bool DENSE = false>
__device__ __forceinline__ void dit_tile(uint64_t *a, const uint64_t *tw, int logM) {
    const NttRounds rounds(logM);
    const int r1 = rounds.log(0);
    // first round: s0 == 0 is a compile-time fact, which lets PLAIN transforms skip the unit twiddles
    dit_round_dispatch<T, RS, PLAIN, INV, RMAX, IN_CANON, true, DENSE>(r1, a, tw, 0, logM, r1);
    __syncthreads();
    int s0 = r1;
    for (int i = 1; i < rounds.count; i++) {
        const int R = rounds.log(i);
        dit_round_dispatch<T, RS, false, INV, RMAX>(R, a, tw, s0, logM, r1);
        s0 += R;
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
// TAB: the inter-pass factor F(i1, j2) = c s^j2 w_n^(i1 j2) comes from a full table in that same tile
// order (read through L2); otherwise it is advanced by a running product per thread, which costs a
// second multiplication per element.
// MAXT: largest block the instantiation is launched with.  The 256-thread variants have 128+ registers
// per thread and are the only ones that contain 32-point rounds (64 data registers).
// BULK: tile through the tensor map `tmap` ({n2, n1, ncols} words, box {T, min(n1, 256), 1}), twiddles by a 1-D
// bulk copy, both completing on one mbarrier; one work item per thread in the first round.
// colfast: blockIdx.x enumerates (column, coset) -- column fastest -- and blockIdx.y the tile, so that the
// blocks resident at one time share their slice of the inter-pass table (16 columns read it from L2 for one
// fetch from DRAM) and their source tile (8 cosets); otherwise grid = (tile, coset, column).
template <int T, bool PLAIN, bool INV, bool TAB, int MAXT, bool BULK = false>
__global__ void __launch_bounds__(MAXT) dft_pass1_kernel(const uint64_t *__restrict__ src, uint64_t *__restrict__ tmp,
                                                        const uint64_t *__restrict__ stage1,
                                                        const uint64_t *__restrict__ inter_b,
                                                        const uint64_t *__restrict__ inter_full,
                                                        const uint64_t *__restrict__ wlo,
                                                        const uint64_t *__restrict__ whi, int lo_bits, int log1,
                                                        int log2, size_t src_col_stride, int ncosets, int ncols, int colfast,
                                                        const __grid_constant__ CUtensorMap tmap) {
    constexpr int RS = T + 1;
    extern __shared__ __align__(128) uint64_t smem[];
    const int n1 = 1 << log1;
    const size_t n = (size_t)1 << (log1 + log2);
    const int r1 = NttRounds(log1).log(0);
    uint64_t *a = smem;                                  // the tile (tile_words)
    uint64_t *tw = smem + tile_words<T, RS>(log1, r1);   // n1
    int coset = blockIdx.y, col = blockIdx.z;
    uint32_t j2_0 = blockIdx.x * T;
    if (colfast) {
        col = blockIdx.x % ncols;
        coset = blockIdx.x / ncols;
        j2_0 = blockIdx.y * T;
    }
    const uint64_t *s = src + (size_t)col * src_col_stride;
    const int nchunks = n1 / T;
    const uint64_t *ftab = nullptr;
    unsigned mb = 0;
    if constexpr (BULK) {
        __shared__ __align__(8) uint64_t mbar;
        mb = smem_u32(&mbar);
        if (threadIdx.x == 0) mbar_init(mb, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const int rows = n1 < 256 ? n1 : 256;
            mbar_expect_tx(mb, (unsigned)((n1 * T + n1) * 8));
            for (int r = 0; r < n1; r += rows) tensor3d_g2s(smem_u32(a + r * T), &tmap, (int)j2_0, r, col, mb);
            bulk_g2s(smem_u32(tw), stage1 + (size_t)coset * n1, (unsigned)(n1 * 8), mb);
        }
    }
    if (TAB) {
        // this block's slice of the table: T*T consecutive entries in each of the nchunks tiles
        ftab = inter_full + (size_t)coset * n + (size_t)j2_0 * T;
        constexpr int LINE = 16;  // 128-byte lines, in entries
        constexpr int per_chunk = (T * T + LINE - 1) / LINE;
        for (int i = threadIdx.x; i < nchunks * per_chunk; i += blockDim.x) {
            const uint64_t *q = ftab + (((size_t)(i / per_chunk) << log2) * T) + (size_t)(i % per_chunk) * LINE;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        }
    }
    if constexpr (BULK) {
        mbar_wait(mb, 0);
    } else {
        for (int i = threadIdx.x; i < n1; i += blockDim.x) cp_async8(tw + i, stage1 + (size_t)coset * n1 + i);
        load_tile_bitrev<T, RS>(a, s + j2_0, log1, log2, r1);   // row j1 at s[(j1 << log2) + j2_0 + t]
        cp_async_wait_all();
        __syncthreads();
    }
    dit_tile<T, RS, PLAIN, INV, (MAXT <= 256 ? 5 : 4), false, BULK>(a, tw, log1);
    uint64_t *o = tmp + ((size_t)col * ncosets + coset) * n;
    const int it0 = threadIdx.x;
    const int ii = it0 % T, t = (it0 / T) % T;
    const uint32_t j2 = j2_0 + t;
    const int cstep = blockDim.x / (T * T);   // blockDim.x is a multiple of T*T
    const uint32_t step_i1 = cstep * T;
    int c = it0 / (T * T);
    uint32_t i1 = c * T + ii;
    if (TAB) {
        // table loads of U chunks in flight before the first product needs one
        constexpr int U = 8;
        const size_t cstride = ((size_t)cstep << log2) * T;       // entries between this thread's chunks
        const uint64_t *fp = ftab + ((size_t)c << log2) * T + t * T + ii;
        uint64_t *op = o + ((size_t)c << log2) * T + (size_t)j2 * T + ii;
        if ((step_i1 & ((1u << r1) - 1)) == 0) {  // this thread's rows are whole padding groups apart
            const uint64_t *ap = a + tile_off<T, RS>((int)i1, r1) + t;
            const int astride = step_i1 * RS + tile_pad<T>() * (step_i1 >> r1);
            for (; c + (U - 1) * cstep < nchunks; c += U * cstep) {
                uint64_t f[U];
#pragma unroll
                for (int u = 0; u < U; u++) f[u] = __ldg(fp + u * cstride);
#pragma unroll
                for (int u = 0; u < U; u++) op[u * cstride] = gl::mul_canon(ap[u * astride], f[u]);
                fp += U * cstride;
                op += U * cstride;
                ap += U * astride;
            }
            for (; c < nchunks; c += cstep) {
                *op = gl::mul_canon(*ap, __ldg(fp));
                fp += cstride;
                op += cstride;
                ap += astride;
            }
        } else {  // small transforms
            for (; c < nchunks; c += cstep, i1 += step_i1) {
                *op = gl::mul_canon(a[tile_off<T, RS>((int)i1, r1) + t], __ldg(fp));
                fp += cstride;
                op += cstride;
            }
        }
    } else {
        // F(i1, j2) = b[j2] * w_n^(i1*j2), advanced by a running product per thread
        uint64_t f = gl::mul(__ldg(inter_b + (size_t)coset * ((size_t)1 << log2) + j2), root_pow(wlo, whi, lo_bits, i1 * j2));
        const uint64_t d = root_pow(wlo, whi, lo_bits, step_i1 * j2);
        for (; c < nchunks; c += cstep, i1 += step_i1) {
            const uint64_t v = gl::mul_canon(a[tile_off<T, RS>((int)i1, r1) + t], f);
            o[((size_t)c << log2) * T + (size_t)j2 * T + ii] = v;
            f = gl::mul_any(f, d);
        }
    }
}

// Fills the full inter-pass table of one plan: F[coset][(c << log2)*T + j2*T + ii] for i1 = c*T + ii.
__global__ void inter_table_kernel(uint64_t *__restrict__ out, const uint64_t *__restrict__ inter_b,
                                   const uint64_t *__restrict__ wlo, const uint64_t *__restrict__ whi, int lo_bits,
                                   int log1, int log2, int logT) {
    const size_t n = (size_t)1 << (log1 + log2);
    const int coset = blockIdx.y;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t ii = idx & ((1u << logT) - 1);
        const uint32_t j2 = (idx >> logT) & ((1u << log2) - 1);
        const uint32_t c = idx >> (logT + log2);
        const uint32_t i1 = (c << logT) | ii;
        out[(size_t)coset * n + idx] =
            gl::mul(__ldg(inter_b + ((size_t)coset << log2) + j2), root_pow(wlo, whi, lo_bits, i1 * j2));
    }
}

// ---- pass 2 -------------------------------------------------------------------------------
// grid: (n1/T, ncosets, ncols)
// BULK: the tile is one contiguous block of tmp (T * n2 words): a single 1-D bulk copy, another for the
// twiddles, one mbarrier.  In a three-pass plan (log0 > 0) the rows of the tile are n0 tile columns apart:
// `tmap` describes tmp as {T, n0, n2, blocks} words and the tile arrives in boxes {T, 1, min(n2, 256), 1}.
template <int T, bool INV, int MAXT, bool BULK = false>
__global__ void __launch_bounds__(MAXT) dft_pass2_kernel(const uint64_t *__restrict__ tmp, uint64_t *__restrict__ dst,
                                                        const uint64_t *__restrict__ stage2,
                                                        const uint64_t *__restrict__ post_u,
                                                        const uint64_t *__restrict__ post_v, int log1, int log2,
                                                        size_t dst_col_stride, int ncosets, int deint, int log0,
                                                        const uint64_t *__restrict__ post_j,
                                                        const __grid_constant__ CUtensorMap tmap) {
    // log0 > 0 (three-pass plan): blockIdx.x = (tile of i1) * n0 + j0; the block transforms tile columns
    // q = j2*n0 + j0 of pass 1's output and writes block j0 (n1*n2 words) of the destination; deint == 0.
    constexpr int RS = T + 1;
    extern __shared__ __align__(128) uint64_t smem[];
    const int n2 = 1 << log2;
    const int logn = log1 + log2;          // of one destination block
    const size_t n = (size_t)1 << (logn + log0);
    const int r1 = NttRounds(log2).log(0);
    uint64_t *a = smem;
    uint64_t *tw = smem + tile_words<T, RS>(log2, r1);
    const int coset = blockIdx.y, col = blockIdx.z;
    const uint32_t tile = blockIdx.x >> log0, j0 = blockIdx.x & ((1u << log0) - 1);
    const uint32_t i1_0 = tile * T;
    const uint64_t *s = tmp + ((size_t)col * ncosets + coset) * n + ((size_t)tile << (log2 + log0)) * T + (size_t)j0 * T;
    if constexpr (BULK) {
        __shared__ __align__(8) uint64_t mbar;
        const unsigned mb = smem_u32(&mbar);
        if (threadIdx.x == 0) mbar_init(mb, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(mb, (unsigned)((n2 * T + n2) * 8));
            if (log0 == 0) {
                bulk_g2s(smem_u32(a), s, (unsigned)(n2 * T * 8), mb);
            } else {
                const int rows = n2 < 256 ? n2 : 256;
                const int blk = (int)(((size_t)col * ncosets + coset) * (((size_t)1 << log1) / T) + tile);
                for (int r = 0; r < n2; r += rows) tensor4d_g2s(smem_u32(a + r * T), &tmap, 0, (int)j0, r, blk, mb);
            }
            bulk_g2s(smem_u32(tw), stage2, (unsigned)(n2 * 8), mb);
        }
        mbar_wait(mb, 0);
    } else {
        for (int i = threadIdx.x; i < n2; i += blockDim.x) cp_async8(tw + i, stage2 + i);
        load_tile_bitrev<T, RS>(a, s, log2, (T == 8 ? 3 : 2) + log0, r1);   // row j2 at s[(j2*n0)*T + t]
        cp_async_wait_all();
        __syncthreads();
    }
    dit_tile<T, RS, true, INV, (MAXT <= 256 ? 5 : 4), true, BULK>(a, tw, log2);  // tmp is canonical (pass 1's mul_canon)
    uint64_t *o = dst + (size_t)col * dst_col_stride + (size_t)coset * n + ((size_t)j0 << logn);
    if (post_j) post_v = post_j + ((size_t)j0 << log2);   // w_(n2 n0)^(i2 j0); post_u stays null
    // thread (r0 = tid / T, t = tid % T) stores rows i2 = r0 + k*J of its tile column i1 = i1_0 + t
    const int J = blockDim.x / T;
    const int t = threadIdx.x % T, r0 = threadIdx.x / T;
    const uint32_t i1 = i1_0 + t;
    const int K = n2 / J;
    if ((J & ((1 << r1) - 1)) != 0) {  // small transforms: rows of a thread inside one padding group
        for (int k = 0; k < K; k++) {
            const int i2 = r0 + k * J;
            uint64_t v = a[tile_off<T, RS>(i2, r1) + t];
            if (post_u) v = gl::mul(v, gl::mul(__ldg(post_u + i1), __ldg(post_v + i2)));
            else if (post_v) v = gl::mul(v, __ldg(post_v + i2));
            else v = gl::canon_any(v);
            o[out_index(i1 + ((uint32_t)i2 << log1), logn, deint)] = v;
        }
        return;
    }
    const uint64_t *ap = a + tile_off<T, RS>(r0, r1) + t;
    const int astep = J * RS + tile_pad<T>() * (J >> r1);
    if (deint == 0) {
        uint64_t *op = o + i1 + ((size_t)r0 << log1);
        const size_t ostep = (size_t)J << log1;
        if (post_u) {
            const uint64_t pu = __ldg(post_u + i1);
            const uint64_t *pv = post_v + r0;
#pragma unroll 4
            for (int k = 0; k < K; k++)
                op[k * ostep] = gl::mul(ap[k * astep], gl::mul(pu, __ldg(pv + k * J)));
        } else if (post_v) {
            const uint64_t *pv = post_v + r0;
#pragma unroll 4
            for (int k = 0; k < K; k++) op[k * ostep] = gl::mul(ap[k * astep], __ldg(pv + k * J));
        } else {
#pragma unroll 4
            for (int k = 0; k < K; k++) op[k * ostep] = gl::canon_any(ap[k * astep]);
        }
    } else {
        for (int k = 0; k < K; k++) {
            const int i2 = r0 + k * J;
            uint64_t v = ap[k * astep];
            if (post_u) v = gl::mul(v, gl::mul(__ldg(post_u + i1), __ldg(post_v + i2)));
            else if (post_v) v = gl::mul(v, __ldg(post_v + i2));
            else v = gl::canon_any(v);
            const uint32_t i = i1 + ((uint32_t)i2 << log1);
            o[out_index(i, logn, deint)] = v;
        }
    }
}

// ---- pass 3 (three-pass plans) --------------------------------------------------------------
// In place: for every position i < m = n1*n2 of a (column, coset) block of n = m * 2^R words, the 2^R-point
// DFT over the values at i + j0*m (canonical, written by pass 2), output i0 stored at i + i0*m.
// grid: (m / 256, ncosets, ncols)
template <int R, bool INV>
__global__ void __launch_bounds__(256) dft_pass3_kernel(uint64_t *__restrict__ dst, int logm, size_t dst_col_stride) {
    uint64_t *o = dst + (size_t)blockIdx.z * dst_col_stride + ((size_t)blockIdx.y << (logm + R)) +
                  (size_t)blockIdx.x * 256 + threadIdx.x;
    uint64_t x[1 << R];
    static_for<(1 << R)>([&](auto E) {
        constexpr int e = decltype(E)::value;
        x[e] = o[(size_t)dft_bitrev(e, R) << logm];
    });
    dft_network<R, INV>(x);
    static_for<(1 << R)>([&](auto E) {
        constexpr int e = decltype(E)::value;
        o[(size_t)e << logm] = gl::canon_any(x[e]);
    });
}

// ---- single pass (n <= 2^11) --------------------------------------------------------------
// grid: (1, ncosets, ncols)
template <bool PLAIN, bool INV>
__global__ void __launch_bounds__(256) dft_single_kernel(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst,
                                                         const uint64_t *__restrict__ stage,
                                                         const uint64_t *__restrict__ post_u, uint64_t scale, int logn,
                                                         size_t src_col_stride, size_t dst_col_stride, int deint) {
    extern __shared__ __align__(128) uint64_t smem[];
    const int n = 1 << logn;
    const int r1 = NttRounds(logn).log(0);
    uint64_t *a = smem;
    uint64_t *tw = smem + tile_words<1, 1>(logn, r1);
    const int coset = blockIdx.y, col = blockIdx.z;
    const uint64_t *s = src + (size_t)col * src_col_stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        cp_async8(tw + i, stage + (size_t)coset * n + i);
        cp_async8(a + tile_off<1, 1>((int)bitrev(i, logn), r1), s + i);
    }
    cp_async_wait_all();
    __syncthreads();
    dit_tile<1, 1, PLAIN, INV, 5>(a, tw, logn);
    uint64_t *o = dst + (size_t)col * dst_col_stride + (size_t)coset * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        uint64_t v = a[tile_off<1, 1>(i, r1)];
        if (post_u) v = gl::mul(v, __ldg(post_u + i));
        else if (scale != 1) v = gl::mul(v, scale);
        else v = gl::canon_any(v);
        o[out_index(i, logn, deint)] = v;
    }
}

int dft_tile_width(int log1, int log2) {
    // tile width T: 2^10-point passes with T = 8 and 2^11-point passes with T = 4 both leave room for
    // two blocks per SM (82 / 96 KB each); measured 4.2e11 butterflies/s either way, against 3.0e11
    // with one 512-thread block per SM
    static const int forced = [] {  // experiment hook: AERO_NTT_TILE=4|8 forces the width for passes <= 2^10 points
        const char *e = getenv("AERO_NTT_TILE");
        return e ? atoi(e) : 0;
    }();
    const int big = log1 > log2 ? log1 : log2;
    if (big <= 10 && (forced == 4 || forced == 8)) return forced;
    return big <= 10 ? 8 : 4;
}

void dft_fill_inter_table(const DftTables &t, uint64_t *out, cudaStream_t s) {
    const int T = dft_tile_width(t.log1, t.log2);
    const size_t n = (size_t)1 << t.logn;
    unsigned bx = (unsigned)((n + 255) / 256);
    if (bx > 148 * 16) bx = 148 * 16;
    AERO_COUNT_LAUNCH(1);
    inter_table_kernel<<<dim3(bx, t.ncosets), 256, 0, s>>>(out, t.inter_b, t.wlo, t.whi, t.lo_bits, t.log1, t.log2 + t.log0,
                                                          T == 8 ? 3 : 2);
}

// ---- tensor maps for the BULK pass 1 ------------------------------------------------------------------
// cuTensorMapEncodeTiled is host-side arithmetic on the descriptor; it is resolved through the runtime so
// that the library does not link against libcuda.
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                           const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeTiledFn tensor_map_encoder() {
    static const TensorMapEncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (TensorMapEncodeTiledFn)p;
    }();
    return fn;
}
// experiment hooks: AERO_NTT_BULK=0 keeps the LDGSTS tile copies, AERO_NTT_COLFAST=0 the (tile, coset, column) grid
static int env_flag(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
static bool ntt_use_bulk() {
    static const bool on = env_flag("AERO_NTT_BULK", 1) != 0;
    return on;
}
static bool ntt_colfast() {
    static const bool on = env_flag("AERO_NTT_COLFAST", 1) != 0;
    return on;
}
// {n2, n1, ncols} words with strides {8, 8*n2, 8*col_stride} bytes; box {T, min(n1, 256), 1}
static bool encode_pass1_map(CUtensorMap *m, const uint64_t *src, int log1, int log2, int ncols, size_t col_stride, int T) {
    const TensorMapEncodeTiledFn fn = tensor_map_encoder();
    if (!fn || ((uintptr_t)src & 15)) return false;
    const cuuint64_t n1 = 1ULL << log1, n2 = 1ULL << log2;
    if (ncols > 1 && ((col_stride & 1) || col_stride < n1 * n2)) return false;
    const cuuint64_t dims[3] = {n2, n1, (cuuint64_t)ncols};
    const cuuint64_t strides[2] = {n2 * 8, (ncols > 1 ? (cuuint64_t)col_stride : n1 * n2) * 8};
    const cuuint32_t box[3] = {(cuuint32_t)T, (cuuint32_t)(n1 < 256 ? n1 : 256), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void *)src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int T, bool PLAIN, bool INV, bool TAB, int MAXT, bool BULK>
static void launch_pass1_impl(const DftTables &t, const DftLaunch &l, int nc, int threads, size_t smem, const CUtensorMap &map,
                              cudaStream_t s) {
    static DeviceOnce once;  // function attributes are per device
    once.run([] {
        cudaFuncSetAttribute(dft_pass1_kernel<T, PLAIN, INV, TAB, MAXT, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(dft_pass1_kernel<T, PLAIN, INV, TAB, MAXT, BULK>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    });
    const int logq = t.log2 + t.log0;   // pass 1 sees n2 * n0 tile columns
    const int n1 = 1 << t.log1, nq = 1 << logq;
    const size_t n = (size_t)1 << t.logn;
    // column-fastest block order needs (columns x cosets) in grid.x and the tiles in grid.y (<= 65535)
    const bool colfast = ntt_colfast() && nq / T <= 65535;
    const dim3 g1 = colfast ? dim3((unsigned)l.ncols * nc, nq / T, 1) : dim3(nq / T, nc, l.ncols);
    dft_pass1_kernel<T, PLAIN, INV, TAB, MAXT, BULK><<<g1, threads, smem, s>>>(
        l.src, l.tmp, t.stage1 + (size_t)l.coset_begin * n1, t.inter_b + (size_t)l.coset_begin * nq,
        TAB ? t.inter_full + (size_t)l.coset_begin * n : nullptr, t.wlo, t.whi, t.lo_bits, t.log1, logq,
        l.src_col_stride, nc, l.ncols, colfast ? 1 : 0, map);
}
template <int T, bool PLAIN, bool INV, bool TAB, int MAXT>
static void launch_pass1(const DftTables &t, const DftLaunch &l, int nc, int threads, size_t smem, cudaStream_t s) {
    CUtensorMap map;
    memset(&map, 0, sizeof map);
    if constexpr (MAXT == 256 && T == 8) {
        // one work item per thread in the first (32-point) round is what the dense tile needs
        const int items = ((1 << t.log1) >> 5) * T;
        if (ntt_use_bulk() && NttRounds(t.log1).log(0) == 5 && items == threads &&
            encode_pass1_map(&map, l.src, t.log1, t.log2 + t.log0, l.ncols, l.src_col_stride, T)) {
            launch_pass1_impl<T, PLAIN, INV, TAB, MAXT, true>(t, l, nc, threads, smem, map, s);
            return;
        }
    }
    launch_pass1_impl<T, PLAIN, INV, TAB, MAXT, false>(t, l, nc, threads, smem, map, s);
}
// pass 2 of a three-pass plan: tmp as {T, n0, n2, ncols * nc * n1 / T} words; box {T, 1, min(n2, 256), 1}
static bool encode_pass2_map(CUtensorMap *m, const uint64_t *tmp, int log1, int log2, int log0, int T, size_t nblocks) {
    const TensorMapEncodeTiledFn fn = tensor_map_encoder();
    if (!fn || ((uintptr_t)tmp & 15) || nblocks > 0x7fffffffULL) return false;
    const cuuint64_t n2 = 1ULL << log2, n0 = 1ULL << log0;
    (void)log1;
    const cuuint64_t dims[4] = {(cuuint64_t)T, n0, n2, (cuuint64_t)nblocks};
    const cuuint64_t strides[3] = {(cuuint64_t)T * 8, n0 * T * 8, n2 * n0 * T * 8};
    const cuuint32_t box[4] = {(cuuint32_t)T, 1, (cuuint32_t)(n2 < 256 ? n2 : 256), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, (void *)tmp, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int T, bool INV, int MAXT, bool BULK>
static void launch_pass2_impl(const DftTables &t, const DftLaunch &l, int nc, int threads, size_t smem, const CUtensorMap &map,
                              cudaStream_t s) {
    static DeviceOnce once;
    once.run([] {
        cudaFuncSetAttribute(dft_pass2_kernel<T, INV, MAXT, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(dft_pass2_kernel<T, INV, MAXT, BULK>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    });
    const int n1 = 1 << t.log1;
    dim3 g2((n1 / T) << t.log0, nc, l.ncols);
    dft_pass2_kernel<T, INV, MAXT, BULK><<<g2, threads, smem, s>>>(l.tmp, l.dst, t.stage2, t.post_u, t.post_v, t.log1, t.log2,
                                                             l.dst_col_stride, nc, l.deinterleave_log, t.log0, t.post_j, map);
}
template <int T, bool INV, int MAXT>
static void launch_pass2(const DftTables &t, const DftLaunch &l, int nc, int threads, size_t smem, cudaStream_t s) {
    CUtensorMap map;
    memset(&map, 0, sizeof map);
    if constexpr (MAXT == 256 && T == 8) {
        const int items = ((1 << t.log2) >> 5) * T;
        // the tile is T * n2 contiguous words of tmp, 16-byte aligned whenever tmp is; three-pass plans go
        // through a tensor map
        if (ntt_use_bulk() && NttRounds(t.log2).log(0) == 5 && items == threads && ((uintptr_t)l.tmp & 15) == 0 &&
            (t.log0 == 0 || encode_pass2_map(&map, l.tmp, t.log1, t.log2, t.log0, T, (size_t)l.ncols * nc * ((1u << t.log1) / T)))) {
            launch_pass2_impl<T, INV, MAXT, true>(t, l, nc, threads, smem, map, s);
            return;
        }
    }
    launch_pass2_impl<T, INV, MAXT, false>(t, l, nc, threads, smem, map, s);
}

template <bool INV>
static void launch_pass3(const DftTables &t, const DftLaunch &l, int nc, cudaStream_t s) {
    const int logm = t.log1 + t.log2;
    const dim3 g(1u << (logm - 8), nc, l.ncols);
    switch (t.log0) {
    case 1: dft_pass3_kernel<1, INV><<<g, 256, 0, s>>>(l.dst, logm, l.dst_col_stride); break;
    case 2: dft_pass3_kernel<2, INV><<<g, 256, 0, s>>>(l.dst, logm, l.dst_col_stride); break;
    case 3: dft_pass3_kernel<3, INV><<<g, 256, 0, s>>>(l.dst, logm, l.dst_col_stride); break;
    default: dft_pass3_kernel<4, INV><<<g, 256, 0, s>>>(l.dst, logm, l.dst_col_stride); break;
    }
}

template <int T>
static void launch_two_pass(const DftTables &t, const DftLaunch &l, cudaStream_t s) {
    const int n1 = 1 << t.log1, n2 = 1 << t.log2;
    const int nc = l.coset_count ? l.coset_count : t.ncosets;
    const size_t smem1 = (size_t)(tile_words<T, T + 1>(t.log1, NttRounds(t.log1).log(0)) + n1) * 8;
    const size_t smem2 = (size_t)(tile_words<T, T + 1>(t.log2, NttRounds(t.log2).log(0)) + n2) * 8;
    // 512 threads while two blocks fit the 227 KB of shared memory of an SM; a 2^12-point pass
    // (192 KB, one block per SM) runs 1024 threads instead so the SM still holds 32 warps
    // a schedule with a 32-point round needs the 256-thread (128-register) instantiation
    auto wide = [](int logm) { return NttRounds(logm).log(0) >= 5; };
    auto threads_for = [&](int logm, size_t smem) {
        const int cap = wide(logm) ? 256 : (smem > 113 * 1024 ? 1024 : 512);
        int items = ((1 << logm) >> NttRounds(logm).log(0)) * T;  // work items of the widest round
        int th = items < 64 ? 64 : (items > cap ? cap : items);
        th = (th / (T * T)) * (T * T);
        return th < T * T ? T * T : th;
    };
    int th1 = threads_for(t.log1, smem1), th2 = threads_for(t.log2, smem2);
    // the tile copies give every thread whole rows: at most one thread row per tile row
    while (th1 / T > n1 && th1 > T * T) th1 /= 2;
    while (th2 / T > n2 && th2 > T * T) th2 /= 2;
    AERO_COUNT_LAUNCH(2);
    const bool tab = t.inter_full != nullptr;
    const bool w1 = wide(t.log1), w2 = wide(t.log2);
#define AERO_P1M(PLAIN, INV, TAB)                                                    \
    (w1 ? launch_pass1<T, PLAIN, INV, TAB, 256>(t, l, nc, th1, smem1, s)             \
        : launch_pass1<T, PLAIN, INV, TAB, 1024>(t, l, nc, th1, smem1, s))
#define AERO_P1(PLAIN, INV) (tab ? AERO_P1M(PLAIN, INV, true) : AERO_P1M(PLAIN, INV, false))
    if (t.plain) {
        if (t.inverse) AERO_P1(true, true); else AERO_P1(true, false);
    } else {
        if (t.inverse) AERO_P1(false, true); else AERO_P1(false, false);
    }
#undef AERO_P1
#undef AERO_P1M
#define AERO_P2(INV)                                                                 \
    (w2 ? launch_pass2<T, INV, 256>(t, l, nc, th2, smem2, s) : launch_pass2<T, INV, 1024>(t, l, nc, th2, smem2, s))
    if (t.inverse) AERO_P2(true); else AERO_P2(false);
#undef AERO_P2
    if (t.log0 > 0) {
        AERO_COUNT_LAUNCH(1);
        if (t.inverse) launch_pass3<true>(t, l, nc, s); else launch_pass3<false>(t, l, nc, s);
    }
}

template <bool PLAIN, bool INV>
static void launch_single(const DftTables &t, const DftLaunch &l, cudaStream_t s) {
    const int n = 1 << t.logn;
    const int items = n >> NttRounds(t.logn).log(0);
    const int th = items < 32 ? 32 : (items > 256 ? 256 : items);
    const int nc = l.coset_count ? l.coset_count : t.ncosets;
    dim3 g(1, nc, l.ncols);
    const size_t smem = (size_t)(tile_words<1, 1>(t.logn, NttRounds(t.logn).log(0)) + n) * 8;
    dft_single_kernel<PLAIN, INV><<<g, th, smem, s>>>(l.src, l.dst, t.stage2 + (size_t)l.coset_begin * n, t.post_u,
                                                               t.single_scale, t.logn, l.src_col_stride,
                                                               l.dst_col_stride, l.deinterleave_log);
}

void dft_run(const DftTables &t, const DftLaunch &l, cudaStream_t s) {
    if (t.log1 == 0) {
        AERO_COUNT_LAUNCH(1);
        if (t.plain) {
            if (t.inverse) launch_single<true, true>(t, l, s); else launch_single<true, false>(t, l, s);
        } else {
            if (t.inverse) launch_single<false, true>(t, l, s); else launch_single<false, false>(t, l, s);
        }
        return;
    }
    if (dft_tile_width(t.log1, t.log2) == 8) launch_two_pass<8>(t, l, s);
    else launch_two_pass<4>(t, l, s);
}

}  // namespace aero
