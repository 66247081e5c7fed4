// BLAKE2s-256 (RFC 7693; unkeyed, 32-byte digest) for sm_100a, specialised for the fork-specific
// element layout of Blake2s_256::hash_elements (winterfell/crypto/src/hash/blake2s/mod.rs:52-77):
// every field element is 8 canonical LE bytes followed by 24 zero bytes, so each 64-byte block
// carries two elements in message words {0,1} and {8,9}; the other twelve words are zero and their
// additions are folded away at compile time (all ten rounds are fully unrolled with a constexpr
// sigma).  32-bit ALU bound: 12 INT ops per G, 80 G per compression.
#pragma once
#include <cstdint>

namespace b2s {

#if defined(__CUDACC__)
#define B2S_HD __host__ __device__ __forceinline__
#else
#define B2S_HD inline
#endif

constexpr uint32_t IV0 = 0x6A09E667u, IV1 = 0xBB67AE85u, IV2 = 0x3C6EF372u, IV3 = 0xA54FF53Au,
                   IV4 = 0x510E527Fu, IV5 = 0x9B05688Cu, IV6 = 0x1F83D9ABu, IV7 = 0x5BE0CD19u;
constexpr uint32_t H0_INIT = IV0 ^ 0x01010020u;  // digest_length=32, key_length=0, fanout=depth=1

B2S_HD uint32_t rotr16(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x1032);
#else
    return (x >> 16) | (x << 16);
#endif
}
B2S_HD uint32_t rotr8(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0321);
#else
    return (x >> 8) | (x << 24);
#endif
}
B2S_HD uint32_t rotr12(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, 12);
#else
    return (x >> 12) | (x << 20);
#endif
}
B2S_HD uint32_t rotr7(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, 7);
#else
    return (x >> 7) | (x << 25);
#endif
}

// Message-word additions go to the FMA pipe: the kernels built on this are bound by the ALU pipe
// (LOP3/SHF/PRMT: 8 per G), and ptxas otherwise folds `a + b + m` into one three-input IADD3 -- an
// ALU-pipe instruction -- while the FMA pipe idles.  `m * one + a` with `one` read from constant
// memory (opaque to ptxas) is a plain IMAD; a compile-time-zero word still costs nothing.
#if defined(__CUDACC__)
static __constant__ uint32_t k_one = 1u;
#endif
#if defined(__CUDA_ARCH__)
template <bool HAS>
__device__ __forceinline__ uint32_t add_msg(uint32_t a, uint32_t m) {
    if (HAS) asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a) : "r"(m), "r"(k_one));
    return a;
}
#else
template <bool HAS>
inline uint32_t add_msg(uint32_t a, uint32_t m) { return HAS ? a + m : a; }
#endif

// HX / HY: whether message words x / y can be non-zero (false: known-zero padding words)
template <bool HX, bool HY>
B2S_HD void G(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d, uint32_t x, uint32_t y) {
    a = add_msg<HX>(a + b, x);
    d = rotr16(d ^ a);
    c = c + d;
    b = rotr12(b ^ c);
    a = add_msg<HY>(a + b, y);
    d = rotr8(d ^ a);
    c = c + d;
    b = rotr7(b ^ c);
}

// One round with a compile-time message schedule row (s0..s15).  MASK bit i set = word i may be non-zero.
#define B2S_NZ(s) (((MASK) >> (s)) & 1u)
#define B2S_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    G<B2S_NZ(s0), B2S_NZ(s1)>(v0, v4, v8, v12, m[s0], m[s1]);                           \
    G<B2S_NZ(s2), B2S_NZ(s3)>(v1, v5, v9, v13, m[s2], m[s3]);                           \
    G<B2S_NZ(s4), B2S_NZ(s5)>(v2, v6, v10, v14, m[s4], m[s5]);                          \
    G<B2S_NZ(s6), B2S_NZ(s7)>(v3, v7, v11, v15, m[s6], m[s7]);                          \
    G<B2S_NZ(s8), B2S_NZ(s9)>(v0, v5, v10, v15, m[s8], m[s9]);                          \
    G<B2S_NZ(s10), B2S_NZ(s11)>(v1, v6, v11, v12, m[s10], m[s11]);                      \
    G<B2S_NZ(s12), B2S_NZ(s13)>(v2, v7, v8, v13, m[s12], m[s13]);                       \
    G<B2S_NZ(s14), B2S_NZ(s15)>(v3, v4, v9, v14, m[s14], m[s15]);

// Compression F(h, m, t, last).  Words outside MASK must be zero (they are not even read).
template <uint32_t MASK = 0xFFFFu>
B2S_HD void compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, bool last) {
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = IV0, v9 = IV1, v10 = IV2, v11 = IV3;
    uint32_t v12 = IV4 ^ t0, v13 = IV5, v14 = last ? ~IV6 : IV6, v15 = IV7;
    B2S_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    h[0] ^= v0 ^ v8;
    h[1] ^= v1 ^ v9;
    h[2] ^= v2 ^ v10;
    h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12;
    h[5] ^= v5 ^ v13;
    h[6] ^= v6 ^ v14;
    h[7] ^= v7 ^ v15;
}

// Note (measured, round 1): the row-hash kernel runs at 96-97 % of the ALU pipe (LOP3/SHF/PRMT:
// 640 per compression is the floor).  Moving rotations to the FMA pipe as IMAD.WIDE (x * 2^(32-n),
// lo + hi) balanced the static instruction mix (569 ALU / 538 FMA) but ran 1.5 % SLOWER on B200:
// tools/int_peak.cu shows IMAD.WIDE / IMAD.HI issue at ~0.4x the IMAD rate AND block the ALU issue
// port meanwhile, so only plain IMAD (the additions) is worth moving.

B2S_HD void init(uint32_t h[8]) {
    h[0] = H0_INIT; h[1] = IV1; h[2] = IV2; h[3] = IV3;
    h[4] = IV4; h[5] = IV5; h[6] = IV6; h[7] = IV7;
}

// One 64-byte block holding elements e0 (words 0,1) and e1 (words 8,9), zero elsewhere.
B2S_HD void compress_pair(uint32_t h[8], uint64_t e0, uint64_t e1, uint32_t t0, bool last) {
    const uint32_t m[16] = {(uint32_t)e0, (uint32_t)(e0 >> 32), 0, 0, 0, 0, 0, 0,
                            (uint32_t)e1, (uint32_t)(e1 >> 32), 0, 0, 0, 0, 0, 0};
    compress<0x0303u>(h, m, t0, last);
}

// merge(a, b) = BLAKE2s(a || b): one final 64-byte block (blake2s/mod.rs:37-39).
B2S_HD void merge(const uint32_t a[8], const uint32_t b[8], uint32_t out[8]) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = a[i]; m[8 + i] = b[i]; }
    init(out);
    compress(out, m, 64u, true);
}

// merge_with_int(seed, v) = BLAKE2s(seed || LE64(v)): one 40-byte final block (blake2s/mod.rs:41-46).
B2S_HD void merge_with_int(const uint32_t seed[8], uint64_t v, uint32_t out[8]) {
    uint32_t m[16] = {seed[0], seed[1], seed[2], seed[3], seed[4], seed[5], seed[6], seed[7],
                      (uint32_t)v, (uint32_t)(v >> 32), 0, 0, 0, 0, 0, 0};
    init(out);
    compress<0x03FFu>(out, m, 40u, true);
}

}  // namespace b2s
