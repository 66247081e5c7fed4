// Goldilocks field arithmetic for sm_100a kernels and for the host-side driver.
//
// p = 2^64 - 2^32 + 1.  The reference (winterfell/math/src/field/f64/mod.rs:37-61) keeps elements
// in Montgomery form x*2^64 mod p; on the device we keep CANONICAL values in [0, p) because the
// special form of p makes the plain 128-bit product cheap to reduce (2^64 = 2^32-1, 2^96 = -1
// mod p) and because hashing (blake2s/mod.rs:52-77) needs canonical bytes anyway.  Conversion
// happens once at the C-ABI boundary (mont_to_canon / canon_to_mont below, replacing
// f64/mod.rs:59-61 `new` and :234 `as_int`).  All results are exact, so they equal the
// reference's canonical values bit for bit.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

namespace gl {

constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t EPS = 0xFFFFFFFFULL;            // 2^64 mod p
constexpr uint64_t GENERATOR = 7;                  // f64/mod.rs:218 (also the LDE/FRI coset offset)
constexpr uint64_t TWO_ADIC_ROOT = 1753635133440165772ULL;  // f64/mod.rs:43, order 2^32
constexpr uint64_t MONT_R_INV = 18446744065119617025ULL;    // 2^-64 mod p
// (2^-64 = 2^128 since 2^192 = 1;  2^128 = (2^96)*(2^32) = -2^32 = p - 2^32)

// ---- device fast paths (32-bit carry chains in PTX) -------------------------------------------
// "any" = a u64 congruent to the value mod p but not necessarily < p; "canonical" = in [0, p).
// The compiler's 64-bit compare/select sequences cost ~2x the INT-pipe work of these chains, and
// the NTT kernels are INT-pipe bound (profiles/r01_*): ptxas turns the mad.cc chain into
// IMAD.WIDE.U32 with predicate carry-out, the rest into IADD3/IADD3.X.
#if defined(__CUDACC__)
// any x any -> any.  {hh:hl:lo} = a*b ;  lo - hh (borrow: -EPS) + hl*EPS (carry: +EPS)
__device__ __forceinline__ uint64_t mul_any(uint64_t a, uint64_t b) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    uint32_t x0, x1;
    asm("{\n\t"
        ".reg .u32 r0, r1, r2, r3, m, c, t0, t1;\n\t"
        "mul.lo.u32 r0, %2, %4;\n\t"
        "mul.hi.u32 r1, %2, %4;\n\t"
        "mad.lo.cc.u32 r1, %2, %5, r1;\n\t"
        "madc.hi.u32 r2, %2, %5, 0;\n\t"
        "mad.lo.cc.u32 r1, %3, %4, r1;\n\t"
        "madc.hi.cc.u32 r2, %3, %4, r2;\n\t"
        "addc.u32 r3, 0, 0;\n\t"
        "mad.lo.cc.u32 r2, %3, %5, r2;\n\t"
        "madc.hi.u32 r3, %3, %5, r3;\n\t"
        "sub.cc.u32 %0, r0, r3;\n\t"
        "subc.cc.u32 %1, r1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"          // borrow ? 0xffffffff (= EPS) : 0
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t"
        "mul.lo.u32 t0, r2, 0xffffffff;\n\t"
        "mul.hi.u32 t1, r2, 0xffffffff;\n\t"
        "add.cc.u32 %0, %0, t0;\n\t"
        "addc.cc.u32 %1, %1, t1;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "neg.s32 c, c;\n\t"              // carry ? EPS : 0
        "add.cc.u32 %0, %0, c;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(x0), "=&r"(x1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((uint64_t)x1 << 32) | x0;
}
// any x any -> canonical.  Same product and first fix-up as mul_any; the last fix-up and the
// canonicalisation merge: with r = x + hl*EPS (carry c) and s = r + EPS (carry c2), the result is s
// when c | c2 (c: r wrapped, and r + EPS < p; c2: r >= p), else r.  Three instructions fewer than
// canon_any(mul_any()).
__device__ __forceinline__ uint64_t mul_canon(uint64_t a, uint64_t b) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    uint32_t x0, x1;
    asm("{\n\t"
        ".reg .u32 r0, r1, r2, r3, m, c, t0, t1, s0, s1;\n\t"
        ".reg .pred q;\n\t"
        "mul.lo.u32 r0, %2, %4;\n\t"
        "mul.hi.u32 r1, %2, %4;\n\t"
        "mad.lo.cc.u32 r1, %2, %5, r1;\n\t"
        "madc.hi.u32 r2, %2, %5, 0;\n\t"
        "mad.lo.cc.u32 r1, %3, %4, r1;\n\t"
        "madc.hi.cc.u32 r2, %3, %4, r2;\n\t"
        "addc.u32 r3, 0, 0;\n\t"
        "mad.lo.cc.u32 r2, %3, %5, r2;\n\t"
        "madc.hi.u32 r3, %3, %5, r3;\n\t"
        "sub.cc.u32 %0, r0, r3;\n\t"
        "subc.cc.u32 %1, r1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t"
        "mul.lo.u32 t0, r2, 0xffffffff;\n\t"
        "mul.hi.u32 t1, r2, 0xffffffff;\n\t"
        "add.cc.u32 %0, %0, t0;\n\t"
        "addc.cc.u32 %1, %1, t1;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "add.cc.u32 s0, %0, 0xffffffff;\n\t"
        "addc.cc.u32 s1, %1, 0;\n\t"
        "addc.u32 c, c, 0;\n\t"
        "setp.ne.u32 q, c, 0;\n\t"
        "selp.u32 %0, s0, %0, q;\n\t"
        "selp.u32 %1, s1, %1, q;\n\t"
        "}"
        : "=&r"(x0), "=&r"(x1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((uint64_t)x1 << 32) | x0;
}
// any -> canonical:  x >= p  <=>  x + EPS carries out of 64 bits (p + EPS = 2^64)
__device__ __forceinline__ uint64_t canon_any(uint64_t x) {
    uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32), y0, y1;
    asm("{\n\t"
        ".reg .u32 t0, t1, c;\n\t"
        ".reg .pred q;\n\t"
        "add.cc.u32 t0, %2, 0xffffffff;\n\t"
        "addc.cc.u32 t1, %3, 0;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "setp.ne.u32 q, c, 0;\n\t"
        "selp.u32 %0, t0, %2, q;\n\t"
        "selp.u32 %1, t1, %3, q;\n\t"
        "}"
        : "=r"(y0), "=r"(y1)
        : "r"(x0), "r"(x1));
    return ((uint64_t)y1 << 32) | y0;
}
// u any, v canonical -> any.  u + v < 2^64 + p, so one +EPS fix-up cannot carry again.
__device__ __forceinline__ uint64_t add_ac(uint64_t u, uint64_t v) {
    uint32_t u0 = (uint32_t)u, u1 = (uint32_t)(u >> 32), v0 = (uint32_t)v, v1 = (uint32_t)(v >> 32), y0, y1;
    asm("{\n\t"
        ".reg .u32 c;\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.cc.u32 %1, %3, %5;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "neg.s32 c, c;\n\t"
        "add.cc.u32 %0, %0, c;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(y0), "=&r"(y1)
        : "r"(u0), "r"(u1), "r"(v0), "r"(v1));
    return ((uint64_t)y1 << 32) | y0;
}
// u any, v canonical -> any (canonical when u is).  u - v > -p, so one -EPS fix-up cannot borrow again.
__device__ __forceinline__ uint64_t sub_ac(uint64_t u, uint64_t v) {
    uint32_t u0 = (uint32_t)u, u1 = (uint32_t)(u >> 32), v0 = (uint32_t)v, v1 = (uint32_t)(v >> 32), y0, y1;
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"
        "subc.cc.u32 %1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(y0), "=&r"(y1)
        : "r"(u0), "r"(u1), "r"(v0), "r"(v1));
    return ((uint64_t)y1 << 32) | y0;
}
// u any, v canonical -> canonical.  Like add_ac, but the +EPS fix-up and the final reduction merge
// (the same select as in mul_canon): r = u + v (carry c), s = r + EPS (carry c2); c | c2 ? s : r.
__device__ __forceinline__ uint64_t add_cc(uint64_t u, uint64_t v) {
    uint32_t u0 = (uint32_t)u, u1 = (uint32_t)(u >> 32), v0 = (uint32_t)v, v1 = (uint32_t)(v >> 32), y0, y1;
    asm("{\n\t"
        ".reg .u32 c, s0, s1;\n\t"
        ".reg .pred q;\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.cc.u32 %1, %3, %5;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "add.cc.u32 s0, %0, 0xffffffff;\n\t"
        "addc.cc.u32 s1, %1, 0;\n\t"
        "addc.u32 c, c, 0;\n\t"
        "setp.ne.u32 q, c, 0;\n\t"
        "selp.u32 %0, s0, %0, q;\n\t"
        "selp.u32 %1, s1, %1, q;\n\t"
        "}"
        : "=&r"(y0), "=&r"(y1)
        : "r"(u0), "r"(u1), "r"(v0), "r"(v1));
    return ((uint64_t)y1 << 32) | y0;
}

// x any -> x * 2^S canonical, for a compile-time 0 < S < 96 that is not a multiple of 32.
// 2 generates the 192 = 3*64 element subgroup of the field (2^96 = -1), so every 64th root of unity
// is a power of two (w_64 = 2^39 for the reference's TWO_ADIC_ROOT): the twiddles INSIDE a radix-4 /
// 8 / 16 butterfly are shifts, which cost three funnel shifts and a short carry chain instead of the
// four IMAD.WIDE of a general product.  With t = 2^32 (t^2 = t - 1, t^3 = -1 mod p), S = 32a + b and
// {y2:y1:y0} = x << b (y2 < 2^b):
//   a = 0:  {y1:y0} + y2*EPS                       (the tail of mul_canon)
//   a = 1:  y0*t + y1*(t-1) - y2 = ({y0:0} - y2) - (p - y1*EPS)     (two canonical subtractions)
//   a = 2:  y0*(t-1) - y1 - y2*t = y0*EPS - {y2:y1}                 (one canonical subtraction)
template <int S>
__device__ __forceinline__ uint64_t mul_pow2(uint64_t x) {
    static_assert(S > 0 && S < 96 && S % 32 != 0, "shift must be in (0,96) and not a multiple of 32");
    constexpr int a = S / 32, b = S % 32;
    const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32);
    const uint32_t y0 = x0 << b, y1 = __funnelshift_l(x0, x1, b), y2 = x1 >> (32 - b);
    uint32_t r0, r1;
    if (a == 0) {
        asm("{\n\t"
            ".reg .u32 c, t0, t1, s0, s1;\n\t"
            ".reg .pred q;\n\t"
            "mul.lo.u32 t0, %4, 0xffffffff;\n\t"
            "mul.hi.u32 t1, %4, 0xffffffff;\n\t"
            "add.cc.u32 %0, %2, t0;\n\t"
            "addc.cc.u32 %1, %3, t1;\n\t"
            "addc.u32 c, 0, 0;\n\t"
            "add.cc.u32 s0, %0, 0xffffffff;\n\t"
            "addc.cc.u32 s1, %1, 0;\n\t"
            "addc.u32 c, c, 0;\n\t"
            "setp.ne.u32 q, c, 0;\n\t"
            "selp.u32 %0, s0, %0, q;\n\t"
            "selp.u32 %1, s1, %1, q;\n\t"
            "}"
            : "=&r"(r0), "=&r"(r1)
            : "r"(y0), "r"(y1), "r"(y2));
    } else if (a == 1) {
        asm("{\n\t"
            ".reg .u32 m, w0, w1, n0, n1, ny;\n\t"
            "sub.cc.u32 w0, 0, %4;\n\t"       // W = {y0:0} - y2  (+p on borrow)
            "subc.cc.u32 w1, %2, 0;\n\t"
            "subc.u32 m, 0, 0;\n\t"
            "sub.cc.u32 w0, w0, m;\n\t"
            "subc.u32 w1, w1, 0;\n\t"
            "not.b32 ny, %3;\n\t"             // N = p - y1*EPS = {~y1 : y1 + 1}  (<= p)
            "add.cc.u32 n0, %3, 1;\n\t"
            "addc.u32 n1, ny, 0;\n\t"
            "sub.cc.u32 %0, w0, n0;\n\t"      // W - N  (+p on borrow)
            "subc.cc.u32 %1, w1, n1;\n\t"
            "subc.u32 m, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, m;\n\t"
            "subc.u32 %1, %1, 0;\n\t"
            "}"
            : "=&r"(r0), "=&r"(r1)
            : "r"(y0), "r"(y1), "r"(y2));
    } else {
        asm("{\n\t"
            ".reg .u32 m, t0, t1;\n\t"
            "mul.lo.u32 t0, %2, 0xffffffff;\n\t"  // y0*EPS <= (2^32-1)^2 < p
            "mul.hi.u32 t1, %2, 0xffffffff;\n\t"
            "sub.cc.u32 %0, t0, %3;\n\t"
            "subc.cc.u32 %1, t1, %4;\n\t"
            "subc.u32 m, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, m;\n\t"
            "subc.u32 %1, %1, 0;\n\t"
            "}"
            : "=&r"(r0), "=&r"(r1)
            : "r"(y0), "r"(y1), "r"(y2));
    }
    return ((uint64_t)r1 << 32) | r0;
}
// Lazily reduced sum of products: a 160-bit accumulator takes up to 2^32 full 64x64-bit products
// (13 instructions each: four wide multiply-adds and the carry tail) and is reduced ONCE, instead of a
// 25-instruction mul_canon plus a 9-instruction modular add per term.  Used by the dot-product
// shaped kernels (DEEP accumulation, OOD chunk evaluation).
struct Acc160 {
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
    __device__ __forceinline__ void mac(uint64_t a, uint64_t b) {
        const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
            "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %8, %2;\n\t"
            "madc.hi.cc.u32 %3, %6, %8, %3;\n\t"
            "addc.u32 %4, %4, 0;\n\t"
            "mad.lo.cc.u32 %1, %5, %8, %1;\n\t"
            "madc.hi.cc.u32 %2, %5, %8, %2;\n\t"
            "addc.cc.u32 %3, %3, 0;\n\t"
            "addc.u32 %4, %4, 0;\n\t"
            "mad.lo.cc.u32 %1, %6, %7, %1;\n\t"
            "madc.hi.cc.u32 %2, %6, %7, %2;\n\t"
            "addc.cc.u32 %3, %3, 0;\n\t"
            "addc.u32 %4, %4, 0;\n\t"
            : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4)
            : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    }
    // canonical value of the sum: 2^64 = EPS, 2^96 = -1, 2^128 = -2^32 (mod p)
    __device__ __forceinline__ uint64_t reduce() const;
};
#endif

GL_HD uint64_t add(uint64_t a, uint64_t b) {  // a, b canonical -> canonical  (f64/mod.rs:273)
#if defined(__CUDA_ARCH__)
    return canon_any(add_ac(a, b));
#else
    uint64_t s = a + b;
    bool wrap = (s < a) | (s >= P);
    return s + (wrap ? EPS : 0ULL);  // s - p == s + EPS (mod 2^64)
#endif
}
GL_HD uint64_t sub(uint64_t a, uint64_t b) {  // f64/mod.rs:293
#if defined(__CUDA_ARCH__)
    return sub_ac(a, b);
#else
    uint64_t d = a - b;
    return d - ((a < b) ? EPS : 0ULL);  // d + p == d - EPS (mod 2^64)
#endif
}
GL_HD uint64_t neg(uint64_t a) { return a ? P - a : 0ULL; }

// 128-bit product {hi,lo} -> canonical.  hi = hh*2^32 + hl:  x = lo - hh + hl*(2^32-1) (mod p)
GL_HD uint64_t reduce128(uint64_t lo, uint64_t hi) {
    uint64_t hh = hi >> 32, hl = hi & EPS;
    uint64_t t0 = lo - hh;
    if (lo < hh) t0 -= EPS;              // borrow: +p
    uint64_t t1 = (hl << 32) - hl;       // hl * (2^32-1) < 2^64
    uint64_t r = t0 + t1;
    if (r < t0) r += EPS;                // carry: 2^64 = EPS (cannot carry twice: see DESIGN.md)
    return r >= P ? r - P : r;
}

GL_HD void mul_wide(uint64_t a, uint64_t b, uint64_t &lo, uint64_t &hi) {
#if defined(__CUDA_ARCH__)
    // four IMAD.WIDE.U32 with 64-bit accumulate; no partial sum can overflow 64 bits
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    uint64_t p00 = (uint64_t)a0 * b0;
    uint64_t t = (uint64_t)a0 * b1 + (p00 >> 32);
    uint64_t u = (uint64_t)a1 * b0 + (t & EPS);
    hi = (uint64_t)a1 * b1 + (t >> 32) + (u >> 32);
    lo = (p00 & EPS) | (u << 32);
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    lo = (uint64_t)x;
    hi = (uint64_t)(x >> 64);
#endif
}

GL_HD uint64_t mul(uint64_t a, uint64_t b) {  // f64/mod.rs:311
#if defined(__CUDA_ARCH__)
    return mul_canon(a, b);
#else
    uint64_t lo, hi;
    mul_wide(a, b, lo, hi);
    return reduce128(lo, hi);
#endif
}
GL_HD uint64_t sqr(uint64_t a) { return mul(a, a); }

GL_HD uint64_t pow(uint64_t b, uint64_t e) {  // f64/mod.rs:103 (exp)
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = mul(r, b);
        b = mul(b, b);
        e >>= 1;
    }
    return r;
}
GL_HD uint64_t inv(uint64_t a) { return pow(a, P - 2); }  // f64/mod.rs:120

// StarkField::get_root_of_unity (math/src/field/traits.rs:224-233)
GL_HD uint64_t root_of_unity(uint32_t log_order) { return pow(TWO_ADIC_ROOT, 1ULL << (32 - log_order)); }

// Boundary conversions.  Rust memory may hold any u64 < 2^64 (f64/mod.rs:56 "stored in the range
// [0, 2^64)"), so reduce first.
GL_HD uint64_t canon(uint64_t x) { return x >= P ? x - P : x; }
GL_HD uint64_t mont_to_canon(uint64_t x) { return mul(canon(x), MONT_R_INV); }
GL_HD uint64_t canon_to_mont(uint64_t x) { return mul(x, EPS); }

#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t Acc160::reduce() const {
    const uint64_t x = reduce128(((uint64_t)r1 << 32) | r0, ((uint64_t)r3 << 32) | r2);
    return sub(x, canon((uint64_t)r4 << 32));
}
#endif

}  // namespace gl
