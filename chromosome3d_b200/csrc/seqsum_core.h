// seqsum_core.h — the arithmetic behind the parallel evaluation of a LEFT-TO-RIGHT fp64 sum (seqsum.cuh).
//
// The Perl's `$sum += ...` loops (C3D:132-139 the mean of IF**alpha, C3D:590-598 the sum of deviations) round after every
// addition, so their result depends on the order: a tree sum does not reproduce them.  A serial chain of DADDs does, at 4.4 ns
// per term on a B200 (0.9 ms for the 207,025 cells of chr1_500kb — the largest single item between the device-timed ensemble
// and the end-to-end call).  This header states what makes the serial sum parallel WITHOUT changing a bit:
//
//   While the running sum s stays in one binade [2^e, 2^(e+1)) it is an integer multiple S of u = 2^(e-52), and adding a term
//   0 <= a < 2^(e+1) with round-to-nearest gives  S' = S + q + [r > u/2],  a = q*u + r, 0 <= r < u  — as long as r != u/2 (a tie
//   goes to the even neighbour: it depends on the parity of S) and S' < 2^53 (otherwise the sum leaves the binade and u changes).
//   So inside a binade each term contributes an INTEGER that does not depend on s, integer addition is associative, and a
//   prefix scan finds the first term that is a tie, negative, non-finite, or carries the sum out of the binade.  Everything
//   before it is consumed at once; that one term is added with a real fp64 add (which is the definition); repeat.
//
// Plain C++ on purpose (no CUDA types): tests/seqsum_host_check.cpp compiles the same functions with g++ and checks them
// against a scalar loop on adversarial inputs (ties, binade crossings, zeros, denormals, negative and non-finite terms).
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SEQSUM_HD __host__ __device__ __forceinline__
#else
#define SEQSUM_HD inline
#endif

namespace chr3d {

SEQSUM_HD uint64_t seqsum_bits(double x) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t b; memcpy(&b, &x, 8); return b;
#endif
}
SEQSUM_HD double seqsum_from_bits(uint64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double x; memcpy(&x, &b, 8); return x;
#endif
}

#define SEQSUM_ONE52 (1ull << 52)
#define SEQSUM_TOP (1ull << 53)

// The running sum as (biased exponent, 53-bit integer significand).  ok: s is a positive normal number whose unit in the last
// place is a normal number as well; anything else (zero, denormal, negative, inf, nan) takes the one-term-at-a-time path.
struct SeqSumFrame { int ef; uint64_t S; bool ok; };
SEQSUM_HD SeqSumFrame seqsum_frame(double s) {
    const uint64_t b = seqsum_bits(s);
    SeqSumFrame f;
    f.ef = (int)((b >> 52) & 0x7ffull);
    f.S = (b & (SEQSUM_ONE52 - 1)) | SEQSUM_ONE52;
    f.ok = (b >> 63) == 0 && f.ef >= 64 && f.ef <= 2046;
    return f;
}
// the sum back from its frame (2^52 <= S < 2^53)
SEQSUM_HD double seqsum_value(int ef, uint64_t S) { return seqsum_from_bits(((uint64_t)ef << 52) | (S & (SEQSUM_ONE52 - 1))); }

// Contribution of the term a, in units of the frame's last place.  stop = true: this term cannot be folded into the integer
// scan (tie, negative, denormal, non-finite, at least 2^(e+1), or no valid frame) and has to be added with an fp64 add.
// The return value never exceeds 2^53 - 1.
SEQSUM_HD uint64_t seqsum_contrib(double a, int ef, bool frame_ok, bool& stop) {
    const uint64_t b = seqsum_bits(a);
    stop = false;
    if ((b << 1) == 0) return 0;                   // +0 / -0: s + 0 = s (s is never -0: the sum starts at +0)
    if (!frame_ok) { stop = true; return 0; }
    const int efa = (int)((b >> 52) & 0x7ffull);
    if ((b >> 63) != 0 || efa == 0 || efa == 0x7ff || efa > ef) { stop = true; return 0; }
    const int sh = ef - efa;
    const uint64_t m = (b & (SEQSUM_ONE52 - 1)) | SEQSUM_ONE52;
    if (sh == 0) return m;                         // same binade: exact in units of u (and S + m >= 2^53: the caller sees the carry)
    if (sh >= 54) return 0;                        // a < u/2
    const uint64_t half = 1ull << (sh - 1), r = m & ((half << 1) - 1);
    if (r == half) { stop = true; return 0; }
    return (m >> sh) + (r > half ? 1ull : 0ull);
}

// The same contribution through three fp64 additions instead of 64-bit integer shifts (what the kernel runs: a fifth of the
// instructions).  B = 2^e is the bottom of the frame's binade — an even multiple of u — so for 0 < a < B/2
//     x = B + a        rounds a to a multiple of u exactly as s + a does (ties aside), and stays below 1.5 B,
//     cu = x - B       is exact (Sterbenz), the contribution times u; its integer value is the low 52 bits of x,
//     r = a - cu       is exact as well (|r| <= u/2, a multiple of a's own last place): |r| = u/2 is the tie.
// Terms of B/2 and more (a quarter of the running sum: only the first few terms of a sum) are left to the fp64 add, like
// negative and non-finite ones; a denormal term adds 0 here (it is below u/2 in every valid frame).
struct SeqSumFpFrame { double B, halfB, halfu; };
SEQSUM_HD SeqSumFpFrame seqsum_fp_frame(const SeqSumFrame& f) {
    SeqSumFpFrame g;
    const uint64_t ef = (uint64_t)(f.ok ? f.ef : 1023);
    g.B = seqsum_from_bits(ef << 52); g.halfB = seqsum_from_bits((ef - 1) << 52); g.halfu = seqsum_from_bits((ef - 53) << 52);
    return g;
}
SEQSUM_HD double seqsum_add(double x, double y) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(x, y);         // never re-associated or fused
#else
    volatile double t = x + y; return t;
#endif
}
SEQSUM_HD uint64_t seqsum_contrib_fp(double a, const SeqSumFpFrame& g, bool frame_ok, bool& stop) {
    stop = false;
    if (a == 0.0) return 0;                                                   // +0 / -0
    if (!frame_ok || !(a > 0.0) || !(a < g.halfB)) { stop = true; return 0; }  // a NaN fails both comparisons
    const double x = seqsum_add(g.B, a);
    const double cu = seqsum_add(x, -g.B);
    const double r = seqsum_add(a, -cu);
    if (r == g.halfu || r == -g.halfu) { stop = true; return 0; }
    return seqsum_bits(x) & (SEQSUM_ONE52 - 1);
}

}  // namespace chr3d
