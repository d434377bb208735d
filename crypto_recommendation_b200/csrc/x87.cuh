// x87.cuh -- the reference's extended-precision inner product and cosine, emulated with doubles.
// Compiles for the device (nvcc: round-to-nearest intrinsics, no FMA contraction) and for the host (g++ with
// -ffp-contract=off), so that tests/test_x87_cpu.py can check it against real `long double` arithmetic.
#pragma once
#include <cmath>
#include <cstring>

#ifdef __CUDACC__
#define CRX_X87_FN __device__ __forceinline__
#define CRX_DADD(a, b) __dadd_rn((a), (b))
#define CRX_DSUB(a, b) __dsub_rn((a), (b))
#define CRX_DMUL(a, b) __dmul_rn((a), (b))
#define CRX_DDIV(a, b) __ddiv_rn((a), (b))
#define CRX_DSQRT(a) __dsqrt_rn((a))
#define CRX_FMA(a, b, c) __fma_rn((a), (b), (c))
#define CRX_BITS(x) __double_as_longlong((x))
#define CRX_FROM_BITS(b) __longlong_as_double((b))
#else
#define CRX_X87_FN static inline
#define CRX_DADD(a, b) ((a) + (b))
#define CRX_DSUB(a, b) ((a) - (b))
#define CRX_DMUL(a, b) ((a) * (b))
#define CRX_DDIV(a, b) ((a) / (b))
#define CRX_DSQRT(a) std::sqrt((a))
#define CRX_FMA(a, b, c) std::fma((a), (b), (c))
static inline long long crx_bits_(double x) { long long b; std::memcpy(&b, &x, 8); return b; }
static inline double crx_from_bits_(long long b) { double x; std::memcpy(&x, &b, 8); return x; }
#define CRX_BITS(x) crx_bits_((x))
#define CRX_FROM_BITS(b) crx_from_bits_((b))
#endif

// error-free sum: s + e == a + b exactly
CRX_X87_FN void two_sum(double a, double b, double& s, double& e) {
    s = CRX_DADD(a, b);
    double bb = CRX_DSUB(s, a);
    e = CRX_DADD(CRX_DSUB(a, CRX_DSUB(s, bb)), CRX_DSUB(b, bb));
}

// cust_vector.hpp:107-121 accumulates the double-rounded products in an x87 `long double`: every addition is rounded
// to a 64-bit mantissa (round to nearest even).  Such a value is held here as h + l with h a double and l a multiple
// of ulp64 = 2^-11 ulp(h): after each exact addition (two_sum) the low part is rounded to that grid with the
// add-and-subtract-a-constant trick.  cust_vector.hpp:160-174 then divides in extended precision and converts the
// quotient to double: two roundings, reproduced by cos_sim_x87.  (Inputs whose exact sum sits within 2^-105 of a
// rounding tie are the only cases that can differ: probability ~2^-40 per operation.)
struct X87 {
    double h, l;
};
// v rounded to the 64-bit-mantissa grid of the value s + v (|v| <= ulp(s))
CRX_X87_FN double x87_round_low(double s, double v) {
    long long b = CRX_BITS(s);
    long long eb = b & 0x7ff0000000000000LL;
    if (eb < (13LL << 52)) return v;                    // zero / tiny: nothing to round in any realistic input
    bool pow2 = (b & 0x000fffffffffffffLL) == 0;        // s = +-2^e and v pulls the value below it: one binade down
    bool opposite = v != 0.0 && ((v < 0.0) != (s < 0.0));
    long long me = eb - (11LL << 52) - ((pow2 && opposite) ? (1LL << 52) : 0LL);
    double M = CRX_FROM_BITS(me | 0x0008000000000000LL);   // 1.5 * 2^(e-11): ulp(M) = 2^(e-63) = ulp64
#ifdef __CUDACC__
    return CRX_DSUB(CRX_DADD(v, M), M);
#else
    volatile double t = v + M;   // keep the compiler from folding (v + M) - M
    return t - M;
#endif
}
CRX_X87_FN void x87_add(X87& acc, double p) {   // acc = round64(acc + p)
    double s, e, s2, v2;
    two_sum(acc.h, p, s, e);
    two_sum(s, CRX_DADD(acc.l, e), s2, v2);
    acc.h = s2;
    acc.l = x87_round_low(s2, v2);
}
CRX_X87_FN double x87_to_double(const X87& a) { return CRX_DADD(a.h, a.l); }
// double(inner_product / denom), denom = sqrt(na) * sqrt(nb) in double (cust_vector.hpp:171-173)
// the same from the square roots of the two sums of squares (callers that evaluate many pairs of the same rows)
CRX_X87_FN double cos_sim_x87_roots(const X87& ip, double ra, double rb) {
    double denom = CRX_DMUL(ra, rb);
    double q1 = CRX_DDIV(ip.h, denom);
    double r = CRX_FMA(-q1, denom, ip.h);
    double q2 = CRX_DDIV(CRX_DADD(r, ip.l), denom);
    double s, v;
    two_sum(q1, q2, s, v);
    return CRX_DADD(s, x87_round_low(s, v));
}
CRX_X87_FN double cos_sim_x87(const X87& ip, double na, double nb) {
    double denom = CRX_DMUL(CRX_DSQRT(na), CRX_DSQRT(nb));
    double q1 = CRX_DDIV(ip.h, denom);
    double r = CRX_FMA(-q1, denom, ip.h);               // exact remainder of the first quotient digit block
    double q2 = CRX_DDIV(CRX_DADD(r, ip.l), denom);
    double s, v;
    two_sum(q1, q2, s, v);
    return CRX_DADD(s, x87_round_low(s, v));            // round to 64 bits, then to double
}
