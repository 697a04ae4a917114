/*
 * jxpow.h -- a deterministic x^y for x > 0 built only from IEEE-754 binary64
 * +, -, *, /, fma, rint and bit moves, so that the SAME sequence of correctly
 * rounded operations runs on the host (gcc, -ffp-contract=off) and on the
 * device (nvcc; explicit __dmul_rn/__dadd_rn/__fma_rn so -fmad cannot contract).
 * Result: identical bits on CPU and GPU, error < 0.7 ulp (measured against
 * an exact rational reference in tests/test_jxpow.py).
 *
 * Why: Jexpresso's equation of state is P = C0 (ρθ)^γ (src/kernel/physics/
 * constitutiveLaw.jl:22-24), evaluated with Julia's `^` (< 1 ulp).  libm pow,
 * CUDA pow and Julia pow are three different < 1..2 ulp implementations, so a
 * kernel using CUDA's pow can only match a CPU restatement to ~1e-16 relative
 * in P -- which hydrostatic cancellation amplifies to ~1e-12 in the RHS.  With
 * this header both sides compute the same bits and the element kernels can be
 * checked for EXACT equality against the oracle.
 *
 * Method: x = 2^e m, m in [sqrt(1/2), sqrt(2)];  log m = 2 atanh(s), s = (m-1)/(m+1)
 * carried as a double-double head plus a double tail series;  y*log x in
 * double-double;  exp by k = rint(p/ln2), r = p - k ln2 (two-part ln2), degree-14
 * Taylor tail added to the exact sum 1 + r.
 */
#ifndef JXPOW_H
#define JXPOW_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define JXP_FN __device__ __forceinline__
#define JXP_MUL(a, b) __dmul_rn((a), (b))
#define JXP_ADD(a, b) __dadd_rn((a), (b))
#define JXP_SUB(a, b) __dsub_rn((a), (b))
#define JXP_FMA(a, b, c) __fma_rn((a), (b), (c))
#define JXP_DIV(a, b) __ddiv_rn((a), (b))
#define JXP_RINT(a) rint(a)
#define JXP_D2LL(a) __double_as_longlong(a)
#define JXP_LL2D(a) __longlong_as_double(a)
#else
#if defined(__CUDACC__)
#define JXP_FN __host__ inline
#else
#define JXP_FN static inline
#endif
#define JXP_MUL(a, b) ((a) * (b))
#define JXP_ADD(a, b) ((a) + (b))
#define JXP_SUB(a, b) ((a) - (b))
#define JXP_FMA(a, b, c) fma((a), (b), (c))
#define JXP_DIV(a, b) ((a) / (b))
#define JXP_RINT(a) rint(a)
static inline int64_t jxp_d2ll(double a) { int64_t r; memcpy(&r, &a, 8); return r; }
static inline double jxp_ll2d(int64_t a) { double r; memcpy(&r, &a, 8); return r; }
#define JXP_D2LL(a) jxp_d2ll(a)
#define JXP_LL2D(a) jxp_ll2d(a)
#endif

JXP_FN double jx_pow(double x, double y) {
    int64_t bits = JXP_D2LL(x);
    int64_t ebits = (bits >> 52) & 0x7ff;
    if (bits <= 0 || ebits == 0 || ebits == 0x7ff || !(y == y)) return pow(x, y); /* outside the EOS domain */
    int64_t e = ebits - 1023;
    double m = JXP_LL2D((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL); /* [1,2) */
    if (m > 1.4142135623730951) { m = JXP_MUL(m, 0.5); e += 1; }
    const double de = (double)e;

    /* s = (m-1)/(m+1) as s_hi + s_lo */
    const double num = JXP_SUB(m, 1.0);                   /* exact (Sterbenz) */
    const double den = JXP_ADD(m, 1.0);
    const double bb = JXP_SUB(den, m);                    /* TwoSum error of m + 1 */
    const double den_lo = JXP_ADD(JXP_SUB(m, JXP_SUB(den, bb)), JXP_SUB(1.0, bb));
    const double s_hi = JXP_DIV(num, den);
    const double rem = JXP_FMA(-s_hi, den, num);          /* exact remainder */
    const double s_lo = JXP_DIV(JXP_FMA(-s_hi, den_lo, rem), den);

    /* atanh tail: 2 s^3 (1/3 + z/5 + z^2/7 + ...), z = s^2 */
    const double z = JXP_MUL(s_hi, s_hi);
    double P = 1.0 / 27.0;
    P = JXP_FMA(P, z, 1.0 / 25.0);
    P = JXP_FMA(P, z, 1.0 / 23.0);
    P = JXP_FMA(P, z, 1.0 / 21.0);
    P = JXP_FMA(P, z, 1.0 / 19.0);
    P = JXP_FMA(P, z, 1.0 / 17.0);
    P = JXP_FMA(P, z, 1.0 / 15.0);
    P = JXP_FMA(P, z, 1.0 / 13.0);
    P = JXP_FMA(P, z, 1.0 / 11.0);
    P = JXP_FMA(P, z, 1.0 / 9.0);
    P = JXP_FMA(P, z, 1.0 / 7.0);
    P = JXP_FMA(P, z, 1.0 / 5.0);
    P = JXP_FMA(P, z, 1.0 / 3.0);
    const double tail = JXP_MUL(2.0, JXP_FMA(JXP_MUL(s_hi, z), P, s_lo));
    const double L1 = JXP_MUL(2.0, s_hi);

    /* log x = e ln2 + log m  as (H, l) */
    const double ln2_hi = 6.93147180369123816490e-01;     /* 0x3fe62e42fee00000: low 21 bits zero */
    const double ln2_lo = 1.90821492927058770002e-10;
    const double A = JXP_MUL(de, ln2_hi);                 /* exact */
    const double H0 = JXP_ADD(A, L1);
    const double bv = JXP_SUB(H0, A);
    const double h0 = JXP_ADD(JXP_SUB(A, JXP_SUB(H0, bv)), JXP_SUB(L1, bv));
    const double l0 = JXP_ADD(h0, JXP_FMA(de, ln2_lo, tail));
    const double H = JXP_ADD(H0, l0);
    const double l = JXP_SUB(l0, JXP_SUB(H, H0));

    /* p = y * log x  as (p_hi, p_lo) */
    const double p_hi = JXP_MUL(y, H);
    const double p_lo = JXP_FMA(y, l, JXP_FMA(y, H, -p_hi));

    /* exp(p) */
    const double k = JXP_RINT(JXP_MUL(p_hi, 1.4426950408889634));
    if (!(k > -1000.0 && k < 1000.0)) return pow(x, y);
    const double r_hi = JXP_FMA(-k, ln2_hi, p_hi);
    const double r_lo = JXP_FMA(-k, ln2_lo, p_lo);
    const double r = JXP_ADD(r_hi, r_lo);
    const double c = JXP_ADD(JXP_SUB(r_hi, r), r_lo);
    double Q = 1.0 / 87178291200.0;                       /* 1/14! */
    Q = JXP_FMA(Q, r, 1.0 / 6227020800.0);
    Q = JXP_FMA(Q, r, 1.0 / 479001600.0);
    Q = JXP_FMA(Q, r, 1.0 / 39916800.0);
    Q = JXP_FMA(Q, r, 1.0 / 3628800.0);
    Q = JXP_FMA(Q, r, 1.0 / 362880.0);
    Q = JXP_FMA(Q, r, 1.0 / 40320.0);
    Q = JXP_FMA(Q, r, 1.0 / 5040.0);
    Q = JXP_FMA(Q, r, 1.0 / 720.0);
    Q = JXP_FMA(Q, r, 1.0 / 120.0);
    Q = JXP_FMA(Q, r, 1.0 / 24.0);
    Q = JXP_FMA(Q, r, 1.0 / 6.0);
    Q = JXP_FMA(Q, r, 0.5);
    const double small = JXP_FMA(JXP_MUL(r, r), Q, c);
    const double S = JXP_ADD(1.0, r);
    const double bs = JXP_SUB(S, 1.0);
    const double s_err = JXP_ADD(JXP_SUB(1.0, JXP_SUB(S, bs)), JXP_SUB(r, bs));
    const double er = JXP_ADD(S, JXP_ADD(s_err, small));
    const double two_k = JXP_LL2D(((int64_t)k + 1023) << 52);
    return JXP_MUL(er, two_k);
}

#endif /* JXPOW_H */
