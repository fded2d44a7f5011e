/*
 * gnx_math.h -- deterministic scalar math shared by the CUDA kernels (device) and
 * the CPU oracle (host C).  Every function is a fixed sequence of IEEE-754
 * correctly-rounded operations (add, mul, fma, div, rint, conversions), so the same
 * inputs give the same BITS on the GPU and on the CPU.  That is what lets the
 * parity tests demand bit-exact float32 outputs from the smoother kernels instead
 * of a tolerance: libm's exp()/expf() and CUDA's differ in the last ulp.
 *
 * Host build: compile with -ffp-contract=off (the Makefile under oracle/ does).
 * Device build: the *_rn intrinsics are never contracted by nvcc.
 *
 * What these replace in the reference's third-party stack:
 *   gnx_exp      scipy.special.expit's exp (LogisticRegression._predict_proba_lr,
 *                reference call site src/Base/base.py:174)
 *   gnx_expf_cr  expf() inside xgboost's Softmax (common/math.h), reference call
 *                site src/Smooth/smooth.py:46
 */
#ifndef GNX_MATH_H_
#define GNX_MATH_H_

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define GNX_HD __host__ __device__ __forceinline__
#define GNX_MUL(a, b) __dmul_rn((a), (b))
#define GNX_ADD(a, b) __dadd_rn((a), (b))
#define GNX_SUB(a, b) __dsub_rn((a), (b))
#define GNX_FMA(a, b, c) __fma_rn((a), (b), (c))
#define GNX_DIV(a, b) __ddiv_rn((a), (b))
#define GNX_FSUB(a, b) __fsub_rn((a), (b))
#define GNX_FADD(a, b) __fadd_rn((a), (b))
#define GNX_FDIV(a, b) __fdiv_rn((a), (b))
#elif defined(__CUDACC__)
#define GNX_HD __host__ __device__ inline
#define GNX_MUL(a, b) ((a) * (b))
#define GNX_ADD(a, b) ((a) + (b))
#define GNX_SUB(a, b) ((a) - (b))
#define GNX_FMA(a, b, c) fma((a), (b), (c))
#define GNX_DIV(a, b) ((a) / (b))
#define GNX_FSUB(a, b) ((a) - (b))
#define GNX_FADD(a, b) ((a) + (b))
#define GNX_FDIV(a, b) ((a) / (b))
#else
#define GNX_HD static inline
#define GNX_MUL(a, b) ((a) * (b))
#define GNX_ADD(a, b) ((a) + (b))
#define GNX_SUB(a, b) ((a) - (b))
#define GNX_FMA(a, b, c) fma((a), (b), (c))
#define GNX_DIV(a, b) ((a) / (b))
#define GNX_FSUB(a, b) ((a) - (b))
#define GNX_FADD(a, b) ((a) + (b))
#define GNX_FDIV(a, b) ((a) / (b))
#endif

GNX_HD double gnx_bits_to_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
#endif
}

/* 2^k for -1022 <= k <= 1023 */
GNX_HD double gnx_pow2i(int k) { return gnx_bits_to_double((uint64_t)(k + 1023) << 52); }

/* exp(x), ~1 ulp, bit-reproducible across host and device. */
GNX_HD double gnx_exp(double x) {
    if (x != x) return x;
    if (x > 709.782712893384) return gnx_bits_to_double(0x7ff0000000000000ULL);
    if (x < -745.2) return 0.0;
    const double LOG2E = 1.4426950408889634074;
    const double LN2_HI = 6.93147180369123816490e-01; /* 0x3fe62e42fee00000 */
    const double LN2_LO = 1.90821492927058770002e-10; /* 0x3dea39ef35793c76 */
    double kf = rint(GNX_MUL(x, LOG2E));
    int k = (int)kf;
    double r = GNX_FMA(-kf, LN2_HI, x);
    r = GNX_FMA(-kf, LN2_LO, r);
    /* Taylor to degree 13 on |r| <= 0.3466: truncation 4e-18 relative */
    double p = 1.6059043836821613e-10;              /* 1/13! */
    p = GNX_FMA(p, r, 2.08767569878681e-09);        /* 1/12! */
    p = GNX_FMA(p, r, 2.505210838544172e-08);       /* 1/11! */
    p = GNX_FMA(p, r, 2.755731922398589e-07);       /* 1/10! */
    p = GNX_FMA(p, r, 2.7557319223985893e-06);      /* 1/9!  */
    p = GNX_FMA(p, r, 2.48015873015873e-05);        /* 1/8!  */
    p = GNX_FMA(p, r, 0.0001984126984126984);       /* 1/7!  */
    p = GNX_FMA(p, r, 0.001388888888888889);        /* 1/6!  */
    p = GNX_FMA(p, r, 0.008333333333333333);        /* 1/5!  */
    p = GNX_FMA(p, r, 0.041666666666666664);        /* 1/4!  */
    p = GNX_FMA(p, r, 0.16666666666666666);         /* 1/3!  */
    p = GNX_FMA(p, r, 0.5);
    p = GNX_FMA(p, r, 1.0);
    p = GNX_FMA(p, r, 1.0);
    if (k < -1021) return GNX_MUL(GNX_MUL(p, gnx_pow2i(k + 1000)), gnx_pow2i(-1000));
    if (k > 1022) return GNX_MUL(GNX_MUL(p, gnx_pow2i(k - 2)), 4.0);
    return GNX_MUL(p, gnx_pow2i(k));
}

/* float exp computed in double and rounded once: (with overwhelming probability)
 * the correctly rounded expf, and bit-identical on host and device. */
GNX_HD float gnx_expf_cr(float x) { return (float)gnx_exp((double)x); }

/* scipy.special.expit for float64: 1/(1+exp(-x)) */
GNX_HD double gnx_expit(double x) { return GNX_DIV(1.0, GNX_ADD(1.0, gnx_exp(-x))); }

/* numpy's float64 add.reduce over a contiguous run of n <= 128 values
 * (numpy/core/src/umath/loops_utils.h pairwise sum: <8 sequential, else 8 lanes). */
GNX_HD double gnx_np_sum(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; i++) res = GNX_ADD(res, a[i]);
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] = GNX_ADD(r[j], a[i + j]);
    double res = GNX_ADD(GNX_ADD(GNX_ADD(r[0], r[1]), GNX_ADD(r[2], r[3])),
                         GNX_ADD(GNX_ADD(r[4], r[5]), GNX_ADD(r[6], r[7])));
    for (; i < n; i++) res = GNX_ADD(res, a[i]);
    return res;
}

#endif /* GNX_MATH_H_ */
