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


/* ---------------------------------------------------------------------------------------
 * Device-only exact replacements for operations that the B200 executes on its (very slow
 * for 64-bit operands) conversion/transcendental unit: int64 -> double, double -> float,
 * float -> double and the reciprocal seed of the double division.  Each returns exactly
 * what the plain C operator returns on the host (IEEE round-to-nearest-even), so the
 * oracle keeps using the plain operators; `gnx_selftest_math` (include/gnx.h) checks the
 * equivalence on the device over random operands.
 * ------------------------------------------------------------------------------------- */
#if defined(__CUDACC__)
/* (double)v, correctly rounded: exact halves, one rounding in the fma */
__device__ __forceinline__ double gnx_ll2d(long long v) {
    const int hi = (int)(v >> 32);
    const unsigned lo = (unsigned)v;
    const double dhi = __dsub_rn(__hiloint2double(0x43300000, hi ^ 0x80000000), 4503601774854144.0); /* 2^52 + 2^31 */
    const double dlo = __dsub_rn(__hiloint2double(0x43300000, (int)lo), 4503599627370496.0);          /* 2^52 */
    return __fma_rn(dhi, 4294967296.0, dlo);
}
/* (float)p for doubles whose result is a normal float; anything else takes the builtin */
__device__ __forceinline__ float gnx_d2f(double p) {
    const int hi = __double2hiint(p);
    const unsigned lo = (unsigned)__double2loint(p);
    const unsigned e = ((unsigned)hi >> 20) & 0x7ffu;
    if (e - 897u >= 253u) return __double2float_rn(p);
    const unsigned mant = (((unsigned)hi & 0xfffffu) << 3) | (lo >> 29);
    const unsigned rest = lo & 0x1fffffffu;
    unsigned f = ((unsigned)hi & 0x80000000u) | ((e - 896u) << 23) | mant;
    f += (rest > 0x10000000u || (rest == 0x10000000u && (mant & 1u))) ? 1u : 0u;
    return __uint_as_float(f);
}
/* (double)x, exact, for normal floats and zero */
__device__ __forceinline__ double gnx_f2d(float x) {
    const unsigned b = __float_as_uint(x);
    const unsigned e = (b >> 23) & 0xffu;
    if (e - 1u >= 254u) return (b << 1) == 0u ? __hiloint2double((int)(b & 0x80000000u), 0) : (double)x;
    return __hiloint2double((int)((b & 0x80000000u) | ((e + 896u) << 20) | ((b & 0x7fffffu) >> 3)), (int)((b & 7u) << 29));
}
/* a / b, correctly rounded; fast path for operands well inside the exponent range */
__device__ __forceinline__ double gnx_ddiv(double a, double b) {
    const int hb = __double2hiint(b), ha = __double2hiint(a);
    const unsigned eb = ((unsigned)hb >> 20) & 0x7ffu, ea = ((unsigned)ha >> 20) & 0x7ffu;
    if (eb - 923u > 200u || ea - 200u > 1600u) return __ddiv_rn(a, b);
    /* single-precision reciprocal of b truncated to float as the Newton seed */
    const unsigned lob = (unsigned)__double2loint(b);
    const float bf = __uint_as_float(((unsigned)hb & 0x80000000u) | ((eb - 896u) << 23) | (((unsigned)hb & 0xfffffu) << 3) | (lob >> 29));
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(bf));
    const unsigned rb = __float_as_uint(rf);
    double y = __hiloint2double((int)((rb & 0x80000000u) | ((((rb >> 23) & 0xffu) + 896u) << 20) | ((rb & 0x7fffffu) >> 3)),
                                (int)((rb & 7u) << 29));
    double e = __fma_rn(-b, y, 1.0);
    y = __fma_rn(y, e, y);
    e = __fma_rn(-b, y, 1.0);
    y = __fma_rn(y, e, y);
    e = __fma_rn(-b, y, 1.0);
    y = __fma_rn(y, e, y);
    double q = __dmul_rn(a, y);
    const double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
    return q;
}
#endif
#if defined(__CUDA_ARCH__)
#undef GNX_DIV
#define GNX_DIV(a, b) gnx_ddiv((a), (b))
#define GNX_LL2D(v) gnx_ll2d(v)
#define GNX_D2F(p) gnx_d2f(p)
#define GNX_F2D(x) gnx_f2d(x)
#else
#define GNX_LL2D(v) ((double)(v))
#define GNX_D2F(p) ((float)(p))
#define GNX_F2D(x) ((double)(x))
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

GNX_HD int gnx_double_lo_int(double d) {
#if defined(__CUDA_ARCH__)
    return __double2loint(d);
#else
    uint64_t u;
    memcpy(&u, &d, sizeof u);
    return (int)(uint32_t)u;
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
    /* round-to-nearest-even integer of x*log2(e) by the 1.5*2^52 trick (|t| < 2^31): identical
     * to rint() and free of the slow 64-bit convert instructions on the device */
    const double kd = GNX_ADD(GNX_MUL(x, LOG2E), 6755399441055744.0);
    const int k = gnx_double_lo_int(kd);
    const double kf = GNX_SUB(kd, 6755399441055744.0);
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
GNX_HD float gnx_expf_cr(float x) { return GNX_D2F(gnx_exp(GNX_F2D(x))); }

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
