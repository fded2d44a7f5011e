// selftest.cu -- device-side equivalence checks of the exact math helpers in
// include/gnx_math.h against the builtin IEEE operations they stand in for.
#include "common.cuh"

namespace gnx {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
// double with a random mantissa and an exponent drawn uniformly from [elo, ehi]
__device__ __forceinline__ double rnd_double(uint64_t r, int elo, int ehi, bool neg_ok) {
    const uint64_t mant = r & 0xfffffffffffffull;
    const int e = elo + (int)((r >> 52) % (uint64_t)(ehi - elo + 1));
    const uint64_t sign = (neg_ok && ((r >> 63) & 1)) ? 0x8000000000000000ull : 0ull;
    return __longlong_as_double((long long)(sign | ((uint64_t)(e + 1023) << 52) | mant));
}
// the pre-magic formulation of gnx_exp's range reduction, for comparison
__device__ double exp_rint_form(double x) {
    if (x != x) return x;
    if (x > 709.782712893384) return __longlong_as_double(0x7ff0000000000000ll);
    if (x < -745.2) return 0.0;
    const double kf = rint(__dmul_rn(x, 1.4426950408889634074));
    const int k = (int)kf;
    double r = __fma_rn(-kf, 6.93147180369123816490e-01, x);
    r = __fma_rn(-kf, 1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    const double c[13] = {2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
                          2.48015873015873e-05, 0.0001984126984126984, 0.001388888888888889, 0.008333333333333333,
                          0.041666666666666664, 0.16666666666666666, 0.5, 1.0, 1.0};
    for (int i = 0; i < 13; i++) p = __fma_rn(p, r, c[i]);
    if (k < -1021) return __dmul_rn(__dmul_rn(p, gnx_pow2i(k + 1000)), gnx_pow2i(-1000));
    if (k > 1022) return __dmul_rn(__dmul_rn(p, gnx_pow2i(k - 2)), 4.0);
    return __dmul_rn(p, gnx_pow2i(k));
}

__global__ void selftest_kernel(int64_t n, uint64_t seed, unsigned long long* bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t r0 = mix64(seed + 4 * (uint64_t)i), r1 = mix64(r0), r2 = mix64(r1), r3 = mix64(r2);
        // 0: division, in the shapes the kernels use (1/(1+e), p/s) and generic operands
        {
            double a, b;
            switch (i & 3) {
                case 0: a = 1.0; b = __dadd_rn(1.0, rnd_double(r0, -60, 90, false)); break;
                case 1: a = rnd_double(r0, -200, 0, false); b = rnd_double(r1, -3, 3, false); break;
                case 2: a = rnd_double(r0, -700, 700, true); b = rnd_double(r1, -99, 99, true); break;
                default: a = rnd_double(r0, -1000, 1000, true); b = rnd_double(r1, -300, 300, true); break;
            }
            const double q = gnx_ddiv(a, b), w = __ddiv_rn(a, b);
            if (__double_as_longlong(q) != __double_as_longlong(w)) atomicAdd(bad + 0, 1ull);
        }
        // 1: int64 -> double
        {
            const long long v = (long long)r2 >> (r3 & 63);
            if (__double_as_longlong(gnx_ll2d(v)) != __double_as_longlong(__ll2double_rn(v))) atomicAdd(bad + 1, 1ull);
        }
        // 2: double -> float, including exact ties
        {
            double p = rnd_double(r1, -160, 130, true);
            if ((i & 7) == 0) {  // force a tie or near-tie: 23 mantissa bits + exactly half an ulp (+- 1 bit)
                uint64_t u = (uint64_t)__double_as_longlong(p) & ~0x1fffffffull;
                u |= 0x10000000ull;
                if ((i & 8) == 0) u += (int64_t)((r3 & 3)) - 1;
                p = __longlong_as_double((long long)u);
            }
            if (__float_as_uint(gnx_d2f(p)) != __float_as_uint(__double2float_rn(p))) atomicAdd(bad + 2, 1ull);
        }
        // 3: float -> double
        {
            const float x = __uint_as_float((unsigned)r3);
            if (x == x && __double_as_longlong(gnx_f2d(x)) != __double_as_longlong((double)x)) atomicAdd(bad + 3, 1ull);
        }
        // 4: exp range reduction by the 1.5*2^52 trick vs rint()/(int)
        {
            double x = rnd_double(r2, -20, 9, true);
            if ((i & 15) == 0) x = __dmul_rn((double)((int)(r3 % 2001) - 1000) + 0.5, 0.6931471805599453);  // near k + 1/2
            if (__double_as_longlong(gnx_exp(x)) != __double_as_longlong(exp_rint_form(x))) atomicAdd(bad + 4, 1ull);
        }
    }
}

}  // namespace gnx

extern "C" int gnx_selftest_math(int64_t n, uint64_t seed, int64_t* mismatches /*[5] host*/) {
    GNX_REQUIRE(n > 0 && mismatches, "gnx_selftest_math: bad arguments");
    if (gnx::require_blackwell()) return 1;
    unsigned long long* d = nullptr;
    GNX_CUDA(cudaMalloc((void**)&d, 5 * sizeof(unsigned long long)));
    GNX_CUDA(cudaMemset(d, 0, 5 * sizeof(unsigned long long)));
    gnx::selftest_kernel<<<gnx::sm_count() * 8, 256>>>(n, seed, d);
    unsigned long long h[5];
    cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    GNX_CUDA(e);
    for (int i = 0; i < 5; i++) mismatches[i] = (int64_t)h[i];
    return 0;
}
