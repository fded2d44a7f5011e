// lr_base.cuh -- model handle + shared epilogue of the logistic-regression base (K1)
#pragma once

#include <vector>

#include "common.cuh"

namespace gnx {

constexpr int LR_NCOLS = 64;    // MMA N: fixed-point limb columns per window
constexpr int LR_KC = 128;      // SNPs (bytes) per chunk == one 128B swizzle atom
constexpr int LR_TILE_BYTES = LR_NCOLS * LR_KC;

// Device-side view of the packed model, passed by value to the kernels.
struct LrDev {
    int A, Ar, L, apad, s, W;
    int64_t C, M, ctx;
    int n_chunks;
    int raw;                   // 1: a class group of a split model (A > 8): store expit(d) as float64, stride A, no normalisation
    int dbg;                   // profiling switches (GNX_LR_DBG): 1 = no MMA, 2 = no MMA + no epilogue math, 4 = no weight loads
    const int8_t* wt;          // [n_tiles][LR_NCOLS][LR_KC] int8 limb planes; column = limb*apad + class
    const double* bias;        // [W][Ar]
    const int32_t* k0;         // [W] first chunk of window w
    const int32_t* kend;       // [W] one past the last chunk of window w
    const int32_t* tile_off;   // [W] tile index of (w, k0[w])
    const int32_t* chunk_w0;   // [n_chunks] first window covering chunk k
    const int32_t* chunk_wn;   // [n_chunks] number of windows covering chunk k
    const uint4* chunk_sched;  // [n_chunks] i-th covering window: tile << 3 | valid << 2 | last chunk << 1 | first chunk
};

}  // namespace gnx

struct gnx_lr {
    gnx::LrDev d;
    int device;
    int kernel_sel;  // 0 tcgen05, 1 dp4a
    int n_tiles;
    std::vector<int32_t> h_k0, h_kend, h_tile_off, h_chunk_w0, h_chunk_wn;
    std::vector<uint32_t> h_chunk_sched;
    void* d_blob;    // single allocation holding every table
    // TMA descriptor of the weight tensor (128 bytes, CUtensorMap) -- built lazily
    alignas(64) unsigned char tmap_w[128];
    bool tmap_w_ready;
    // A > 8 with the full 7 limbs: two class groups of <= 8 classes (each a model of its own, `raw` output, the SAME
    // fixed-point scale), run one after the other and normalised together (lr_base.cu); this handle then only carries
    // the geometry in `d`
    gnx_lr* sub[2];
};

namespace gnx {

int lr_launch_tc(const gnx_lr* m, const int8_t* X, int64_t N, int64_t ldX, void* B, bool f64, cudaStream_t st);
int lr_launch_dp4a_any(const gnx_lr* m, const int8_t* X, int64_t N, int64_t ldX, void* B, bool f64, cudaStream_t st);

template <typename OutT>
__device__ __forceinline__ OutT lr_out(double v);
template <>
__device__ __forceinline__ double lr_out<double>(double v) { return v; }
template <>
__device__ __forceinline__ float lr_out<float>(double v) { return GNX_D2F(v); }

// numpy add.reduce order over the A class probabilities of a row (gnx_np_sum): sequential below 8 terms, 8-lane
// pairwise from 8 on
template <int APAD>
__device__ __forceinline__ double lr_np_sum(const double (&p)[APAD], int A) {
    double s;
    if (A < 8) {
        s = 0.0;
#pragma unroll
        for (int a = 0; a < (APAD < 8 ? APAD : 8); a++)
            if (a < A) s = GNX_ADD(s, p[a]);
    } else {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = p[j];
        if (APAD > 8) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (A >= 16) r[j] = GNX_ADD(r[j], p[(8 + j) % APAD]);
        }
        s = GNX_ADD(GNX_ADD(GNX_ADD(r[0], r[1]), GNX_ADD(r[2], r[3])), GNX_ADD(GNX_ADD(r[4], r[5]), GNX_ADD(r[6], r[7])));
        const int done = (A >= 16) ? 16 : 8;
#pragma unroll
        for (int a = 8; a < APAD; a++)
            if (a >= done && a < A) s = GNX_ADD(s, p[a]);
    }
    return s;
}

// (hap, window) epilogue shared by both kernels: limb recombination (exact int64),
// scale, intercept, expit, row-normalise (sklearn _predict_proba_lr), store.
//   acc: LR_NCOLS int32 accumulators of this haplotype, column = limb*APAD + class
template <int APAD, typename OutT>
__device__ __forceinline__ void lr_epilogue_store(const int32_t (&acc)[LR_NCOLS], const LrDev& m, int w,
                                                  OutT* __restrict__ out /* &B[n][w][0] */) {
    constexpr int LMAX = LR_NCOLS / APAD;
    double d[APAD];
    const double scale = gnx_pow2i(-m.s);
#pragma unroll
    for (int a = 0; a < APAD; a++) {
        // exact 64-bit Horner over the limbs: tot = sum_l acc_l * 256^l.  No overflow: the scale is
        // chosen at model create so that 2 * sum |q| < 2^62 (|x| <= 2), and the int32 accumulators
        // hold for windows up to 131000 SNPs
        long long tot = 0;
#pragma unroll
        for (int l = LMAX - 1; l >= 0; l--)
            if (l < m.L) tot = tot * 256 + (long long)acc[l * APAD + a];
        d[a] = 0.0;
        if (a < m.Ar) d[a] = GNX_ADD(GNX_MUL(GNX_LL2D(tot), scale), __ldg(m.bias + (int64_t)w * m.Ar + a));
    }
    if (m.raw) {   // class group of a split model: un-normalised sigmoids, float64 (OutT is double on this path)
#pragma unroll
        for (int a = 0; a < APAD; a++)
            if (a < m.Ar) out[a] = lr_out<OutT>(gnx_expit(d[a]));
        return;
    }
    if (m.A == 2) {
        double p = gnx_expit(d[0]);
        out[0] = lr_out<OutT>(GNX_SUB(1.0, p));
        out[1] = lr_out<OutT>(p);
        return;
    }
    double p[APAD];
#pragma unroll
    for (int a = 0; a < APAD; a++) p[a] = (a < m.A) ? gnx_expit(d[a]) : 0.0;
    const double s = lr_np_sum<APAD>(p, m.A);
#pragma unroll
    for (int a = 0; a < APAD; a++)
        if (a < m.A) out[a] = lr_out<OutT>(GNX_DIV(p[a], s));
}

}  // namespace gnx
