// lr_base.cu -- K1: LogisticRegressionBase.predict_proba for all windows.
//
// Replaces src/Base/base.py:146-180 + sklearn LogisticRegression.predict_proba
// (src/Base/models.py:12-21).  The float64 weights are turned into exact fixed
// point: q = rint(w * 2^s), reflect pads folded onto the SNPs they mirror (in
// integers), q split into L signed base-256 limbs.  X is int8 in {0,1,2}, so every
// limb plane is an int8 x int8 -> int32 contraction with an exact result; the
// epilogue recombines the limbs in int64 and finishes in float64.
//
// This file: host-side model packing, the dp4a CUDA-core kernel (cross-check path)
// and the C ABI.  The tcgen05/TMA/TMEM kernel lives in lr_base_tc.cu and consumes
// the same packed tiles.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "lr_base.cuh"

namespace gnx {

static inline int64_t pad_to_orig(int64_t p, int64_t C, int64_t ctx) {
    if (p < ctx) return ctx - 1 - p;
    if (p >= ctx + C) return 2 * C + ctx - 1 - p;
    return p - ctx;
}

// ------------------------------------------------------------------ dp4a kernel
// grid (W, ceil(N/128)), 128 threads; thread = one haplotype x one window with all
// 64 limb-column accumulators in registers; weights broadcast from shared memory.
template <int APAD, typename OutT>
__global__ void __launch_bounds__(128) lr_dp4a_kernel(LrDev m, const int8_t* __restrict__ X, int64_t N,
                                                      int64_t ldX, OutT* __restrict__ B) {
    __shared__ uint32_t xs[128][33];
    __shared__ __align__(16) uint32_t ws[32][68];
    const int w = blockIdx.x;
    const int64_t hap0 = (int64_t)blockIdx.y * 128;
    const int tid = threadIdx.x;
    int32_t acc[LR_NCOLS];
#pragma unroll
    for (int c = 0; c < LR_NCOLS; c++) acc[c] = 0;
    const int k0 = m.k0[w], kend = m.kend[w];
    const bool aligned4 = ((ldX & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 3) == 0);
    for (int k = k0; k < kend; k++) {
        const int64_t j0 = (int64_t)k * LR_KC;
        __syncthreads();
        // stage X chunk [128 haps][128 SNPs]
        for (int i = 0; i < 32; i++) {
            const int row = i * 4 + (tid >> 5), word = tid & 31;
            const int64_t n = hap0 + row, j = j0 + word * 4;
            uint32_t v = 0;
            if (n < N) {
                const int8_t* p = X + n * ldX + j;
                if (aligned4 && j + 4 <= m.C) {
                    v = *reinterpret_cast<const uint32_t*>(p);
                } else {
#pragma unroll
                    for (int b = 0; b < 4; b++)
                        if (j + b < m.C) v |= (uint32_t)(uint8_t)p[b] << (8 * b);
                }
            }
            xs[row][word] = v;
        }
        // stage weight tile, transposed to [word][col]
        const uint32_t* wt = reinterpret_cast<const uint32_t*>(m.wt + (int64_t)(m.tile_off[w] + (k - k0)) * LR_TILE_BYTES);
        for (int i = 0; i < 16; i++) {
            const int idx = i * 128 + tid;
            ws[idx & 31][idx >> 5] = wt[idx];
        }
        __syncthreads();
#pragma unroll 4
        for (int q = 0; q < 32; q++) {
            const int xw = (int)xs[tid][q];
#pragma unroll
            for (int c4 = 0; c4 < LR_NCOLS / 4; c4++) {
                const uint4 wv = *reinterpret_cast<const uint4*>(&ws[q][c4 * 4]);
                acc[c4 * 4 + 0] = __dp4a(xw, (int)wv.x, acc[c4 * 4 + 0]);
                acc[c4 * 4 + 1] = __dp4a(xw, (int)wv.y, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = __dp4a(xw, (int)wv.z, acc[c4 * 4 + 2]);
                acc[c4 * 4 + 3] = __dp4a(xw, (int)wv.w, acc[c4 * 4 + 3]);
            }
        }
    }
    const int64_t n = hap0 + tid;
    if (n < N) lr_epilogue_store<APAD, OutT>(acc, m, w, B + (n * m.W + w) * m.A);
}

template <typename OutT>
static int lr_launch_dp4a(const gnx_lr* m, const int8_t* X, int64_t N, int64_t ldX, OutT* B, cudaStream_t st) {
    dim3 grid((unsigned)m->d.W, (unsigned)ceil_div(N, 128));
    if (m->d.apad == 8)
        lr_dp4a_kernel<8, OutT><<<grid, 128, 0, st>>>(m->d, X, N, ldX, B);
    else
        lr_dp4a_kernel<16, OutT><<<grid, 128, 0, st>>>(m->d, X, N, ldX, B);
    GNX_CUDA(cudaGetLastError());
    return 0;
}

int lr_launch_dp4a_any(const gnx_lr* m, const int8_t* X, int64_t N, int64_t ldX, void* B, bool f64, cudaStream_t st) {
    return f64 ? lr_launch_dp4a<double>(m, X, N, ldX, (double*)B, st) : lr_launch_dp4a<float>(m, X, N, ldX, (float*)B, st);
}

// split model (A > 8): row-normalise the un-normalised sigmoids of the two class groups (sklearn _predict_proba_lr:
// p /= p.sum(axis=1), numpy's summation order) into B [rows, na0 + na1]
template <typename OutT>
__global__ void lr_split_normalise_kernel(const double* __restrict__ p0, const double* __restrict__ p1, int na0, int na1, int64_t rows,
                                          OutT* __restrict__ B) {
    const int A = na0 + na1;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        double p[16];
#pragma unroll
        for (int a = 0; a < 16; a++) p[a] = 0.0;
#pragma unroll
        for (int a = 0; a < 8; a++)
            if (a < na0) p[a] = p0[r * na0 + a];
#pragma unroll
        for (int a = 0; a < 8; a++)
            if (a < na1) p[8 + a] = p1[r * na1 + a];
        const double s = lr_np_sum<16>(p, A);
#pragma unroll
        for (int a = 0; a < 16; a++)
            if (a < A) B[r * A + a] = lr_out<OutT>(GNX_DIV(p[a], s));
    }
}

}  // namespace gnx

using namespace gnx;

extern "C" {

// One packed model.  A = classes of the output row (its stride), Ar = coefficient rows per window (1 for the binary layout),
// raw = 1 for a class group of a split model (un-normalised float64 sigmoids out), s_forced >= 0 fixes the fixed-point
// exponent (the class groups of a split model share the exponent chosen over ALL classes, so every logit is the one a
// single 7-limb model would produce).
static int lr_create_one(gnx_lr_t** out, int A, int Ar, int raw, int s_forced, int64_t C, int64_t M, int64_t ctx, const double* coef,
                         const double* intercept, int limbs, int* s_chosen) {
    const int apad = (Ar <= 8) ? 8 : 16;
    const int Lmax = LR_NCOLS / apad;
    int L = limbs == 0 ? 7 : limbs;
    if (L > Lmax) L = Lmax;
    GNX_REQUIRE(L >= 2, "gnx_lr_model_create: limbs=%d too small", L);
    const int64_t W = C / M, rem = C - M * W, M_ = M + 2 * ctx;
    GNX_REQUIRE(W < (1 << 30), "too many windows");
    GNX_REQUIRE(M_ + rem <= 131000, "gnx_lr_model_create: windows of %lld SNPs exceed the 131000 the int32 limb accumulators allow", (long long)(M_ + rem));

    // padded window ranges, folded original ranges
    std::vector<int64_t> lo(W), len(W), s0(W), e0(W), coff(W);
    int64_t off = 0, maxlen = 0;
    for (int64_t w = 0; w < W; w++) {
        lo[w] = (w == W - 1) ? (C + 2 * ctx - (M_ + rem)) : w * M;
        len[w] = (w == W - 1) ? (M_ + rem) : M_;
        s0[w] = std::max<int64_t>(0, w * M - ctx);
        e0[w] = (w == W - 1) ? C : std::min<int64_t>(C, w * M + M + ctx);
        coff[w] = off;
        off += (int64_t)Ar * len[w];
        maxlen = std::max(maxlen, e0[w] - s0[w]);
    }
    // pass 1: scale selection from the float64-folded weights
    double amax = 0.0, ssum = 0.0;
    {
        std::vector<double> f(maxlen);
        for (int64_t w = 0; w < W; w++)
            for (int a = 0; a < Ar; a++) {
                std::fill(f.begin(), f.begin() + (e0[w] - s0[w]), 0.0);
                const double* cf = coef + coff[w] + (int64_t)a * len[w];
                for (int64_t j = 0; j < len[w]; j++) {
                    GNX_REQUIRE(std::isfinite(cf[j]), "gnx_lr_model_create: non-finite coefficient (window %lld)", (long long)w);
                    f[pad_to_orig(lo[w] + j, C, ctx) - s0[w]] += cf[j];
                }
                double sa = 0.0;
                for (int64_t j = 0; j < e0[w] - s0[w]; j++) {
                    amax = std::max(amax, fabs(f[j]));
                    sa += fabs(f[j]);
                }
                ssum = std::max(ssum, sa);
            }
    }
    int s;
    if (amax == 0.0) {
        s = 8 * L - 2;
    } else {
        int e1, e2;
        frexp(amax, &e1);
        frexp(ssum, &e2);
        s = std::min(8 * L - 2 - e1, 60 - e2);
    }
    if (s_forced >= 0) s = s_forced;
    GNX_REQUIRE(s >= 8 && s <= 1000, "gnx_lr_model_create: weights out of range for fixed point (s=%d)", s);
    if (s_chosen) *s_chosen = s;
    if (out == nullptr) return 0;   // scale selection only
    const double scale = ldexp(1.0, s);

    gnx_lr* m = new gnx_lr();
    m->sub[0] = m->sub[1] = nullptr;
    m->kernel_sel = 0;
    m->tmap_w_ready = false;
    m->d_blob = nullptr;
    cudaGetDevice(&m->device);
    const int n_chunks = (int)ceil_div(C, LR_KC);
    m->h_k0.resize(W);
    m->h_kend.resize(W);
    m->h_tile_off.resize(W);
    int n_tiles = 0;
    for (int64_t w = 0; w < W; w++) {
        m->h_k0[w] = (int32_t)(s0[w] / LR_KC);
        m->h_kend[w] = (int32_t)ceil_div(e0[w], LR_KC);
        m->h_tile_off[w] = n_tiles;
        n_tiles += m->h_kend[w] - m->h_k0[w];
    }
    m->n_tiles = n_tiles;
    m->h_chunk_w0.assign(n_chunks, 0);
    m->h_chunk_wn.assign(n_chunks, 0);
    for (int64_t w = W - 1; w >= 0; w--)
        for (int k = m->h_k0[w]; k < m->h_kend[w]; k++) {
            m->h_chunk_w0[k] = (int32_t)w;
            m->h_chunk_wn[k]++;
        }

    m->h_chunk_sched.assign((size_t)n_chunks * 4, 0u);
    for (int k = 0; k < n_chunks; k++)
        for (int i = 0; i < std::min(4, m->h_chunk_wn[k]); i++) {
            const int w = m->h_chunk_w0[k] + i;
            const uint32_t tile = (uint32_t)(m->h_tile_off[w] + (k - m->h_k0[w]));
            m->h_chunk_sched[(size_t)k * 4 + i] = (tile << 3) | 4u | ((k == m->h_kend[w] - 1) ? 2u : 0u) | ((k == m->h_k0[w]) ? 1u : 0u);
        }

    // pass 2: quantise, fold (int64), split into limbs, scatter into tiles
    std::vector<int8_t> tiles((size_t)n_tiles * LR_TILE_BYTES, 0);
    {
        std::vector<long long> q(maxlen);
        for (int64_t w = 0; w < W; w++)
            for (int a = 0; a < Ar; a++) {
                const int64_t n = e0[w] - s0[w];
                std::fill(q.begin(), q.begin() + n, 0LL);
                const double* cf = coef + coff[w] + (int64_t)a * len[w];
                for (int64_t j = 0; j < len[w]; j++)
                    q[pad_to_orig(lo[w] + j, C, ctx) - s0[w]] += llrint(cf[j] * scale);
                for (int64_t j = 0; j < n; j++) {
                    const int64_t snp = s0[w] + j;
                    const int64_t tile = m->h_tile_off[w] + (snp / LR_KC - m->h_k0[w]);
                    int8_t* t = tiles.data() + tile * LR_TILE_BYTES + (snp % LR_KC);
                    long long v = q[j];
                    for (int l = 0; l < L; l++) {
                        const int8_t dgt = (int8_t)(v & 0xFF);
                        t[(size_t)(l * apad + a) * LR_KC] = dgt;
                        v = (v - dgt) >> 8;
                    }
                    if (v != 0) {
                        delete m;
                        set_error("gnx_lr_model_create: fixed-point overflow (internal)");
                        return 3;
                    }
                }
            }
    }

    // one device blob: tiles | bias | k0 | kend | tile_off | chunk_w0 | chunk_wn
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t o_tiles = 0, o_bias = al(tiles.size()), o_k0 = o_bias + al(sizeof(double) * W * Ar),
                 o_kend = o_k0 + al(4 * W), o_toff = o_kend + al(4 * W), o_cw0 = o_toff + al(4 * W),
                 o_cwn = o_cw0 + al(4 * (size_t)n_chunks), o_sch = o_cwn + al(4 * (size_t)n_chunks),
                 total = o_sch + al(16 * (size_t)n_chunks);
    char* blob = nullptr;
    cudaError_t e = cudaMalloc((void**)&blob, total);
    if (e != cudaSuccess) {
        delete m;
        set_error("gnx_lr_model_create: cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e));
        return 1;
    }
    m->d_blob = blob;
    bool ok = true;
    ok &= cudaMemcpy(blob + o_tiles, tiles.data(), tiles.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_bias, intercept, sizeof(double) * W * Ar, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_k0, m->h_k0.data(), 4 * W, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_kend, m->h_kend.data(), 4 * W, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_toff, m->h_tile_off.data(), 4 * W, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_cw0, m->h_chunk_w0.data(), 4 * (size_t)n_chunks, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_cwn, m->h_chunk_wn.data(), 4 * (size_t)n_chunks, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_sch, m->h_chunk_sched.data(), 16 * (size_t)n_chunks, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        cudaFree(blob);
        delete m;
        set_error("gnx_lr_model_create: H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    LrDev& d = m->d;
    d.A = A; d.Ar = Ar; d.L = L; d.apad = apad; d.s = s; d.W = (int)W;
    d.C = C; d.M = M; d.ctx = ctx; d.n_chunks = n_chunks; d.dbg = 0; d.raw = raw;
    d.wt = reinterpret_cast<const int8_t*>(blob + o_tiles);
    d.bias = reinterpret_cast<const double*>(blob + o_bias);
    d.k0 = reinterpret_cast<const int32_t*>(blob + o_k0);
    d.kend = reinterpret_cast<const int32_t*>(blob + o_kend);
    d.tile_off = reinterpret_cast<const int32_t*>(blob + o_toff);
    d.chunk_w0 = reinterpret_cast<const int32_t*>(blob + o_cw0);
    d.chunk_wn = reinterpret_cast<const int32_t*>(blob + o_cwn);
    d.chunk_sched = reinterpret_cast<const uint4*>(blob + o_sch);
    *out = m;
    return 0;
}

int gnx_lr_model_create(gnx_lr_t** out, int A, int64_t C, int64_t M, int64_t ctx, const double* coef,
                        const double* intercept, int limbs) {
    GNX_REQUIRE(out != nullptr, "gnx_lr_model_create: out is NULL");
    *out = nullptr;
    GNX_REQUIRE(A >= 2 && A <= 16, "gnx_lr_model_create: A=%d unsupported (2..16)", A);
    GNX_REQUIRE(C > 0 && M > 0 && M <= C && ctx >= 0 && ctx <= C, "gnx_lr_model_create: bad geometry C=%lld M=%lld ctx=%lld",
                (long long)C, (long long)M, (long long)ctx);
    GNX_REQUIRE(coef && intercept, "gnx_lr_model_create: NULL weights");
    if (require_blackwell()) return 1;
    const int Ar = (A == 2) ? 1 : A;
    const int L = limbs == 0 ? 7 : limbs;
    if (Ar <= 8 || L <= LR_NCOLS / 16) return lr_create_one(out, A, Ar, 0, -1, C, M, ctx, coef, intercept, limbs, nullptr);
    // A > 8 with more than 4 limbs: 16-column limb groups leave room for 4 limbs in the 64 accumulator columns of a
    // window, so the classes are split into two groups of <= 8, each packed with the full limb count under ONE
    // exponent, run one after the other and normalised together (lr_split_normalise_kernel)
    int s = 0;
    {
        // exponent over all classes for L limbs: the selection pass of a 16-column model with its limb clamp lifted
        const int64_t W = C / M, rem = C - M * W, M_ = M + 2 * ctx;
        double amax = 0.0, ssum = 0.0;
        std::vector<double> f;
        int64_t off = 0;
        for (int64_t w = 0; w < W; w++) {
            const int64_t lo = (w == W - 1) ? (C + 2 * ctx - (M_ + rem)) : w * M, len = (w == W - 1) ? (M_ + rem) : M_;
            const int64_t s0 = std::max<int64_t>(0, w * M - ctx), e0 = (w == W - 1) ? C : std::min<int64_t>(C, w * M + M + ctx);
            f.assign((size_t)(e0 - s0), 0.0);
            for (int a = 0; a < Ar; a++) {
                std::fill(f.begin(), f.end(), 0.0);
                const double* cf = coef + off + (int64_t)a * len;
                for (int64_t j = 0; j < len; j++) {
                    GNX_REQUIRE(std::isfinite(cf[j]), "gnx_lr_model_create: non-finite coefficient (window %lld)", (long long)w);
                    f[pad_to_orig(lo + j, C, ctx) - s0] += cf[j];
                }
                double sa = 0.0;
                for (double v : f) {
                    amax = std::max(amax, fabs(v));
                    sa += fabs(v);
                }
                ssum = std::max(ssum, sa);
            }
            off += (int64_t)Ar * len;
        }
        if (amax == 0.0) {
            s = 8 * L - 2;
        } else {
            int e1, e2;
            frexp(amax, &e1);
            frexp(ssum, &e2);
            s = std::min(8 * L - 2 - e1, 60 - e2);
        }
        GNX_REQUIRE(s >= 8 && s <= 1000, "gnx_lr_model_create: weights out of range for fixed point (s=%d)", s);
    }
    gnx_lr* m = new gnx_lr();
    m->sub[0] = m->sub[1] = nullptr;
    m->kernel_sel = 0;
    m->tmap_w_ready = false;
    m->d_blob = nullptr;
    m->n_tiles = 0;
    cudaGetDevice(&m->device);
    m->d = LrDev{};
    m->d.A = A; m->d.Ar = Ar; m->d.L = L; m->d.apad = 16; m->d.s = s; m->d.W = (int)(C / M);
    m->d.C = C; m->d.M = M; m->d.ctx = ctx;
    const int64_t W = C / M, rem = C - M * W, M_ = M + 2 * ctx;
    const int grp[3] = {0, 8, Ar};
    for (int g = 0; g < 2; g++) {
        const int a0 = grp[g], na = grp[g + 1] - grp[g];
        std::vector<double> cs, is((size_t)W * na);
        int64_t off = 0;
        for (int64_t w = 0; w < W; w++) {
            const int64_t len = (w == W - 1) ? (M_ + rem) : M_;
            cs.insert(cs.end(), coef + off + (int64_t)a0 * len, coef + off + (int64_t)(a0 + na) * len);
            for (int a = 0; a < na; a++) is[(size_t)w * na + a] = intercept[(size_t)w * Ar + a0 + a];
            off += (int64_t)Ar * len;
        }
        const int rc = lr_create_one(&m->sub[g], na, na, 1, s, C, M, ctx, cs.data(), is.data(), L, nullptr);
        if (rc) {
            gnx_lr_model_destroy(m);
            return rc;
        }
    }
    *out = m;
    return 0;
}

void gnx_lr_model_destroy(gnx_lr_t* m) {
    if (!m) return;
    for (int g = 0; g < 2; g++)
        if (m->sub[g]) gnx_lr_model_destroy(m->sub[g]);
    if (m->d_blob) cudaFree(m->d_blob);
    delete m;
}

int gnx_lr_model_scale(const gnx_lr_t* m) { return m ? m->d.s : -1; }
int gnx_lr_model_windows(const gnx_lr_t* m) { return m ? m->d.W : -1; }

int gnx_lr_set_kernel(gnx_lr_t* m, int which) {
    GNX_REQUIRE(m != nullptr, "gnx_lr_set_kernel: NULL model");
    GNX_REQUIRE(which == 0 || which == 1, "gnx_lr_set_kernel: unknown kernel %d", which);
    m->kernel_sel = which;
    for (int g = 0; g < 2; g++)
        if (m->sub[g]) m->sub[g]->kernel_sel = which;
    return 0;
}

static int lr_predict_any(const gnx_lr_t* m, const int8_t* X, int64_t N, int64_t ldX, void* B, bool f64, void* stream) {
    GNX_REQUIRE(m != nullptr, "gnx_lr_predict: NULL model");
    GNX_REQUIRE(N >= 0 && ldX >= m->d.C, "gnx_lr_predict: bad shape N=%lld ldX=%lld (C=%lld)", (long long)N, (long long)ldX, (long long)m->d.C);
    if (N == 0) return 0;
    GNX_REQUIRE(X && B, "gnx_lr_predict: NULL buffer");
    GNX_REQUIRE(ceil_div(N, 128) <= 65535, "gnx_lr_predict: N too large for one call (%lld)", (long long)N);
    cudaStream_t st = (cudaStream_t)stream;
    if (m->sub[0]) {
        // split model (A > 8): the two class groups write expit(d) as float64 [N, W, 8] / [N, W, A - 8]; one pass
        // normalises the A values of a row in numpy's summation order
        const int A = m->d.A, na0 = m->sub[0]->d.A, na1 = m->sub[1]->d.A;
        const int64_t rows = N * m->d.W;
        double* p0 = nullptr;
        GNX_CUDA(cudaMallocAsync((void**)&p0, sizeof(double) * (size_t)rows * (na0 + na1), st));
        double* p1 = p0 + (size_t)rows * na0;
        int rc = lr_predict_any(m->sub[0], X, N, ldX, p0, true, stream);
        if (!rc) rc = lr_predict_any(m->sub[1], X, N, ldX, p1, true, stream);
        if (!rc) {
            const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(rows, 256), (int64_t)sm_count() * 16);
            if (f64) lr_split_normalise_kernel<double><<<grid, 256, 0, st>>>(p0, p1, na0, na1, rows, (double*)B);
            else lr_split_normalise_kernel<float><<<grid, 256, 0, st>>>(p0, p1, na0, na1, rows, (float*)B);
            if (cudaGetLastError() != cudaSuccess) rc = 1;
            (void)A;
        }
        cudaFreeAsync(p0, st);
        if (rc == 1 && !gnx_last_error()[0]) set_error("gnx_lr_predict: split-model launch failed");
        return rc;
    }
    if (m->kernel_sel == 0) return lr_launch_tc(m, X, N, ldX, B, f64, st);
    return lr_launch_dp4a_any(m, X, N, ldX, B, f64, st);
}

int gnx_lr_predict(const gnx_lr_t* m, const int8_t* X_dev, int64_t N, int64_t ldX, float* B_dev, void* stream) {
    return lr_predict_any(m, X_dev, N, ldX, B_dev, false, stream);
}

int gnx_lr_predict_f64(const gnx_lr_t* m, const int8_t* X_dev, int64_t N, int64_t ldX, double* B_dev, void* stream) {
    return lr_predict_any(m, X_dev, N, ldX, B_dev, true, stream);
}

}  // extern "C"
