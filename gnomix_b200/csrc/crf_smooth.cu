// crf_smooth.cu -- K5: CRF_Smoother.predict_proba / predict.
//
// Replaces src/Smooth/crf.py:17-67 (npy2crf dict marshalling -> sklearn_crfsuite
// predict_marginals -> crf2npy) behind src/Smooth/smooth.py:40-65.  The arithmetic is
// CRFsuite's crf1d scaled forward-backward (crf1d_context.c: alpha_score, beta_score,
// marginal_point) with every attribute/label pair as a state feature and every label
// pair as a transition, no BOS/EOS terms -- restated in oracle/gnx_oracle.c
// (orc_crf_smooth), whose operation order this kernel follows bit for bit (explicit
// round-to-nearest mul/add, no contraction, gnx_exp for exp).
//
// One thread walks one haplotype: the chain is strictly sequential in W, the work per
// step is 2*L*L multiply-adds in float64, and haplotypes are independent.  The forward
// pass parks the unnormalised alpha row in the output buffer; the backward pass re-derives
// the scale factor from it (same sum, same order) and overwrites it with the marginal.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace gnx {

constexpr int CRF_MAX = 16;

struct CrfDev {
    int A, L;
    const double* state_w;  // [A][L]
    const double* exp_trans;  // [L][L] = exp(trans_w), i -> j
};

template <int AT, int LT>
__global__ void __launch_bounds__(64)
crf_smooth_kernel(CrfDev m, const double* __restrict__ B, int64_t N, int W, double* __restrict__ proba,
                  int32_t* __restrict__ label) {
    __shared__ double s_sw[CRF_MAX * CRF_MAX];
    __shared__ double s_et[CRF_MAX * CRF_MAX];
    const int A = AT ? AT : m.A, L = LT ? LT : m.L;
    constexpr int AM = AT ? AT : CRF_MAX, LM = LT ? LT : CRF_MAX;
    for (int i = threadIdx.x; i < A * L; i += blockDim.x) s_sw[i] = m.state_w[i];
    for (int i = threadIdx.x; i < L * L; i += blockDim.x) s_et[i] = m.exp_trans[i];
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double* b = B + n * (int64_t)W * A;
    double* out = proba + n * (int64_t)W * L;

    auto exp_state = [&](int t, double* es) {
        double bv[AM];
#pragma unroll
        for (int a = 0; a < AM; a++) bv[a] = (a < A) ? __ldg(b + (int64_t)t * A + a) : 0.0;
#pragma unroll
        for (int y = 0; y < LM; y++) {
            if (y < L) {
                double st = 0.0;
#pragma unroll
                for (int a = 0; a < AM; a++)
                    if (a < A) st = GNX_ADD(st, GNX_MUL(bv[a], s_sw[a * L + y]));
                es[y] = gnx_exp(st);
            }
        }
    };
    auto row_scale = [&](const double* cur) {
        double sum = 0.0;
#pragma unroll
        for (int y = 0; y < LM; y++)
            if (y < L) sum = GNX_ADD(sum, cur[y]);
        return (sum != 0.0) ? GNX_DIV(1.0, sum) : 1.0;
    };

    // ---- forward: alpha_t = (alpha_{t-1} . exp(T)) (*) exp(state_t), scaled to sum 1
    double al[LM], es[LM], cur[LM];
    for (int t = 0; t < W; t++) {
        exp_state(t, es);
        if (t == 0) {
#pragma unroll
            for (int y = 0; y < LM; y++)
                if (y < L) cur[y] = es[y];
        } else {
#pragma unroll
            for (int y = 0; y < LM; y++) cur[y] = 0.0;
#pragma unroll
            for (int i = 0; i < LM; i++)
                if (i < L) {
#pragma unroll
                    for (int y = 0; y < LM; y++)
                        if (y < L) cur[y] = GNX_ADD(cur[y], GNX_MUL(al[i], s_et[i * L + y]));
                }
#pragma unroll
            for (int y = 0; y < LM; y++)
                if (y < L) cur[y] = GNX_MUL(cur[y], es[y]);
        }
        const double sc = row_scale(cur);
#pragma unroll
        for (int y = 0; y < LM; y++)
            if (y < L) {
                al[y] = GNX_MUL(cur[y], sc);
                out[(int64_t)t * L + y] = cur[y];
            }
    }

    // ---- backward + marginals
    double bt[LM], es_next[LM];
    for (int t = W - 1; t >= 0; t--) {
#pragma unroll
        for (int y = 0; y < LM; y++)
            if (y < L) cur[y] = out[(int64_t)t * L + y];
        const double sc = row_scale(cur);
        if (t == W - 1) {
#pragma unroll
            for (int y = 0; y < LM; y++) bt[y] = sc;
        } else {
            double row[LM], nb[LM];
#pragma unroll
            for (int y = 0; y < LM; y++)
                if (y < L) row[y] = GNX_MUL(bt[y], es_next[y]);
#pragma unroll
            for (int i = 0; i < LM; i++)
                if (i < L) {
                    double acc = 0.0;
#pragma unroll
                    for (int y = 0; y < LM; y++)
                        if (y < L) acc = GNX_ADD(acc, GNX_MUL(s_et[i * L + y], row[y]));
                    nb[i] = acc;
                }
#pragma unroll
            for (int y = 0; y < LM; y++)
                if (y < L) bt[y] = GNX_MUL(nb[y], sc);
        }
        if (t > 0) exp_state(t, es_next);
        int best = 0;
        double pb = 0.0;
#pragma unroll
        for (int y = 0; y < LM; y++)
            if (y < L) {
                const double p = GNX_DIV(GNX_MUL(GNX_MUL(cur[y], sc), bt[y]), sc);
                out[(int64_t)t * L + y] = p;
                if (y == 0 || p > pb) {
                    pb = p;
                    best = y;
                }
            }
        if (label) label[n * W + t] = best;
    }
}

// Lane-parallel variant (L, A <= 8): 8 lanes per haplotype, lane y owns label y.  Every output element is
// produced by exactly the operation sequence of the kernel above (sums run over i / y / a in the same
// order, operands fetched from the owning lanes with shuffles), so results are bit-identical; the chain
// per step shrinks from L*L to L multiply-adds and one exp, and 8x more threads hide the latency.
// Global memory is touched in bursts of CRF_TB steps per haplotype (staged in shared memory): a haplotype's
// rows are contiguous in t, but one 56-byte row per haplotype per step across 20 000 concurrent haplotypes
// is a DRAM row miss every time (the first version of this kernel ran no faster than thread-per-haplotype).
constexpr int CRF_TB = 16;
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__global__ void __launch_bounds__(128)
crf_smooth_lanes_kernel(CrfDev m, const double* __restrict__ B, int64_t N, int W, double* __restrict__ proba,
                        int32_t* __restrict__ label) {
    __shared__ double s_b[16][CRF_TB * 8];     // base probabilities of the chunk, [tt][a] dense (stride A)
    __shared__ double s_o[16][CRF_TB * 8];     // alpha rows / marginals of the chunk, [tt][y] dense (stride L)
    __shared__ int32_t s_l[16][CRF_TB];
    const int A = m.A, L = m.L;
    const int lane = threadIdx.x & 31, y = lane & 7, g0 = lane & ~7, grp = threadIdx.x >> 3;
    const int64_t n_raw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool live = n_raw < N;
    const int64_t n = live ? n_raw : N - 1;          // idle groups shadow the last haplotype, stores masked
    const int yy = (y < L) ? y : L - 1;               // idle lanes shadow the last label
    const double* b = B + n * (int64_t)W * A;
    double* out = proba + n * (int64_t)W * L;
    double* sb = s_b[grp];
    double* so = s_o[grp];
    int32_t* sl = s_l[grp];
    double sw[8], etc[8], etr[8];                     // state_w[:, y], exp_trans[:, y], exp_trans[y, :]
#pragma unroll
    for (int k = 0; k < 8; k++) {
        sw[k] = (k < A) ? m.state_w[k * L + yy] : 0.0;
        etc[k] = (k < L) ? m.exp_trans[k * L + yy] : 0.0;
        etr[k] = (k < L) ? m.exp_trans[yy * L + k] : 0.0;
    }
    auto exp_state = [&](int tt) {
        double st = 0.0;
#pragma unroll
        for (int a = 0; a < 8; a++)
            if (a < A) st = GNX_ADD(st, GNX_MUL(sb[tt * A + a], sw[a]));
        return gnx_exp(st);
    };
    auto row_scale = [&](double mine) {              // 1 / sum_y cur[y], summed in label order
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double v = shfl_d(mine, g0 + k);
            if (k < L) sum = GNX_ADD(sum, v);
        }
        return (sum != 0.0) ? GNX_DIV(1.0, sum) : 1.0;
    };

    // ---- forward
    double al = 0.0;
    for (int t0 = 0; t0 < W; t0 += CRF_TB) {
        const int nt = min(CRF_TB, W - t0);
        for (int i = y; i < nt * A; i += 8) sb[i] = __ldg(b + (int64_t)t0 * A + i);
        __syncwarp();
        for (int tt = 0; tt < nt; tt++) {
            const double es = exp_state(tt);
            double cur;
            if (t0 + tt == 0) {
                cur = es;
            } else {
                cur = 0.0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const double ai = shfl_d(al, g0 + i);
                    if (i < L) cur = GNX_ADD(cur, GNX_MUL(ai, etc[i]));
                }
                cur = GNX_MUL(cur, es);
            }
            const double sc = row_scale(cur);
            al = GNX_MUL(cur, sc);
            if (y < L) so[tt * L + y] = cur;
        }
        __syncwarp();
        if (live)
            for (int i = y; i < nt * L; i += 8) out[(int64_t)t0 * L + i] = so[i];
        __syncwarp();
    }

    // ---- backward + marginals (chunks in reverse; each thread re-reads only what its own group wrote)
    double bt = 0.0, es_next = 0.0;
    const int last0 = ((W - 1) / CRF_TB) * CRF_TB;
    for (int t0 = last0; t0 >= 0; t0 -= CRF_TB) {
        const int nt = min(CRF_TB, W - t0);
        for (int i = y; i < nt * A; i += 8) sb[i] = __ldg(b + (int64_t)t0 * A + i);
        // (idle groups shadow the last haplotype: they do not read rows its own group may still be writing)
        for (int i = y; i < nt * L; i += 8) so[i] = live ? out[(int64_t)t0 * L + i] : 1.0;
        __syncwarp();
        for (int tt = nt - 1; tt >= 0; tt--) {
            const int t = t0 + tt;
            const double cur = so[tt * L + yy];
            const double sc = row_scale(cur);
            if (t == W - 1) {
                bt = sc;
            } else {
                const double row = GNX_MUL(bt, es_next);
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const double rk = shfl_d(row, g0 + k);
                    if (k < L) acc = GNX_ADD(acc, GNX_MUL(etr[k], rk));
                }
                bt = GNX_MUL(acc, sc);
            }
            if (t > 0) es_next = exp_state(tt);
            const double p = GNX_DIV(GNX_MUL(GNX_MUL(cur, sc), bt), sc);
            __syncwarp();                                // every lane has read so[tt] before it is overwritten
            if (y < L) so[tt * L + y] = p;
            int best = 0;
            double pb = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double pk = shfl_d(p, g0 + k);
                if (k < L && (k == 0 || pk > pb)) {
                    pb = pk;
                    best = k;
                }
            }
            if (y == 0) sl[tt] = best;
        }
        __syncwarp();
        if (live) {
            for (int i = y; i < nt * L; i += 8) out[(int64_t)t0 * L + i] = so[i];
            if (label)
                for (int i = y; i < nt; i += 8) label[n * W + t0 + i] = sl[i];
        }
        __syncwarp();
    }
}

}  // namespace gnx

struct gnx_crf {
    gnx::CrfDev d;
    int device;
    void* d_blob;
};

namespace gnx {
void crf_dims(const gnx_crf* m, int* A, int* L) { *A = m->d.A; *L = m->d.L; }
}  // namespace gnx

using namespace gnx;

extern "C" {

int gnx_crf_model_create(gnx_crf_t** out, int A, int L, const double* state_w, const double* trans_w) {
    GNX_REQUIRE(out != nullptr, "gnx_crf_model_create: out is NULL");
    *out = nullptr;
    GNX_REQUIRE(A >= 1 && A <= CRF_MAX && L >= 2 && L <= CRF_MAX, "gnx_crf_model_create: A=%d L=%d unsupported (<= %d)", A, L, CRF_MAX);
    GNX_REQUIRE(state_w && trans_w, "gnx_crf_model_create: NULL weights");
    if (require_blackwell()) return 1;
    std::vector<double> et((size_t)L * L);
    for (int i = 0; i < L * L; i++) {
        GNX_REQUIRE(isfinite(trans_w[i]), "gnx_crf_model_create: non-finite transition weight");
        et[i] = gnx_exp(trans_w[i]);
    }
    for (int i = 0; i < A * L; i++) GNX_REQUIRE(isfinite(state_w[i]), "gnx_crf_model_create: non-finite state weight");
    char* blob = nullptr;
    const size_t sb = sizeof(double) * A * L, tb = sizeof(double) * L * L;
    GNX_CUDA(cudaMalloc((void**)&blob, sb + tb));
    bool ok = cudaMemcpy(blob, state_w, sb, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + sb, et.data(), tb, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        cudaFree(blob);
        set_error("gnx_crf_model_create: H2D copy failed");
        return 1;
    }
    gnx_crf* m = new gnx_crf();
    cudaGetDevice(&m->device);
    m->d_blob = blob;
    m->d = CrfDev{A, L, reinterpret_cast<const double*>(blob), reinterpret_cast<const double*>(blob + sb)};
    *out = m;
    return 0;
}

void gnx_crf_model_destroy(gnx_crf_t* m) {
    if (!m) return;
    if (m->d_blob) cudaFree(m->d_blob);
    delete m;
}

int gnx_crf_smooth(const gnx_crf_t* m, const double* B_dev, int64_t N, int W, double* proba_dev, int32_t* label_dev,
                   void* stream) {
    GNX_REQUIRE(m != nullptr, "gnx_crf_smooth: NULL model");
    GNX_REQUIRE(N >= 0 && W >= 1, "gnx_crf_smooth: bad shape N=%lld W=%d", (long long)N, W);
    if (N == 0) return 0;
    GNX_REQUIRE(B_dev && (proba_dev || label_dev), "gnx_crf_smooth: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    double* work = proba_dev;
    if (!work) GNX_CUDA(cudaMallocAsync((void**)&work, sizeof(double) * (size_t)N * W * m->d.L, st));
    // lane-parallel kernel for the reference's sizes (A = L <= 8); GNX_CRF_KERNEL=0 forces the
    // thread-per-haplotype kernel (cross-check: same bits)
    const char* ek = getenv("GNX_CRF_KERNEL");
    if (m->d.A <= 8 && m->d.L <= 8 && !(ek && ek[0] == '0')) {
        crf_smooth_lanes_kernel<<<(int)ceil_div(N * 8, 128), 128, 0, st>>>(m->d, B_dev, N, W, work, label_dev);
        cudaError_t e = cudaGetLastError();
        if (!proba_dev) cudaFreeAsync(work, st);
        GNX_CUDA(e);
        return 0;
    }
    const int grid = (int)ceil_div(N, 64);
    if (m->d.A == 7 && m->d.L == 7)
        crf_smooth_kernel<7, 7><<<grid, 64, 0, st>>>(m->d, B_dev, N, W, work, label_dev);
    else
        crf_smooth_kernel<0, 0><<<grid, 64, 0, st>>>(m->d, B_dev, N, W, work, label_dev);
    cudaError_t e = cudaGetLastError();
    if (!proba_dev) cudaFreeAsync(work, st);
    GNX_CUDA(e);
    return 0;
}

}  // extern "C"
