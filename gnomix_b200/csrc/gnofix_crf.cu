// gnofix_crf.cu -- K6c: Gnofix with the CRF smoother.  AN EXTENSION: the reference has no such path
// (src/model.py:194 asserts `smooth.gnofix`, which only XGB_Smoother sets; src/Gnofix/gnofix.py:157 feeds flattened
// S-window rows to `smoother.model.predict_proba`, which a chain CRF does not have), so there is NO REFERENCE ORACLE
// for it.  What is defined here is the reference's gnofix control flow
// (src/Gnofix/gnofix.py:58-208 with its default arguments, exactly as gnofix.cu follows it) with the two smoother
// plug-ins replaced by their CRF counterparts:
//   smoother.predict(B)                  -> argmax of the CRF marginals of the whole chain (K5, what CRF_Smoother.predict does)
//   smoother.model.predict_proba(scope)  -> the CRF marginal AT THE CENTRE of the S-window scope treated as a chain of its own
// The checker is oracle/np_oracle.py::gnofix_crf_extension (the oracle's restatement of the reference control flow
// with the oracle's CRF); results are bit-identical to it because every number comes out of K5 (gnx_crf_smooth),
// whose float64 marginals are bit-exact against the oracle.
//
// A CRF is global: a switch anywhere changes every marginal of the pair, so (unlike the tree smoother) re-smoothing
// is a full forward-backward pass over both haplotypes.  That makes one check + switch a few thousand sequential
// steps, and the work is organised in ROUNDS over all individuals instead of one long-running team per individual:
//   scan     warp per individual: next window with a label discontinuity (and no remembered rejection), or the
//            end-of-iteration bookkeeping (X_m seen before / max_it), exactly as gnofix.cu
//   build    the 4 candidate scopes of every individual that has a check this round -> [4k, S, A] rows
//   K5       marginals of those 4k chains of length S
//   decide   centre marginals -> accept / reject, tracker bits, rejection memo
//   gather   the current pair of every individual that switched -> [2k', W, A]; K5; labels scattered back
// A round takes up to GFC_DEPTH checks of an individual at once: the scan collects its next discontinuities, all of
// them are scored against the CURRENT pair, and the decision walks them in order -- a rejected check changes nothing,
// so the ones before the first accepted switch saw exactly the state the one-at-a-time loop would have shown them; the
// ones behind it are discarded (the scan resumes behind the switch).  Most checks are rejected, so this cuts the number
// of rounds (each of which pays the latency of a full K5 pass) about threefold.
// The pair is never swapped in memory: one tracker bit per window says which original haplotype the current "m"
// row comes from; X and B are permuted once at the end.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

struct gnx_crf;
namespace gnx {
void crf_dims(const gnx_crf* m, int* A, int* L);

struct GfcState {
    const double* B;        // [2n][W][A] original base probabilities
    int32_t* Y;             // [2n][W] labels of the current pair
    uint32_t* trk;          // [n][nw] tracker bits
    uint32_t* rej;          // [n][nw] check at w rejected, scope unchanged since
    const uint32_t* dif;    // [n][nw] original haplotypes differ in SNP block j
    uint32_t* hist;         // [n][max_it][nw] tracker at the start of every outer iteration
    int* it;                // [n] outer iterations started
    int* wcur;              // [n] next window to scan
    int* cbase;             // [n] first entry of the individual's checks of this round in cand_w / cand_i
    int* ncand;             // [n] number of them (<= GFC_DEPTH)
    unsigned char* ended;   // [n] the scan of this round reached the last window
    unsigned char* done;    // [n]
    int* act;               // compact list: individuals with checks this round
    int* cand_w;            // compact list: window of every check of this round
    int* cand_i;            // ... and its individual
    int* sw;                // compact list: individuals that switched this round
    int* counts;            // [0] |act|, [1] |sw|, [2] checks scored this round, [3] checks the one-at-a-time loop would have run
    int64_t n;
    int W, A, S, nw, max_it, memo;
};

constexpr int GFC_DEPTH = 4;   // checks of one individual scored per round

__device__ __forceinline__ uint32_t gfc_bit(const uint32_t* v, int j) { return (v[j >> 5] >> (j & 31)) & 1u; }

__global__ void gfc_scan_kernel(GfcState g) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= g.n || g.done[i]) return;
    const int W = g.W, nw = g.nw;
    const int32_t* y0 = g.Y + (2 * i) * W;
    const int32_t* y1 = y0 + W;
    uint32_t* trk = g.trk + i * nw;
    const uint32_t* rej = g.rej + i * nw;
    const uint32_t* dif = g.dif + i * nw;
    uint32_t* hist = g.hist + i * (int64_t)g.max_it * nw;
    int wcur = g.wcur[i], it = g.it[i];
    for (;;) {
        // the next (up to GFC_DEPTH) windows with a discontinuity in either haplotype (gnofix.py:116-118) and no
        // remembered rejection
        int cw[GFC_DEPTH];
        int c = 0;
        bool ended = true;
        for (int base = wcur; base < W; base += 32) {
            const int ww = base + lane;
            const bool hit = ww < W && (y0[ww] != y0[ww - 1] || y1[ww] != y1[ww - 1]) && !gfc_bit(rej, ww);
            unsigned b = __ballot_sync(0xffffffffu, hit);
            while (b && c < GFC_DEPTH) {
                cw[c++] = base + __ffs(b) - 1;
                b &= b - 1;
            }
            if (c == GFC_DEPTH) {
                ended = false;   // (windows may remain behind the last one taken)
                break;
            }
        }
        if (c > 0) {
            if (lane == 0) {
                const int cb = atomicAdd(&g.counts[2], c);
                for (int j = 0; j < c; j++) {
                    g.cand_w[cb + j] = cw[j];
                    g.cand_i[cb + j] = (int)i;
                }
                g.cbase[i] = cb;
                g.ncand[i] = c;
                g.ended[i] = ended ? 1 : 0;
                g.it[i] = it;
                g.act[atomicAdd(&g.counts[0], 1)] = (int)i;
            }
            return;
        }
        // the scan of this outer iteration is over: stop at max_it or when X_m was seen before (gnofix.py:104-113)
        it++;
        bool stop = it >= g.max_it;
        for (int p = 0; p < it && !stop; p++) {
            bool same = true;
            for (int j = lane; j < nw; j += 32) same &= (((hist[(int64_t)p * nw + j] ^ trk[j]) & dif[j]) == 0u);
            if (__all_sync(0xffffffffu, same)) stop = true;
        }
        if (stop) {
            if (lane == 0) {
                g.done[i] = 1;
                g.ncand[i] = 0;
                g.it[i] = it;
            }
            return;
        }
        for (int j = lane; j < nw; j += 32) hist[(int64_t)it * nw + j] = trk[j];
        __syncwarp();
        wcur = 1;
    }
}

// rows[(4k + r)][s][:]: r = 0, 1 the pair as it is, r = 2, 3 with the tails swapped at w (gnofix.py:140-157)
__global__ void gfc_build_kernel(GfcState g, double* __restrict__ rows, int ncand) {
    const int S = g.S, A = g.A, W = g.W;
    const int64_t total = (int64_t)ncand * 4 * S;
    const int half = (S - 1) / 2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx / (4 * S)), rem = (int)(idx - (int64_t)k * 4 * S);
        const int r = rem / S, s = rem - r * S;
        const int64_t i = g.cand_i[k];
        const int w = g.cand_w[k];
        const int center = min(max(w, half), W - S + half);
        const int j = center - half + s;
        const int h = r & 1;
        const int hh = (r >= 2 && j >= w) ? 1 - h : h;
        const int src = hh ^ (int)gfc_bit(g.trk + i * g.nw, j);
        const double* from = g.B + ((2 * i + src) * (int64_t)W + j) * A;
        double* to = rows + idx * A;
        for (int a = 0; a < A; a++) to[a] = from[a];
    }
}

// one thread per individual with checks this round: its checks in scan order, up to the first accepted switch
__global__ void gfc_decide_kernel(GfcState g, const double* __restrict__ marg, int L, int nact) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nact) return;
    const int S = g.S, W = g.W, nw = g.nw;
    const int64_t i = g.act[k];
    const int c = (S - 1) / 2;
    uint32_t* trk = g.trk + i * nw;
    uint32_t* rej = g.rej + i * nw;
    const int cb = g.cbase[i], nc = g.ncand[i];
    for (int j = 0; j < nc; j++) {
        const int w = g.cand_w[cb + j];
        double p[4];
        for (int r = 0; r < 4; r++) {
            const double* mrow = marg + (((int64_t)4 * (cb + j) + r) * S + c) * L;
            double best = mrow[0];
            for (int y = 1; y < L; y++) best = (mrow[y] > best) ? mrow[y] : best;
            p[r] = best;
        }
        const double p_orig = (p[1] > p[0]) ? p[1] : p[0], p_sw = (p[3] > p[2]) ? p[3] : p[2];
        if (!(p_sw * 0.5 > p_orig * 0.5)) {   // gnofix.py:171, prior_switch_prob = 0.5
            if (g.memo) rej[w >> 5] |= 1u << (w & 31);
            continue;
        }
        // accepted: the checks behind it were scored against a pair that no longer exists -- the scan resumes at w + 1
        for (int ww = max(1, w - S); ww <= min(W - 1, w + S); ww++) rej[ww >> 5] &= ~(1u << (ww & 31));
        for (int q = w >> 5; q < nw; q++) {
            uint32_t mask = 0xffffffffu;
            if (q == (w >> 5)) mask <<= (w & 31);
            if (q == nw - 1 && (W & 31)) mask &= (1u << (W & 31)) - 1u;
            trk[q] ^= mask;
        }
        g.wcur[i] = w + 1;
        g.sw[atomicAdd(&g.counts[1], 1)] = (int)i;
        atomicAdd(&g.counts[3], j + 1);
        return;
    }
    // every check of the round rejected: resume behind the last one (or end the outer iteration)
    g.wcur[i] = g.ended[i] ? W : g.cand_w[cb + nc - 1] + 1;
    atomicAdd(&g.counts[3], nc);
}

// the current pair of switched individuals sw[off .. off + cnt): pairs[(2k + h)][j][:] = B[2i + (h ^ trk[j])][j][:]
__global__ void gfc_gather_kernel(GfcState g, double* __restrict__ pairs, int off, int cnt) {
    const int W = g.W, A = g.A;
    const int64_t total = (int64_t)cnt * 2 * W;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx / (2 * W)), rem = (int)(idx - (int64_t)k * 2 * W);
        const int h = rem / W, j = rem - h * W;
        const int64_t i = g.sw[off + k];
        const int src = h ^ (int)gfc_bit(g.trk + i * g.nw, j);
        const double* from = g.B + ((2 * i + src) * (int64_t)W + j) * A;
        double* to = pairs + idx * A;
        for (int a = 0; a < A; a++) to[a] = from[a];
    }
}

__global__ void gfc_scatter_kernel(GfcState g, const int32_t* __restrict__ lab, int off, int cnt) {
    const int W = g.W;
    const int64_t total = (int64_t)cnt * 2 * W;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx / (2 * W)), rem = (int)(idx - (int64_t)k * 2 * W);
        g.Y[(2 * (int64_t)g.sw[off + k]) * W + rem] = lab[idx];
    }
}

__global__ void gfc_init_kernel(GfcState g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    g.it[i] = 0;
    g.wcur[i] = 1;
    g.ncand[i] = 0;
    g.ended[i] = 0;
    g.done[i] = 0;
}

// diff bit j of individual i: rows 2i and 2i+1 of X differ somewhere in SNP block j (as gnofix.cu)
__global__ void gfc_diff_kernel(const int8_t* __restrict__ X, int64_t ldX, int64_t C, int W, int nw, int64_t ws, uint32_t* __restrict__ diff) {
    const int64_t ind = blockIdx.x;
    const int8_t* a = X + (2 * ind) * ldX;
    const int8_t* b = a + ldX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < W; j += nwarps) {
        const int64_t s = (int64_t)j * ws, e = (j == W - 1) ? C : s + ws;
        bool d = false;
        for (int64_t p = s + lane; p < e; p += 32) d |= (a[p] != b[p]);
        const unsigned any = __ballot_sync(0xffffffffu, d);
        if (lane == 0 && any) atomicOr(diff + ind * nw + (j >> 5), 1u << (j & 31));
    }
}

// final permutation (X tails by SNP block, float64 B by window) and the tracker rows
__global__ void gfc_apply_kernel(int8_t* __restrict__ X, int64_t ldX, int64_t C, double* __restrict__ B, int W, int A, int nw, int64_t ws,
                                 const uint32_t* __restrict__ trk, int32_t* __restrict__ tracker) {
    const int64_t ind = blockIdx.x;
    const uint32_t* t = trk + ind * nw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < W; j += nwarps) {
        const int bit = (int)gfc_bit(t, j);
        if (tracker && lane == 0) {
            tracker[(2 * ind) * W + j] = bit;
            tracker[(2 * ind + 1) * W + j] = 1 - bit;
        }
        if (!bit) continue;
        if (X) {
            int8_t* a = X + (2 * ind) * ldX;
            int8_t* b = a + ldX;
            const int64_t s = (int64_t)j * ws, e = (j == W - 1) ? C : s + ws;
            for (int64_t p = s + lane; p < e; p += 32) {
                const int8_t v = a[p];
                a[p] = b[p];
                b[p] = v;
            }
        }
        double* b0 = B + ((2 * ind) * W + j) * A;
        double* b1 = B + ((2 * ind + 1) * W + j) * A;
        if (lane < A) {
            const double v = b0[lane];
            b0[lane] = b1[lane];
            b1[lane] = v;
        }
    }
}

static int64_t g_gfc_stats[4] = {0, 0, 0, 0};   // rounds, checks, accepted switches, outer iterations (last call)

}  // namespace gnx

using namespace gnx;

extern "C" int gnx_gnofix_crf(const gnx_crf_t* m, int S, int8_t* X_dev, int64_t ldX, int64_t C, double* B_dev, int64_t n_ind, int W,
                              int max_it, int32_t* Y_dev, int32_t* tracker_dev, void* stream) {
    GNX_REQUIRE(m != nullptr, "gnx_gnofix_crf: NULL CRF model");
    int A = 0, L = 0;
    crf_dims(m, &A, &L);
    GNX_REQUIRE(A == L, "gnx_gnofix_crf: the CRF has %d attributes and %d labels; the smoother's CRF has one of each per ancestry", A, L);
    GNX_REQUIRE(S >= 1 && (S & 1), "gnx_gnofix_crf: smoother width S=%d must be odd", S);
    GNX_REQUIRE(n_ind >= 0 && W >= 2 && W >= S && C >= W && ldX >= C, "gnx_gnofix_crf: bad shape n_ind=%lld W=%d S=%d C=%lld ldX=%lld",
                (long long)n_ind, W, S, (long long)C, (long long)ldX);
    GNX_REQUIRE(max_it >= 1 && max_it <= 64, "gnx_gnofix_crf: max_it=%d outside 1..64", max_it);
    if (n_ind == 0) return 0;
    GNX_REQUIRE(B_dev && Y_dev, "gnx_gnofix_crf: NULL buffer (X_dev may be NULL to skip the SNP-level swap)");
    if (require_blackwell()) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    const int nw = (W + 31) / 32;
    const int64_t ws = C / W;
    const int64_t cap = std::min<int64_t>(n_ind, 4096);   // individuals re-smoothed per K5 call (bounds the scratch)

    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t b_bits = al((size_t)n_ind * nw * 4), b_hist = al((size_t)n_ind * max_it * nw * 4), b_int = al((size_t)n_ind * 4);
    const size_t b_cand = al((size_t)n_ind * GFC_DEPTH * 4), b_byte = al((size_t)n_ind);
    const size_t b_rows = al((size_t)n_ind * GFC_DEPTH * 4 * S * A * 8), b_pairs = al((size_t)cap * 2 * W * A * 8), b_lab = al((size_t)cap * 2 * W * 4);
    const size_t total = 3 * b_bits + b_hist + 6 * b_int + 2 * b_cand + 2 * b_byte + 256 + 2 * b_rows + 2 * b_pairs + b_lab;
    char* scratch = nullptr;
    GNX_CUDA(cudaMallocAsync((void**)&scratch, total, st));
    char* q = scratch;
    auto take = [&](size_t b) { char* r = q; q += b; return r; };
    GfcState g{};
    g.B = B_dev;
    g.Y = Y_dev;
    g.trk = reinterpret_cast<uint32_t*>(take(b_bits));
    g.rej = reinterpret_cast<uint32_t*>(take(b_bits));
    uint32_t* dif = reinterpret_cast<uint32_t*>(take(b_bits));
    g.dif = dif;
    g.hist = reinterpret_cast<uint32_t*>(take(b_hist));
    g.it = reinterpret_cast<int*>(take(b_int));
    g.wcur = reinterpret_cast<int*>(take(b_int));
    g.cbase = reinterpret_cast<int*>(take(b_int));
    g.ncand = reinterpret_cast<int*>(take(b_int));
    g.act = reinterpret_cast<int*>(take(b_int));
    g.sw = reinterpret_cast<int*>(take(b_int));
    g.cand_w = reinterpret_cast<int*>(take(b_cand));
    g.cand_i = reinterpret_cast<int*>(take(b_cand));
    g.ended = reinterpret_cast<unsigned char*>(take(b_byte));
    g.done = reinterpret_cast<unsigned char*>(take(b_byte));
    g.counts = reinterpret_cast<int*>(take(256));
    double* rows = reinterpret_cast<double*>(take(b_rows));
    double* marg_rows = reinterpret_cast<double*>(take(b_rows));
    double* pairs = reinterpret_cast<double*>(take(b_pairs));
    double* marg_pairs = reinterpret_cast<double*>(take(b_pairs));
    int32_t* lab_pairs = reinterpret_cast<int32_t*>(take(b_lab));
    g.n = n_ind;
    g.W = W; g.A = A; g.S = S; g.nw = nw; g.max_it = max_it;
    const char* em = getenv("GNX_GNOFIX_MEMO");
    g.memo = (em && em[0] == '0') ? 0 : 1;

    int rc = 0;
    int* h_counts = nullptr;
    auto fail = [&](int code) {
        if (h_counts) cudaFreeHost(h_counts);
        cudaFreeAsync(scratch, st);
        return code;
    };
    if (cudaMallocHost((void**)&h_counts, 4 * sizeof(int)) != cudaSuccess) {
        set_error("gnx_gnofix_crf: pinned allocation failed");
        return fail(1);
    }
#define GFC_CUDA(call)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));         \
            return fail(1);                                                                          \
        }                                                                                            \
    } while (0)
    // state: tracker 0, no rejections, hist[0] = tracker (zeros), difference bits
    GFC_CUDA(cudaMemsetAsync(g.trk, 0, 2 * b_bits, st));
    GFC_CUDA(cudaMemsetAsync(g.hist, 0, b_hist, st));
    if (X_dev) {
        GFC_CUDA(cudaMemsetAsync(dif, 0, b_bits, st));
        gfc_diff_kernel<<<(unsigned)n_ind, 256, 0, st>>>(X_dev, ldX, C, W, nw, ws, dif);
    } else {
        GFC_CUDA(cudaMemsetAsync(dif, 0xff, b_bits, st));   // no X: every block counts as differing (tracker equality)
    }
    gfc_init_kernel<<<(unsigned)ceil_div(n_ind, 256), 256, 0, st>>>(g);
    GFC_CUDA(cudaGetLastError());
    // initial labels of the pair as it is: smoother.predict(B) (gnofix.py:80)
    for (int64_t off = 0; off < n_ind; off += cap) {
        const int64_t cnt = std::min(cap, n_ind - off);
        rc = gnx_crf_smooth(m, B_dev + off * 2 * W * A, 2 * cnt, W, marg_pairs, Y_dev + off * 2 * W, stream);
        if (rc) return fail(rc);
    }
    int64_t rounds = 0, checks = 0, accepts = 0;
    for (;;) {
        GFC_CUDA(cudaMemsetAsync(g.counts, 0, 4 * sizeof(int), st));
        gfc_scan_kernel<<<(unsigned)ceil_div(n_ind * 32, 256), 256, 0, st>>>(g);
        GFC_CUDA(cudaGetLastError());
        GFC_CUDA(cudaMemcpyAsync(h_counts, g.counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        GFC_CUDA(cudaStreamSynchronize(st));
        const int nact = h_counts[0], ncand = h_counts[2];
        if (nact == 0) break;   // every individual is done
        rounds++;
        gfc_build_kernel<<<(unsigned)std::min<int64_t>(ceil_div((int64_t)ncand * 4 * S, 256), (int64_t)sm_count() * 8), 256, 0, st>>>(g, rows, ncand);
        GFC_CUDA(cudaGetLastError());
        rc = gnx_crf_smooth(m, rows, (int64_t)4 * ncand, S, marg_rows, nullptr, stream);
        if (rc) return fail(rc);
        gfc_decide_kernel<<<(unsigned)ceil_div(nact, 128), 128, 0, st>>>(g, marg_rows, L, nact);
        GFC_CUDA(cudaGetLastError());
        GFC_CUDA(cudaMemcpyAsync(h_counts, g.counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        GFC_CUDA(cudaStreamSynchronize(st));
        const int nsw = h_counts[1];
        checks += h_counts[3];
        accepts += nsw;
        // full re-smoothing of the pairs that switched (Smoother.predict, gnofix.py:190)
        for (int off = 0; off < nsw; off += (int)cap) {
            const int cnt = (int)std::min<int64_t>(cap, nsw - off);
            const unsigned grid = (unsigned)std::min<int64_t>(ceil_div((int64_t)cnt * 2 * W, 256), (int64_t)sm_count() * 8);
            gfc_gather_kernel<<<grid, 256, 0, st>>>(g, pairs, off, cnt);
            GFC_CUDA(cudaGetLastError());
            rc = gnx_crf_smooth(m, pairs, (int64_t)2 * cnt, W, marg_pairs, lab_pairs, stream);
            if (rc) return fail(rc);
            gfc_scatter_kernel<<<grid, 256, 0, st>>>(g, lab_pairs, off, cnt);
            GFC_CUDA(cudaGetLastError());
        }
    }
    gfc_apply_kernel<<<(unsigned)n_ind, 256, 0, st>>>(X_dev, ldX, C, B_dev, W, A, nw, ws, g.trk, tracker_dev);
    GFC_CUDA(cudaGetLastError());
    // outer iterations (for the stats): summed on the host after the run
    {
        std::vector<int> its((size_t)n_ind);
        GFC_CUDA(cudaMemcpyAsync(its.data(), g.it, (size_t)n_ind * sizeof(int), cudaMemcpyDeviceToHost, st));
        GFC_CUDA(cudaStreamSynchronize(st));
        int64_t s = 0;
        for (int v : its) s += v;
        g_gfc_stats[3] = s;
    }
    g_gfc_stats[0] = rounds;
    g_gfc_stats[1] = checks;
    g_gfc_stats[2] = accepts;
#undef GFC_CUDA
    cudaFreeHost(h_counts);
    GNX_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

/* counters of the last gnx_gnofix_crf call of this process: rounds, candidate checks, accepted switches, outer iterations */
extern "C" int gnx_gnofix_crf_last_stats(int64_t* out4) {
    GNX_REQUIRE(out4 != nullptr, "gnx_gnofix_crf_last_stats: NULL");
    for (int i = 0; i < 4; i++) out4[i] = gnx::g_gfc_stats[i];
    return 0;
}
