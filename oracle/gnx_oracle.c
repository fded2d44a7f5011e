/*
 * gnx_oracle.c -- CPU restatement (plain C + OpenMP) of the gnomix inference hot
 * path.  TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs; never by the product.
 *
 * Every function cites the reference file:line (under /root/reference) it follows;
 * oracle/np_oracle.py holds the readable NumPy twin of each function and the tests
 * check the two against each other and against tests/golden/ (vectors produced by
 * the reference's own Python, see oracle/make_golden.py).
 *
 * Build: oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../include/gnx_math.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_A 64

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

double orc_exp(double x) { return gnx_exp(x); }
float orc_expf_cr(float x) { return gnx_expf_cr(x); }

/* ---- Row A: Base.pad index map (src/Base/base.py:41-44) ------------------ */
static inline int64_t pad_to_orig(int64_t p, int64_t C, int64_t ctx) {
    if (p < ctx) return ctx - 1 - p;
    if (p >= ctx + C) return 2 * C + ctx - 1 - p;
    return p - ctx;
}

/* sklearn _predict_proba_lr epilogue for one (haplotype, window):
 * expit, then prob /= prob.sum(axis=1); binary: [1-p, p]. */
static inline void lr_epilogue(const double* d, int A, double* out) {
    if (A == 2) {
        double p = gnx_expit(d[0]);
        out[0] = 1.0 - p;
        out[1] = p;
        return;
    }
    double p[ORC_MAX_A];
    for (int a = 0; a < A; a++) p[a] = gnx_expit(d[a]);
    double s = gnx_np_sum(p, A);
    for (int a = 0; a < A; a++) out[a] = p[a] / s;
}

/* ---- Row B: LR base, float64 restatement of Base.predict_proba_vectorized
 * (src/Base/base.py:146-180) with sklearn LR windows.  coef: windows concatenated,
 * window w is [A_rows, Mw] row-major at coef_off[w]; intercept [W, A_rows].
 * A_rows = A (A>2) or 1 (A==2).  Output Bout float64 [N, W, A]. */
void orc_lr_f64(const int8_t* X, int64_t N, int64_t ldX, int64_t C, int64_t M, int64_t ctx, int A,
                const double* coef, const int64_t* coef_off, const double* intercept, double* Bout) {
    int64_t W = C / M, rem = C - M * W, M_ = M + 2 * ctx;
    int Ar = (A == 2) ? 1 : A;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t n = 0; n < N; n++) {
        const int8_t* x = X + n * ldX;
        for (int64_t w = 0; w < W; w++) {
            int64_t lo = (w == W - 1) ? (C + 2 * ctx - (M_ + rem)) : w * M;
            int64_t len = (w == W - 1) ? (M_ + rem) : M_;
            const double* cf = coef + coef_off[w];
            double d[ORC_MAX_A];
            for (int a = 0; a < Ar; a++) {
                double acc = 0.0;
                for (int64_t j = 0; j < len; j++)
                    acc += (double)x[pad_to_orig(lo + j, C, ctx)] * cf[a * len + j];
                d[a] = acc + intercept[w * Ar + a];
            }
            lr_epilogue(d, A, Bout + (n * W + w) * A);
        }
    }
}

/* ---- Row B, fixed-point form: exact integer dot with folded int64 weights
 * (np_oracle.lr_quantize_fold); qf: windows concatenated, window w is
 * [A_rows, e_w-s_w] at q_off[w].  Writes float32 (Bf) and/or float64 (Bd). */
void orc_lr_fixed(const int8_t* X, int64_t N, int64_t ldX, int64_t C, int64_t M, int64_t ctx, int A,
                  const int64_t* qf, const int64_t* q_off, int s, const double* intercept,
                  float* Bf, double* Bd) {
    int64_t W = C / M;
    int Ar = (A == 2) ? 1 : A;
    double scale = gnx_pow2i(-s);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t n = 0; n < N; n++) {
        const int8_t* x = X + n * ldX;
        for (int64_t w = 0; w < W; w++) {
            int64_t s0 = w * M - ctx;
            if (s0 < 0) s0 = 0;
            int64_t e0 = (w == W - 1) ? C : w * M + M + ctx;
            if (e0 > C) e0 = C;
            int64_t len = e0 - s0;
            const int64_t* q = qf + q_off[w];
            double d[ORC_MAX_A], out[ORC_MAX_A];
            for (int a = 0; a < Ar; a++) {
                int64_t acc = 0;
                for (int64_t j = 0; j < len; j++) acc += (int64_t)x[s0 + j] * q[a * len + j];
                d[a] = (double)acc * scale + intercept[w * Ar + a];
            }
            lr_epilogue(d, A, out);
            for (int a = 0; a < A; a++) {
                if (Bf) Bf[(n * W + w) * A + a] = (float)out[a];
                if (Bd) Bd[(n * W + w) * A + a] = out[a];
            }
        }
    }
}

/* ---- Row D: slide_window (src/Smooth/utils.py:4-29) index map -------------
 * padded window j of a W-long row -> original window. pad=(S+1)/2 */
static inline int64_t spad_to_orig(int64_t j, int64_t W, int64_t pad) {
    if (j < pad) return pad - 1 - j;
    if (j >= pad + W) return W - 1 - (j - pad - W);
    return j - pad;
}

void orc_slide_window(const float* B, int64_t N, int64_t W, int A, int S, float* out) {
    int64_t pad = (S + 1) / 2;
#pragma omp parallel for
    for (int64_t n = 0; n < N; n++)
        for (int64_t w = 0; w < W; w++)
            for (int64_t s = 0; s < S; s++) {
                int64_t o = spad_to_orig(w + s, W, pad);
                for (int a = 0; a < A; a++)
                    out[((n * W + w) * S + s) * A + a] = B[(n * W + o) * A + a];
            }
}

/* ---- Row E: xgboost multi:softprob predictor ------------------------------
 * (src/Smooth/models.py:14-20; xgboost==1.1.1 is absent from the reference tree:
 * restated from its published CPU predictor + common/math.h Softmax). */
typedef struct {
    int A, n_trees;
    const int32_t *feat, *left, *right, *tree_offsets;
    const float *thr, *leaf, *base_margin;
    const uint8_t* default_left;
} orc_gbt_t;

static inline void gbt_row(const orc_gbt_t* m, const float* row, int64_t stride, float* proba) {
    float psum[ORC_MAX_A];
    int A = m->A;
    for (int c = 0; c < A; c++) psum[c] = 0.0f;
    for (int t = 0; t < m->n_trees; t++) {
        int o = m->tree_offsets[t];
        int nid = 0;
        while (m->feat[o + nid] >= 0) {
            float x = row[(int64_t)m->feat[o + nid] * stride];
            int go_left = (x != x) ? (m->default_left[o + nid] != 0) : (x < m->thr[o + nid]);
            nid = go_left ? m->left[o + nid] : m->right[o + nid];
        }
        psum[t % A] = psum[t % A] + m->leaf[o + nid];
    }
    float mg[ORC_MAX_A];
    for (int c = 0; c < A; c++) mg[c] = m->base_margin[c] + psum[c];
    float wmax = mg[0];
    for (int c = 1; c < A; c++) wmax = fmaxf(mg[c], wmax);
    double wsum = 0.0;
    for (int c = 0; c < A; c++) {
        mg[c] = gnx_expf_cr(mg[c] - wmax);
        wsum += mg[c];
    }
    float ws = (float)wsum;
    for (int c = 0; c < A; c++) proba[c] = mg[c] / ws;
}

void orc_gbt_rows(int A, int n_trees, const int32_t* feat, const float* thr, const int32_t* left,
                  const int32_t* right, const uint8_t* default_left, const float* leaf,
                  const int32_t* tree_offsets, const float* base_margin, const float* rows,
                  int64_t k, int64_t F, float* proba) {
    orc_gbt_t m = {A, n_trees, feat, left, right, tree_offsets, thr, leaf, base_margin, default_left};
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < k; r++) gbt_row(&m, rows + r * F, 1, proba + r * A);
}

/* Smoother.predict_proba + predict for the XGB smoother without materialising
 * X_slide: row (n,w) = Bpad[n, w:w+S, :] (src/Smooth/utils.py:14-23);
 * label = first argmax (src/Smooth/smooth.py:61). */
void orc_gbt_smooth(int A, int S, int n_trees, const int32_t* feat, const float* thr,
                    const int32_t* left, const int32_t* right, const uint8_t* default_left,
                    const float* leaf, const int32_t* tree_offsets, const float* base_margin,
                    const float* B, int64_t N, int64_t W, float* proba, int32_t* label) {
    orc_gbt_t m = {A, n_trees, feat, left, right, tree_offsets, thr, leaf, base_margin, default_left};
    int64_t pad = (S + 1) / 2;
#pragma omp parallel
    {
        float* bp = (float*)malloc(sizeof(float) * (size_t)(W + 2 * pad) * A);
#pragma omp for schedule(dynamic, 1)
        for (int64_t n = 0; n < N; n++) {
            for (int64_t j = 0; j < W + 2 * pad; j++)
                memcpy(bp + j * A, B + (n * W + spad_to_orig(j, W, pad)) * A, sizeof(float) * A);
            for (int64_t w = 0; w < W; w++) {
                float pr[ORC_MAX_A];
                gbt_row(&m, bp + w * A, 1, pr);
                int best = 0;
                for (int c = 0; c < A; c++) {
                    if (proba) proba[(n * W + w) * A + c] = pr[c];
                    if (pr[c] > pr[best]) best = c;
                }
                if (label) label[n * W + w] = best;
            }
        }
        free(bp);
    }
}

/* ---- Row F: CRF marginals (src/Smooth/crf.py:62-67 -> CRFsuite crf1d_context.c
 * crf1dc_alpha_score / crf1dc_beta_score / crf1dc_marginal_point; library absent
 * from the reference tree, restated from the published algorithm).
 * B float64 [N, W, A]; state_w [A, L]; trans_w [L, L]; exp() is gnx_exp so the
 * kernel can match bit for bit. */
void orc_crf_smooth(const double* B, int64_t N, int64_t W, int A, int L, const double* state_w,
                    const double* trans_w, double* proba, int32_t* label) {
    double et[ORC_MAX_A * ORC_MAX_A];
    for (int i = 0; i < L * L; i++) et[i] = gnx_exp(trans_w[i]);
#pragma omp parallel
    {
        double* es = (double*)malloc(sizeof(double) * W * L);
        double* al = (double*)malloc(sizeof(double) * W * L);
        double* sc = (double*)malloc(sizeof(double) * W);
#pragma omp for schedule(dynamic, 1)
        for (int64_t n = 0; n < N; n++) {
            const double* b = B + n * W * A;
            for (int64_t t = 0; t < W; t++)
                for (int y = 0; y < L; y++) {
                    double st = 0.0;
                    for (int a = 0; a < A; a++) st += b[t * A + a] * state_w[a * L + y];
                    es[t * L + y] = gnx_exp(st);
                }
            double cur[ORC_MAX_A], row[ORC_MAX_A], bt[ORC_MAX_A];
            for (int64_t t = 0; t < W; t++) {
                if (t == 0) {
                    for (int y = 0; y < L; y++) cur[y] = es[y];
                } else {
                    for (int y = 0; y < L; y++) cur[y] = 0.0;
                    for (int i = 0; i < L; i++)
                        for (int y = 0; y < L; y++) cur[y] += al[(t - 1) * L + i] * et[i * L + y];
                    for (int y = 0; y < L; y++) cur[y] *= es[t * L + y];
                }
                double sum = 0.0;
                for (int y = 0; y < L; y++) sum += cur[y];
                sc[t] = (sum != 0.0) ? 1.0 / sum : 1.0;
                for (int y = 0; y < L; y++) al[t * L + y] = cur[y] * sc[t];
            }
            for (int y = 0; y < L; y++) bt[y] = sc[W - 1];
            for (int64_t t = W - 1; t >= 0; t--) {
                if (t < W - 1) {
                    for (int y = 0; y < L; y++) row[y] = bt[y] * es[(t + 1) * L + y];
                    for (int i = 0; i < L; i++) {
                        double acc = 0.0;
                        for (int y = 0; y < L; y++) acc += et[i * L + y] * row[y];
                        cur[i] = acc;
                    }
                    for (int y = 0; y < L; y++) bt[y] = cur[y] * sc[t];
                }
                int best = 0;
                double pb = 0.0;
                for (int y = 0; y < L; y++) {
                    double p = al[t * L + y] * bt[y] / sc[t];
                    if (proba) proba[(n * W + t) * L + y] = p;
                    if (y == 0 || p > pb) { pb = p; best = y; }
                }
                if (label) label[n * W + t] = best;
            }
        }
        free(es);
        free(al);
        free(sc);
    }
}

/* ---- Row C: CovRSK string kernel (src/Base/string_kernel.py:91-101) --------
 * Direct DP restatement: tri = current match-run length, cov_tri = #{m in Ms:
 * m <= tri}, K += cov_tri.  ms_ohe[t] = 1 iff t in Ms (length Mlen+1). */
void orc_covrsk(const int8_t* X, int64_t nx, const int8_t* Y, int64_t ny, int64_t Mlen,
                const uint8_t* ms_ohe, int64_t* K) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < nx; i++)
        for (int64_t j = 0; j < ny; j++) {
            const int8_t *x = X + i * Mlen, *y = Y + j * Mlen;
            int64_t tri = 0, cov = 0, k = 0;
            for (int64_t p = 0; p < Mlen; p++) {
                if (x[p] == y[p]) {
                    tri++;
                    cov += ms_ohe[tri];
                    k += cov;
                } else {
                    tri = 0;
                    cov = 0;
                }
            }
            K[i * ny + j] = k;
        }
}

/* libsvm sigmoid_predict + multiclass_probability (svm.cpp), as called by
 * sklearn.svm.SVC(probability=True).predict_proba (src/Base/models.py:213-215).
 * K [n, nSV] kernel values against the support vectors grouped by class. */
static void multiclass_probability(int k, const double* r /*[k,k]*/, double* p) {
    int t, j, iter = 0, max_iter = k > 100 ? k : 100;
    double Q[ORC_MAX_A * ORC_MAX_A], Qp[ORC_MAX_A], pQp, eps = 0.005 / k;
    for (t = 0; t < k; t++) {
        p[t] = 1.0 / k;
        Q[t * k + t] = 0;
        for (j = 0; j < t; j++) {
            Q[t * k + t] += r[j * k + t] * r[j * k + t];
            Q[t * k + j] = Q[j * k + t];
        }
        for (j = t + 1; j < k; j++) {
            Q[t * k + t] += r[j * k + t] * r[j * k + t];
            Q[t * k + j] = -r[j * k + t] * r[t * k + j];
        }
    }
    for (iter = 0; iter < max_iter; iter++) {
        pQp = 0;
        for (t = 0; t < k; t++) {
            Qp[t] = 0;
            for (j = 0; j < k; j++) Qp[t] += Q[t * k + j] * p[j];
            pQp += p[t] * Qp[t];
        }
        double max_error = 0;
        for (t = 0; t < k; t++) {
            double error = fabs(Qp[t] - pQp);
            if (error > max_error) max_error = error;
        }
        if (max_error < eps) break;
        for (t = 0; t < k; t++) {
            double diff = (-Qp[t] + pQp) / Q[t * k + t];
            p[t] += diff;
            pQp = (pQp + diff * (diff * Q[t * k + t] + 2 * Qp[t])) / (1 + diff) / (1 + diff);
            for (j = 0; j < k; j++) {
                Qp[j] = (Qp[j] + diff * Q[t * k + j]) / (1 + diff);
                p[j] /= (1 + diff);
            }
        }
    }
}

void orc_svc_proba(const int64_t* K, int64_t n, int64_t nSV, int k, const int32_t* n_support,
                   const double* dual_coef /*[k-1,nSV]*/, const double* intercept,
                   const double* probA, const double* probB, double* out) {
    int start[ORC_MAX_A];
    start[0] = 0;
    for (int i = 1; i < k; i++) start[i] = start[i - 1] + n_support[i - 1];
#pragma omp parallel for
    for (int64_t r = 0; r < n; r++) {
        const int64_t* kv = K + r * nSV;
        double pair[ORC_MAX_A * ORC_MAX_A];
        int p = 0;
        for (int i = 0; i < k; i++)
            for (int j = i + 1; j < k; j++) {
                double sum = 0;
                for (int t = 0; t < n_support[i]; t++)
                    sum += dual_coef[(int64_t)(j - 1) * nSV + start[i] + t] * (double)kv[start[i] + t];
                for (int t = 0; t < n_support[j]; t++)
                    sum += dual_coef[(int64_t)i * nSV + start[j] + t] * (double)kv[start[j] + t];
                double dec = sum + intercept[p];
                double fApB = dec * probA[p] + probB[p], v;
                if (fApB >= 0)
                    v = gnx_exp(-fApB) / (1.0 + gnx_exp(-fApB));
                else
                    v = 1.0 / (1 + gnx_exp(fApB));
                if (v < 1e-7) v = 1e-7;
                if (v > 1 - 1e-7) v = 1 - 1e-7;
                pair[i * k + j] = v;
                pair[j * k + i] = 1 - v;
                p++;
            }
        if (k == 2) {
            out[r * 2] = pair[1];
            out[r * 2 + 1] = pair[2];
        } else
            multiclass_probability(k, pair, out + r * k);
    }
}
