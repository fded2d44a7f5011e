/*
 * gnx.h -- C ABI of libgnx.so, the B200 (sm_100a) implementation of the gnomix
 * local-ancestry inference hot path:
 *
 *     Base.predict_proba  ->  Smoother.predict_proba / predict  ->  [gnofix]
 *
 * Each entry point replaces one call site of the reference (file:line under
 * AI-sandbox/gnomix @ 43bec05).  The reference is pure Python and has no FFI of its
 * own; INTEGRATION.md shows the ctypes stub a gnomix maintainer would add at each
 * site.  Conventions:
 *   - every function returns 0 on success, non-zero on failure; gnx_last_error()
 *     returns a thread-local message.  Nothing throws, nothing prints.
 *   - `*_dev` pointers are device pointers on the CURRENT cuda device and are owned
 *     by the caller; the library owns only the opaque model handles (which live on
 *     the device that was current when they were created).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Device-pointer entry points only enqueue work; they do not synchronise.
 *   - There is NO CPU fallback: without a usable sm_100 device every compute entry
 *     point fails with an error.
 *   - Haplotype matrix X: int8, row-major [N, ldX] with ldX >= C, values {0,1,2}
 *     (2 = missing, src/utils.py:150-153).  Rows 2i, 2i+1 are individual i.
 *   - Probability tensors: row-major [N, W, A].
 */
#ifndef GNX_H_
#define GNX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNX_VERSION 120 /* 0.2.0: tile smoother kernel, whole-pipeline host entry points */

typedef struct gnx_lr gnx_lr_t;   /* per-window logistic-regression base (K1) */
typedef struct gnx_gbt gnx_gbt_t; /* gradient-boosted-tree smoother (K4)      */
typedef struct gnx_crf gnx_crf_t; /* linear-chain CRF smoother (K5)           */
typedef struct gnx_svc gnx_svc_t; /* CovRSK string-kernel SVC base (K2+K3)    */
typedef struct gnx_cal gnx_cal_t; /* per-class isotonic calibrator (K7)        */
typedef struct gnx_vcf gnx_vcf_t; /* parsed VCF (host side)                    */

int gnx_version(void);
const char* gnx_last_error(void);
/* number of visible CUDA devices with compute capability 10.x; 0 if none */
int gnx_device_count(void);
/* Device self-test of the exact math helpers of gnx_math.h (division, int64->double,
 * double<->float, exp range reduction) against the builtin IEEE operations over n random
 * operand sets; mismatches[0..4] (host) receive the number of differing results. */
int gnx_selftest_math(int64_t n, uint64_t seed, int64_t* mismatches);

/* ---------------------------------------------------------------------------
 * K1  LogisticRegressionBase.predict_proba for all W windows
 * replaces: src/Base/base.py:146-180 (Base.predict_proba_vectorized) with the
 *           per-window models of src/Base/models.py:12-21 (sklearn
 *           LogisticRegression.predict_proba -> _predict_proba_lr).
 * coef:      windows concatenated; window w is row-major [A_rows, M_w] float64 with
 *            M_w = M+2*ctx (w < W-1) or M+2*ctx+rem (w = W-1), feature order = the
 *            reference's padded window (reflect pad included), A_rows = A (A > 2)
 *            or 1 (A == 2, sklearn binary layout).
 * intercept: [W, A_rows] float64.
 * limbs:     signed base-256 digits per fixed-point weight (0 = default 7).  A > 8 with more
 *            than 4 limbs is packed as two class groups of <= 8 under one exponent, run one
 *            after the other (X is read twice) and normalised together -- the same exact
 *            logits as for A <= 8; limbs <= 4 keeps the single-pass 16-column model
 *            (|logit error| ~1e-6).  Windows up to 131000 SNPs.  The
 *            fixed-point scale is chosen so that the int64 totals cannot overflow for
 *            |x| <= 2 (the reference's matrix only holds {0,1,2}).
 * Environment GNX_LR_DBG (bit mask, profiling only): 1 skip MMAs, 2 skip the
 * epilogue, 4 skip weight loads -- results are garbage when set.
 * W is implied: W = C / M.  Output B float32 (or float64) [N, W, A].
 * ------------------------------------------------------------------------- */
int gnx_lr_model_create(gnx_lr_t** out, int A, int64_t C, int64_t M, int64_t ctx,
                        const double* coef, const double* intercept, int limbs);
void gnx_lr_model_destroy(gnx_lr_t* m);
/* fixed-point exponent s chosen at create time: q = rint(w * 2^s) */
int gnx_lr_model_scale(const gnx_lr_t* m);
int gnx_lr_model_windows(const gnx_lr_t* m);
/* kernel selector: 0 = tcgen05/TMA/TMEM tensor-core kernel (default),
 *                  1 = dp4a CUDA-core kernel (cross-check; same exact integer math) */
int gnx_lr_set_kernel(gnx_lr_t* m, int which);
int gnx_lr_predict(const gnx_lr_t* m, const int8_t* X_dev, int64_t N, int64_t ldX,
                   float* B_dev, void* stream);
int gnx_lr_predict_f64(const gnx_lr_t* m, const int8_t* X_dev, int64_t N, int64_t ldX,
                       double* B_dev, void* stream);

/* ---------------------------------------------------------------------------
 * K4  XGB_Smoother: slide_window + XGBClassifier.predict_proba + argmax
 * replaces: src/Smooth/utils.py:4-29 (slide_window), src/Smooth/smooth.py:40-65
 *           (Smoother.predict_proba / predict) with the model of
 *           src/Smooth/models.py:14-20 (xgboost multi:softprob).
 * Forest layout (xgboost dump order): node arrays concatenated over trees, tree t
 * owns nodes [tree_offsets[t], tree_offsets[t+1]); left/right are node indices
 * relative to the tree's first node; feat < 0 marks a leaf whose value is leaf[];
 * split test `x[feat] < thr` -> left, NaN -> default_left; tree t adds to class
 * t % A; margin_c = base_margin[c] + sum (float32, tree order); softmax float32.
 * Feature f of row (n,w) is Bpad[n, w + f / A, f % A], Bpad = reflect pad (S+1)/2.
 * ------------------------------------------------------------------------- */
int gnx_gbt_model_create(gnx_gbt_t** out, int A, int S, int n_trees, const int32_t* feat,
                         const float* thr, const int32_t* left, const int32_t* right,
                         const uint8_t* default_left, const float* leaf,
                         const int32_t* tree_offsets, const float* base_margin);
void gnx_gbt_model_destroy(gnx_gbt_t* m);
/* kernel selector (same results from every one of them):
 *   0  = default: rank-form kernels for depth <= 4 forests -- the tile kernel (lanes = 32 haplotypes, warps =
 *        windows, conflict-free lane-interleaved rank tiles behind a u16 rank pass) for batches of >= 24
 *        haplotypes, the row kernel (lanes = windows of one haplotype) below that; the generic kernel otherwise;
 *   1  = generic float traversal (cross-check; any depth <= 8);
 *   10 = row kernel, one-word nodes and index arithmetic (the walk gnofix uses);
 *   14 = row kernel, block layout with one accumulating byte offset per walk;
 *   16 = tile kernel whatever the batch size (falls back to the row kernel when the shape does not fit).
 * Environment GNX_GBT_VARIANT=0|4|6 picks 10|14|16 at model-create time (profiling). */
int gnx_gbt_set_kernel(gnx_gbt_t* m, int which);
/* Profiling: with `on`, the tile path of gnx_gbt_smooth records CUDA events around its two launches (K4a rank pass,
 * K4b tile walk) on the caller's stream; gnx_gbt_last_phase_ms waits for the last call and returns their durations. */
int gnx_gbt_set_profile(gnx_gbt_t* m, int on);
int gnx_gbt_last_phase_ms(const gnx_gbt_t* m, float* rank_ms, float* walk_ms);
/* proba_dev [N,W,A] float32 and label_dev [N,W] int32; either may be NULL */
int gnx_gbt_smooth(const gnx_gbt_t* m, const float* B_dev, int64_t N, int W, float* proba_dev,
                   int32_t* label_dev, void* stream);
/* smoother.model.predict_proba(rows[k, S*A]) as gnofix calls it
 * (src/Gnofix/gnofix.py:157): rows are already-flattened scopes, no sliding. */
int gnx_gbt_rows(const gnx_gbt_t* m, const float* rows_dev, int64_t k, float* proba_dev,
                 void* stream);

/* ---------------------------------------------------------------------------
 * K5  CRF_Smoother.predict_proba (CRF.predict_marginals) + argmax
 * replaces: src/Smooth/crf.py:62-67 via src/Smooth/smooth.py:40-65.
 * state_w [A, L] (attribute a -> label y), trans_w [L, L] (i -> j), float64.
 * ------------------------------------------------------------------------- */
int gnx_crf_model_create(gnx_crf_t** out, int A, int L, const double* state_w,
                         const double* trans_w);
void gnx_crf_model_destroy(gnx_crf_t* m);
int gnx_crf_smooth(const gnx_crf_t* m, const double* B_dev, int64_t N, int W, double* proba_dev,
                   int32_t* label_dev, void* stream);

/* ---------------------------------------------------------------------------
 * K2+K3  CovRSKBase.predict_proba: string kernel against the training matrix of
 * every window + libsvm probability epilogue.
 * replaces: src/Base/string_kernel.py:91-123 and sklearn.svm.SVC.predict_proba as
 *           configured at src/Base/models.py:195-215, driven by base.py:146-180.
 * sv:        support vectors of all windows concatenated; window w holds
 *            nsv_total[w] rows of M_w int8 features (padded-window feature order),
 *            grouped by class with n_support[w*A + c] rows for class c.
 * dual_coef: per window [A-1, nsv_total[w]] float64 (SVC._dual_coef_), concatenated.
 * intercept, probA, probB: [W, A*(A-1)/2] float64 (SVC._intercept_, probA_, probB_).
 * Ms:        CovSample(M_w, 0.6, 1.0, seed=37) prefix-stable list, length n_ms.
 * ------------------------------------------------------------------------- */
int gnx_svc_model_create(gnx_svc_t** out, int A, int64_t C, int64_t M, int64_t ctx,
                         const int8_t* sv, const int32_t* n_support, const double* dual_coef,
                         const double* intercept, const double* probA, const double* probB,
                         const int32_t* Ms, int n_ms);
void gnx_svc_model_destroy(gnx_svc_t* m);
int gnx_svc_predict(const gnx_svc_t* m, const int8_t* X_dev, int64_t N, int64_t ldX,
                    double* B_dev, void* stream);
/* raw kernel values of window w against its support vectors: K [N, nsv_total[w]] int32 */
int gnx_svc_kernel_window(const gnx_svc_t* m, int w, const int8_t* X_dev, int64_t N, int64_t ldX,
                          int32_t* K_dev, void* stream);

/* ---------------------------------------------------------------------------
 * K6  Gnomix.phase / gnofix with the reference's default arguments
 * replaces: src/model.py:188-214 and src/Gnofix/gnofix.py:58-208 (+ phasing.py:
 *           182-198) for all individuals in one launch.
 * X_dev [2n, ldX] int8 (nullable: skips the SNP-level swap) and B_dev [2n, W, A]
 * float32 are updated IN PLACE (tails swapped); Y_dev [2n, W] int32 receives the
 * final labels; tracker_dev [2n, W] int32 (nullable) the gnofix_tracker rows.
 * Needs W >= 2*S (as XGB_Smoother asserts), max_it <= 64, a forest of depth <= 4
 * with <= 65535 distinct thresholds, and B without NaN: a pair whose base
 * probabilities hold a NaN is REFUSED -- its labels come back as -1, X / B / tracker
 * untouched -- because the rank form of this kernel cannot follow the default child
 * of a node the way the smoother kernels do (the Python plugin raises on it).
 * ------------------------------------------------------------------------- */
int gnx_gnofix(const gnx_gbt_t* m, int8_t* X_dev, int64_t ldX, int64_t C, float* B_dev,
               int64_t n_ind, int W, int max_it, int32_t* Y_dev, int32_t* tracker_dev,
               void* stream);

/* profiling counters of the last gnx_gnofix call (synchronises): outer iterations, scans,
 * candidate checks, accepted switches, summed over individuals.  Cross-check knobs (same
 * results): GNX_GNOFIX_MEMO=0 disables the rejected-check memo, GNX_GNOFIX_SPLIT=0 re-smooths
 * with one thread per row instead of (row, class) chain tasks, GNX_GNOFIX_TEAMS=1..4 bounds the
 * teams per SM. */
int gnx_gnofix_last_stats(int64_t* out4);

/* ---------------------------------------------------------------------------
 * K6c  Gnofix with the CRF smoother -- AN EXTENSION WITHOUT A REFERENCE ORACLE.
 * The reference refuses this combination (src/model.py:194 asserts `smooth.gnofix`, set by
 * XGB_Smoother only; src/Gnofix/gnofix.py:157 calls `smoother.model.predict_proba` on flattened
 * S-window rows, which a chain CRF does not have).  Defined as the
 * reference's gnofix control flow (src/Gnofix/gnofix.py:58-208, default arguments) with
 *   smoother.predict(B)                 := argmax of the CRF marginals of the whole chain,
 *   smoother.model.predict_proba(scope) := the CRF marginal at the centre of the S-window scope
 *                                          run as a chain of its own.
 * X_dev [2n, ldX] int8 (nullable) and B_dev [2n, W, A] FLOAT64 (what the CRF smoother reads) are
 * updated in place; Y_dev [2n, W] int32 the final labels; tracker_dev [2n, W] int32 nullable.
 * S odd, S <= W, max_it <= 64.  Checked against oracle/np_oracle.py::gnofix_crf_extension.
 * ------------------------------------------------------------------------- */
int gnx_gnofix_crf(const gnx_crf_t* m, int S, int8_t* X_dev, int64_t ldX, int64_t C, double* B_dev,
                   int64_t n_ind, int W, int max_it, int32_t* Y_dev, int32_t* tracker_dev,
                   void* stream);
/* counters of the last gnx_gnofix_crf call: rounds, candidate checks, accepted switches, outer iterations */
int gnx_gnofix_crf_last_stats(int64_t* out4);

/* ---------------------------------------------------------------------------
 * K7  Calibrator.transform on smoother probabilities (+ argmax)
 * replaces: src/Smooth/Calibration.py:57-69 (per-class IsotonicRegression(out_of_bounds=
 *           "clip").transform) and the normalisation of lines 24-39, as called from
 *           src/Smooth/smooth.py:48-52.
 * n_thr [A] thresholds per class; x_thr / y_thr the classes' X_thresholds_ / y_thresholds_
 * concatenated (as float64); is_f32 = 1 if the fitted thresholds are float32 (a model
 * fitted on the XGB smoother's float32 probabilities: scikit-learn then evaluates in
 * float32), 0 for float64 (CRF smoother).  proba_dev [rows, A] float32 (in_is_f32) or
 * float64; out_dev [rows, A] float64 and label_dev [rows] int32 (first maximum), either
 * may be NULL.
 * ------------------------------------------------------------------------- */
int gnx_cal_model_create(gnx_cal_t** out, int A, int is_f32, const int32_t* n_thr,
                         const double* x_thr, const double* y_thr);
void gnx_cal_model_destroy(gnx_cal_t* m);
int gnx_calibrate(const gnx_cal_t* m, const void* proba_dev, int in_is_f32, int64_t rows,
                  double* out_dev, int32_t* label_dev, void* stream);

/* ---------------------------------------------------------------------------
 * Host-buffer pipeline: Gnomix.predict_proba / predict on a numpy-style host
 * matrix (src/model.py:169-179, gnomix.py:55-58).  Streams haplotype chunks
 * through two device slots (H2D, K1, K4, D2H overlapped on two streams); part of
 * every chunk crosses PCIe as 2-bit planes packed by the host cores (see below).
 * X_host may be pageable or pinned.  proba_host [N,W,A] float32 may be NULL.
 * Synchronous: returns when label_host / proba_host are complete.
 * Environment: GNX_HOST_PACK=0 (no packing), GNX_HOST_PACK_FRAC=f (packed fraction of
 * each chunk instead of the calibrated one), GNX_HOST_THREADS=n.
 * ------------------------------------------------------------------------- */
int gnx_infer_host(const gnx_lr_t* lr, const gnx_gbt_t* gbt, const int8_t* X_host, int64_t N,
                   int64_t ldX, float* proba_host, int32_t* label_host, int64_t chunk_haps);

/* The same host-buffer pipeline for every plugin combination of run_inference (gnomix.py:48-72):
 *   base      exactly one of lr (LogisticRegressionBase) / svc (CovRSKBase);
 *   smoother  exactly one of gbt (XGB_Smoother) / crf (CRF_Smoother);
 *   cal       optional Calibrator applied to the smoother's probabilities (src/Smooth/smooth.py:48-52);
 *   phase     != 0: Gnomix.phase (Gnofix, tree smoother only, N even): label_host receives gnofix's labels,
 *             X_phased_host (int8 [N, C], may be NULL) the phased haplotypes, and proba_host -- when asked for --
 *             the probabilities of the phased haplotypes run through the model again (gnomix.py:60-72);
 *             max_it <= 0 means the reference's default 50;
 *   x_packed  != 0: X_host holds rows already packed to 2-bit planes (the gnx_pack_rows_host layout, as
 *             gnx_vcf_to_haplotypes_packed writes them) and ldX is the row pitch in 64-bit words: a quarter of
 *             the bytes cross PCIe and no host core packs anything.
 * Dtypes follow the reference's hand-offs: the CRF reads the base's float64 probabilities, the tree smoother
 * float32 (a float64 string-kernel base output is rounded to float32, as slide_window's float32 matrix does);
 * proba_host is float32 [N,W,A] for the tree smoother without calibrator, float64 otherwise. */
typedef struct gnx_pipeline {
    const gnx_lr_t* lr;
    const gnx_svc_t* svc;
    const gnx_gbt_t* gbt;
    const gnx_crf_t* crf;
    const gnx_cal_t* cal;
    int phase;
    int max_it;
    int x_packed;
    int crf_phase_S;   /* > 0 with phase and crf: the CRF + Gnofix EXTENSION (gnx_gnofix_crf, no reference behaviour) with
                          this smoother width; 0 keeps the reference's refusal (src/model.py:194) */
} gnx_pipeline_t;
int gnx_infer_host_ex(const gnx_pipeline_t* p, const void* X_host, int64_t N, int64_t ldX, void* proba_host,
                      int32_t* label_host, int8_t* X_phased_host, int64_t chunk_haps);

/* ---------------------------------------------------------------------------
 * Packed transfer of the haplotype matrix (the reference's int8 matrix,
 * src/utils.py:104-159, carries 2 bits per byte): gnx_infer_host packs on the host
 * cores while the previous chunk is in flight and unpacks on the device, so a quarter
 * of the bytes cross PCIe (GNX_HOST_PACK=0 disables it).  The two halves are exported
 * for callers that stage their own transfers.
 * Packed row: group g (SNPs 64g..64g+63) = { u64 plane0, u64 plane1 }, bit i of
 * planeK = bit K of X[row][64g+i]; pitch_words >= 2*ceil(C/64) 64-bit words per row.
 * gnx_pack_rows_host: pure host code (no device needed); *out_of_range = 1 if some
 * value was outside 0..3 (packed form unusable).  threads <= 0: gnx_host_threads().
 * gnx_unpack_dev: X_dev [n, ldX] int8, ldX % 16 == 0, columns [0, min(ldX, 32*pitch_words))
 * are written (SNPs >= C as 0).
 * ------------------------------------------------------------------------- */
int gnx_pack_rows_host(const int8_t* X_host, int64_t n, int64_t ldX, int64_t C, uint64_t* packed_host,
                       int64_t pitch_words, int threads, int* out_of_range);
int gnx_unpack_dev(const uint64_t* packed_dev, int64_t n, int64_t pitch_words, int64_t C,
                   int8_t* X_dev, int64_t ldX, void* stream);
/* Host int8 matrix [N, ldX] -> device int8 matrix [N, ld_dev] (ld_dev a multiple of 128 >= C,
 * the row pitch the base kernels address) through the same packed transfer; synchronous.
 * This is how the Python plugins upload a numpy haplotype matrix (Base.predict_proba, phase). */
int gnx_upload_haplotypes(const int8_t* X_host, int64_t N, int64_t ldX, int64_t C, int8_t* X_dev,
                          int64_t ld_dev);
/* frees the device slots and pinned staging buffers the two calls above keep between calls */
int gnx_release_workspace(void);
/* rates measured by gnx_infer_host's one-off calibration (0 before it ran): host pack
 * rate in GB/s of int8 input, pinned H2D rate in GB/s */
int gnx_infer_host_rates(double* pack_gbs, double* h2d_gbs);
/* what the last gnx_infer_host call moved: packed fraction of each chunk's rows, bytes
 * copied host->device and device->host */
int gnx_infer_host_last_transfer(double* frac, int64_t* h2d_bytes, int64_t* d2h_bytes);
/* ---------------------------------------------------------------------------
 * Host-side output of run_inference: body of the .fb file
 * replaces: the per-window formatting loop of src/postprocess.py:100-126 (write_fb).
 * Appends (append != 0) or writes W lines to `path`: line l = prefixes[l] followed by the
 * tab-separated str(proba[n, l, a]) (n outer, a inner; numpy's shortest round-trip float
 * formatting, float32 or float64 per is_f64) and a newline.  proba row-major [N, W, A] host
 * memory.  Lines are formatted on `threads` host threads (<= 0: gnx_host_threads()).
 * gnx_format_floats is the formatter alone (newline-separated; returns the length or -1).
 * ------------------------------------------------------------------------- */
int gnx_write_fb_body(const char* path, int append, const void* proba, int is_f64, int64_t N,
                      int64_t W, int64_t A, const char* const* prefixes, int threads);
int64_t gnx_format_floats(const void* values, int is_f64, int64_t n, char* out, int64_t cap);
/* Records of the phased VCF (replaces the per-record loop of npy_to_vcf, src/utils.py:247-329):
 * record j = CHROM POS ID REF ALT QUAL PASS . GT then "a|b" per sample with one character
 * '0' + value per haplotype.  String columns are blobs of n_rec newline-terminated fields,
 * hap is int8 [n_hap][ld] host memory (column j = record j, rows 2i / 2i+1 = sample i). */
int gnx_write_vcf_body(const char* path, int append, int64_t n_rec, int64_t n_hap, const int8_t* hap,
                       int64_t ld, const int64_t* pos, const char* chrom, int64_t chrom_len,
                       const char* id, int64_t id_len, const char* ref, int64_t ref_len,
                       const char* alt, int64_t alt_len, const char* qual, int64_t qual_len,
                       int threads);
/* ---------------------------------------------------------------------------
 * Host-side input of run_inference: VCF(.gz) -> genotype calls
 * replaces: allel.read_vcf behind read_vcf (src/utils.py:55-81) for the fields the reference
 *           reads (calldata/GT, variants/POS|REF|ALT|CHROM|ID|QUAL, samples).
 * gnx_vcf_open reads / inflates the file (plain text; bgzip members in parallel; other gzip
 * files through one zlib stream), indexes its lines and selects
 * the records whose CHROM equals `chm` (NULL: all); `threads` <= 0: gnx_host_threads().
 * gnx_vcf_copy parses them in parallel straight into the caller's arrays: gt [records]
 * [samples][2] int8 (-1 = missing or haploid second allele), pos [records] int32, qual
 * [records] float32 (NaN for '.'); any pointer may be NULL.
 * gnx_vcf_strings: field 0 CHROM, 1 ID, 2 REF, 3 ALT (comma-separated as in the file),
 * 4 sample names; every string followed by a newline is written into buf when cap suffices;
 * returns the bytes needed (call with buf = NULL first).
 * ------------------------------------------------------------------------- */
int gnx_vcf_open(gnx_vcf_t** out, const char* path, const char* chm, int threads);
void gnx_vcf_close(gnx_vcf_t* v);
int64_t gnx_vcf_num_records(const gnx_vcf_t* v);
int64_t gnx_vcf_num_samples(const gnx_vcf_t* v);
int gnx_vcf_copy(gnx_vcf_t* v, int8_t* gt, int32_t* pos, float* qual);
int64_t gnx_vcf_strings(gnx_vcf_t* v, int field, char* buf, int64_t cap);
/* vcf_to_npy after the SNP intersection (src/utils.py:121-153): X[2s+h][fmt_idx[k]] =
 * gt[vcf_idx[k]][s][h], flipped 0 <-> 1 where swap[k] (nullable), everything that is not 0 / 1 and
 * every column outside fmt_idx = miss_fill.  gt [R][S][2] int8, X [2S][ldX] int8 host memory. */
int gnx_vcf_to_haplotypes(const int8_t* gt, int64_t R, int64_t S, const int64_t* vcf_idx,
                          const int64_t* fmt_idx, const uint8_t* swap, int64_t n_idx, int64_t C,
                          int miss_fill, int8_t* X, int64_t ldX, int threads);
/* The same gather written straight into 2-bit planes (gnx_pack_rows_host layout, pitch_words >= 2*ceil(C/64)
 * 64-bit words per haplotype row, `packed` preferably pinned): what gnx_infer_host_ex takes with x_packed, so
 * the driver path VCF -> inference never materialises (or packs) the int8 matrix.  miss_fill must be in 0..3. */
int gnx_vcf_to_haplotypes_packed(const int8_t* gt, int64_t R, int64_t S, const int64_t* vcf_idx,
                                 const int64_t* fmt_idx, const uint8_t* swap, int64_t n_idx, int64_t C,
                                 int miss_fill, uint64_t* packed, int64_t pitch_words, int threads);
/* pinned (page-locked) host memory for callers without a CUDA runtime of their own (numpy drivers) */
int gnx_host_alloc_pinned(void** out, int64_t bytes);
int gnx_host_free_pinned(void* p);
/* host threads the library uses (cores this process may run on, or GNX_HOST_THREADS) */
int gnx_host_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* GNX_H_ */
