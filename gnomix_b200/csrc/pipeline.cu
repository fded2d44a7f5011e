// pipeline.cu -- host-buffer entry point: Gnomix.predict / predict_proba on a
// numpy-style host matrix (src/model.py:169-179, gnomix.py:55-58).
//
// Haplotype chunks stream through two device slots on two streams so that the H2D
// copy of chunk i+1 overlaps K1/K4 of chunk i and the D2H copy of chunk i-1.
//
// PCIe is the bound of this path (int8 matrix in, labels out), and the matrix only
// carries 2 bits per byte, so the host cores pack rows into bit planes (host_pack.cpp,
// all cores, AVX-512/AVX2) and unpack_kernel (pack.cu) restores the int8 tile in HBM.
// The DMA engine and the cores work at the same time: of every chunk the first
// fraction f of the rows crosses the bus packed (a quarter of the bytes, but it costs
// core time) and the rest raw (no core time), with f = 1 / (R/P + 3/4) balancing the
// two for the measured pack rate P and H2D rate R (calibrated once per workspace).
// Pageable input is always packed (the cores read it faster than a pageable cudaMemcpy
// stages it).  A chunk holding values outside 0..3 is shipped raw.
// GNX_HOST_PACK=0 disables packing, GNX_HOST_PACK_FRAC=f pins the fraction,
// GNX_HOST_THREADS bounds the host threads.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>

#include "gbt_smooth.cuh"
#include "host_pack.h"
#include "lr_base.cuh"

namespace gnx {

int unpack_rows(const uint64_t* packed_dev, int64_t n, int64_t pitch_words, int64_t C, int8_t* X_dev, int64_t ldX,
                cudaStream_t st);

constexpr int kStage = 3;  // pinned staging buffers of packed rows (pack i+1 while i and i-1 are in flight)

// device slot buffers (two of each) and pinned host staging of the outputs
enum { D_X, D_B, D_B32, D_P, D_Q, D_L, D_PK, D_COUNT };   // int8 X | base out | float32 copy of a float64 base out |
                                                         // smoother proba | calibrated proba | labels | packed rows
enum { H_L, H_P, H_X, H_COUNT };                          // labels | final proba | phased X

struct Workspace {
    int device = -1;
    size_t dbytes[D_COUNT] = {}, hbytes[H_COUNT] = {}, stage_bytes = 0;
    double pack_gbs = 0.0, h2d_gbs = 0.0;      // calibrated host pack / pinned H2D rates (GB/s of int8 / of bytes)
    int calib_threads = 0;
    double frac_hint = -1.0;                   // packed fraction the controller of the last call settled on (< 0: none)
    double frac_step = 0.0;
    double last_frac = 0.0;                    // what the last gnx_infer_host call did (average over its chunks)
    int64_t last_h2d = 0, last_d2h = 0;
    void* dev[D_COUNT][2] = {};
    void* host[H_COUNT][2] = {};
    uint64_t* stage[kStage] = {};              // pinned host staging of packed rows
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t h2d_done[kStage] = {};
    cudaEvent_t out_done[2] = {};
    int8_t* X(int s) const { return static_cast<int8_t*>(dev[D_X][s]); }
    uint64_t* PK(int s) const { return static_cast<uint64_t*>(dev[D_PK][s]); }
    void release() {
        for (int i = 0; i < 2; i++) {
            for (int k = 0; k < D_COUNT; k++) {
                if (dev[k][i]) cudaFree(dev[k][i]);
                dev[k][i] = nullptr;
            }
            for (int k = 0; k < H_COUNT; k++) {
                if (host[k][i]) cudaFreeHost(host[k][i]);
                host[k][i] = nullptr;
            }
        }
        for (int i = 0; i < kStage; i++) {
            if (stage[i]) cudaFreeHost(stage[i]);
            stage[i] = nullptr;
        }
        for (int k = 0; k < D_COUNT; k++) dbytes[k] = 0;
        for (int k = 0; k < H_COUNT; k++) hbytes[k] = 0;
        stage_bytes = 0;
    }
};

static Workspace g_ws;
static std::mutex g_ws_mu;

// grows (never shrinks) the slot buffers to at least the requested sizes
static int ensure(Workspace& ws, const size_t (&db)[D_COUNT], const size_t (&hb)[H_COUNT], size_t stage_b) {
    int dev = 0;
    GNX_CUDA(cudaGetDevice(&dev));
    if (ws.device != dev) {
        ws.release();
        for (int i = 0; i < 2; i++) {
            if (ws.st[i]) cudaStreamDestroy(ws.st[i]);
            GNX_CUDA(cudaStreamCreateWithFlags(&ws.st[i], cudaStreamNonBlocking));
            if (ws.out_done[i]) cudaEventDestroy(ws.out_done[i]);
            GNX_CUDA(cudaEventCreateWithFlags(&ws.out_done[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < kStage; i++) {
            if (ws.h2d_done[i]) cudaEventDestroy(ws.h2d_done[i]);
            GNX_CUDA(cudaEventCreateWithFlags(&ws.h2d_done[i], cudaEventDisableTiming));
        }
        ws.device = dev;
    }
    for (int k = 0; k < D_COUNT; k++) {
        if (db[k] <= ws.dbytes[k]) continue;
        for (int i = 0; i < 2; i++) {
            if (ws.dev[k][i]) cudaFree(ws.dev[k][i]);
            ws.dev[k][i] = nullptr;
            GNX_CUDA(cudaMalloc(&ws.dev[k][i], db[k]));
        }
        ws.dbytes[k] = db[k];
    }
    for (int k = 0; k < H_COUNT; k++) {
        if (hb[k] <= ws.hbytes[k]) continue;
        for (int i = 0; i < 2; i++) {
            if (ws.host[k][i]) cudaFreeHost(ws.host[k][i]);
            ws.host[k][i] = nullptr;
            GNX_CUDA(cudaHostAlloc(&ws.host[k][i], hb[k], cudaHostAllocDefault));
        }
        ws.hbytes[k] = hb[k];
    }
    if (stage_b > ws.stage_bytes) {
        for (int i = 0; i < kStage; i++) {
            if (ws.stage[i]) cudaFreeHost(ws.stage[i]);
            ws.stage[i] = nullptr;
            GNX_CUDA(cudaHostAlloc((void**)&ws.stage[i], stage_b, cudaHostAllocDefault));
        }
        ws.stage_bytes = stage_b;
    }
    return 0;
}

static bool env_pack_enabled() {
    const char* e = getenv("GNX_HOST_PACK");
    return !(e && e[0] == '0');
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// One-off measurement of the host pack rate (on the caller's own rows) and of the pinned H2D rate.
static int calibrate(Workspace& ws, const int8_t* X_host, int64_t n, int64_t ldX, int64_t C, int64_t pitch_words, int threads) {
    const int64_t rows = std::min<int64_t>(n, std::max<int64_t>(32, (int64_t(1) << 28) / std::max<int64_t>(C, 1)));
    const int64_t fit = (int64_t)(ws.stage_bytes / ((size_t)pitch_words * 8));
    const int64_t r = std::min(rows, fit);
    if (r <= 0) return 0;
    pack_rows(X_host, r, ldX, C, ws.stage[0], pitch_words, threads, nullptr);
    const double t0 = now_s();
    pack_rows(X_host, r, ldX, C, ws.stage[0], pitch_words, threads, nullptr);
    const double tp = std::max(now_s() - t0, 1e-7);
    ws.pack_gbs = (double)r * (double)C / tp * 1e-9;
    const size_t bytes = std::min<size_t>(std::min(ws.stage_bytes, ws.dbytes[D_PK]), size_t(1) << 28);
    cudaEvent_t a, b;
    GNX_CUDA(cudaEventCreate(&a));
    GNX_CUDA(cudaEventCreate(&b));
    GNX_CUDA(cudaMemcpyAsync(ws.PK(0), ws.stage[0], bytes, cudaMemcpyHostToDevice, ws.st[0]));
    GNX_CUDA(cudaEventRecord(a, ws.st[0]));
    GNX_CUDA(cudaMemcpyAsync(ws.PK(0), ws.stage[0], bytes, cudaMemcpyHostToDevice, ws.st[0]));
    GNX_CUDA(cudaEventRecord(b, ws.st[0]));
    GNX_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    GNX_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    ws.h2d_gbs = (double)bytes / std::max((double)ms, 1e-3) * 1e-6;
    ws.calib_threads = threads;
    return 0;
}

// accessors defined next to the opaque structs
void svc_dims(const gnx_svc* m, int64_t* C, int* W, int* A);
void crf_dims(const gnx_crf* m, int* A, int* L);
int cal_classes(const gnx_cal* m);

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __double2float_rn(in[i]);   // round to nearest even, as numpy's astype(float32) (slide_window's float32 X_slide)
}

// Packed fraction of each chunk: measured, not modelled.  The controller times windows of chunks of the running job
// (host clock between chunk starts; in steady state that is the bottleneck resource, whichever it is -- cores, bus
// or host DRAM shared with the other ranks of the box) and hill-climbs f; f = 0 (everything raw, no core time) is
// probed at the start of every call whose packed rate is not clearly above what the bus carries raw, so that the
// packed path is not slower than the unpacked one (8 ranks on one host: it was, 129 k against 150 k haplotypes/s,
// before the probe was repeated per call -- the ranks of a job start their calls together, so they probe together).
struct FracCtl {
    static constexpr int kSkip = 2, kMeasure = 3;   // chunks after a change of f that are not timed / that are
    double f = 0.0, best_f = 0.0, best_rate = 0.0, step = 0.1;
    int dir = +1, tried_other = 0, probe_zero = 0, settled = 0;
    int in_win = 0;
    double t_start = 0.0;
    int64_t rows = 0;
    bool first = true;
    double raw_bound = 0.0;   // rows/s the bus alone could carry unpacked (calibrated H2D rate / row bytes)
    void begin(double f0, double step0, bool probe0, double raw_rows_per_s) {
        f = best_f = f0;
        step = step0 > 0.0 ? step0 : 0.1;
        probe_zero = probe0 ? 1 : 0;
        raw_bound = raw_rows_per_s;
    }
    // called at the start of every chunk with the rows the previous chunk held
    void tick(double now, int64_t rows_prev) {
        if (settled) return;
        if (in_win > kSkip) rows += rows_prev;
        if (in_win == kSkip) { t_start = now; rows = 0; }
        in_win++;
        if (in_win <= kSkip + kMeasure) return;
        const double rate = (double)rows / std::max(now - t_start, 1e-9);
        in_win = 0;
        decide(rate);
    }
    void decide(double rate) {
        if (first) {   // rate at the starting point
            first = false;
            best_rate = rate;
            best_f = f;
            // f = 0 is worth a window only where it could win: when the packed path runs well above what the bus
            // carries raw (one GPU with all the host's cores) the probe would just cost; when several ranks share the
            // host's memory system the packed path can fall BELOW the raw one (packing moves 1.5 bytes of host DRAM
            // traffic per SNP, a raw copy 1), and every call checks
            if (probe_zero == 1 && f > 0.0 && (raw_bound <= 0.0 || rate < 1.3 * raw_bound)) { probe_zero = 2; f = 0.0; return; }
            probe_zero = 0;
            next_candidate();
            return;
        }
        if (probe_zero == 2) {   // this window ran raw
            probe_zero = 0;
            if (rate >= best_rate) { best_rate = rate; best_f = 0.0; dir = +1; }
            f = best_f;
            next_candidate();
            return;
        }
        if (rate > best_rate * 1.02) {
            best_rate = rate;
            best_f = f;
            tried_other = 0;
        } else {
            if (!tried_other) { dir = -dir; tried_other = 1; }
            else { step *= 0.5; tried_other = 0; }
            best_rate = 0.98 * best_rate + 0.02 * rate;   // let a stale best decay slowly
        }
        next_candidate();
    }
    void next_candidate() {
        if (step < 0.03) { f = best_f; settled = 1; return; }
        double c = best_f + dir * step;
        if (c < 0.0 || c > 1.0) {
            dir = -dir;
            c = best_f + dir * step;
        }
        f = std::min(1.0, std::max(0.0, c));
    }
};
}  // namespace gnx

using namespace gnx;

extern "C" int gnx_infer_host_ex(const gnx_pipeline_t* p, const void* X_host_v, int64_t N, int64_t ldX, void* proba_host_v,
                                 int32_t* label_host, int8_t* X_phased_host, int64_t chunk_haps) {
    GNX_REQUIRE(p != nullptr, "gnx_infer_host_ex: NULL pipeline");
    GNX_REQUIRE((p->lr != nullptr) != (p->svc != nullptr), "gnx_infer_host_ex: exactly one base model (lr or svc) is needed");
    GNX_REQUIRE((p->gbt != nullptr) != (p->crf != nullptr), "gnx_infer_host_ex: exactly one smoother (gbt or crf) is needed");
    GNX_REQUIRE(!p->phase || p->gbt || (p->crf && p->crf_phase_S > 0),
                "gnx_infer_host_ex: Gnofix needs the tree smoother (src/model.py:194); the CRF extension is opted into with crf_phase_S");
    int64_t C = 0;
    int W = 0, A = 0;
    if (p->lr) { C = p->lr->d.C; W = p->lr->d.W; A = p->lr->d.A; }
    else svc_dims(p->svc, &C, &W, &A);
    if (p->gbt) GNX_REQUIRE(p->gbt->d.A == A, "gnx_infer_host_ex: base has A=%d, smoother A=%d", A, p->gbt->d.A);
    if (p->crf) {
        int ca = 0, cl = 0;
        crf_dims(p->crf, &ca, &cl);
        GNX_REQUIRE(ca == A && cl == A, "gnx_infer_host_ex: base has A=%d, CRF %d attributes / %d labels", A, ca, cl);
    }
    if (p->cal) GNX_REQUIRE(cal_classes(p->cal) == A, "gnx_infer_host_ex: calibrator has %d classes, base A=%d", cal_classes(p->cal), A);
    const int64_t pitch = (C + 127) & ~int64_t(127);
    const int64_t pitch_words = pitch / 32;  // 2 x u64 per 64 SNPs
    const bool x_packed = p->x_packed != 0;
    GNX_REQUIRE(N >= 0 && (x_packed ? ldX >= 2 * ((C + 63) / 64) : ldX >= C), "gnx_infer_host_ex: bad shape (N=%lld, ldX=%lld, C=%lld)",
                (long long)N, (long long)ldX, (long long)C);
    GNX_REQUIRE(!p->phase || N % 2 == 0, "gnx_infer_host_ex: Gnofix works on haplotype pairs, N=%lld is odd", (long long)N);
    if (N == 0) return 0;
    GNX_REQUIRE(X_host_v && label_host, "gnx_infer_host_ex: NULL buffer");
    if (require_blackwell()) return 1;
    const int8_t* X_host = static_cast<const int8_t*>(X_host_v);
    const uint64_t* XP_host = static_cast<const uint64_t*>(X_host_v);
    // dtypes along the pipeline, as the reference's stages hand them over
    const bool base_f64 = p->svc != nullptr || p->crf != nullptr;   // libsvm / the CRF's input are float64
    const bool need_b32 = base_f64 && p->gbt != nullptr;            // string-kernel base into the tree smoother
    const bool sm_f64 = p->crf != nullptr;
    const bool out_f64 = sm_f64 || p->cal != nullptr;               // what proba_host holds
    const size_t out_elt = out_f64 ? 8 : 4;
    const bool want_proba = proba_host_v != nullptr;
    const bool pack = env_pack_enabled() && !x_packed;
    // ~2 GB of X per slot unpacked; ~0.6 GB when packing (finer chunks: the host pack of chunk i+1 is what
    // overlaps the transfer + kernels of chunk i, and the first pack / last kernels are not overlapped at all;
    // measured on chr1 x 16 384 haplotypes: 83.0k hap/s at 256-768 haplotypes per chunk, 79.8k at 1024-2048)
    const int64_t target = (pack || x_packed) ? (int64_t(5) << 27) : (int64_t(1) << 31);
    int64_t chunk = chunk_haps > 0 ? (chunk_haps + 1) / 2 * 2 : std::max<int64_t>(256, target / pitch / 256 * 256);
    chunk = std::min<int64_t>(chunk, (N + 255) / 256 * 256);
    std::lock_guard<std::mutex> lock(g_ws_mu);
    Workspace& ws = g_ws;
    const bool lab_direct = is_pinned(label_host), proba_direct = want_proba && is_pinned(proba_host_v);
    const bool xph_direct = X_phased_host && is_pinned(X_phased_host);
    const bool x_pinned = is_pinned(X_host_v);
    {
        const size_t wa = (size_t)chunk * W * A;
        size_t db[D_COUNT] = {}, hb[H_COUNT] = {};
        db[D_X] = (size_t)chunk * pitch;
        db[D_B] = wa * (base_f64 ? 8 : 4);
        db[D_B32] = need_b32 ? wa * 4 : 0;
        db[D_P] = (want_proba || p->cal) ? wa * (sm_f64 ? 8 : 4) : 0;
        db[D_Q] = (p->cal && want_proba) ? wa * 8 : 0;
        db[D_L] = (size_t)chunk * W * 4;
        db[D_PK] = (pack || x_packed) ? (size_t)chunk * pitch_words * 8 : 0;
        hb[H_L] = lab_direct ? 0 : (size_t)chunk * W * 4;
        hb[H_P] = (want_proba && !proba_direct) ? wa * out_elt : 0;
        hb[H_X] = (X_phased_host && !xph_direct) ? (size_t)chunk * C : 0;
        const size_t stage_b = (pack || (x_packed && !x_pinned)) ? (size_t)chunk * pitch_words * 8 : 0;
        if (ensure(ws, db, hb, stage_b)) return 1;
    }
    const int threads = host_threads_default();
    // fraction of each chunk's rows that is packed by the cores (the rest crosses the bus raw)
    double frac = 0.0;
    FracCtl ctl;
    bool adapt = false;
    if (pack) {
        frac = 1.0;
        const char* fe = getenv("GNX_HOST_PACK_FRAC");
        if (fe && *fe) {
            frac = std::min(1.0, std::max(0.0, atof(fe)));
        } else if (x_pinned) {
            bool fresh = false;
            if (ws.pack_gbs <= 0.0 || ws.calib_threads != threads) {
                if (calibrate(ws, X_host, N, ldX, C, pitch_words, threads)) return 1;
                ws.frac_hint = -1.0;
                ws.frac_step = 0.0;
                fresh = true;
            }
            if (ws.pack_gbs > 0.0 && ws.h2d_gbs > 0.0 && ws.pack_gbs < ws.h2d_gbs) {
                // The cores (with the memory system behind them) are scarcer than the bus: many ranks on one host.
                // Measured on the 8-GPU box (4 host threads per rank; rates calibrated while every rank calibrates:
                // pack 9.4 GB/s, H2D 15.9 GB/s): everything raw 150 k haplotypes/s, the packed fractions the per-rank
                // controllers drift to 129-137 k -- a rank that packs takes memory bandwidth from the other ranks'
                // copies, so greedy per-rank hill-climbing settles above the joint optimum.  Ship raw.
                frac = 0.0;
            } else if (ws.pack_gbs > 0.0 && ws.h2d_gbs > 0.0) {
                // starting point: balance of the calibrated rates (the cores lose some memory bandwidth to the
                // concurrent DMA reads: derate P); from there the controller follows what it measures
                frac = ws.frac_hint >= 0.0 ? ws.frac_hint : 1.0 / (ws.h2d_gbs / (0.85 * ws.pack_gbs) + 0.75);
                (void)fresh;
                ctl.begin(frac, ws.frac_step, true, ws.h2d_gbs * 1e9 / (double)C);
                adapt = true;
            }
        }
    }
    // Results: D2H straight into the caller's buffers when they are pinned; otherwise into pinned staging,
    // copied out by the worker pool once the chunk's D2H has finished (a single-threaded memcpy of the labels
    // would be ~15 % of the per-chunk host time).
    char* proba_host = static_cast<char*>(proba_host_v);
    int64_t pend_n0[2] = {0, 0}, pend_n[2] = {0, 0};
    auto copy_out = [&](char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, int64_t rows) {
        if (dpitch == width && spitch == width) {
            const size_t bytes = width * (size_t)rows, blk = size_t(1) << 20;
            const int64_t nb = (int64_t)((bytes + blk - 1) / blk);
            parallel_for(nb, threads, [&](int64_t b) {
                const size_t o = (size_t)b * blk;
                memcpy(dst + o, src + o, std::min(blk, bytes - o));
            });
        } else {
            parallel_for(rows, threads, [&](int64_t r) { memcpy(dst + (size_t)r * dpitch, src + (size_t)r * spitch, width); });
        }
    };
    auto drain = [&](int s) -> int {
        if (pend_n[s] == 0) return 0;
        GNX_CUDA(cudaEventSynchronize(ws.out_done[s]));
        if (!lab_direct)
            copy_out(reinterpret_cast<char*>(label_host + pend_n0[s] * W), (size_t)W * 4, static_cast<const char*>(ws.host[H_L][s]),
                     (size_t)W * 4, (size_t)W * 4, pend_n[s]);
        if (want_proba && !proba_direct)
            copy_out(proba_host + (size_t)pend_n0[s] * W * A * out_elt, (size_t)W * A * out_elt, static_cast<const char*>(ws.host[H_P][s]),
                     (size_t)W * A * out_elt, (size_t)W * A * out_elt, pend_n[s]);
        if (X_phased_host && !xph_direct)
            copy_out(reinterpret_cast<char*>(X_phased_host + pend_n0[s] * C), (size_t)C, static_cast<const char*>(ws.host[H_X][s]), (size_t)C,
                     (size_t)C, pend_n[s]);
        pend_n[s] = 0;
        return 0;
    };
    // base stage of one slot (also used for the second pass over the phased haplotypes)
    auto run_base = [&](int s, int64_t n, cudaStream_t st) -> int {
        if (p->lr) {
            if (base_f64) return gnx_lr_predict_f64(p->lr, ws.X(s), n, pitch, static_cast<double*>(ws.dev[D_B][s]), st);
            return gnx_lr_predict(p->lr, ws.X(s), n, pitch, static_cast<float*>(ws.dev[D_B][s]), st);
        }
        if (gnx_svc_predict(p->svc, ws.X(s), n, pitch, static_cast<double*>(ws.dev[D_B][s]), st)) return 1;
        return 0;
    };
    auto smoother_input = [&](int s, int64_t n, cudaStream_t st) -> const float* {   // float32 B for the tree smoother
        if (!need_b32) return static_cast<const float*>(ws.dev[D_B][s]);
        const int64_t cnt = n * W * A;
        f64_to_f32_kernel<<<(unsigned)std::min<int64_t>(ceil_div(cnt, 256), 148 * 16), 256, 0, st>>>(
            static_cast<const double*>(ws.dev[D_B][s]), static_cast<float*>(ws.dev[D_B32][s]), cnt);
        return static_cast<const float*>(ws.dev[D_B32][s]);
    };
    bool stage_used[kStage] = {};
    int it = 0;
    ws.last_frac = 0.0;
    ws.last_h2d = ws.last_d2h = 0;
    int64_t rows_packed = 0, rows_prev = 0;
    // ramp: nothing overlaps the pack / H2D of the first chunk, so it is a quarter of a chunk and the second one a half
    const bool ramp = chunk >= 256 && N > 2 * chunk;
    int64_t n = 0;
    for (int64_t n0 = 0; n0 < N; n0 += n, it++) {
        const int s = it & 1, hs = it % kStage;
        n = std::min((ramp && it < 2) ? ((it == 0 ? chunk / 4 : chunk / 2) & ~int64_t(1)) : chunk, N - n0);
        cudaStream_t st = ws.st[s];
        if (adapt) {
            ctl.tick(now_s(), rows_prev);
            frac = ctl.f;
        }
        rows_prev = n;
        if (x_packed) {
            // the caller's rows are 2-bit planes already (gnx_pack_rows_host layout): a quarter of the bytes, no core time
            const uint64_t* src = XP_host + n0 * ldX;
            if (!x_pinned || ldX != pitch_words) {
                if (stage_used[hs]) GNX_CUDA(cudaEventSynchronize(ws.h2d_done[hs]));
                copy_out(reinterpret_cast<char*>(ws.stage[hs]), (size_t)pitch_words * 8, reinterpret_cast<const char*>(src), (size_t)ldX * 8,
                         (size_t)std::min(ldX, pitch_words) * 8, n);
                src = ws.stage[hs];
            }
            if (drain(s)) return 1;
            GNX_CUDA(cudaMemcpyAsync(ws.PK(s), src, (size_t)n * pitch_words * 8, cudaMemcpyHostToDevice, st));
            if (src == ws.stage[hs]) {
                GNX_CUDA(cudaEventRecord(ws.h2d_done[hs], st));
                stage_used[hs] = true;
            }
            ws.last_h2d += n * pitch_words * 8;
            rows_packed += n;
            if (unpack_rows(ws.PK(s), n, pitch_words, C, ws.X(s), pitch, st)) return 1;
        } else {
            const int64_t np = frac >= 1.0 ? n : (int64_t)(frac * (double)n);  // rows [0, np) packed, [np, n) raw
            // the raw part of slot s may only be overwritten once chunk it-2 has left it: stream order covers the device
            // side; raw part first, so that the DMA engine moves it while the cores pack the rest
            if (np < n) {
                ws.last_h2d += (n - np) * C;
                GNX_CUDA(cudaMemcpy2DAsync(ws.X(s) + np * pitch, (size_t)pitch, X_host + (n0 + np) * ldX, (size_t)ldX, (size_t)C,
                                           (size_t)(n - np), cudaMemcpyHostToDevice, st));
            }
            bool packed = false;
            if (np > 0) {
                if (stage_used[hs]) GNX_CUDA(cudaEventSynchronize(ws.h2d_done[hs]));
                packed = pack_rows(X_host + n0 * ldX, np, ldX, C, ws.stage[hs], pitch_words, threads, nullptr) == 0;
                if (packed) rows_packed += np;
            }
            if (drain(s)) return 1;  // the output staging of slot s is about to be reused
            if (packed) {
                GNX_CUDA(cudaMemcpyAsync(ws.PK(s), ws.stage[hs], (size_t)np * pitch_words * 8, cudaMemcpyHostToDevice, st));
                GNX_CUDA(cudaEventRecord(ws.h2d_done[hs], st));
                stage_used[hs] = true;
                ws.last_h2d += np * pitch_words * 8;
                if (unpack_rows(ws.PK(s), np, pitch_words, C, ws.X(s), pitch, st)) return 1;
            } else if (np > 0) {
                ws.last_h2d += np * C;
                GNX_CUDA(cudaMemcpy2DAsync(ws.X(s), (size_t)pitch, X_host + n0 * ldX, (size_t)ldX, (size_t)C, (size_t)np,
                                           cudaMemcpyHostToDevice, st));
            }
        }
        // ---- Base -> [Gnofix] -> Smoother -> [Calibrator] on slot s
        if (run_base(s, n, st)) return 1;
        int32_t* lab_dev = static_cast<int32_t*>(ws.dev[D_L][s]);
        bool labels_done = false;
        if (p->phase) {
            // Gnomix.phase (src/model.py:188-214): X and B swapped in place, labels from gnofix; the probabilities the
            // driver writes are those of the phased haplotypes run through the model again (gnomix.py:72)
            if (p->gbt) {
                float* b32 = const_cast<float*>(smoother_input(s, n, st));
                if (gnx_gnofix(p->gbt, ws.X(s), pitch, C, b32, n / 2, W, p->max_it > 0 ? p->max_it : 50, lab_dev, nullptr, st)) return 1;
            } else {
                // the CRF + Gnofix extension (no reference behaviour): float64 base probabilities, rounds with host syncs
                if (gnx_gnofix_crf(p->crf, p->crf_phase_S, ws.X(s), pitch, C, static_cast<double*>(ws.dev[D_B][s]), n / 2, W,
                                   p->max_it > 0 ? p->max_it : 50, lab_dev, nullptr, st))
                    return 1;
            }
            labels_done = true;
            if (X_phased_host) {
                GNX_CUDA(cudaMemcpy2DAsync(xph_direct ? (void*)(X_phased_host + n0 * C) : ws.host[H_X][s], (size_t)C, ws.X(s), (size_t)pitch,
                                           (size_t)C, (size_t)n, cudaMemcpyDeviceToHost, st));
                ws.last_d2h += n * C;
            }
            if (want_proba && run_base(s, n, st)) return 1;
        }
        const void* final_proba = nullptr;
        if (!labels_done || want_proba) {
            void* P = (want_proba || p->cal) ? ws.dev[D_P][s] : nullptr;
            int32_t* L = (labels_done || p->cal) ? nullptr : lab_dev;
            if (p->gbt) {
                const float* b32 = smoother_input(s, n, st);
                if (gnx_gbt_smooth(p->gbt, b32, n, W, static_cast<float*>(P), L, st)) return 1;
            } else {
                if (gnx_crf_smooth(p->crf, static_cast<const double*>(ws.dev[D_B][s]), n, W, static_cast<double*>(P), L, st)) return 1;
            }
            final_proba = P;
            if (p->cal) {   // Smoother.predict_proba with calibrate (src/Smooth/smooth.py:48-52): float64 out, labels from it
                void* Q = want_proba ? ws.dev[D_Q][s] : nullptr;
                if (gnx_calibrate(p->cal, P, sm_f64 ? 0 : 1, n * (int64_t)W, static_cast<double*>(Q), labels_done ? nullptr : lab_dev, st))
                    return 1;
                final_proba = Q;
            }
        }
        GNX_CUDA(cudaMemcpyAsync(lab_direct ? (void*)(label_host + n0 * W) : ws.host[H_L][s], lab_dev, (size_t)n * W * 4, cudaMemcpyDeviceToHost, st));
        ws.last_d2h += n * W * 4;
        if (want_proba) {
            GNX_CUDA(cudaMemcpyAsync(proba_direct ? (void*)(proba_host + (size_t)n0 * W * A * out_elt) : ws.host[H_P][s], final_proba,
                                     (size_t)n * W * A * out_elt, cudaMemcpyDeviceToHost, st));
            ws.last_d2h += n * W * A * (int64_t)out_elt;
        }
        GNX_CUDA(cudaEventRecord(ws.out_done[s], st));
        pend_n0[s] = n0;
        pend_n[s] = n;
    }
    ws.last_frac = (double)rows_packed / (double)N;
    if (adapt && it >= 2 * (FracCtl::kSkip + FracCtl::kMeasure + 1)) {
        ws.frac_hint = ctl.settled ? ctl.best_f : ctl.best_f;
        ws.frac_step = ctl.settled ? 0.06 : std::max(ctl.step, 0.03);   // keep probing a little on later calls
    }
    // the older pending chunk first
    if (drain(it & 1)) return 1;
    if (drain((it & 1) ^ 1)) return 1;
    return 0;
}

extern "C" int gnx_infer_host(const gnx_lr_t* lr, const gnx_gbt_t* gbt, const int8_t* X_host, int64_t N, int64_t ldX,
                              float* proba_host, int32_t* label_host, int64_t chunk_haps) {
    GNX_REQUIRE(lr && gbt, "gnx_infer_host: NULL model");
    gnx_pipeline_t p = {};
    p.lr = lr;
    p.gbt = gbt;
    return gnx_infer_host_ex(&p, X_host, N, ldX, proba_host, label_host, nullptr, chunk_haps);
}

extern "C" int gnx_host_alloc_pinned(void** out, int64_t bytes) {
    GNX_REQUIRE(out != nullptr && bytes >= 0, "gnx_host_alloc_pinned: bad arguments");
    *out = nullptr;
    if (bytes == 0) return 0;
    GNX_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return 0;
}

extern "C" int gnx_host_free_pinned(void* p) {
    if (p) GNX_CUDA(cudaFreeHost(p));
    return 0;
}

/* last calibration of the host-buffer pipeline: pack rate (GB/s of int8 input), pinned H2D rate (GB/s) */
extern "C" int gnx_infer_host_rates(double* pack_gbs, double* h2d_gbs) {
    std::lock_guard<std::mutex> lock(g_ws_mu);
    if (pack_gbs) *pack_gbs = g_ws.pack_gbs;
    if (h2d_gbs) *h2d_gbs = g_ws.h2d_gbs;
    return 0;
}

/* what the last gnx_infer_host call moved: packed fraction of each chunk, bytes H2D, bytes D2H */
extern "C" int gnx_infer_host_last_transfer(double* frac, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    std::lock_guard<std::mutex> lock(g_ws_mu);
    if (frac) *frac = g_ws.last_frac;
    if (h2d_bytes) *h2d_bytes = g_ws.last_h2d;
    if (d2h_bytes) *d2h_bytes = g_ws.last_d2h;
    return 0;
}

/* Host int8 matrix -> device int8 matrix [N, ld_dev] (ld_dev a multiple of 128 >= C, the pitch K1 / K2 address),
 * through the same packed transfer as gnx_infer_host: the cores pack chunk i+1 to 2-bit planes while chunk i
 * crosses the bus, unpack_kernel restores the bytes in place.  Pinned input splits every chunk between packed
 * and raw rows like gnx_infer_host; chunks with values outside 0..3 go raw.  Synchronous. */
extern "C" int gnx_upload_haplotypes(const int8_t* X_host, int64_t N, int64_t ldX, int64_t C, int8_t* X_dev, int64_t ld_dev) {
    GNX_REQUIRE(N >= 0 && C >= 1 && ldX >= C, "gnx_upload_haplotypes: bad shape");
    GNX_REQUIRE(ld_dev % 128 == 0 && ld_dev >= C && ((uintptr_t)X_dev & 15) == 0, "gnx_upload_haplotypes: ld_dev must be a multiple of 128 >= C and X_dev 16-byte aligned");
    if (N == 0) return 0;
    GNX_REQUIRE(X_host && X_dev, "gnx_upload_haplotypes: NULL buffer");
    if (require_blackwell()) return 1;
    const int64_t pitch = (C + 127) & ~int64_t(127);
    const int64_t pitch_words = pitch / 32;
    if (!env_pack_enabled()) {
        GNX_CUDA(cudaMemcpy2D(X_dev, (size_t)ld_dev, X_host, (size_t)ldX, (size_t)C, (size_t)N, cudaMemcpyHostToDevice));
        return 0;
    }
    int64_t chunk = std::max<int64_t>(256, (int64_t(5) << 27) / pitch / 256 * 256);
    chunk = std::min<int64_t>(chunk, (N + 255) / 256 * 256);
    std::lock_guard<std::mutex> lock(g_ws_mu);
    Workspace& ws = g_ws;
    {
        size_t db[D_COUNT] = {}, hb[H_COUNT] = {};
        db[D_PK] = (size_t)chunk * pitch_words * 8;
        if (ensure(ws, db, hb, db[D_PK])) return 1;
    }
    const int threads = host_threads_default();
    double frac = 1.0;
    const char* fe = getenv("GNX_HOST_PACK_FRAC");
    if (fe && *fe) {
        frac = std::min(1.0, std::max(0.0, atof(fe)));
    } else if (is_pinned(X_host)) {
        if (ws.pack_gbs <= 0.0 || ws.calib_threads != threads) {
            if (calibrate(ws, X_host, N, ldX, C, pitch_words, threads)) return 1;
            ws.frac_hint = -1.0;
        }
        if (ws.pack_gbs > 0.0 && ws.h2d_gbs > 0.0)
            frac = ws.pack_gbs < ws.h2d_gbs ? 0.0   // many ranks on one host: ship raw (see gnx_infer_host_ex)
                                            : (ws.frac_hint >= 0.0 ? ws.frac_hint : 1.0 / (ws.h2d_gbs / (0.85 * ws.pack_gbs) + 0.75));
    }
    bool stage_used[kStage] = {};
    int it = 0;
    for (int64_t n0 = 0; n0 < N; n0 += chunk, it++) {
        const int s = it & 1, hs = it % kStage;
        const int64_t n = std::min(chunk, N - n0);
        cudaStream_t st = ws.st[s];
        int8_t* dst = X_dev + n0 * ld_dev;
        const int64_t np = frac >= 1.0 ? n : (int64_t)(frac * (double)n);
        if (np < n)
            GNX_CUDA(cudaMemcpy2DAsync(dst + np * ld_dev, (size_t)ld_dev, X_host + (n0 + np) * ldX, (size_t)ldX, (size_t)C,
                                       (size_t)(n - np), cudaMemcpyHostToDevice, st));
        if (np == 0) continue;
        if (stage_used[hs]) GNX_CUDA(cudaEventSynchronize(ws.h2d_done[hs]));
        if (pack_rows(X_host + n0 * ldX, np, ldX, C, ws.stage[hs], pitch_words, threads, nullptr) == 0) {
            GNX_CUDA(cudaMemcpyAsync(ws.PK(s), ws.stage[hs], (size_t)np * pitch_words * 8, cudaMemcpyHostToDevice, st));
            GNX_CUDA(cudaEventRecord(ws.h2d_done[hs], st));
            stage_used[hs] = true;
            if (unpack_rows(ws.PK(s), np, pitch_words, C, dst, ld_dev, st)) return 1;
        } else {
            GNX_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld_dev, X_host + n0 * ldX, (size_t)ldX, (size_t)C, (size_t)np,
                                       cudaMemcpyHostToDevice, st));
        }
    }
    GNX_CUDA(cudaStreamSynchronize(ws.st[0]));
    GNX_CUDA(cudaStreamSynchronize(ws.st[1]));
    return 0;
}

/* Frees the device slots and pinned staging buffers gnx_infer_host / gnx_upload_haplotypes keep between calls
 * (they are re-created on demand); the calibration is kept. */
extern "C" int gnx_release_workspace(void) {
    std::lock_guard<std::mutex> lock(g_ws_mu);
    {   // the smoother's rank scratch lives in the stream-ordered pool (gbt_tile.cu keeps freed blocks there)
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(pool, 0);
        }
    }
    if (g_ws.device >= 0) {
        int cur = 0;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != g_ws.device) {
            cudaSetDevice(g_ws.device);
            g_ws.release();
            cudaSetDevice(cur);
            return 0;
        }
    }
    g_ws.release();
    return 0;
}
