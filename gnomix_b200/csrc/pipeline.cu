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

struct Workspace {
    int device = -1;
    size_t x_bytes = 0, b_bytes = 0, p_bytes = 0, l_bytes = 0, pk_bytes = 0;
    double pack_gbs = 0.0, h2d_gbs = 0.0;      // calibrated host pack / pinned H2D rates (GB/s of int8 / of bytes)
    int calib_threads = 0;
    double frac_hint = 0.0;                    // packed fraction the feedback loop of the last call settled on
    double last_frac = 0.0;                    // what the last gnx_infer_host call did (average over its chunks)
    int64_t last_h2d = 0, last_d2h = 0;
    int8_t* X[2] = {nullptr, nullptr};
    float* B[2] = {nullptr, nullptr};
    float* P[2] = {nullptr, nullptr};
    int32_t* Lb[2] = {nullptr, nullptr};
    uint64_t* PK[2] = {nullptr, nullptr};      // packed rows on the device
    uint64_t* stage[kStage] = {};              // pinned host staging of packed rows
    int32_t* Lh[2] = {nullptr, nullptr};       // pinned host staging of labels
    float* Ph[2] = {nullptr, nullptr};         // pinned host staging of probabilities
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t h2d_done[kStage] = {};
    cudaEvent_t out_done[2] = {};
    void release() {
        for (int i = 0; i < 2; i++) {
            if (X[i]) cudaFree(X[i]);
            if (B[i]) cudaFree(B[i]);
            if (P[i]) cudaFree(P[i]);
            if (Lb[i]) cudaFree(Lb[i]);
            if (PK[i]) cudaFree(PK[i]);
            if (Lh[i]) cudaFreeHost(Lh[i]);
            if (Ph[i]) cudaFreeHost(Ph[i]);
            X[i] = nullptr; B[i] = nullptr; P[i] = nullptr; Lb[i] = nullptr; PK[i] = nullptr; Lh[i] = nullptr; Ph[i] = nullptr;
        }
        for (int i = 0; i < kStage; i++) {
            if (stage[i]) cudaFreeHost(stage[i]);
            stage[i] = nullptr;
        }
        x_bytes = b_bytes = p_bytes = l_bytes = pk_bytes = 0;
    }
};

static Workspace g_ws;
static std::mutex g_ws_mu;

static int ensure(Workspace& ws, size_t xb, size_t bb, size_t pb, size_t lb, size_t pkb) {
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
    if (xb > ws.x_bytes || bb > ws.b_bytes || pb > ws.p_bytes || lb > ws.l_bytes || pkb > ws.pk_bytes) {
        xb = std::max(xb, ws.x_bytes); bb = std::max(bb, ws.b_bytes); pb = std::max(pb, ws.p_bytes);
        lb = std::max(lb, ws.l_bytes); pkb = std::max(pkb, ws.pk_bytes);
        ws.release();
        for (int i = 0; i < 2; i++) {
            if (xb) GNX_CUDA(cudaMalloc((void**)&ws.X[i], xb));
            if (bb) GNX_CUDA(cudaMalloc((void**)&ws.B[i], bb));
            if (pb) GNX_CUDA(cudaMalloc((void**)&ws.P[i], pb));
            if (pb) GNX_CUDA(cudaHostAlloc((void**)&ws.Ph[i], pb, cudaHostAllocDefault));
            if (lb) GNX_CUDA(cudaMalloc((void**)&ws.Lb[i], lb));
            if (lb) GNX_CUDA(cudaHostAlloc((void**)&ws.Lh[i], lb, cudaHostAllocDefault));
            if (pkb) GNX_CUDA(cudaMalloc((void**)&ws.PK[i], pkb));
        }
        if (pkb)
            for (int i = 0; i < kStage; i++) GNX_CUDA(cudaHostAlloc((void**)&ws.stage[i], pkb, cudaHostAllocDefault));
        ws.x_bytes = xb; ws.b_bytes = bb; ws.p_bytes = pb; ws.l_bytes = lb; ws.pk_bytes = pkb;
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
    const int64_t fit = (int64_t)(ws.pk_bytes / ((size_t)pitch_words * 8));
    const int64_t r = std::min(rows, fit);
    if (r <= 0) return 0;
    pack_rows(X_host, r, ldX, C, ws.stage[0], pitch_words, threads, nullptr);
    const double t0 = now_s();
    pack_rows(X_host, r, ldX, C, ws.stage[0], pitch_words, threads, nullptr);
    const double tp = std::max(now_s() - t0, 1e-7);
    ws.pack_gbs = (double)r * (double)C / tp * 1e-9;
    const size_t bytes = std::min<size_t>(ws.pk_bytes, size_t(1) << 28);
    cudaEvent_t a, b;
    GNX_CUDA(cudaEventCreate(&a));
    GNX_CUDA(cudaEventCreate(&b));
    GNX_CUDA(cudaMemcpyAsync(ws.PK[0], ws.stage[0], bytes, cudaMemcpyHostToDevice, ws.st[0]));
    GNX_CUDA(cudaEventRecord(a, ws.st[0]));
    GNX_CUDA(cudaMemcpyAsync(ws.PK[0], ws.stage[0], bytes, cudaMemcpyHostToDevice, ws.st[0]));
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

}  // namespace gnx

using namespace gnx;

extern "C" int gnx_infer_host(const gnx_lr_t* lr, const gnx_gbt_t* gbt, const int8_t* X_host, int64_t N, int64_t ldX,
                              float* proba_host, int32_t* label_host, int64_t chunk_haps) {
    GNX_REQUIRE(lr && gbt, "gnx_infer_host: NULL model");
    GNX_REQUIRE(lr->d.A == gbt->d.A, "gnx_infer_host: base has A=%d, smoother A=%d", lr->d.A, gbt->d.A);
    GNX_REQUIRE(N >= 0 && ldX >= lr->d.C, "gnx_infer_host: bad shape");
    if (N == 0) return 0;
    GNX_REQUIRE(X_host && label_host, "gnx_infer_host: NULL buffer");
    const int64_t C = lr->d.C;
    const int W = lr->d.W, A = lr->d.A;
    const int64_t pitch = (C + 127) & ~int64_t(127);
    const bool pack = env_pack_enabled();
    // ~2 GB of X per slot unpacked; ~0.6 GB when packing (finer chunks: the host pack of chunk i+1 is what
    // overlaps the transfer + kernels of chunk i, and the first pack / last kernels are not overlapped at all;
    // measured on chr1 x 16 384 haplotypes: 83.0k hap/s at 256-768 haplotypes per chunk, 79.8k at 1024-2048)
    const int64_t target = pack ? (int64_t(5) << 27) : (int64_t(1) << 31);
    int64_t chunk = chunk_haps > 0 ? chunk_haps : std::max<int64_t>(256, target / pitch / 256 * 256);
    chunk = std::min<int64_t>(chunk, (N + 255) / 256 * 256);
    const int64_t pitch_words = pitch / 32;  // 2 x u64 per 64 SNPs
    std::lock_guard<std::mutex> lock(g_ws_mu);
    Workspace& ws = g_ws;
    if (ensure(ws, (size_t)chunk * pitch, (size_t)chunk * W * A * sizeof(float), proba_host ? (size_t)chunk * W * A * sizeof(float) : 0,
               (size_t)chunk * W * sizeof(int32_t), pack ? (size_t)chunk * pitch_words * 8 : 0))
        return 1;
    const int threads = host_threads_default();
    // fraction of each chunk's rows that is packed by the cores (the rest crosses the bus raw)
    double frac = 0.0;
    bool adapt = false;  // feedback on the fraction: pack more when the host waits for the bus, less when it never does
    if (pack) {
        frac = 1.0;
        const char* fe = getenv("GNX_HOST_PACK_FRAC");
        if (fe && *fe) {
            frac = std::min(1.0, std::max(0.0, atof(fe)));
        } else if (is_pinned(X_host)) {
            if (ws.pack_gbs <= 0.0 || ws.calib_threads != threads) {
                if (calibrate(ws, X_host, N, ldX, C, pitch_words, threads)) return 1;
                ws.frac_hint = 0.0;
            }
            if (ws.pack_gbs > 0.0 && ws.h2d_gbs > 0.0) {
                // the cores lose some memory bandwidth to the concurrent DMA reads: derate P
                frac = ws.frac_hint > 0.0 ? ws.frac_hint : 1.0 / (ws.h2d_gbs / (0.85 * ws.pack_gbs) + 0.75);
                adapt = true;
            }
        }
    }
    // Results: D2H straight into the caller's buffers when they are pinned; otherwise into pinned staging,
    // copied out by the worker pool once the chunk's D2H has finished (a single-threaded memcpy of the labels
    // would be ~15 % of the per-chunk host time).
    const bool lab_direct = is_pinned(label_host), proba_direct = proba_host && is_pinned(proba_host);
    int64_t pend_n0[2] = {0, 0}, pend_n[2] = {0, 0};
    auto copy_out = [&](char* dst, const char* src, size_t bytes) {
        const size_t blk = size_t(1) << 20;
        const int64_t nb = (int64_t)((bytes + blk - 1) / blk);
        parallel_for(nb, threads, [&](int64_t b) {
            const size_t o = (size_t)b * blk;
            memcpy(dst + o, src + o, std::min(blk, bytes - o));
        });
    };
    auto drain = [&](int s) -> int {
        if (pend_n[s] == 0) return 0;
        GNX_CUDA(cudaEventSynchronize(ws.out_done[s]));
        if (!lab_direct)
            copy_out(reinterpret_cast<char*>(label_host + pend_n0[s] * W), reinterpret_cast<const char*>(ws.Lh[s]),
                     (size_t)pend_n[s] * W * sizeof(int32_t));
        if (proba_host && !proba_direct)
            copy_out(reinterpret_cast<char*>(proba_host + pend_n0[s] * W * A), reinterpret_cast<const char*>(ws.Ph[s]),
                     (size_t)pend_n[s] * W * A * sizeof(float));
        pend_n[s] = 0;
        return 0;
    };
    bool stage_used[kStage] = {};
    int it = 0;
    ws.last_frac = 0.0;
    ws.last_h2d = ws.last_d2h = 0;
    int64_t rows_packed = 0;
    for (int64_t n0 = 0; n0 < N; n0 += chunk, it++) {
        const int s = it & 1, hs = it % kStage;
        const int64_t n = std::min(chunk, N - n0);
        cudaStream_t st = ws.st[s];
        int64_t np = frac >= 1.0 ? n : (int64_t)(frac * (double)n);  // rows [0, np) packed, [np, n) raw
        double wait_s = 0.0, pack_s = 0.0, t0;
        // raw part first: the DMA engine moves it while the cores pack the rest
        if (np < n) ws.last_h2d += (n - np) * C;
        if (np < n)
            GNX_CUDA(cudaMemcpy2DAsync(ws.X[s] + np * pitch, (size_t)pitch, X_host + (n0 + np) * ldX, (size_t)ldX, (size_t)C,
                                       (size_t)(n - np), cudaMemcpyHostToDevice, st));
        bool packed = false;
        if (np > 0) {
            t0 = now_s();
            if (stage_used[hs]) GNX_CUDA(cudaEventSynchronize(ws.h2d_done[hs]));
            wait_s += now_s() - t0;
            t0 = now_s();
            packed = pack_rows(X_host + n0 * ldX, np, ldX, C, ws.stage[hs], pitch_words, threads, nullptr) == 0;
            pack_s = now_s() - t0;
            if (packed) rows_packed += np;
        }
        t0 = now_s();
        if (drain(s)) return 1;  // slot s is about to be overwritten (stream order covers the device side)
        wait_s += now_s() - t0;
        if (adapt && it >= 2) {
            if (wait_s > 0.15 * pack_s) frac = std::min(1.0, frac + 0.04);
            else if (wait_s < 0.03 * pack_s) frac = std::max(0.05, frac - 0.02);
        }
        if (packed) {
            GNX_CUDA(cudaMemcpyAsync(ws.PK[s], ws.stage[hs], (size_t)np * pitch_words * 8, cudaMemcpyHostToDevice, st));
            GNX_CUDA(cudaEventRecord(ws.h2d_done[hs], st));
            stage_used[hs] = true;
            ws.last_h2d += np * pitch_words * 8;
            if (unpack_rows(ws.PK[s], np, pitch_words, C, ws.X[s], pitch, st)) return 1;
        } else if (np > 0) {
            ws.last_h2d += np * C;
            GNX_CUDA(cudaMemcpy2DAsync(ws.X[s], (size_t)pitch, X_host + n0 * ldX, (size_t)ldX, (size_t)C, (size_t)np,
                                       cudaMemcpyHostToDevice, st));
        }
        if (gnx_lr_predict(lr, ws.X[s], n, pitch, ws.B[s], st)) return 1;
        if (gnx_gbt_smooth(gbt, ws.B[s], n, W, proba_host ? ws.P[s] : nullptr, ws.Lb[s], st)) return 1;
        GNX_CUDA(cudaMemcpyAsync(lab_direct ? label_host + n0 * W : ws.Lh[s], ws.Lb[s], (size_t)n * W * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, st));
        if (proba_host)
            GNX_CUDA(cudaMemcpyAsync(proba_direct ? proba_host + n0 * W * A : ws.Ph[s], ws.P[s], (size_t)n * W * A * sizeof(float),
                                     cudaMemcpyDeviceToHost, st));
        GNX_CUDA(cudaEventRecord(ws.out_done[s], st));
        ws.last_d2h += n * W * (int64_t)sizeof(int32_t) + (proba_host ? n * W * A * (int64_t)sizeof(float) : 0);
        pend_n0[s] = n0;
        pend_n[s] = n;
    }
    ws.last_frac = (double)rows_packed / (double)N;
    if (adapt && it >= 8) ws.frac_hint = frac;
    // the older pending chunk first
    if (drain(it & 1)) return 1;
    if (drain((it & 1) ^ 1)) return 1;
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
    if (ensure(ws, 0, 0, 0, 0, (size_t)chunk * pitch_words * 8)) return 1;
    const int threads = host_threads_default();
    double frac = 1.0;
    const char* fe = getenv("GNX_HOST_PACK_FRAC");
    if (fe && *fe) {
        frac = std::min(1.0, std::max(0.0, atof(fe)));
    } else if (is_pinned(X_host)) {
        if (ws.pack_gbs <= 0.0 || ws.calib_threads != threads) {
            if (calibrate(ws, X_host, N, ldX, C, pitch_words, threads)) return 1;
            ws.frac_hint = 0.0;
        }
        if (ws.pack_gbs > 0.0 && ws.h2d_gbs > 0.0)
            frac = ws.frac_hint > 0.0 ? ws.frac_hint : 1.0 / (ws.h2d_gbs / (0.85 * ws.pack_gbs) + 0.75);
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
            GNX_CUDA(cudaMemcpyAsync(ws.PK[s], ws.stage[hs], (size_t)np * pitch_words * 8, cudaMemcpyHostToDevice, st));
            GNX_CUDA(cudaEventRecord(ws.h2d_done[hs], st));
            stage_used[hs] = true;
            if (unpack_rows(ws.PK[s], np, pitch_words, C, dst, ld_dev, st)) return 1;
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
