// pipeline.cu -- host-buffer entry point: Gnomix.predict / predict_proba on a
// numpy-style host matrix (src/model.py:169-179, gnomix.py:55-58).
//
// Haplotype chunks stream through two device slots on two streams so that the H2D
// copy of chunk i+1 overlaps K1/K4 of chunk i and the D2H copy of chunk i-1.
#include <algorithm>
#include <mutex>

#include "gbt_smooth.cuh"
#include "lr_base.cuh"

namespace gnx {

struct Workspace {
    int device = -1;
    size_t x_bytes = 0, b_bytes = 0, p_bytes = 0, l_bytes = 0;
    int8_t* X[2] = {nullptr, nullptr};
    float* B[2] = {nullptr, nullptr};
    float* P[2] = {nullptr, nullptr};
    int32_t* Lb[2] = {nullptr, nullptr};
    cudaStream_t st[2] = {nullptr, nullptr};
    void release() {
        for (int i = 0; i < 2; i++) {
            if (X[i]) cudaFree(X[i]);
            if (B[i]) cudaFree(B[i]);
            if (P[i]) cudaFree(P[i]);
            if (Lb[i]) cudaFree(Lb[i]);
            X[i] = nullptr; B[i] = nullptr; P[i] = nullptr; Lb[i] = nullptr;
        }
        x_bytes = b_bytes = p_bytes = l_bytes = 0;
    }
};

static Workspace g_ws;
static std::mutex g_ws_mu;

static int ensure(Workspace& ws, size_t xb, size_t bb, size_t pb, size_t lb) {
    int dev = 0;
    GNX_CUDA(cudaGetDevice(&dev));
    if (ws.device != dev) {
        ws.release();
        for (int i = 0; i < 2; i++) {
            if (ws.st[i]) cudaStreamDestroy(ws.st[i]);
            GNX_CUDA(cudaStreamCreateWithFlags(&ws.st[i], cudaStreamNonBlocking));
        }
        ws.device = dev;
    }
    if (xb > ws.x_bytes || bb > ws.b_bytes || pb > ws.p_bytes || lb > ws.l_bytes) {
        ws.release();
        for (int i = 0; i < 2; i++) {
            GNX_CUDA(cudaMalloc((void**)&ws.X[i], xb));
            GNX_CUDA(cudaMalloc((void**)&ws.B[i], bb));
            if (pb) GNX_CUDA(cudaMalloc((void**)&ws.P[i], pb));
            GNX_CUDA(cudaMalloc((void**)&ws.Lb[i], lb));
        }
        ws.x_bytes = xb; ws.b_bytes = bb; ws.p_bytes = pb; ws.l_bytes = lb;
    }
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
    int64_t chunk = chunk_haps > 0 ? chunk_haps : std::max<int64_t>(256, (int64_t(1) << 31) / pitch / 256 * 256);  // ~2 GB of X per slot
    chunk = std::min<int64_t>(chunk, (N + 255) / 256 * 256);
    std::lock_guard<std::mutex> lock(g_ws_mu);
    Workspace& ws = g_ws;
    if (ensure(ws, (size_t)chunk * pitch, (size_t)chunk * W * A * sizeof(float), proba_host ? (size_t)chunk * W * A * sizeof(float) : 0,
               (size_t)chunk * W * sizeof(int32_t)))
        return 1;
    int it = 0;
    for (int64_t n0 = 0; n0 < N; n0 += chunk, it++) {
        const int s = it & 1;
        const int64_t n = std::min(chunk, N - n0);
        cudaStream_t st = ws.st[s];
        GNX_CUDA(cudaMemcpy2DAsync(ws.X[s], (size_t)pitch, X_host + n0 * ldX, (size_t)ldX, (size_t)C, (size_t)n, cudaMemcpyHostToDevice, st));
        if (gnx_lr_predict(lr, ws.X[s], n, pitch, ws.B[s], st)) return 1;
        if (gnx_gbt_smooth(gbt, ws.B[s], n, W, proba_host ? ws.P[s] : nullptr, ws.Lb[s], st)) return 1;
        GNX_CUDA(cudaMemcpyAsync(label_host + n0 * W, ws.Lb[s], (size_t)n * W * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (proba_host)
            GNX_CUDA(cudaMemcpyAsync(proba_host + n0 * W * A, ws.P[s], (size_t)n * W * A * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    GNX_CUDA(cudaStreamSynchronize(ws.st[0]));
    GNX_CUDA(cudaStreamSynchronize(ws.st[1]));
    return 0;
}
