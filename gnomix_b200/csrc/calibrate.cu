// calibrate.cu -- K7: Calibrator.transform of the smoother stage on the device.
//
// Replaces src/Smooth/Calibration.py:57-69 (per-class IsotonicRegression(out_of_bounds="clip")
// .transform) and the normalisation of lines 24-39, as applied by Smoother.predict_proba
// (src/Smooth/smooth.py:48-52).  scikit-learn evaluates the fitted step function in the dtype
// of its thresholds: float64 thresholds go through numpy.interp (slope form, float64), float32
// thresholds through scipy interp1d's two-weight formula in float32; both are followed
// operation by operation (explicit round-to-nearest intrinsics, no contraction), so the
// float64 output is bit-identical to the CPU path.  Thread = one (haplotype, window) row.
#include <vector>

#include "common.cuh"

struct gnx_cal {
    int A, is_f32, device;
    int total;
    int* d_off;      // [A+1] first threshold of class c
    double* d_x64;   // thresholds as float64 (exact for float32 models)
    double* d_y64;
    float* d_x32;    // float32 copies (float32 models only)
    float* d_y32;
    double inv_A;
};

namespace gnx {

int cal_classes(const gnx_cal* m) { return m->A; }

constexpr int CAL_MAX_A = 16;

__device__ __forceinline__ double cal_interp64(const double* __restrict__ xp, const double* __restrict__ fp, int n, double x) {
    if (n == 1) return fp[0];
    // np.clip then np.interp: j with xp[j] <= x < xp[j+1]
    x = fmin(fmax(x, xp[0]), xp[n - 1]);
    int lo = 0, hi = n;  // #{xp <= x}
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid + 1; else hi = mid;
    }
    const int j = lo - 1;
    if (j >= n - 1) return fp[n - 1];
    if (xp[j] == x) return fp[j];
    const double slope = __ddiv_rn(__dsub_rn(fp[j + 1], fp[j]), __dsub_rn(xp[j + 1], xp[j]));
    double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[j])), fp[j]);
    if (r != r) {
        r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[j + 1])), fp[j + 1]);
        if (r != r && fp[j] == fp[j + 1]) r = fp[j];
    }
    return r;
}

__device__ __forceinline__ float cal_interp32(const float* __restrict__ xp, const float* __restrict__ fp, int n, float x) {
    if (n == 1) return fp[0];
    x = fminf(fmaxf(x, xp[0]), xp[n - 1]);
    int lo = 0, hi = n;  // #{xp < x}  (searchsorted side='left')
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] < x) lo = mid + 1; else hi = mid;
    }
    const int h = min(max(lo, 1), n - 1), l = h - 1;
    const float d = __fsub_rn(xp[h], xp[l]);
    const float a = __fdiv_rn(__fsub_rn(x, xp[l]), d), b = __fdiv_rn(__fsub_rn(xp[h], x), d);
    return __fadd_rn(__fmul_rn(a, fp[h]), __fmul_rn(b, fp[l]));
}

// numpy's pairwise sum of a contiguous run of n <= 16 doubles (DOUBLE_pairwise_sum)
__device__ __forceinline__ double np_sum_small(const double* v, int n) {
    if (n < 8) {
        double s = 0.0;
        for (int i = 0; i < n; i++) s = __dadd_rn(s, v[i]);
        return s;
    }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = v[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], v[i + j]);
    double s = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; i++) s = __dadd_rn(s, v[i]);
    return s;
}

template <typename TIN>
__global__ void __launch_bounds__(256)
calibrate_kernel(gnx_cal m, const TIN* __restrict__ in, int64_t rows, double* __restrict__ out, int32_t* __restrict__ label) {
    extern __shared__ __align__(16) unsigned char smem[];
    // thresholds in shared memory: x64 | y64 | x32 | y32 | off
    double* x64 = reinterpret_cast<double*>(smem);
    double* y64 = x64 + m.total;
    float* x32 = reinterpret_cast<float*>(y64 + m.total);
    float* y32 = x32 + m.total;
    int* off = reinterpret_cast<int*>(y32 + m.total);
    for (int i = threadIdx.x; i < m.total; i += blockDim.x) {
        x64[i] = m.d_x64[i];
        y64[i] = m.d_y64[i];
        if (m.is_f32) {
            x32[i] = m.d_x32[i];
            y32[i] = m.d_y32[i];
        }
    }
    for (int i = threadIdx.x; i <= m.A; i += blockDim.x) off[i] = m.d_off[i];
    __syncthreads();
    const int A = m.A;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        double v[CAL_MAX_A];
        for (int c = 0; c < A; c++) {
            const int o = off[c], n = off[c + 1] - o;
            const TIN t = in[r * A + c];
            if (m.is_f32) {
                const float tf = (float)t;  // check_array(T, dtype=float32): round to nearest
                v[c] = (tf != tf) ? (double)tf : (double)cal_interp32(x32 + o, y32 + o, n, tf);
            } else {
                const double td = (double)t;
                v[c] = (td != td) ? td : cal_interp64(x64 + o, y64 + o, n, td);
            }
        }
        if (A == 2) {
            v[0] = __dsub_rn(1.0, v[1]);
        } else {
            const double s = np_sum_small(v, A);
            for (int c = 0; c < A; c++) v[c] = __ddiv_rn(v[c], s);
        }
        int best = 0;
        for (int c = 0; c < A; c++) {
            double p = v[c];
            if (p != p) p = m.inv_A;
            if (1.0 < p && p <= 1.0 + 1e-5) p = 1.0;
            v[c] = p;
            if (p > v[best]) best = c;
            if (out) out[r * A + c] = p;
        }
        if (label) label[r] = best;
    }
}

}  // namespace gnx

using namespace gnx;

extern "C" {

int gnx_cal_model_create(gnx_cal_t** out, int A, int is_f32, const int32_t* n_thr, const double* x_thr, const double* y_thr) {
    GNX_REQUIRE(out != nullptr, "gnx_cal_model_create: out is NULL");
    *out = nullptr;
    GNX_REQUIRE(A >= 2 && A <= CAL_MAX_A, "gnx_cal_model_create: A=%d unsupported (2..%d)", A, CAL_MAX_A);
    GNX_REQUIRE(n_thr && x_thr && y_thr, "gnx_cal_model_create: NULL array");
    if (require_blackwell()) return 1;
    std::vector<int> off(A + 1, 0);
    for (int c = 0; c < A; c++) {
        GNX_REQUIRE(n_thr[c] >= 1, "gnx_cal_model_create: class %d has no thresholds", c);
        off[c + 1] = off[c] + n_thr[c];
    }
    const int total = off[A];
    for (int c = 0; c < A; c++)
        for (int i = off[c] + 1; i < off[c + 1]; i++)
            GNX_REQUIRE(x_thr[i] > x_thr[i - 1], "gnx_cal_model_create: thresholds of class %d are not strictly increasing", c);
    const size_t smem = (size_t)total * 24 + (size_t)(A + 1) * 4;
    GNX_REQUIRE(smem <= 200 * 1024, "gnx_cal_model_create: %d thresholds do not fit shared memory", total);
    std::vector<float> x32(total), y32(total);
    for (int i = 0; i < total; i++) {
        x32[i] = (float)x_thr[i];
        y32[i] = (float)y_thr[i];
        if (is_f32)
            GNX_REQUIRE((double)x32[i] == x_thr[i] && (double)y32[i] == y_thr[i], "gnx_cal_model_create: is_f32 set but threshold %d is not a float32 value", i);
    }
    gnx_cal* m = new gnx_cal();
    m->A = A; m->is_f32 = is_f32 ? 1 : 0; m->total = total; m->inv_A = 1.0 / A;
    m->d_off = nullptr; m->d_x64 = m->d_y64 = nullptr; m->d_x32 = m->d_y32 = nullptr;
    cudaGetDevice(&m->device);
    bool ok = cudaMalloc((void**)&m->d_off, sizeof(int) * (A + 1)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&m->d_x64, sizeof(double) * total) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&m->d_y64, sizeof(double) * total) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&m->d_x32, sizeof(float) * total) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&m->d_y32, sizeof(float) * total) == cudaSuccess;
    ok = ok && cudaMemcpy(m->d_off, off.data(), sizeof(int) * (A + 1), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(m->d_x64, x_thr, sizeof(double) * total, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(m->d_y64, y_thr, sizeof(double) * total, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(m->d_x32, x32.data(), sizeof(float) * total, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(m->d_y32, y32.data(), sizeof(float) * total, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        gnx_cal_model_destroy(m);
        set_error("gnx_cal_model_create: device allocation / copy failed");
        return 1;
    }
    *out = m;
    return 0;
}

void gnx_cal_model_destroy(gnx_cal_t* m) {
    if (!m) return;
    cudaFree(m->d_off); cudaFree(m->d_x64); cudaFree(m->d_y64); cudaFree(m->d_x32); cudaFree(m->d_y32);
    delete m;
}

int gnx_calibrate(const gnx_cal_t* m, const void* proba_dev, int in_is_f32, int64_t rows, double* out_dev,
                  int32_t* label_dev, void* stream) {
    GNX_REQUIRE(m != nullptr, "gnx_calibrate: NULL model");
    GNX_REQUIRE(rows >= 0, "gnx_calibrate: bad row count");
    if (rows == 0) return 0;
    GNX_REQUIRE(proba_dev && (out_dev || label_dev), "gnx_calibrate: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)m->total * 24 + (size_t)(m->A + 1) * 4;
    const int grid = (int)std::min<int64_t>(ceil_div(rows, 256), (int64_t)sm_count() * 4);
    if (in_is_f32) {
        GNX_CUDA(cudaFuncSetAttribute(calibrate_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        calibrate_kernel<float><<<grid, 256, smem, st>>>(*m, static_cast<const float*>(proba_dev), rows, out_dev, label_dev);
    } else {
        GNX_CUDA(cudaFuncSetAttribute(calibrate_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        calibrate_kernel<double><<<grid, 256, smem, st>>>(*m, static_cast<const double*>(proba_dev), rows, out_dev, label_dev);
    }
    GNX_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
