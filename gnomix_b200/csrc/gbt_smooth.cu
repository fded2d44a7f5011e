// gbt_smooth.cu -- K4: XGB_Smoother.predict_proba / predict.
//
// Replaces slide_window (src/Smooth/utils.py:4-29) + XGBClassifier.predict_proba
// (src/Smooth/models.py:14-20 via src/Smooth/smooth.py:40-56) + argmax (smooth.py:61).
// X_slide [N*W, S*A] (150 GB at chr1 x 50k haplotypes) is never built: row (n,w) is
// the contiguous slice Bpad[n][w*A : w*A + S*A] of the reflect-padded base
// probabilities of haplotype n, which one CTA keeps in shared memory next to the
// whole forest.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "gbt_smooth.cuh"

namespace gnx {

// One CTA per haplotype (grid-stride).  Shared memory: [forest (optional)] [Bpad row].
template <int AT, bool FOREST_SMEM>
__global__ void __launch_bounds__(512, 1)
gbt_smooth_kernel(GbtDev m, size_t forest_bytes, const float* __restrict__ B, int64_t N, int W,
                  float* __restrict__ proba, int32_t* __restrict__ label) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int A = AT ? AT : m.A;
    const uint2* nodes = m.nodes;
    const float* leaves = m.leaves;
    float* bp = reinterpret_cast<float*>(smem);
    if (FOREST_SMEM) {
        const uint4* src = reinterpret_cast<const uint4*>(m.nodes);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (size_t i = threadIdx.x; i < forest_bytes / 16; i += blockDim.x) dst[i] = src[i];
        nodes = reinterpret_cast<const uint2*>(smem);
        leaves = reinterpret_cast<const float*>(smem + (size_t)m.T * m.n_split * sizeof(uint2));
        bp = reinterpret_cast<float*>(smem + forest_bytes);
    }
    const int pad = (m.S + 1) / 2;
    const int Wp = W + 2 * pad;
    for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
        __syncthreads();
        const float* bn = B + n * (int64_t)W * A;
        for (int idx = threadIdx.x; idx < Wp * A; idx += blockDim.x) {
            const int j = idx / A, a = idx - j * A;
            bp[idx] = __ldg(bn + (int64_t)spad_to_orig(j, W, pad) * A + a);
        }
        __syncthreads();
        for (int w = threadIdx.x; w < W; w += blockDim.x) {
            float psum[AT ? AT : GBT_MAX_A];
            gbt_eval_row<AT>(m, nodes, leaves, bp + (size_t)w * A, psum);
            gbt_finish<AT>(m, psum, proba ? proba + (n * W + w) * A : nullptr, label ? label + n * W + w : nullptr);
        }
    }
}


// ---------------------------------------------------------------- rank-form fast path
// Depth-4 forests.  Thread = one (haplotype, window) row; all lanes of a warp walk the
// same tree, so the tree's top three nodes are a broadcast load and every further step
// is: node word -> (byte offset of the feature in the row, threshold index) -> one
// integer compare against the pre-ranked input.  Rank rows hold rank << 16, node words
// are (k << 16 | byte offset), so `x < thr`  <=>  !(rank_word > node_word).
template <int AT, int VAR, typename TOPT>
__global__ void __launch_bounds__(RK_THREADS, 1)
gbt_smooth_rank_kernel(const __grid_constant__ TOPT topc, GbtDev m, const unsigned char* __restrict__ forest_img, size_t forest_bytes,
                       const float* __restrict__ B, int64_t N, int W, int G, int Lseg,
                       float* __restrict__ proba, int32_t* __restrict__ label) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int A = AT ? AT : m.A;
    constexpr int AMAX = AT ? AT : GBT_MAX_A;
    {
        const uint4* src = reinterpret_cast<const uint4*>(forest_img);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (size_t i = threadIdx.x; i < forest_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    // VAR 2 (wide): lower uint2 [T][12] | leaves [T][16];  VAR 0/1: lower u32 [T][12] | leaves | top uint4 [T]
    // VAR 4 (accumulating offset): block u32 [T][16] | leaves [T][16]
    constexpr int LOWER_B = (VAR == 2) ? RK_LOWER * 8 : (VAR == 4) ? RK_BLOCK * 4 : RK_LOWER * 4;
    const uint32_t* lower_s = reinterpret_cast<const uint32_t*>(smem);
    const float* leaves_s = reinterpret_cast<const float*>(smem + (size_t)m.T * LOWER_B);
    const uint4* top_s = reinterpret_cast<const uint4*>(smem + (size_t)m.T * (RK_LOWER + RK_LEAVES) * 4);
    (void)top_s;
    uint32_t* rk = reinterpret_cast<uint32_t*>(smem + forest_bytes);
    const int pad = (m.S + 1) / 2;
    const int ast = m.astride;
    // A unit is (haplotype, segment of Lseg windows); it needs the padded slots [w0, w0 + Lseg + S - 1).
    // Whole chromosomes that fit shared memory are one segment (nseg == 1).
    const int nseg = (W + Lseg - 1) / Lseg;
    const int Lslots = Lseg + m.S - 1;
    const int unit_words = Lslots * ast;
    const int64_t units = N * nseg;
    int* nan_flag = reinterpret_cast<int*>(rk + (size_t)G * unit_words);
    const int rounds = m.T / A;

    for (int64_t g0 = (int64_t)blockIdx.x * G; g0 < units; g0 += (int64_t)gridDim.x * G) {
        const int gn = (int)min((int64_t)G, units - g0);
        __syncthreads();
        if (threadIdx.x == 0) *nan_flag = 0;
        __syncthreads();
        bool saw_nan = false;
        for (int idx = threadIdx.x; idx < gn * Lslots * A; idx += blockDim.x) {
            const int h = idx / (Lslots * A), rem = idx - h * (Lslots * A);
            const int jl = rem / A, a = rem - jl * A;
            const int64_t u = g0 + h, n = u / nseg;
            const int j = (int)(u - n * nseg) * Lseg + jl;
            float x = 0.f;
            if (j < W + m.S - 1) x = __ldg(B + (n * W + spad_to_orig(j, W, pad)) * A + a);
            saw_nan |= (x != x);
            rk[h * unit_words + jl * ast + a] = gbt_rank_of(m.thr_table, m.K, x) << 16;
        }
        if (saw_nan) *nan_flag = 1;
        __syncthreads();
        const bool slow = (*nan_flag != 0);
        if (slow) {
            // NaN inputs follow each node's default child: generic traversal on float rows
            __syncthreads();
            float* bp = reinterpret_cast<float*>(rk);
            for (int idx = threadIdx.x; idx < gn * Lslots * A; idx += blockDim.x) {
                const int h = idx / (Lslots * A), rem = idx - h * (Lslots * A);
                const int jl = rem / A, a = rem - jl * A;
                const int64_t u = g0 + h, n = u / nseg;
                const int j = (int)(u - n * nseg) * Lseg + jl;
                float x = 0.f;
                if (j < W + m.S - 1) x = __ldg(B + (n * W + spad_to_orig(j, W, pad)) * A + a);
                bp[h * unit_words + jl * A + a] = x;
            }
            __syncthreads();
        }
        for (int r = threadIdx.x; r < gn * Lseg; r += blockDim.x) {
            const int h = r / Lseg, wl = r - h * Lseg;
            const int64_t u = g0 + h, n = u / nseg;
            const int w = (int)(u - n * nseg) * Lseg + wl;
            if (w >= W) continue;
            float psum[AMAX];
            if (slow) {
                gbt_eval_row<AT>(m, m.nodes, m.leaves, reinterpret_cast<const float*>(rk) + h * unit_words + wl * A, psum);
            } else {
                const unsigned char* row = reinterpret_cast<const unsigned char*>(rk + h * unit_words + wl * ast);
                if constexpr (VAR == 2)
                    gbt_rank_walk_w<AT>(A, row, topc, reinterpret_cast<const unsigned char*>(lower_s),
                                        reinterpret_cast<const unsigned char*>(leaves_s), rounds, psum);
                else if constexpr (VAR == 4)
                    gbt_rank_walk_o<AT>(A, row, topc, lower_s, leaves_s, rounds, psum);
                else if constexpr (VAR == 1)
                    gbt_rank_walk_c<AT>(A, row, topc, lower_s, leaves_s, rounds, psum);
                else
                    gbt_rank_walk<AT>(A, row, top_s, lower_s, leaves_s, rounds, psum);
            }
            gbt_finish<AT>(m, psum, proba ? proba + (n * W + w) * A : nullptr, label ? label + n * W + w : nullptr);
        }
    }
}

// ---------------------------------------------------------------- tile variant (two kernels)
// K4a: exact rank transform of the base probabilities, B f32 [n] -> R u16 [n]; NaN -> 0xFFFF.
__global__ void __launch_bounds__(512)
gbt_rank_u16_kernel(const float* __restrict__ thr, int K, int table_in_smem, const float* __restrict__ B, int64_t count,
                    uint16_t* __restrict__ R) {
    extern __shared__ __align__(16) unsigned char smem[];
    const float* tab = thr;
    if (table_in_smem) {
        float* t = reinterpret_cast<float*>(smem);
        for (int i = threadIdx.x; i < K; i += blockDim.x) t[i] = __ldg(thr + i);
        __syncthreads();
        tab = t;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = __ldg(B + i);
        int lo = 0, hi = K;  // #{j : tab[j] <= x}
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (tab[mid] <= x) lo = mid + 1; else hi = mid;
        }
        R[i] = (x != x) ? (uint16_t)0xFFFFu : (uint16_t)lo;
    }
}

// K4b: tile = 32 haplotypes (lane = haplotype) x Lseg windows (warp = window); the rank tile is
// lane-interleaved in shared memory (see gbt_rank_walk_t), the forest sits next to it.
template <int AT, bool BLOCK, typename TOPT>
__global__ void __launch_bounds__(RK_THREADS, 1)
gbt_smooth_tile_kernel(const __grid_constant__ TOPT topc, GbtDev m, const unsigned char* __restrict__ forest_img,
                       size_t forest_bytes, const uint16_t* __restrict__ R, const float* __restrict__ B, int64_t N, int W,
                       int nseg, int Lseg, float* __restrict__ proba, int32_t* __restrict__ label) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int A = AT ? AT : m.A;
    constexpr int AMAX = AT ? AT : GBT_MAX_A;
    {
        const uint4* src = reinterpret_cast<const uint4*>(forest_img);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (size_t i = threadIdx.x; i < forest_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    const uint32_t* lower_s = reinterpret_cast<const uint32_t*>(smem);
    const float* leaves_s = reinterpret_cast<const float*>(smem + (size_t)m.T * (BLOCK ? RK_BLOCK : RK_LOWER) * 4);
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem + forest_bytes);
    const int pad = (m.S + 1) / 2;
    const int Lslots = Lseg + m.S - 1;
    const int rounds = m.T / A;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nhb = (N + 31) / 32;
    for (int64_t t = blockIdx.x; t < nhb * nseg; t += gridDim.x) {
        const int64_t hb = t / nseg;
        const int sg = (int)(t - hb * nseg);
        const int64_t n = hb * 32 + lane;
        const int w0 = sg * Lseg;
        __syncthreads();
        int saw_nan = 0;
        for (int e = warp; e < Lslots * A; e += RK_THREADS / 32) {
            const int jl = e / A, a = e - jl * A;
            const int j = w0 + jl;
            uint32_t r = 0;
            if (n < N && j < W + m.S - 1) r = __ldg(R + (n * W + spad_to_orig(j, W, pad)) * A + a);
            saw_nan |= (r == 0xFFFFu);
            tile[e * 32 + lane] = r << 16;
        }
        const int slow = __syncthreads_or(saw_nan);
        if (slow) {
            // NaN inputs follow each node's default child: generic float traversal, one haplotype of the
            // tile at a time, its padded float row staged where the rank tile was
            float* bp = reinterpret_cast<float*>(tile);
            for (int h = 0; h < 32; h++) {
                const int64_t nh = hb * 32 + h;
                if (nh >= N) break;
                __syncthreads();
                for (int e = threadIdx.x; e < Lslots * A; e += blockDim.x) {
                    const int jl = e / A, a = e - jl * A;
                    const int j = w0 + jl;
                    bp[e] = (j < W + m.S - 1) ? __ldg(B + (nh * W + spad_to_orig(j, W, pad)) * A + a) : 0.f;
                }
                __syncthreads();
                for (int wl = threadIdx.x; wl < Lseg && w0 + wl < W; wl += blockDim.x) {
                    float psum[AMAX];
                    gbt_eval_row<AT>(m, m.nodes, m.leaves, bp + (size_t)wl * A, psum);
                    const int64_t row = nh * W + w0 + wl;
                    gbt_finish<AT>(m, psum, proba ? proba + row * A : nullptr, label ? label + row : nullptr);
                }
            }
            continue;
        }
        for (int wl = warp; wl < Lseg; wl += RK_THREADS / 32) {
            const int w = w0 + wl;
            if (w >= W) break;
            const uint32_t row = (uint32_t)__cvta_generic_to_shared(tile) + (uint32_t)(wl * A * 128 + lane * 4);
            float psum[AMAX];
            if constexpr (BLOCK) gbt_rank_walk_to<AT>(A, row, topc, lower_s, leaves_s, rounds, psum);
            else if constexpr (std::is_same<TOPT, GbtTopW>::value) gbt_rank_walk_t<AT>(A, row, topc, lower_s, leaves_s, rounds, psum);
            if (n < N) gbt_finish<AT>(m, psum, proba ? proba + (n * W + w) * A : nullptr, label ? label + n * W + w : nullptr);
        }
    }
}

// smoother.model.predict_proba(rows[k, F]) -- rows straight from global memory
template <int AT>
__global__ void gbt_rows_kernel(GbtDev m, const float* __restrict__ rows, int64_t k, float* __restrict__ proba) {
    const int A = AT ? AT : m.A;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= k) return;
    float psum[AT ? AT : GBT_MAX_A];
    gbt_eval_row<AT>(m, m.nodes, m.leaves, rows + r * m.F, psum);
    gbt_finish<AT>(m, psum, proba + r * A, nullptr);
}

static void fill_heap(const int32_t* feat, const float* thr, const int32_t* left, const int32_t* right,
                      const uint8_t* dl, const float* leaf, int node, int h, int depth, int D, uint2* nodes,
                      float* leaves, int n_split) {
    if (depth == D) {
        leaves[h - n_split] = leaf[node];
        return;
    }
    if (feat[node] < 0) {  // leaf above the bottom level: always-left filler, same leaf on both sides
        float inf = INFINITY;
        uint32_t bits;
        memcpy(&bits, &inf, 4);
        nodes[h] = make_uint2(0x80000000u, bits);
        fill_heap(feat, thr, left, right, dl, leaf, node, 2 * h + 1, depth + 1, D, nodes, leaves, n_split);
        fill_heap(feat, thr, left, right, dl, leaf, node, 2 * h + 2, depth + 1, D, nodes, leaves, n_split);
        return;
    }
    uint32_t bits;
    memcpy(&bits, &thr[node], 4);
    nodes[h] = make_uint2((uint32_t)feat[node] | ((dl && dl[node]) ? 0x80000000u : 0u), bits);
    fill_heap(feat, thr, left, right, dl, leaf, left[node], 2 * h + 1, depth + 1, D, nodes, leaves, n_split);
    fill_heap(feat, thr, left, right, dl, leaf, right[node], 2 * h + 2, depth + 1, D, nodes, leaves, n_split);
}

static int tree_depth(const int32_t* feat, const int32_t* left, const int32_t* right, int node, int n_nodes, int guard) {
    if (node < 0 || node >= n_nodes || guard > 64) return -1000;
    if (feat[node] < 0) return 0;
    int a = tree_depth(feat, left, right, left[node], n_nodes, guard + 1);
    int b = tree_depth(feat, left, right, right[node], n_nodes, guard + 1);
    return 1 + std::max(a, b);
}

}  // namespace gnx

using namespace gnx;

extern "C" {

int gnx_gbt_model_create(gnx_gbt_t** out, int A, int S, int n_trees, const int32_t* feat, const float* thr,
                         const int32_t* left, const int32_t* right, const uint8_t* default_left,
                         const float* leaf, const int32_t* tree_offsets, const float* base_margin) {
    GNX_REQUIRE(out != nullptr, "gnx_gbt_model_create: out is NULL");
    *out = nullptr;
    GNX_REQUIRE(A >= 2 && A <= GBT_MAX_A, "gnx_gbt_model_create: A=%d unsupported (2..%d)", A, GBT_MAX_A);
    GNX_REQUIRE(S >= 1 && (S & 1), "gnx_gbt_model_create: S=%d must be odd and positive", S);
    GNX_REQUIRE(n_trees > 0 && n_trees % A == 0, "gnx_gbt_model_create: n_trees=%d must be a positive multiple of A=%d", n_trees, A);
    GNX_REQUIRE(feat && thr && left && right && leaf && tree_offsets && base_margin, "gnx_gbt_model_create: NULL array");
    if (require_blackwell()) return 1;
    const int F = S * A;
    int D = 0;
    for (int t = 0; t < n_trees; t++) {
        const int o = tree_offsets[t], nn = tree_offsets[t + 1] - o;
        GNX_REQUIRE(nn >= 1, "gnx_gbt_model_create: tree %d is empty", t);
        const int d = tree_depth(feat + o, left + o, right + o, 0, nn, 0);
        GNX_REQUIRE(d >= 0, "gnx_gbt_model_create: tree %d is malformed", t);
        D = std::max(D, d);
        for (int i = 0; i < nn; i++)
            GNX_REQUIRE(feat[o + i] < F, "gnx_gbt_model_create: tree %d uses feature %d >= S*A=%d", t, feat[o + i], F);
    }
    if (D < 4) D = 4;  // shallower forests are padded with always-left fillers (enables the rank-form kernel)
    GNX_REQUIRE(D <= GBT_MAX_DEPTH, "gnx_gbt_model_create: depth %d > %d", D, GBT_MAX_DEPTH);
    const int n_split = (1 << D) - 1, n_leaf = 1 << D;
    std::vector<uint2> nodes((size_t)n_trees * n_split);
    std::vector<float> leaves((size_t)n_trees * n_leaf);
    for (int t = 0; t < n_trees; t++) {
        const int o = tree_offsets[t];
        fill_heap(feat + o, thr + o, left + o, right + o, default_left ? default_left + o : nullptr, leaf + o, 0, 0, 0, D,
                  nodes.data() + (size_t)t * n_split, leaves.data() + (size_t)t * n_leaf, n_split);
    }
    const size_t nb = nodes.size() * sizeof(uint2), lb = leaves.size() * sizeof(float);
    const size_t forest = (nb + lb + 15) & ~size_t(15);

    // rank form (see gbt_smooth.cuh): sorted distinct thresholds of the real splits
    std::vector<float> tab;
    for (int t = 0; t < n_trees; t++)
        for (int i = tree_offsets[t]; i < tree_offsets[t + 1]; i++)
            if (feat[i] >= 0) {
                GNX_REQUIRE(thr[i] == thr[i], "gnx_gbt_model_create: NaN split threshold in tree %d", t);
                tab.push_back(thr[i]);
            }
    std::sort(tab.begin(), tab.end());
    tab.erase(std::unique(tab.begin(), tab.end()), tab.end());
    const int K = (int)tab.size();
    const int astride = A | 1;
    const bool rank_ok = (D == 4) && K <= 65535 && (size_t)S * astride * 4 < 65536;
    std::vector<uint32_t> rimg;  // lower [T][12] | leaves [T][16] | top [T][4]
    if (rank_ok) {
        rimg.assign((size_t)n_trees * (RK_LOWER + RK_LEAVES + 4), 0u);
        uint32_t* lower = rimg.data();
        float* rleaves = reinterpret_cast<float*>(rimg.data() + (size_t)n_trees * RK_LOWER);
        uint32_t* top = rimg.data() + (size_t)n_trees * (RK_LOWER + RK_LEAVES);
        for (int t = 0; t < n_trees; t++) {
            for (int h = 0; h < n_split; h++) {
                const uint2 nd = nodes[(size_t)t * n_split + h];
                float th;
                memcpy(&th, &nd.y, 4);
                const uint32_t f = nd.x & 0x7fffffffu;
                uint32_t k;
                if (th == INFINITY && (nd.x >> 31) && f == 0 && !std::binary_search(tab.begin(), tab.end(), th))
                    k = 0xFFFFu;  // always-left filler
                else if (th == INFINITY && !std::binary_search(tab.begin(), tab.end(), th))
                    k = 0xFFFFu;
                else
                    k = (uint32_t)(std::lower_bound(tab.begin(), tab.end(), th) - tab.begin());
                const uint32_t off = ((f / A) * astride + (f % A)) * 4u;
                const uint32_t word = (k << 16) | off;
                if (h < 3) top[(size_t)t * 4 + h] = word;
                else lower[(size_t)t * RK_LOWER + (h - 3)] = word;
            }
            memcpy(rleaves + (size_t)t * RK_LEAVES, leaves.data() + (size_t)t * n_leaf, sizeof(float) * RK_LEAVES);
        }
    }
    // wide image: lower uint2 [T][12] | leaves [T][16]
    std::vector<uint32_t> wimg;
    if (rank_ok) {
        wimg.assign((size_t)n_trees * (2 * RK_LOWER + RK_LEAVES), 0u);
        const uint32_t* lower = rimg.data();
        for (size_t i = 0; i < (size_t)n_trees * RK_LOWER; i++) {
            wimg[2 * i] = lower[i] & 0xffff0000u;
            wimg[2 * i + 1] = lower[i] & 0xffffu;
        }
        memcpy(wimg.data() + (size_t)n_trees * 2 * RK_LOWER, rimg.data() + (size_t)n_trees * RK_LOWER,
               sizeof(uint32_t) * (size_t)n_trees * RK_LEAVES);
    }
    const size_t wb = (wimg.size() * 4 + 15) & ~size_t(15);
    const size_t rb = rimg.size() * 4, tb = ((size_t)K * 4 + 15) & ~size_t(15);
    char* blob = nullptr;
    GNX_CUDA(cudaMalloc((void**)&blob, forest + 256 + rb + tb + wb + 16));
    bool ok = cudaMemcpy(blob, nodes.data(), nb, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + nb, leaves.data(), lb, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + forest, base_margin, sizeof(float) * A, cudaMemcpyHostToDevice) == cudaSuccess;
    if (rb) ok &= cudaMemcpy(blob + forest + 256, rimg.data(), rb, cudaMemcpyHostToDevice) == cudaSuccess;
    if (K) ok &= cudaMemcpy(blob + forest + 256 + rb, tab.data(), (size_t)K * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (wb) ok &= cudaMemcpy(blob + forest + 256 + rb + tb, wimg.data(), wimg.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        cudaFree(blob);
        set_error("gnx_gbt_model_create: H2D copy failed");
        return 1;
    }
    gnx_gbt* m = new gnx_gbt();
    cudaGetDevice(&m->device);
    m->d_blob = blob;
    m->forest_bytes = forest;
    m->d = GbtDev{A, S, n_trees, D, F, n_split, n_leaf, reinterpret_cast<const uint2*>(blob),
                  reinterpret_cast<const float*>(blob + nb), reinterpret_cast<const float*>(blob + forest),
                  rank_ok ? 1 : 0, K, astride, reinterpret_cast<const float*>(blob + forest + 256 + rb),
                  reinterpret_cast<const uint4*>(blob + forest + 256 + (size_t)n_trees * (RK_LOWER + RK_LEAVES) * 4),
                  reinterpret_cast<const uint32_t*>(blob + forest + 256)};
    m->rank_forest_bytes = rb;
    m->rank_forest = reinterpret_cast<const unsigned char*>(blob + forest + 256);
    m->use_rank = 1;
    m->h_topc = nullptr;
    if (rank_ok && n_trees <= GBT_TOPC_MAX_T) {
        m->h_topc = new GbtTopC();
        const uint32_t* top = rimg.data() + (size_t)n_trees * (RK_LOWER + RK_LEAVES);
        for (int t = 0; t < n_trees; t++)
            for (int k = 0; k < 3; k++) m->h_topc->w[3 * t + k] = top[(size_t)t * 4 + k];
    }
    // tile image: the narrow image with the feature index in the low half of every node word
    m->h_topt = nullptr;
    m->h_toptn = nullptr;
    m->tile_forest = nullptr;
    m->tile_forest_bytes = 0;
    m->tblock_forest = nullptr;
    m->tblock_forest_bytes = 0;
    if (rank_ok && n_trees <= GBT_TOPW_MAX_T && K <= 65534 && F <= 65535) {
        std::vector<uint32_t> timg((size_t)n_trees * (RK_LOWER + RK_LEAVES), 0u);
        m->h_topt = new GbtTopW();
        m->h_toptn = new GbtTopC();
        auto conv = [&](uint32_t word) -> uint32_t {  // (k << 16 | byte offset in a row) -> (k << 16 | feature index)
            const uint32_t off = (word & 0xffffu) / 4u, slot = off / (uint32_t)astride, a = off % (uint32_t)astride;
            return (word & 0xffff0000u) | (slot * (uint32_t)A + a);
        };
        const uint32_t* lower = rimg.data();
        const uint32_t* top = rimg.data() + (size_t)n_trees * (RK_LOWER + RK_LEAVES);
        for (size_t i = 0; i < (size_t)n_trees * RK_LOWER; i++) timg[i] = conv(lower[i]);
        memcpy(timg.data() + (size_t)n_trees * RK_LOWER, rimg.data() + (size_t)n_trees * RK_LOWER, sizeof(uint32_t) * (size_t)n_trees * RK_LEAVES);
        for (int t = 0; t < n_trees; t++)
            for (int k = 0; k < 3; k++) {
                const uint32_t wd = conv(top[(size_t)t * 4 + k]);
                m->h_topt->w[3 * t + k] = make_uint2(wd & 0xffff0000u, (wd & 0xffffu) * 128u);
                m->h_toptn->w[3 * t + k] = wd;
            }
        void* d_t = nullptr;
        if (cudaMalloc(&d_t, timg.size() * 4) != cudaSuccess || cudaMemcpy(d_t, timg.data(), timg.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("gnx_gbt_model_create: tile image allocation failed");
            return 1;
        }
        m->tile_forest = static_cast<const unsigned char*>(d_t);
        m->tile_forest_bytes = timg.size() * 4;
        std::vector<uint32_t> tb((size_t)n_trees * (RK_BLOCK + RK_LEAVES), 0u);
        for (int t = 0; t < n_trees; t++) {
            for (int i2 = 0; i2 < 4; i2++) tb[(size_t)t * RK_BLOCK + 4 * i2] = timg[(size_t)t * RK_LOWER + i2];
            for (int i3 = 0; i3 < 8; i3++) tb[(size_t)t * RK_BLOCK + 2 * i3 + 1] = timg[(size_t)t * RK_LOWER + 4 + i3];
        }
        memcpy(tb.data() + (size_t)n_trees * RK_BLOCK, timg.data() + (size_t)n_trees * RK_LOWER, sizeof(uint32_t) * (size_t)n_trees * RK_LEAVES);
        void* d_tb = nullptr;
        if (cudaMalloc(&d_tb, tb.size() * 4) != cudaSuccess || cudaMemcpy(d_tb, tb.data(), tb.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("gnx_gbt_model_create: tile block image allocation failed");
            return 1;
        }
        m->tblock_forest = static_cast<const unsigned char*>(d_tb);
        m->tblock_forest_bytes = tb.size() * 4;
    }
    // block image (accumulating-offset variant): per tree 16 words -- level-2 node i2 at word 4 * i2, level-3 node
    // i3 at word 2 * i3 + 1 -- then the leaves
    m->block_forest = nullptr;
    m->block_forest_bytes = 0;
    if (rank_ok && n_trees <= GBT_TOPC_MAX_T) {
        std::vector<uint32_t> bimg((size_t)n_trees * (RK_BLOCK + RK_LEAVES), 0u);
        const uint32_t* lower = rimg.data();
        for (int t = 0; t < n_trees; t++) {
            for (int i2 = 0; i2 < 4; i2++) bimg[(size_t)t * RK_BLOCK + 4 * i2] = lower[(size_t)t * RK_LOWER + i2];
            for (int i3 = 0; i3 < 8; i3++) bimg[(size_t)t * RK_BLOCK + 2 * i3 + 1] = lower[(size_t)t * RK_LOWER + 4 + i3];
        }
        memcpy(bimg.data() + (size_t)n_trees * RK_BLOCK, rimg.data() + (size_t)n_trees * RK_LOWER, sizeof(uint32_t) * (size_t)n_trees * RK_LEAVES);
        void* d_b = nullptr;
        if (cudaMalloc(&d_b, bimg.size() * 4) != cudaSuccess || cudaMemcpy(d_b, bimg.data(), bimg.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("gnx_gbt_model_create: block image allocation failed");
            return 1;
        }
        m->block_forest = static_cast<const unsigned char*>(d_b);
        m->block_forest_bytes = bimg.size() * 4;
    }
    m->h_topw = nullptr;
    m->wide_forest = reinterpret_cast<const unsigned char*>(blob + forest + 256 + rb + tb);
    m->wide_forest_bytes = wb;
    if (rank_ok && n_trees <= GBT_TOPW_MAX_T) {
        m->h_topw = new GbtTopW();
        const uint32_t* top = rimg.data() + (size_t)n_trees * (RK_LOWER + RK_LEAVES);
        for (int t = 0; t < n_trees; t++)
            for (int k = 0; k < 3; k++) m->h_topw->w[3 * t + k] = make_uint2(top[(size_t)t * 4 + k] & 0xffff0000u, top[(size_t)t * 4 + k] & 0xffffu);
    }
    // measured on B200 (chr1 x 20 000 haplotypes, scripts/k4_probe.py): accumulating-offset block layout 22.7 ms,
    // one-word nodes + byte-offset walk 23.7 ms, two-word nodes 25.4 ms (LDS.64 costs two shared-memory wavefronts
    // and the kernel is wavefront / issue co-limited), lane-interleaved tiles 25.7 ms, tiles + block layout 26.8 ms
    m->variant = (m->block_forest && m->h_topc) ? 4 : (m->h_topc ? 1 : 0);
    if (const char* e = getenv("GNX_GBT_VARIANT")) {  // profiling / cross-check switch
        const int v = atoi(e);
        if (v == 0 || (v == 1 && m->h_topc) || (v == 2 && m->h_topw) || (v == 3 && m->h_topt) || (v == 4 && m->block_forest) || ((v == 5 || v == 6) && m->tblock_forest)) m->variant = v;
    }
    *out = m;
    return 0;
}

void gnx_gbt_model_destroy(gnx_gbt_t* m) {
    if (!m) return;
    if (m->d_blob) cudaFree(m->d_blob);
    if (m->tile_forest) cudaFree(const_cast<unsigned char*>(m->tile_forest));
    if (m->block_forest) cudaFree(const_cast<unsigned char*>(m->block_forest));
    if (m->tblock_forest) cudaFree(const_cast<unsigned char*>(m->tblock_forest));
    delete m->h_topt;
    delete m->h_toptn;
    delete m->h_topw;
    delete m->h_topc;
    delete m;
}

#define GBT_DISPATCH_A(A, CALL) \
    switch (A) {                \
        case 2: CALL(2); break; \
        case 3: CALL(3); break; \
        case 4: CALL(4); break; \
        case 5: CALL(5); break; \
        case 6: CALL(6); break; \
        case 7: CALL(7); break; \
        case 8: CALL(8); break; \
        default: CALL(0); break; \
    }

int gnx_gbt_smooth(const gnx_gbt_t* m, const float* B_dev, int64_t N, int W, float* proba_dev, int32_t* label_dev,
                   void* stream) {
    GNX_REQUIRE(m != nullptr, "gnx_gbt_smooth: NULL model");
    GNX_REQUIRE(N >= 0 && W >= 1, "gnx_gbt_smooth: bad shape N=%lld W=%d", (long long)N, W);
    const int pad = (m->d.S + 1) / 2;
    GNX_REQUIRE(W >= pad, "gnx_gbt_smooth: W=%d smaller than the reflect pad %d", W, pad);
    if (N == 0) return 0;
    GNX_REQUIRE(B_dev && (proba_dev || label_dev), "gnx_gbt_smooth: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bp_bytes = (size_t)(W + 2 * pad) * m->d.A * sizeof(float);
    const size_t smem_max = 227 * 1024;
    if (m->d.rank_ok && m->use_rank && (m->variant == 3 || m->variant == 5 || m->variant == 6) && m->h_topt) {
        const bool blockv = (m->variant >= 5);
        const bool narrow_top = (m->variant == 6);
        const unsigned char* timgp = blockv ? m->tblock_forest : m->tile_forest;
        const size_t timgb = blockv ? m->tblock_forest_bytes : m->tile_forest_bytes;
        // tile variant: K4a rank transform into a stream-ordered scratch buffer, K4b tile kernel
        const size_t slot_bytes = (size_t)m->d.A * 128;
        const size_t room = smem_max - timgb - 16;
        const int Lmax = (int)std::min<int64_t>((int64_t)(room / slot_bytes) - (m->d.S - 1), 2 * (RK_THREADS / 32));
        if (Lmax >= 32 || Lmax >= W) {
            const int nseg = (int)ceil_div(W, std::min(Lmax, W));
            const int Lseg = (int)ceil_div(W, nseg);
            const size_t smem = timgb + (size_t)(Lseg + m->d.S - 1) * slot_bytes + 16;
            const int64_t count = N * (int64_t)W * m->d.A;
            uint16_t* R = nullptr;
            GNX_CUDA(cudaMallocAsync((void**)&R, (size_t)count * sizeof(uint16_t), st));
            const int in_smem = (size_t)m->d.K * 4 <= 160 * 1024;
            const size_t rsm = in_smem ? (size_t)m->d.K * 4 : 0;
            GNX_CUDA(cudaFuncSetAttribute(gbt_rank_u16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(160 * 1024)));
            gbt_rank_u16_kernel<<<(int)std::min<int64_t>(ceil_div(count, 512), (int64_t)sm_count() * 4), 512, rsm, st>>>(
                m->d.thr_table, m->d.K, in_smem, B_dev, count, R);
            const int64_t tiles = ceil_div(N, 32) * nseg;
            const int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count());
#define CALLT(AT)                                                                                                              \
    do {                                                                                                                       \
        if (narrow_top) {                                                                                                      \
            GNX_CUDA(cudaFuncSetAttribute(gbt_smooth_tile_kernel<AT, true, GbtTopC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            gbt_smooth_tile_kernel<AT, true, GbtTopC><<<grid, RK_THREADS, smem, st>>>(*m->h_toptn, m->d, timgp, timgb, R, B_dev, N, W, nseg, \
                                                                                      Lseg, proba_dev, label_dev);            \
        } else if (blockv) {                                                                                                   \
            GNX_CUDA(cudaFuncSetAttribute(gbt_smooth_tile_kernel<AT, true, GbtTopW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            gbt_smooth_tile_kernel<AT, true, GbtTopW><<<grid, RK_THREADS, smem, st>>>(*m->h_topt, m->d, timgp, timgb, R, B_dev, N, W, nseg, \
                                                                                      Lseg, proba_dev, label_dev);            \
        } else {                                                                                                               \
            GNX_CUDA(cudaFuncSetAttribute(gbt_smooth_tile_kernel<AT, false, GbtTopW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            gbt_smooth_tile_kernel<AT, false, GbtTopW><<<grid, RK_THREADS, smem, st>>>(*m->h_topt, m->d, timgp, timgb, R, B_dev, N, W, nseg, \
                                                                                       Lseg, proba_dev, label_dev);           \
        }                                                                                                                      \
    } while (0)
            GBT_DISPATCH_A(m->d.A, CALLT)
#undef CALLT
            GNX_CUDA(cudaGetLastError());
            GNX_CUDA(cudaFreeAsync(R, st));
            return 0;
        }
    }
    if (m->d.rank_ok && m->use_rank) {
        // fast path: units of (haplotype, segment of Lseg windows), G units per CTA pass.  One segment
        // per haplotype when the chromosome fits shared memory; (G, Lseg) chosen for the best fit of
        // rows to the 1024 threads.
        const size_t slot_bytes = (size_t)m->d.astride * 4;
        const int var = (m->variant == 3 || m->variant >= 5) ? (m->h_topc ? 1 : 0) : m->variant;  // tile variants not applicable here
        const size_t img_bytes = (var == 2) ? m->wide_forest_bytes : (var == 4) ? m->block_forest_bytes : m->rank_forest_bytes;
        const unsigned char* img = (var == 2) ? m->wide_forest : (var == 4) ? m->block_forest : m->rank_forest;
        const size_t room = smem_max - img_bytes - 16;
        int bestG = 0, bestL = 0;
        double best_eff = 0.0;
        for (int nseg = 1; nseg <= 64 && bestG == 0; nseg++) {
            const int L = (int)ceil_div(W, nseg);
            const size_t unit_bytes = (size_t)(L + m->d.S - 1) * slot_bytes;
            if (unit_bytes > room) continue;
            for (int G = 1; G <= 8; G++) {
                if ((size_t)G * unit_bytes > room) break;
                const int64_t rows = (int64_t)G * L;
                // useful rows per pass over the thread block, discounted by the halo that is re-ranked per segment
                const double eff = (double)rows / (double)(ceil_div(rows, RK_THREADS) * RK_THREADS) * (double)L / (double)(L + m->d.S - 1);
                if (eff > best_eff + 0.02) { best_eff = eff; bestG = G; bestL = L; }
            }
        }
        if (bestG > 0) {
            const size_t smem = img_bytes + (size_t)bestG * (bestL + m->d.S - 1) * slot_bytes + 16;
            const int64_t units = N * ceil_div(W, bestL);
            const int grid = (int)std::min<int64_t>(ceil_div(units, bestG), (int64_t)sm_count());
#define LAUNCHR(AT, VAR, TOPT, TOPV)                                                                                           \
    do {                                                                                                                       \
        GNX_CUDA(cudaFuncSetAttribute(gbt_smooth_rank_kernel<AT, VAR, TOPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gbt_smooth_rank_kernel<AT, VAR, TOPT><<<grid, RK_THREADS, smem, st>>>(TOPV, m->d, img, img_bytes, B_dev, N, W, bestG, bestL, \
                                                                              proba_dev, label_dev);                           \
    } while (0)
#define CALLR(AT)                                                  \
    do {                                                           \
        if (var == 4) {                                            \
            LAUNCHR(AT, 4, GbtTopC, *m->h_topc);                   \
        } else if (var == 2) {                                     \
            LAUNCHR(AT, 2, GbtTopW, *m->h_topw);                   \
        } else if (var == 1) {                                     \
            LAUNCHR(AT, 1, GbtTopC, *m->h_topc);                   \
        } else {                                                   \
            static const GbtTopC none{};                           \
            LAUNCHR(AT, 0, GbtTopC, none);                         \
        }                                                          \
    } while (0)
            GBT_DISPATCH_A(m->d.A, CALLR)
#undef CALLR
#undef LAUNCHR
            GNX_CUDA(cudaGetLastError());
            return 0;
        }
    }
    GNX_REQUIRE(bp_bytes <= smem_max, "gnx_gbt_smooth: W=%d too large for the generic kernel's shared-memory row (forests deeper than 4 levels)", W);
    const bool forest_smem = m->forest_bytes + bp_bytes <= smem_max;
    const size_t smem = bp_bytes + (forest_smem ? m->forest_bytes : 0);
    const int grid = (int)std::min<int64_t>(N, (int64_t)sm_count() * (forest_smem ? 1 : 2));
#define CALL(AT)                                                                                                   \
    do {                                                                                                           \
        if (forest_smem) {                                                                                         \
            GNX_CUDA(cudaFuncSetAttribute(gbt_smooth_kernel<AT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            gbt_smooth_kernel<AT, true><<<grid, 512, smem, st>>>(m->d, m->forest_bytes, B_dev, N, W, proba_dev, label_dev);      \
        } else {                                                                                                   \
            GNX_CUDA(cudaFuncSetAttribute(gbt_smooth_kernel<AT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            gbt_smooth_kernel<AT, false><<<grid, 512, smem, st>>>(m->d, m->forest_bytes, B_dev, N, W, proba_dev, label_dev);     \
        }                                                                                                          \
    } while (0)
    GBT_DISPATCH_A(m->d.A, CALL)
#undef CALL
    GNX_CUDA(cudaGetLastError());
    return 0;
}

int gnx_gbt_set_kernel(gnx_gbt_t* m, int which) {
    GNX_REQUIRE(m != nullptr, "gnx_gbt_set_kernel: NULL model");
    GNX_REQUIRE(which == 0 || which == 1 || (which >= 10 && which <= 16), "gnx_gbt_set_kernel: unknown kernel %d", which);
    if (which >= 10) {  // rank-form flavour: 10 narrow nodes, 11 narrow + parameter-bank tops, 12 wide nodes
        const int v = which - 10;
        m->use_rank = 1;
        if (!m->d.rank_ok) return 0;  // not a rank-form forest: the generic kernel runs whatever the flavour
        GNX_REQUIRE(v == 0 || (v == 1 && m->h_topc) || (v == 2 && m->h_topw) || (v == 3 && m->h_topt) || (v == 4 && m->block_forest) || ((v == 5 || v == 6) && m->tblock_forest), "gnx_gbt_set_kernel: flavour %d not available for this forest", v);
        m->variant = v;
        m->use_rank = 1;
        return 0;
    }
    m->use_rank = (which == 0);
    return 0;
}

int gnx_gbt_rows(const gnx_gbt_t* m, const float* rows_dev, int64_t k, float* proba_dev, void* stream) {
    GNX_REQUIRE(m != nullptr, "gnx_gbt_rows: NULL model");
    GNX_REQUIRE(k >= 0, "gnx_gbt_rows: bad k");
    if (k == 0) return 0;
    GNX_REQUIRE(rows_dev && proba_dev, "gnx_gbt_rows: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)ceil_div(k, 128);
#define CALL(AT) gbt_rows_kernel<AT><<<grid, 128, 0, st>>>(m->d, rows_dev, k, proba_dev)
    GBT_DISPATCH_A(m->d.A, CALL)
#undef CALL
    GNX_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
