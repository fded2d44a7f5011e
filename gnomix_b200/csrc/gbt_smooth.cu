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
    // VAR 0: lower u32 [T][12] | leaves [T][16] | top uint4 [T];  VAR 4 (accumulating offset): block u32 [T][16] | leaves [T][16]
    constexpr int LOWER_B = (VAR == 4) ? RK_BLOCK * 4 : RK_LOWER * 4;
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
                if constexpr (VAR == 4)
                    gbt_rank_walk_o<AT>(A, row, topc, lower_s, leaves_s, rounds, psum);
                else
                    gbt_rank_walk<AT>(A, row, top_s, lower_s, leaves_s, rounds, psum);
            }
            gbt_finish<AT>(m, psum, proba ? proba + (n * W + w) * A : nullptr, label ? label + n * W + w : nullptr);
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
    const size_t rb = rimg.size() * 4, tb = ((size_t)K * 4 + 15) & ~size_t(15);
    char* blob = nullptr;
    GNX_CUDA(cudaMalloc((void**)&blob, forest + 256 + rb + tb + 16));
    bool ok = cudaMemcpy(blob, nodes.data(), nb, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + nb, leaves.data(), lb, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + forest, base_margin, sizeof(float) * A, cudaMemcpyHostToDevice) == cudaSuccess;
    if (rb) ok &= cudaMemcpy(blob + forest + 256, rimg.data(), rb, cudaMemcpyHostToDevice) == cudaSuccess;
    if (K) ok &= cudaMemcpy(blob + forest + 256 + rb, tab.data(), (size_t)K * 4, cudaMemcpyHostToDevice) == cudaSuccess;
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
    m->h_tiletop = nullptr;
    m->rank_lut = nullptr;
    m->rank_tab = nullptr;
    m->profile = 0;
    m->ev[0] = m->ev[1] = m->ev[2] = nullptr;
    m->tile_forest = nullptr;
    m->tile_forest_bytes = 0;
    m->block_forest = nullptr;
    m->block_forest_bytes = 0;
    auto fail = [&](const char* what) {
        set_error("gnx_gbt_model_create: %s", what);
        gnx_gbt_model_destroy(m);
        return 1;
    };
    if (rank_ok) {
        const uint32_t* lower = rimg.data();
        const uint32_t* top = rimg.data() + (size_t)n_trees * (RK_LOWER + RK_LEAVES);
        if (n_trees <= GBT_TOPC_MAX_T) {   // tops for the parameter bank of the row kernel
            m->h_topc = new GbtTopC();
            for (int t = 0; t < n_trees; t++)
                for (int k = 0; k < 3; k++) m->h_topc->w[3 * t + k] = top[(size_t)t * 4 + k];
        }
        // block image (accumulating-offset walk of the row kernel): per tree 16 words -- level-2 node i2 at word
        // 4 * i2, level-3 node i3 at word 2 * i3 + 1 -- then the leaves of all trees
        std::vector<uint32_t> bimg((size_t)n_trees * (RK_BLOCK + RK_LEAVES), 0u);
        for (int t = 0; t < n_trees; t++) {
            for (int i2 = 0; i2 < 4; i2++) bimg[(size_t)t * RK_BLOCK + 4 * i2] = lower[(size_t)t * RK_LOWER + i2];
            for (int i3 = 0; i3 < 8; i3++) bimg[(size_t)t * RK_BLOCK + 2 * i3 + 1] = lower[(size_t)t * RK_LOWER + 4 + i3];
        }
        memcpy(bimg.data() + (size_t)n_trees * RK_BLOCK, rimg.data() + (size_t)n_trees * RK_LOWER, sizeof(uint32_t) * (size_t)n_trees * RK_LEAVES);
        void* d_b = nullptr;
        if (cudaMalloc(&d_b, bimg.size() * 4) != cudaSuccess) return fail("block image allocation failed");
        m->block_forest = static_cast<const unsigned char*>(d_b);
        m->block_forest_bytes = bimg.size() * 4;
        if (cudaMemcpy(d_b, bimg.data(), bimg.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail("block image copy failed");
        // tile image (gbt_tile.cu): node word (k << 17) | (0x8000 + (feature << 7)), so `pair word > node word` is the
        // split test and `word & 0x1ff80` the (biased) byte offset of the feature in a lane-interleaved tile; one 128-byte record per tree:
        // the same 16-word block, then its 16 leaves
        if (K <= GBT_TILE_MAX_K && F <= GBT_TILE_MAX_F && n_trees <= GBT_TILE_MAX_T) {
            auto conv = [&](uint32_t word) -> uint32_t {  // (k << 16 | byte offset in a rank row) -> tile word
                const uint32_t k = word >> 16, off = (word & 0xffffu) / 4u, slot = off / (uint32_t)astride, a = off % (uint32_t)astride;
                const uint32_t kk = (k == 0xFFFFu) ? (uint32_t)GBT_TILE_MAX_K : k;   // always-left filler: no rank exceeds it
                return (kk << 17) | (GBT_TILE_FBIAS + ((slot * (uint32_t)A + a) << 7));
            };
            std::vector<uint32_t> timg((size_t)n_trees * 32, 0u);
            for (int t = 0; t < n_trees; t++) {
                for (int i2 = 0; i2 < 4; i2++) timg[(size_t)t * 32 + 4 * i2] = conv(lower[(size_t)t * RK_LOWER + i2]);
                for (int i3 = 0; i3 < 8; i3++) timg[(size_t)t * 32 + 2 * i3 + 1] = conv(lower[(size_t)t * RK_LOWER + 4 + i3]);
                memcpy(timg.data() + (size_t)t * 32 + 16, rimg.data() + (size_t)n_trees * RK_LOWER + (size_t)t * RK_LEAVES, sizeof(uint32_t) * RK_LEAVES);
            }
            m->h_tiletop = new GbtTileTop();
            for (int t = 0; t < n_trees; t++) {
                const uint32_t t0 = conv(top[(size_t)t * 4]);
                m->h_tiletop->q[t] = make_uint4(t0, conv(top[(size_t)t * 4 + 1]), conv(top[(size_t)t * 4 + 2]), t0 & 0x1ff80u);
            }
            if (gbt_rank_lut_build(tab.data(), K, &m->rank_cells, &m->rank_lut, &m->rank_tab)) {
                gnx_gbt_model_destroy(m);
                return 1;
            }
            void* d_t = nullptr;
            if (cudaMalloc(&d_t, timg.size() * 4) != cudaSuccess) return fail("tile image allocation failed");
            m->tile_forest = static_cast<const unsigned char*>(d_t);
            m->tile_forest_bytes = timg.size() * 4;
            if (cudaMemcpy(d_t, timg.data(), timg.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail("tile image copy failed");
        }
    }
    m->variant = -1;  // chosen per call: tile kernel for batches of haplotypes, row kernel otherwise
    if (const char* e = getenv("GNX_GBT_VARIANT")) {  // profiling / cross-check switch
        const int v = atoi(e);
        if (v == 0 || (v == 4 && m->block_forest && m->h_topc) || (v == 6 && m->tile_forest)) m->variant = v;
    }
    *out = m;
    return 0;
}

void gnx_gbt_model_destroy(gnx_gbt_t* m) {
    if (!m) return;
    if (m->d_blob) cudaFree(m->d_blob);
    if (m->tile_forest) cudaFree(const_cast<unsigned char*>(m->tile_forest));
    if (m->block_forest) cudaFree(const_cast<unsigned char*>(m->block_forest));
    if (m->rank_lut) cudaFree(m->rank_lut);
    if (m->rank_tab) cudaFree(m->rank_tab);
    for (int i = 0; i < 3; i++)
        if (m->ev[i]) cudaEventDestroy(m->ev[i]);
    delete m->h_tiletop;
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
    // tile kernel (gbt_tile.cu): lanes = 32 haplotypes, so it wants a batch; single individuals (and the labels
    // pass of gnofix on few rows) take the row kernel, whose lanes are windows of one haplotype
    if (m->d.rank_ok && m->use_rank && m->tile_forest && (m->variant == 6 || (m->variant < 0 && N >= GBT_TILE_MIN_N))) {
        const int rc = gbt_tile_smooth(m, B_dev, N, W, proba_dev, label_dev, st);
        if (rc >= 0) return rc;
    }
    if (m->d.rank_ok && m->use_rank) {
        // fast path: units of (haplotype, segment of Lseg windows), G units per CTA pass.  One segment
        // per haplotype when the chromosome fits shared memory; (G, Lseg) chosen for the best fit of
        // rows to the 1024 threads.
        const size_t slot_bytes = (size_t)m->d.astride * 4;
        const int var = (m->variant == 0 || !m->block_forest || !m->h_topc) ? 0 : 4;
        const size_t img_bytes = (var == 4) ? m->block_forest_bytes : m->rank_forest_bytes;
        const unsigned char* img = (var == 4) ? m->block_forest : m->rank_forest;
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
#define CALLR(AT) LAUNCHR(AT, 4, GbtTopC, *m->h_topc)
            // the accumulating-offset walk (the production row kernel) is specialised per A; the first image's walk
            // (forests with more trees than the parameter bank holds, cross-checks) runs the generic-A instantiation
            if (var == 4) {
                GBT_DISPATCH_A(m->d.A, CALLR)
            } else {
                static const GbtTopC none{};
                LAUNCHR(0, 0, GbtTopC, none);
            }
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
    CALL(0);   // generic float traversal (forests deeper than 4 levels, cross-checks): one generic-A instantiation
#undef CALL
    GNX_CUDA(cudaGetLastError());
    return 0;
}

int gnx_gbt_set_kernel(gnx_gbt_t* m, int which) {
    GNX_REQUIRE(m != nullptr, "gnx_gbt_set_kernel: NULL model");
    GNX_REQUIRE(which == 0 || which == 1 || which == 10 || which == 14 || which == 16, "gnx_gbt_set_kernel: unknown kernel %d", which);
    m->use_rank = (which != 1);
    if (which == 0) m->variant = -1;
    if (which >= 10 && m->d.rank_ok) {  // not a rank-form forest: the generic kernel runs whatever is asked
        const int v = which - 10;
        GNX_REQUIRE(v == 0 || (v == 4 && m->block_forest && m->h_topc) || (v == 6 && m->tile_forest), "gnx_gbt_set_kernel: kernel %d not available for this forest", which);
        m->variant = v;
    }
    return 0;
}

int gnx_gbt_set_profile(gnx_gbt_t* m, int on) {
    GNX_REQUIRE(m != nullptr, "gnx_gbt_set_profile: NULL model");
    if (on)
        for (int i = 0; i < 3; i++)
            if (!m->ev[i]) GNX_CUDA(cudaEventCreate(&m->ev[i]));
    m->profile = on ? 1 : 0;
    return 0;
}

int gnx_gbt_last_phase_ms(const gnx_gbt_t* m, float* rank_ms, float* walk_ms) {
    GNX_REQUIRE(m != nullptr && m->ev[2], "gnx_gbt_last_phase_ms: profiling was not switched on (gnx_gbt_set_profile)");
    GNX_CUDA(cudaEventSynchronize(m->ev[2]));
    float a = 0.f, b = 0.f;
    GNX_CUDA(cudaEventElapsedTime(&a, m->ev[0], m->ev[1]));
    GNX_CUDA(cudaEventElapsedTime(&b, m->ev[1], m->ev[2]));
    if (rank_ms) *rank_ms = a;
    if (walk_ms) *walk_ms = b;
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
