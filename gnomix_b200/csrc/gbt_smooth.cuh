// gbt_smooth.cuh -- model handle of the tree-ensemble smoother (K4), shared with gnofix (K6)
#pragma once

#include <math.h>

#include "common.cuh"

namespace gnx {

constexpr int GBT_MAX_A = 16;
constexpr int GBT_MAX_DEPTH = 8;
constexpr int RK_THREADS = 1024;
constexpr int RK_LOWER = 12;   // nodes 3..14 of a depth-4 heap
constexpr int RK_LEAVES = 16;
// tile kernel (gbt_tile.cu): node word = (k << 17) | (GBT_TILE_FBIAS + (feature << 7))
constexpr int GBT_TILE_MAX_K = 32767;   // threshold indices (and ranks) must fit 15 bits
constexpr int GBT_TILE_MAX_F = 767;     // GBT_TILE_FBIAS + (feature << 7) must stay below bit 17
constexpr uint32_t GBT_TILE_FBIAS = 0x8000u;   // added to the feature field of a tile node word (gbt_tile.cu: pair words)
constexpr int GBT_TILE_MAX_T = 1792;    // tree tops travel in the kernel parameter bank (28 KB of the 32 KB)
constexpr int GBT_TILE_MIN_N = 24;      // below this many haplotypes the row kernel is used (lanes = haplotypes here)
constexpr int GBT_RANK_CELLS = 8192;    // equal cells over the threshold range (rank pass of the tile kernel)

// Heap-ordered complete forest.  Tree t: (2^D - 1) split nodes then 2^D leaves.
// Shallower subtrees are padded with always-left splits whose both children carry
// the same leaf, so every traversal takes exactly D steps and ends on the value the
// original tree would have returned.
struct GbtDev {
    int A, S, T, D, F;       // classes, smoother width, trees, depth, features (= S*A)
    int n_split, n_leaf;     // 2^D - 1, 2^D
    const uint2* nodes;      // [T][n_split]  .x = feature index | default_left << 31, .y = float bits of split_cond
    const float* leaves;     // [T][n_leaf]
    const float* base;       // [A]
    // ---- rank form (fast path of gbt_smooth, depth-4 forests) ----------------------
    // Every split threshold is replaced by its index k in the sorted table of distinct
    // thresholds and every input value x by rank(x) = #{j : thr_table[j] <= x}; then
    // `x < thr_table[k]`  <=>  `rank(x) <= k` exactly, for every float x that is not NaN.
    int rank_ok;             // 1 if the fast path is usable (D == 4, K <= 65535)
    int K;                   // distinct thresholds
    int astride;             // odd row stride (in words) of the rank rows in shared memory
    const float* thr_table;  // [K] ascending
    const uint4* top;        // [T] .x .y .z = nodes 0,1,2 as (offset << 16 | k), .w unused
    const uint32_t* lower;   // [T][12] nodes 3..14 as (offset << 16 | k)
};

#ifdef __CUDACC__
__device__ __forceinline__ int spad_to_orig(int j, int W, int pad) {
    if (j < pad) return pad - 1 - j;
    if (j >= pad + W) return W - 1 - (j - pad - W);
    return j - pad;
}

// margins (float32, tree order, class = t % A) -> xgboost Softmax -> argmax
template <int AT>
__device__ __forceinline__ void gbt_finish(const GbtDev& m, float* psum, float* __restrict__ proba_out,
                                           int32_t* __restrict__ label_out) {
    const int A = AT ? AT : m.A;
    constexpr int AMAX = AT ? AT : GBT_MAX_A;
    float wmax = -INFINITY;
#pragma unroll
    for (int c = 0; c < AMAX; c++)
        if (c < A) {
            psum[c] = GNX_FADD(__ldg(m.base + c), psum[c]);
            wmax = (c == 0) ? psum[0] : fmaxf(psum[c], wmax);
        }
    double wsum = 0.0;
#pragma unroll
    for (int c = 0; c < AMAX; c++)
        if (c < A) {
            psum[c] = gnx_expf_cr(GNX_FSUB(psum[c], wmax));
            wsum = GNX_ADD(wsum, GNX_F2D(psum[c]));
        }
    const float ws = GNX_D2F(wsum);
    int best = 0;
    float pbest = 0.f;
#pragma unroll
    for (int c = 0; c < AMAX; c++)
        if (c < A) {
            const float p = GNX_FDIV(psum[c], ws);
            if (proba_out) proba_out[c] = p;
            if (c == 0 || p > pbest) {
                pbest = p;
                best = c;
            }
        }
    if (label_out) *label_out = best;
}

template <int AT>
__device__ __forceinline__ void gbt_eval_row(const GbtDev& m, const uint2* __restrict__ nodes,
                                             const float* __restrict__ leaves, const float* __restrict__ row,
                                             float* psum) {
    const int A = AT ? AT : m.A;
    constexpr int AMAX = AT ? AT : GBT_MAX_A;
#pragma unroll
    for (int c = 0; c < AMAX; c++) psum[c] = 0.f;
    const int rounds = m.T / A;
    for (int r = 0; r < rounds; r++) {
#pragma unroll
        for (int c = 0; c < AMAX; c++) {
            if (c < A) {
                const int t = r * A + c;
                const uint2* tn = nodes + (size_t)t * m.n_split;
                int nid = 0;
                for (int d = 0; d < m.D; d++) {
                    const uint2 nd = tn[nid];
                    const float x = row[nd.x & 0x7fffffffu];
                    const bool left = (x != x) ? (nd.x >> 31) : (x < __uint_as_float(nd.y));
                    nid = 2 * nid + (left ? 1 : 2);
                }
                psum[c] = GNX_FADD(psum[c], leaves[(size_t)t * m.n_leaf + (nid - m.n_split)]);
            }
        }
    }
}

__device__ __forceinline__ uint32_t gbt_rank_of(const float* __restrict__ tab, int K, float x) {
    // #{j : tab[j] <= x}
    int lo = 0, hi = K;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(tab + mid) <= x) lo = mid + 1; else hi = mid;
    }
    return (uint32_t)lo;
}


// Rank-form walk of the whole forest for one row (see GbtDev): `row` points at the row's
// first rank word (rank << 16, byte-addressed); top/lower/leaves are the forest image
// (shared memory in the kernels).  psum[c] accumulates class c's leaves in tree order.
template <int AT>
__device__ __forceinline__ void gbt_rank_walk(int A, const unsigned char* __restrict__ row, const uint4* __restrict__ tp,
                                              const uint32_t* __restrict__ lw, const float* __restrict__ lv, int rounds,
                                              float* psum) {
    constexpr int AMAX = AT ? AT : GBT_MAX_A;
#pragma unroll
    for (int c = 0; c < AMAX; c++) psum[c] = 0.f;
#pragma unroll 1
    for (int rd = 0; rd < rounds; rd++) {
#pragma unroll
        for (int c = 0; c < AMAX; c++) {
            if (c < A) {
                const uint4 t4 = tp[c];
                const uint32_t x0 = *reinterpret_cast<const uint32_t*>(row + (t4.x & 0xffffu));
                const bool b0 = x0 > t4.x;
                const uint32_t n1 = b0 ? t4.z : t4.y;
                const uint32_t x1 = *reinterpret_cast<const uint32_t*>(row + (n1 & 0xffffu));
                const bool b1 = x1 > n1;
                const int i2 = (b0 ? 2 : 0) + (b1 ? 1 : 0);
                const uint32_t n2 = lw[c * RK_LOWER + i2];
                const uint32_t x2 = *reinterpret_cast<const uint32_t*>(row + (n2 & 0xffffu));
                const int i3 = 2 * i2 + ((x2 > n2) ? 1 : 0);
                const uint32_t n3 = lw[c * RK_LOWER + 4 + i3];
                const uint32_t x3 = *reinterpret_cast<const uint32_t*>(row + (n3 & 0xffffu));
                const int lf = 2 * i3 + ((x3 > n3) ? 1 : 0);
                psum[c] = GNX_FADD(psum[c], lv[c * RK_LEAVES + lf]);
            }
        }
        tp += A;
        lw += RK_LOWER * A;
        lv += RK_LEAVES * A;
    }
}

// Tree tops in the kernel parameter bank (warp-uniform index -> constant-cache broadcast instead of a
// shared-memory wavefront): one word per node, three per tree.
constexpr int GBT_TOPC_MAX_T = 2048;
struct GbtTopC {
    uint32_t w[3 * GBT_TOPC_MAX_T];
};

// Tile kernel: per tree { node 0, node 1, node 2, node 0's feature offset } in tile node words, one 16-byte load.
struct alignas(16) GbtTileTop {
    uint4 q[GBT_TILE_MAX_T];
};

// Accumulating-offset walk of the row kernel (gnx_gbt_set_kernel 14): nodes 3..14 of a tree sit in one 64-byte block laid out so
// that a single byte offset o = 32 b0 + 16 b1 + 8 b2 + 4 b3 (b = branch bits) addresses every level: level-2 node at
// block + (32 b0 + 16 b1), level-3 node at block + (32 b0 + 16 b1 + 8 b2) + 4, leaf at leaves + o.  Each level then
// costs one compare and one predicated add instead of a select and a shift-add.
constexpr int RK_BLOCK = 16;   // words per tree in the block layout (4 unused)
__device__ __forceinline__ void gnx_add_if_gt(uint32_t& o, uint32_t x, uint32_t n, uint32_t inc) {
    asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, %3;\n\t}" : "+r"(o) : "r"(x), "r"(n), "r"(inc));
}

template <int AT>
__device__ __forceinline__ void gbt_rank_walk_o(int A, const unsigned char* __restrict__ row, const GbtTopC& top,
                                                const uint32_t* __restrict__ blk, const float* __restrict__ lv, int rounds,
                                                float* psum) {
    constexpr int AMAX = AT ? AT : GBT_MAX_A;
#pragma unroll
    for (int c = 0; c < AMAX; c++) psum[c] = 0.f;
    int tbase = 0;
    const unsigned char* bb = reinterpret_cast<const unsigned char*>(blk);
    const unsigned char* lvb = reinterpret_cast<const unsigned char*>(lv);
#pragma unroll 1
    for (int rd = 0; rd < rounds; rd++) {
#pragma unroll
        for (int c = 0; c < AMAX; c++) {
            if (c < A) {
                const uint32_t t0 = top.w[tbase + 3 * c], t1 = top.w[tbase + 3 * c + 1], t2 = top.w[tbase + 3 * c + 2];
                const uint32_t x0 = *reinterpret_cast<const uint32_t*>(row + (t0 & 0xffffu));
                const bool b0 = x0 > t0;
                const uint32_t n1 = b0 ? t2 : t1;
                uint32_t o = b0 ? 32u : 0u;
                gnx_add_if_gt(o, *reinterpret_cast<const uint32_t*>(row + (n1 & 0xffffu)), n1, 16u);
                const uint32_t n2 = *reinterpret_cast<const uint32_t*>(bb + c * (RK_BLOCK * 4) + o);
                gnx_add_if_gt(o, *reinterpret_cast<const uint32_t*>(row + (n2 & 0xffffu)), n2, 8u);
                const uint32_t n3 = *reinterpret_cast<const uint32_t*>(bb + c * (RK_BLOCK * 4) + 4 + o);
                gnx_add_if_gt(o, *reinterpret_cast<const uint32_t*>(row + (n3 & 0xffffu)), n3, 4u);
                psum[c] = GNX_FADD(psum[c], *reinterpret_cast<const float*>(lvb + c * (RK_LEAVES * 4) + o));
            }
        }
        tbase += 3 * A;
        bb += RK_BLOCK * 4 * A;
        lvb += RK_LEAVES * 4 * A;
    }
}

// One tree of the rank-form forest for one row: returns the leaf value.
__device__ __forceinline__ float gbt_rank_tree(const unsigned char* __restrict__ row, const uint4 t4,
                                               const uint32_t* __restrict__ lw, const float* __restrict__ lv) {
    const uint32_t x0 = *reinterpret_cast<const uint32_t*>(row + (t4.x & 0xffffu));
    const bool b0 = x0 > t4.x;
    const uint32_t n1 = b0 ? t4.z : t4.y;
    const uint32_t x1 = *reinterpret_cast<const uint32_t*>(row + (n1 & 0xffffu));
    const int i2 = (b0 ? 2 : 0) + ((x1 > n1) ? 1 : 0);
    const uint32_t n2 = lw[i2];
    const uint32_t x2 = *reinterpret_cast<const uint32_t*>(row + (n2 & 0xffffu));
    const int i3 = 2 * i2 + ((x2 > n2) ? 1 : 0);
    const uint32_t n3 = lw[4 + i3];
    const uint32_t x3 = *reinterpret_cast<const uint32_t*>(row + (n3 & 0xffffu));
    return lv[2 * i3 + ((x3 > n3) ? 1 : 0)];
}
#endif  // __CUDACC__

// monotone cell function of the tile kernel's rank pass (gbt_tile.cu)
struct GbtRankCells {
    float tmin, tmax;
    uint32_t kminA, kminB;   // key(tab[0]), key(1 - tab[K-1])
    int shiftA, shiftB;
};

}  // namespace gnx

struct gnx_gbt {
    gnx::GbtDev d;
    int device;
    void* d_blob;
    size_t forest_bytes;        // nodes + leaves, contiguous (for the shared-memory resident copy of the generic kernel)
    size_t rank_forest_bytes;   // lower | leaves | top, contiguous: image of the one-word row walk (also what gnofix stages)
    const unsigned char* rank_forest;
    int use_rank;               // 1 = rank-form kernels when eligible (default), 0 = generic float traversal
    gnx::GbtTopC* h_topc;       // row kernel: tree tops (k << 16 | byte offset) for the parameter bank (NULL if T too large)
    const unsigned char* block_forest;  // row kernel, accumulating-offset walk: block u32 [T][16] | leaves [T][16]
    size_t block_forest_bytes;
    // tile kernel (gbt_tile.cu): node word = (k << 17) | (feature << 7); per tree one 128-byte record
    // { block u32 [16], leaves f32 [16] }; tree tops in the parameter bank
    gnx::GbtTileTop* h_tiletop;   // 4 words per tree
    uint32_t* rank_lut;           // [GBT_RANK_CELLS] first threshold of the cell | thresholds in it << 16
    gnx::GbtRankCells rank_cells; // monotone cell function of the rank pass (gbt_tile.cu)
    float* rank_tab;              // the threshold table followed by NaN sentinels
    int profile;                  // gnx_gbt_set_profile: record events around the rank pass and the walk
    cudaEvent_t ev[3];
    const unsigned char* tile_forest;
    size_t tile_forest_bytes;
    int variant;                // 0 row kernel / one-word walk, 4 row kernel / block walk, 6 tile kernel; -1 = choose per call
};

namespace gnx {
// gbt_tile.cu: K4a (rank transform into hap-block-interleaved u16 tiles) + K4b (tile walk).  Returns 0 when it ran,
// -1 when the shape does not fit the tile kernel (caller falls back to the row kernel), > 0 on error.
int gbt_rank_lut_build(const float* tab_host, int K, GbtRankCells* cells, uint32_t** lut_dev, float** tab_dev);
int gbt_tile_smooth(const gnx_gbt* m, const float* B_dev, int64_t N, int W, float* proba_dev, int32_t* label_dev,
                    cudaStream_t st);
}  // namespace gnx
