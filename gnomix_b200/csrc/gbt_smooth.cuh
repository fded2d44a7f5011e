// gbt_smooth.cuh -- model handle of the tree-ensemble smoother (K4), shared with gnofix (K6)
#pragma once

#include "common.cuh"

namespace gnx {

constexpr int GBT_MAX_A = 16;
constexpr int GBT_MAX_DEPTH = 8;
constexpr int RK_THREADS = 1024;
constexpr int RK_LOWER = 12;   // nodes 3..14 of a depth-4 heap
constexpr int RK_LEAVES = 16;

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

}  // namespace gnx

struct gnx_gbt {
    gnx::GbtDev d;
    int device;
    void* d_blob;
    size_t forest_bytes;     // nodes + leaves, contiguous (for the shared-memory resident copy)
    size_t rank_forest_bytes;  // lower | leaves | top, contiguous (shared-memory image of the fast path)
    const unsigned char* rank_forest;
    int use_rank;            // 1 = rank-form kernel when eligible (default), 0 = generic float traversal
};
