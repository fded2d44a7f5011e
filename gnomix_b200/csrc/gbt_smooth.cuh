// gbt_smooth.cuh -- model handle of the tree-ensemble smoother (K4), shared with gnofix (K6)
#pragma once

#include "common.cuh"

namespace gnx {

constexpr int GBT_MAX_A = 16;
constexpr int GBT_MAX_DEPTH = 8;

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
};

}  // namespace gnx

struct gnx_gbt {
    gnx::GbtDev d;
    int device;
    void* d_blob;
    size_t forest_bytes;     // nodes + leaves, contiguous (for the shared-memory resident copy)
};
