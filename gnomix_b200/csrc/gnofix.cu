// gnofix.cu -- K6: Gnomix.phase / gnofix with the reference's default arguments, all
// individuals in one launch.
//
// Replaces src/model.py:188-214 (Python loop over individuals) and src/Gnofix/gnofix.py:
// 58-208 (+ track_switch / correct_phase_error, src/Gnofix/phasing.py:182-198) as called
// there: max_it=50, check_criterion="disc_smooth", max_center_offset=0, non_lin_s=0,
// prob_comp="max", prior_switch_prob=0.5, padding=True, no naive switch.
//
// What the reference does per individual (haplotypes m, p):
//   Y = smoother.predict(B); repeat up to max_it times, stopping when X_m repeats:
//     for w = 1..W-1 with a label discontinuity in either haplotype at w:
//       scope = S windows centred on clamp(w); score the pair as it is and with the tails
//       swapped at w by smoother.model.predict_proba on the 4 flattened scopes;
//       score = max over the 2 haplotypes of the max class probability;
//       if swapped > original (strict): swap the tails of B, of the tracker and of X
//       (at SNP w * (C // W)) and re-run smoother.predict on the whole pair.
//
// How this file does it, with identical results:
//   * The base probabilities are rank-transformed once (gbt_smooth.cuh: exact) into a
//     read-only uint16 array; tail swaps are never applied to it.  One bit per window
//     (the tracker: which original haplotype the current "m" row comes from) is the whole
//     state: current_B[h][j] = orig_B[h ^ trk[j]][j].  X and the float B are permuted once
//     at the end by gnofix_apply_kernel.
//   * A check at w stages the 2S-1 padded window slots around w for both haplotypes in
//     shared memory; that one stage serves the 4 candidate scopes and, if the switch is
//     accepted, the re-smoothing.
//   * Re-smoothing after a switch at w only re-evaluates the rows whose receptive field
//     straddles w (at most S-1 rows per haplotype); rows entirely inside the swapped tail
//     exchange their labels, rows entirely before it keep theirs.  Each row's output depends
//     only on its own S*A inputs, so this equals the reference's full smoother.predict(B).
//   * "X_m repeats" is decided from the tracker history and a per-window bit saying whether
//     the two original haplotypes differ anywhere in that window's SNP block
//     (gnofix_diff_kernel): X_m equals an earlier X_m iff the trackers differ only on
//     blocks where the originals are identical.
//   * A rejected check is remembered (one bit per window) until a switch is accepted inside its scope:
//     the outcome of a check depends only on the current pair inside its S-window scope (and is
//     unchanged when the whole scope sits in a swapped tail, the two candidates just trade places),
//     and a rejected check changes no state, so skipping its repeats in later iterations is exact.
//     Most of the reference's up-to-50 iterations re-run the same rejected checks.
//   * A team of 256 threads owns one individual; up to four teams share one CTA and one copy of
//     the forest in shared memory; teams pull individuals from a global counter.
//   * The tree walk is the accumulating-offset walk of the row kernel (gbt_smooth.cuh: one byte offset
//     addresses level-2 node, level-3 node and leaf inside a tree's block), with the three top nodes of a
//     tree read from shared memory as one 16-byte load.  The tracker history (one row of window bits per
//     outer iteration, read once per iteration) lives in global memory, which is what makes room for the
//     fourth team.
#include <stdlib.h>

#include <algorithm>

#include "gbt_smooth.cuh"

namespace gnx {

constexpr int GF_TEAM = 256;
constexpr int GF_MAX_IT = 64;

struct GfArgs {
    const uint16_t* ranks;   // [2n][W][A] rank of the ORIGINAL base probabilities
    const uint32_t* diff;    // [n][nw] bit j: original haplotypes differ in SNP block j
    uint32_t* trk_out;       // [n][nw] final tracker bits
    uint32_t* hist;          // [grid * teams][max_it][nw] tracker history of the individual a team is working on
    const int* nanflag;      // [n] the individual's base probabilities hold a NaN: refused (labels -1, pair left as it is)
    int32_t* Y;              // [2n][W] in: smoother labels of the original pair; out: final labels
    int32_t* tracker;        // [2n][W] or NULL
    int* counter;            // work queue
    int64_t n_ind;
    int W, nw, max_it, teams;
    size_t team_bytes;
    int leaf_words;          // floats of a team's leaf / margin buffer
    unsigned long long* stats;  // [4] iterations, scans, checks, accepted switches (summed over individuals)
    int memo;                // remember rejected checks (exact; GNX_GNOFIX_MEMO=0 re-runs them, for cross-checks)
    int split;               // re-smoothing by (row, class) tasks (GNX_GNOFIX_SPLIT=0: one thread per row, for cross-checks)
};

__device__ __forceinline__ void team_sync(int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(GF_TEAM) : "memory");
}

static unsigned long long* g_gnofix_stats = nullptr;

constexpr int GF_MAX_TEAMS = 4;
constexpr int GF_CH = 5;        // (row, class) chains a thread of the re-smoothing walks at once
constexpr int GF_REC = 144;     // bytes of a tree in shared memory: block u32 [16] | leaves f32 [16] | top uint4

__device__ __forceinline__ uint32_t gf_ld32(const unsigned char* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint32_t gf_lds32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// One tree for one row: the accumulating-offset walk of the row kernel (gbt_smooth.cuh: a single byte offset
// o = 32 b0 + 16 b1 + 8 b2 + 4 b3 addresses level-2 node, level-3 node and leaf inside the tree's record) with the three top
// nodes read from the record as one 16-byte load.  (The row kernel takes the tops from the parameter bank, free there because
// all warps of its CTA are on the same tree; the teams of this kernel are not, the indexed constant loads queue up: measured
// 17 % slower than shared-memory tops.)
__device__ __forceinline__ float gf_tree_t(const unsigned char* __restrict__ row, const unsigned char* __restrict__ rec, const uint4 t4) {
    const bool b0 = gf_ld32(row + (t4.x & 0xffffu)) > t4.x;
    const uint32_t n1 = b0 ? t4.z : t4.y;
    // the accumulating offset starts at the record's (shared) address: node and leaf loads are [o + immediate]
    uint32_t o = (uint32_t)__cvta_generic_to_shared(rec) + (b0 ? 32u : 0u);
    gnx_add_if_gt(o, gf_ld32(row + (n1 & 0xffffu)), n1, 16u);
    const uint32_t n2 = gf_lds32(o);
    gnx_add_if_gt(o, gf_ld32(row + (n2 & 0xffffu)), n2, 8u);
    const uint32_t n3 = gf_lds32(o + 4u);
    gnx_add_if_gt(o, gf_ld32(row + (n3 & 0xffffu)), n3, 4u);
    return __uint_as_float(gf_lds32(o + 64u));
}
__device__ __forceinline__ float gf_tree(const unsigned char* __restrict__ row, const unsigned char* __restrict__ rec) {
    return gf_tree_t(row, rec, *reinterpret_cast<const uint4*>(rec + 128));
}

// Re-smoothing unit of a warp: ONE class and 32 * NCH rows -- lane l walks rows hr0 + 32 k (k < NCH) as independent chains
// through the class's T / A trees.  Every chain of the warp is on the same tree: its record pointer is advanced once and its
// top is one 16-byte broadcast load per tree.  Chains behind the last row walk the last row again (branch-free) and store
// nothing.  hr numbers the rows of both haplotypes (hr = h * nrows + rr).
template <int NCH>
__device__ __forceinline__ void gf_walk_class(const unsigned char* __restrict__ st_b, int hap_bytes, int row_bytes,
                                              const unsigned char* __restrict__ rec, int A, int c, int rounds, int nrows, int hr0,
                                              float* __restrict__ margbuf) {
    const unsigned char* rowp[NCH];
    float ps[NCH];
    const int nrows2 = 2 * nrows;
#pragma unroll
    for (int k = 0; k < NCH; k++) {
        const int hr = min(hr0 + 32 * k, nrows2 - 1);
        const int h = hr >= nrows ? 1 : 0, rr = hr - h * nrows;
        rowp[k] = st_b + h * hap_bytes + rr * row_bytes;
        ps[k] = 0.f;
    }
    const int round_bytes = A * GF_REC;
#pragma unroll 1
    for (int rd = 0; rd < rounds; rd++) {
        const uint4 t4 = *reinterpret_cast<const uint4*>(rec + 128);
#pragma unroll
        for (int k = 0; k < NCH; k++) ps[k] = GNX_FADD(ps[k], gf_tree_t(rowp[k], rec, t4));
        rec += round_bytes;
    }
#pragma unroll
    for (int k = 0; k < NCH; k++)
        if (hr0 + 32 * k < nrows2) margbuf[(hr0 + 32 * k) * A + c] = ps[k];
}

template <int AT>
__global__ void __launch_bounds__(GF_MAX_TEAMS * GF_TEAM, 1)
gnofix_kernel(GbtDev m, const unsigned char* __restrict__ block_img, const uint4* __restrict__ tops_g, GfArgs g) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int A = AT ? AT : m.A;
    constexpr int AMAX = AT ? AT : GBT_MAX_A;
    // the forest as one 144-byte record per tree, from the block image (block u32 [T][16] | leaves f32 [T][16]) and the tops
    {
        const uint4* blk = reinterpret_cast<const uint4*>(block_img);
        const uint4* lvs = reinterpret_cast<const uint4*>(block_img + (size_t)m.T * RK_BLOCK * 4);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < m.T * 9; i += blockDim.x) {
            const int t = i / 9, q = i - t * 9;
            dst[i] = q < 4 ? __ldg(blk + t * 4 + q) : (q < 8 ? __ldg(lvs + t * 4 + (q - 4)) : __ldg(tops_g + t));
        }
    }
    __syncthreads();
    const unsigned char* recs = smem;
    const size_t forest_bytes = (size_t)m.T * GF_REC;

    const int team = threadIdx.x / GF_TEAM, tid = threadIdx.x % GF_TEAM;
    const int W = g.W, S = m.S, T = m.T, ast = m.astride, nw = g.nw;
    const int pad = (S + 1) / 2, half = (S - 1) / 2;
    const int NB = 2 * S - 1;                 // staged padded slots per haplotype
    const bool use_memo = g.memo != 0;
    const int rounds = T / A;

    // ---- team-private shared memory ------------------------------------------------
    unsigned char* tb = smem + forest_bytes + (size_t)team * g.team_bytes;
    uint32_t* st = reinterpret_cast<uint32_t*>(tb);                 // [2][NB][ast]
    uint32_t* cand = st + 2 * NB * ast;                              // [2][S][ast]  (switched m, switched p)
    float* leafbuf = reinterpret_cast<float*>(cand + 2 * S * ast);   // [4][T]
    uint32_t* trk = reinterpret_cast<uint32_t*>(leafbuf + g.leaf_words);   // [nw]
    uint32_t* hist = g.hist + ((size_t)blockIdx.x * g.teams + team) * (size_t)g.max_it * nw;   // [max_it][nw], global
    uint32_t* dif = trk + nw;                                        // [nw]
    uint32_t* rej = dif + nw;                                        // [nw] check at w rejected, scope unchanged since
    float* marg = reinterpret_cast<float*>(rej + nw);                // [4][GBT_MAX_A]
    float* expv = marg + 4 * GBT_MAX_A;                              // [4][GBT_MAX_A]
    float* pmax = expv + 4 * GBT_MAX_A;                              // [4]
    int* ctl = reinterpret_cast<int*>(pmax + 4);                     // [0] next individual, [1] found w, [2] flag
    signed char* Ys = reinterpret_cast<signed char*>(ctl + 4);       // [2][W]

    for (;;) {
        team_sync(team);
        if (tid == 0) ctl[0] = atomicAdd(g.counter, 1);
        team_sync(team);
        const int64_t ind = ctl[0];
        if (ind >= g.n_ind) break;
        if (g.nanflag[ind]) {
            // NaN among the pair's base probabilities: the smoother kernels follow each node's default child, the rank
            // form of this kernel cannot (a NaN has no rank) -- refuse loudly instead of phasing with other semantics
            for (int i = tid; i < 2 * W; i += GF_TEAM) g.Y[(size_t)(2 * ind) * W + i] = -1;
            for (int i = tid; i < nw; i += GF_TEAM) g.trk_out[(size_t)ind * nw + i] = 0u;
            if (g.tracker)
                for (int i = tid; i < W; i += GF_TEAM) {
                    g.tracker[(size_t)(2 * ind) * W + i] = 0;
                    g.tracker[(size_t)(2 * ind + 1) * W + i] = 1;
                }
            continue;
        }
        const uint16_t* rk0 = g.ranks + (size_t)(2 * ind) * W * A;   // original haplotype 0; haplotype 1 follows
        for (int i = tid; i < 2 * W; i += GF_TEAM) Ys[i] = (signed char)g.Y[(size_t)(2 * ind) * W + i];
        for (int i = tid; i < nw; i += GF_TEAM) {
            trk[i] = 0u;
            rej[i] = 0u;
            dif[i] = g.diff[(size_t)ind * nw + i];
        }
        team_sync(team);
        unsigned long long n_it = 0, n_scan = 0, n_check = 0, n_acc = 0;

        for (int it = 0; it < g.max_it; it++) {
            // ---- stop if X_m was seen before (gnofix.py:108-113) ----------------------
            if (tid == 0) ctl[2] = 0;
            team_sync(team);
            if (tid < it) {
                bool same = true;
                for (int i = 0; i < nw; i++) same &= (((hist[(size_t)tid * nw + i] ^ trk[i]) & dif[i]) == 0u);
                if (same) ctl[2] = 1;
            }
            team_sync(team);
            if (ctl[2]) break;
            for (int i = tid; i < nw; i += GF_TEAM) hist[(size_t)it * nw + i] = trk[i];
            n_it++;

            int w_cur = 1;
            for (;;) {
                // ---- next w >= w_cur with a discontinuity in either haplotype (gnofix.py:32)
                int w = W;
                for (int base = w_cur; base < W; base += GF_TEAM) {
                    team_sync(team);
                    if (tid == 0) ctl[1] = W;
                    team_sync(team);
                    const int ww = base + tid;
                    if (ww < W && (Ys[ww] != Ys[ww - 1] || Ys[W + ww] != Ys[W + ww - 1]) && !((rej[ww >> 5] >> (ww & 31)) & 1u))
                        atomicMin(&ctl[1], ww);
                    team_sync(team);
                    w = ctl[1];
                    if (w < W) break;
                }
                n_scan++;
                if (w >= W) break;
                w_cur = w + 1;
                n_check++;

                // ---- stage padded slots [jlo, jlo + NB) of the CURRENT pair --------------
                // rows whose receptive field straddles w (rows < pad also read reflected windows up to pad-1-row)
                const int r_lo = (w <= pad - 1) ? 0 : w - (S - 1 - pad), r_hi = min(W - 1, w + pad - 1);
                const int jlo = r_lo;
                const int nslots = r_hi - r_lo + S;
                for (int idx = tid; idx < 2 * nslots * A; idx += GF_TEAM) {
                    const int h = idx / (nslots * A), rem = idx - h * (nslots * A);
                    const int jj = rem / A, a = rem - jj * A;
                    const int o = spad_to_orig(jlo + jj, W, pad);
                    const int src = h ^ ((trk[o >> 5] >> (o & 31)) & 1u);
                    st[(h * NB + jj) * ast + a] = (uint32_t)__ldg(rk0 + ((size_t)src * W + o) * A + a) << 16;
                }
                // scope of the check: S windows centred on clamp(w) (gnofix.py:122-130), in padded slots
                const int center = min(max(w, half), W - S + half);
                const int lo = center - half;              // first original window of the scope
                const int sj = lo + pad - jlo;             // its slot inside the stage
                team_sync(team);
                // switched candidates: m' = m[lo:w] ++ p[w:lo+S],  p' = p[lo:w] ++ m[w:lo+S]
                for (int idx = tid; idx < 2 * S * A; idx += GF_TEAM) {
                    const int h = idx / (S * A), rem = idx - h * (S * A);
                    const int s = rem / A, a = rem - s * A;
                    const int src = (lo + s < w) ? h : (1 - h);
                    cand[(h * S + s) * ast + a] = st[(src * NB + sj + s) * ast + a];
                }
                team_sync(team);
                // ---- 4 rows x T trees (smoother.model.predict_proba, gnofix.py:157) --------
                // 64 threads per row, four independent trees in flight per thread
                {
                    const int r = tid >> 6, q = tid & 63;
                    const unsigned char* row = reinterpret_cast<const unsigned char*>(
                        r < 2 ? st + (r * NB + sj) * ast : cand + ((r - 2) * S) * ast);
                    for (int t0 = q; t0 < T; t0 += 256) {
                        float v[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) v[j] = gf_tree(row, recs + (size_t)min(t0 + 64 * j, T - 1) * GF_REC);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (t0 + 64 * j < T) leafbuf[r * T + t0 + 64 * j] = v[j];
                    }
                }
                team_sync(team);
                if (tid < 4 * A) {
                    const int r = tid / A, c = tid - r * A;
                    float ps = 0.f;
#pragma unroll 4
                    for (int rd = 0; rd < rounds; rd++) ps = GNX_FADD(ps, leafbuf[r * T + rd * A + c]);
                    marg[r * GBT_MAX_A + c] = GNX_FADD(__ldg(m.base + c), ps);
                }
                team_sync(team);
                if (tid < 4 * A) {
                    const int r = tid / A, c = tid - r * A;
                    float wmax = marg[r * GBT_MAX_A];
                    for (int k = 1; k < A; k++) wmax = fmaxf(marg[r * GBT_MAX_A + k], wmax);
                    expv[r * GBT_MAX_A + c] = gnx_expf_cr(GNX_FSUB(marg[r * GBT_MAX_A + c], wmax));
                }
                team_sync(team);
                if (tid < 4) {
                    double wsum = 0.0;
                    for (int k = 0; k < A; k++) wsum = GNX_ADD(wsum, GNX_F2D(expv[tid * GBT_MAX_A + k]));
                    const float ws = GNX_D2F(wsum);
                    float best = 0.f;
                    for (int k = 0; k < A; k++) {
                        const float p = GNX_FDIV(expv[tid * GBT_MAX_A + k], ws);
                        if (k == 0 || p > best) best = p;
                    }
                    pmax[tid] = best;
                }
                team_sync(team);
                const float p_orig = fmaxf(pmax[0], pmax[1]), p_sw = fmaxf(pmax[2], pmax[3]);
                if (!(p_sw > p_orig)) {                    // gnofix.py:171 (the 0.5 prior cancels)
                    if (tid == 0 && use_memo) rej[w >> 5] |= 1u << (w & 31);
                    continue;
                }
                n_acc++;
                // a switch at w changes the pair inside every scope that contains w: forget those verdicts
                for (int ww = max(1, w - S) + tid; ww <= min(W - 1, w + S); ww += GF_TEAM)
                    atomicAnd(&rej[ww >> 5], ~(1u << (ww & 31)));

                // ---- accept: swap tails at w --------------------------------------------
                for (int i = tid; i < nw; i += GF_TEAM) {
                    const int b0 = i * 32;
                    uint32_t mask = 0u;
                    if (b0 >= w) mask = 0xffffffffu;
                    else if (b0 + 32 > w) mask = 0xffffffffu << (w - b0);
                    if (i == nw - 1 && (W & 31)) mask &= (1u << (W & 31)) - 1u;
                    trk[i] ^= mask;
                }
                // rows entirely inside the tail exchange labels; staged slots inside the tail swap
                for (int ww = w + pad + tid; ww < W; ww += GF_TEAM) {
                    const signed char a0 = Ys[ww];
                    Ys[ww] = Ys[W + ww];
                    Ys[W + ww] = a0;
                }
                for (int idx = tid; idx < nslots * A; idx += GF_TEAM) {
                    const int jj = idx / A, a = idx - jj * A;
                    if (spad_to_orig(jlo + jj, W, pad) >= w) {
                        const uint32_t v0 = st[(jj)*ast + a];
                        st[(jj)*ast + a] = st[(NB + jj) * ast + a];
                        st[(NB + jj) * ast + a] = v0;
                    }
                }
                team_sync(team);
                // rows straddling w: re-evaluate (Smoother.predict, smooth.py:58-61).  The class sums of a row are
                // independent of each other (tree t feeds class t % A), so the work splits into (row, class) tasks of
                // T / A trees: a warp takes one class and up to 32 * GF_CH rows, each lane walking GF_CH rows as
                // independent chains (gf_walk_class) -- seven or eight warps issue, where the row-per-thread form left
                // three of them (and the lanes of a fifth) waiting at the barrier.
                const int nrows = r_hi - r_lo + 1;
                const unsigned char* st_b = reinterpret_cast<const unsigned char*>(st);
                if (g.split) {
                    float* margbuf = leafbuf;                       // [2 * nrows][A]
                    // units = (class, chunk of 32 * nch rows); the fewest chains per lane that fit all units on the 8 warps
                    const int nrows2 = 2 * nrows;
                    int nch = 1;
                    while (nch < GF_CH && A * ((nrows2 + 32 * nch - 1) / (32 * nch)) > GF_TEAM / 32) nch++;
                    const int chunks = (nrows2 + 32 * nch - 1) / (32 * nch);
                    for (int u = tid >> 5; u < A * chunks; u += GF_TEAM / 32) {
                        const int c = u / chunks, j = u - c * chunks;
                        const int hr0 = j * 32 * nch + (tid & 31);
                        const int nl = min(nch, (nrows2 - j * 32 * nch + 31) / 32);   // chains holding a row for some lane
                        const unsigned char* rec = recs + c * GF_REC;
                        switch (nl) {
                            case 5: gf_walk_class<5>(st_b, NB * ast * 4, ast * 4, rec, A, c, rounds, nrows, hr0, margbuf); break;
                            case 4: gf_walk_class<4>(st_b, NB * ast * 4, ast * 4, rec, A, c, rounds, nrows, hr0, margbuf); break;
                            case 3: gf_walk_class<3>(st_b, NB * ast * 4, ast * 4, rec, A, c, rounds, nrows, hr0, margbuf); break;
                            case 2: gf_walk_class<2>(st_b, NB * ast * 4, ast * 4, rec, A, c, rounds, nrows, hr0, margbuf); break;
                            default: gf_walk_class<1>(st_b, NB * ast * 4, ast * 4, rec, A, c, rounds, nrows, hr0, margbuf); break;
                        }
                    }
                    team_sync(team);
                    for (int task = tid; task < 2 * nrows; task += GF_TEAM) {
                        const int h = task / nrows, rr = task - h * nrows;
                        float psum[AMAX];
#pragma unroll
                        for (int c = 0; c < AMAX; c++)
                            if (c < A) psum[c] = margbuf[task * A + c];
                        int32_t lab;
                        gbt_finish<AT>(m, psum, nullptr, &lab);
                        Ys[h * W + r_lo + rr] = (signed char)lab;
                    }
                } else {
                    // one thread per row, its A class sums as A chains (cross-check: GNX_GNOFIX_SPLIT=0)
                    for (int task = tid; task < 2 * nrows; task += GF_TEAM) {
                        const int h = task / nrows, rr = task - h * nrows;
                        const unsigned char* row = st_b + ((h * NB + rr) * ast) * 4;
                        float psum[AMAX];
#pragma unroll
                        for (int c = 0; c < AMAX; c++) psum[c] = 0.f;
                        const unsigned char* rec = recs;
#pragma unroll 1
                        for (int rd = 0; rd < rounds; rd++) {
#pragma unroll
                            for (int c = 0; c < AMAX; c++)
                                if (c < A) psum[c] = GNX_FADD(psum[c], gf_tree(row, rec + c * GF_REC));
                            rec += A * GF_REC;
                        }
                        int32_t lab;
                        gbt_finish<AT>(m, psum, nullptr, &lab);
                        Ys[h * W + r_lo + rr] = (signed char)lab;
                    }
                }
                team_sync(team);
            }
        }
        // ---- results -----------------------------------------------------------------
        team_sync(team);
        if (tid == 0 && g.stats) {
            atomicAdd(g.stats + 0, n_it);
            atomicAdd(g.stats + 1, n_scan);
            atomicAdd(g.stats + 2, n_check);
            atomicAdd(g.stats + 3, n_acc);
        }
        for (int i = tid; i < 2 * W; i += GF_TEAM) g.Y[(size_t)(2 * ind) * W + i] = Ys[i];
        for (int i = tid; i < nw; i += GF_TEAM) g.trk_out[(size_t)ind * nw + i] = trk[i];
        if (g.tracker)
            for (int i = tid; i < W; i += GF_TEAM) {
                const int b = (trk[i >> 5] >> (i & 31)) & 1u;
                g.tracker[(size_t)(2 * ind) * W + i] = b;
                g.tracker[(size_t)(2 * ind + 1) * W + i] = 1 - b;
            }
    }
}

// rank of every original base probability (exact, see gbt_smooth.cuh); NaN ranks above
// every threshold.
__global__ void gnofix_rank_kernel(GbtDev m, const float* __restrict__ B, int64_t count, int64_t per_ind, uint16_t* __restrict__ out,
                                   int* __restrict__ nanflag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = __ldg(B + i);
        if (x != x) nanflag[i / per_ind] = 1;
        out[i] = (x != x) ? (uint16_t)m.K : (uint16_t)gbt_rank_of(m.thr_table, m.K, x);
    }
}

// diff bit j of individual i: rows 2i and 2i+1 of X differ somewhere in SNP block j
// (blocks of ws = C / W SNPs, the last block runs to C -- gnofix.py:74, phasing.py:192-196)
__global__ void gnofix_diff_kernel(const int8_t* __restrict__ X, int64_t ldX, int64_t C, int W, int nw, int64_t ws,
                                   uint32_t* __restrict__ diff) {
    const int64_t ind = blockIdx.x;
    const int8_t* a = X + (2 * ind) * ldX;
    const int8_t* b = a + ldX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < W; j += nwarps) {
        const int64_t s = (int64_t)j * ws, e = (j == W - 1) ? C : s + ws;
        bool d = false;
        for (int64_t p = s + lane; p < e; p += 32) d |= (a[p] != b[p]);
        const unsigned any = __ballot_sync(0xffffffffu, d);
        if (lane == 0 && any) atomicOr(diff + ind * nw + (j >> 5), 1u << (j & 31));
    }
}

// final permutation: X tails (SNP blocks) and float B windows of the two haplotypes
// exchanged wherever the tracker bit is set
__global__ void gnofix_apply_kernel(int8_t* __restrict__ X, int64_t ldX, int64_t C, float* __restrict__ B, int W, int A, int nw,
                                    int64_t ws, const uint32_t* __restrict__ trk) {
    const int64_t ind = blockIdx.x;
    int8_t* a = X + (2 * ind) * ldX;
    int8_t* b = a + ldX;
    const uint32_t* t = trk + ind * nw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < W; j += nwarps) {
        if (!((t[j >> 5] >> (j & 31)) & 1u)) continue;
        if (X) {
            const int64_t s = (int64_t)j * ws, e = (j == W - 1) ? C : s + ws;
            for (int64_t p = s + lane; p < e; p += 32) {
                const int8_t v = a[p];
                a[p] = b[p];
                b[p] = v;
            }
        }
        if (B) {
            float* b0 = B + ((2 * ind) * W + j) * A;
            float* b1 = B + ((2 * ind + 1) * W + j) * A;
            if (lane < A) {
                const float v = b0[lane];
                b0[lane] = b1[lane];
                b1[lane] = v;
            }
        }
    }
}

}  // namespace gnx

using namespace gnx;

extern "C" int gnx_gnofix(const gnx_gbt_t* m, int8_t* X_dev, int64_t ldX, int64_t C, float* B_dev, int64_t n_ind, int W,
                          int max_it, int32_t* Y_dev, int32_t* tracker_dev, void* stream) {
    GNX_REQUIRE(m != nullptr, "gnx_gnofix: NULL smoother model");
    GNX_REQUIRE(n_ind >= 0 && W >= 2 && C >= W && ldX >= C, "gnx_gnofix: bad shape n_ind=%lld W=%d C=%lld ldX=%lld", (long long)n_ind, W,
                (long long)C, (long long)ldX);
    GNX_REQUIRE(max_it >= 1 && max_it <= GF_MAX_IT, "gnx_gnofix: max_it=%d outside 1..%d", max_it, GF_MAX_IT);
    if (n_ind == 0) return 0;
    GNX_REQUIRE(B_dev && Y_dev, "gnx_gnofix: NULL buffer (X_dev may be NULL to skip the SNP-level swap)");
    GNX_REQUIRE(m->d.rank_ok, "gnx_gnofix: needs a forest of depth <= 4 with <= 65535 distinct thresholds (the rank-form image)");
    const int S = m->d.S, A = m->d.A, T = m->d.T;
    GNX_REQUIRE(W >= 2 * S, "gnx_gnofix: W=%d < 2*S=%d (XGB_Smoother asserts W >= 2S, src/Smooth/models.py:13)", W, 2 * S);
    cudaStream_t st = (cudaStream_t)stream;
    const int nw = (W + 31) / 32;
    const int64_t ws = C / W;

    // initial labels of the pair as it is: smoother.predict(B) (gnofix.py:80)
    if (gnx_gbt_smooth(m, B_dev, 2 * n_ind, W, nullptr, Y_dev, stream)) return 1;

    // scratch: ranks u16 [2n][W][A] | diff [n][nw] | trk [n][nw] | counter | NaN flag [n]
    const size_t rb = ((size_t)2 * n_ind * W * A * sizeof(uint16_t) + 255) & ~size_t(255);
    const size_t db = ((size_t)n_ind * nw * 4 + 255) & ~size_t(255);
    const size_t fb = ((size_t)n_ind * 4 + 255) & ~size_t(255);
    char* scratch = nullptr;
    GNX_CUDA(cudaMallocAsync((void**)&scratch, rb + 2 * db + 256 + fb, st));
    uint16_t* ranks = reinterpret_cast<uint16_t*>(scratch);
    uint32_t* diff = reinterpret_cast<uint32_t*>(scratch + rb);
    uint32_t* trk = reinterpret_cast<uint32_t*>(scratch + rb + db);
    int* counter = reinterpret_cast<int*>(scratch + rb + 2 * db);
    int* nanflag = reinterpret_cast<int*>(scratch + rb + 2 * db + 256);
    GNX_CUDA(cudaMemsetAsync(scratch + rb, 0, 2 * db + 256 + fb, st));
    const int64_t count = 2 * n_ind * W * A;
    gnofix_rank_kernel<<<(int)std::min<int64_t>(ceil_div(count, 256), (int64_t)sm_count() * 16), 256, 0, st>>>(m->d, B_dev, count, (int64_t)2 * W * A,
                                                                                                     ranks, nanflag);
    if (X_dev)
        gnofix_diff_kernel<<<(unsigned)n_ind, 256, 0, st>>>(X_dev, ldX, C, W, nw, ws, diff);
    else
        GNX_CUDA(cudaMemsetAsync(diff, 0xff, db, st));  // no X: treat every block as differing (tracker equality)

    const int NB = 2 * S - 1, ast = m->d.astride;
    GNX_REQUIRE(m->block_forest != nullptr, "gnx_gnofix: the forest has no block image");
    const size_t img_bytes = (size_t)T * GF_REC, top_bytes = 0;
    const size_t leaf_words = std::max<size_t>((size_t)4 * T, (size_t)2 * (S - 1) * A);   // check leaves / re-smoothing margins
    size_t team_bytes = (size_t)2 * NB * ast * 4 + (size_t)2 * S * ast * 4 + leaf_words * 4 + (size_t)3 * nw * 4 +
                        (size_t)(8 * GBT_MAX_A + 4) * 4 + 16 + (size_t)2 * W;
    team_bytes = (team_bytes + 15) & ~size_t(15);
    const size_t smem_max = 227 * 1024;
    int teams = GF_MAX_TEAMS;
    {
        const char* et = getenv("GNX_GNOFIX_TEAMS");   // A/B knob
        if (et && atoi(et) >= 1 && atoi(et) <= GF_MAX_TEAMS) teams = atoi(et);
    }
    while (teams > 1 && img_bytes + top_bytes + teams * team_bytes > smem_max) teams--;
    GNX_REQUIRE(img_bytes + top_bytes + teams * team_bytes <= smem_max, "gnx_gnofix: W=%d / forest too large for shared memory", W);
    const size_t smem = img_bytes + top_bytes + teams * team_bytes;
    const int grid = (int)std::min<int64_t>(ceil_div(n_ind, teams), (int64_t)sm_count());
    uint32_t* hist = nullptr;
    GNX_CUDA(cudaMallocAsync((void**)&hist, (size_t)grid * teams * max_it * nw * sizeof(uint32_t), st));
    const char* em = getenv("GNX_GNOFIX_MEMO");
    const char* es = getenv("GNX_GNOFIX_SPLIT");
    static unsigned long long* d_stats = nullptr;   // profiling counters of the last call on this process
    if (!d_stats) GNX_CUDA(cudaMalloc((void**)&d_stats, 4 * sizeof(unsigned long long)));
    GNX_CUDA(cudaMemsetAsync(d_stats, 0, 4 * sizeof(unsigned long long), st));
    g_gnofix_stats = d_stats;
    GfArgs g{ranks, diff, trk, hist, nanflag, Y_dev, tracker_dev, counter, n_ind, W, nw, max_it, teams, team_bytes, (int)leaf_words, d_stats, (em && em[0] == '0') ? 0 : 1, (es && es[0] == '0') ? 0 : 1};
#define CALLG(AT)                                                                                                  \
    do {                                                                                                           \
        GNX_CUDA(cudaFuncSetAttribute(gnofix_kernel<AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gnofix_kernel<AT><<<grid, teams * GF_TEAM, smem, st>>>(m->d, m->block_forest, m->d.top, g);                \
    } while (0)
    switch (A) {
        case 2: CALLG(2); break;
        case 3: CALLG(3); break;
        case 4: CALLG(4); break;
        case 5: CALLG(5); break;
        case 6: CALLG(6); break;
        case 7: CALLG(7); break;
        case 8: CALLG(8); break;
        default: CALLG(0); break;
    }
#undef CALLG
    GNX_CUDA(cudaGetLastError());
    GNX_CUDA(cudaFreeAsync(hist, st));
    if (X_dev)
        gnofix_apply_kernel<<<(unsigned)n_ind, 256, 0, st>>>(X_dev, ldX, C, B_dev, W, A, nw, ws, trk);
    else
        gnofix_apply_kernel<<<(unsigned)n_ind, 256, 0, st>>>(nullptr, 0, 0, B_dev, W, A, nw, 1, trk);
    GNX_CUDA(cudaGetLastError());
    GNX_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

/* profiling counters of the last gnx_gnofix call of this process (synchronises the device):
 * out[0..3] = outer iterations, discontinuity scans, candidate checks, accepted switches, summed over individuals */
extern "C" int gnx_gnofix_last_stats(int64_t* out) {
    GNX_REQUIRE(out != nullptr, "gnx_gnofix_last_stats: NULL");
    out[0] = out[1] = out[2] = out[3] = 0;
    if (!gnx::g_gnofix_stats) return 0;
    unsigned long long h[4];
    GNX_CUDA(cudaMemcpy(h, gnx::g_gnofix_stats, sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; i++) out[i] = (int64_t)h[i];
    return 0;
}
