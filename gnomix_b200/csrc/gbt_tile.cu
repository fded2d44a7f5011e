// gbt_tile.cu -- K4, tile form: the production kernel of XGB_Smoother.predict_proba / predict
// (reference src/Smooth/utils.py:4-29 slide_window + src/Smooth/models.py:14-20 xgboost predict_proba +
// src/Smooth/smooth.py:61 argmax) for batches of haplotypes.
//
// The walk of a depth-4 tree is four data-dependent feature loads, two node loads and one leaf load: seven
// shared-memory wavefronts per warp and tree when a lane owns a row -- six and a half here, because the two
// adjacent windows a warp walks share the load of the root's feature (pair words, below) -- and the kernel is built
// to sit on that floor:
//   * lane = haplotype, warp = window.  The rank tile is lane-interleaved -- word (slot * A + a) * 32 + lane --
//     so a feature load hits bank `lane` whatever node each lane has reached: no bank conflicts under
//     divergence (the row kernel, lanes = windows, pays ~1.2 conflict wavefronts per tree).
//   * tile word of slot s = rank(s) << 17 | rank(s + 1) (ranks are 15 bits), node word = (k << 17) | (0x8000 +
//     (feature << 7)): `tile word > node word` is the split test `!(x < thr)` of the window whose feature sits in
//     slot s -- when rank(s) == k the low halves decide, and the node's (>= 0x8000) is never below the tile word's
//     (<= 0x7fff) -- and `word & 0x1ff80` is the byte offset of the feature row behind `tile - 0x8000`, so a level
//     costs LDS node, LOP3 (mask | lane * 4), LDS feature, ISETP, predicated IADD.  The root of a tree tests the SAME
//     feature in windows w and w + 1, one slot apart: one load serves both (`word << 17` is the second window's).
//   * per tree one 128-byte record { 16-word block, 16 leaves } addressed by ONE accumulating byte offset
//     o = 32 b0 + 16 b1 + 8 b2 + 4 b3 (level-2 node at block + o, level-3 node at block + 4 + o, leaf at
//     leaves + o); lanes of a warp read distinct banks or the same word.
//   * the top three nodes of every tree come from the kernel parameter bank (warp-uniform: constant cache).
//   * a warp walks two adjacent windows at once (rows 2 * warp and 2 * warp + 1 of the tile): the uniform work per
//     tree and the root's feature load are shared and there are 2 * A independent dependency chains in flight.
// K4a turns the float32 base probabilities into their exact ranks (u16) once, written in the order the
// tiles are staged in: R2[haplotype block][padded slot][class][32 lanes], reflect padding materialised.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "gbt_smooth.cuh"

namespace gnx {

__device__ __forceinline__ uint32_t gnx_lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float gnx_lds_f32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}

constexpr int TILE_WARPS = RK_THREADS / 32;
constexpr uint32_t TILE_FMASK = 0x1ff80u;   // feature << 7 field of a node word

// ---------------------------------------------------------------- K4a: rank transform
// B f32 [N, W, A] -> R2 u16 [ceil(N/32)][Wp = W + S - 1][A][32]; NaN -> 0xFFFF, lanes beyond N -> 0.
// rank(x) = #{i : tab[i] <= x}, exactly.  A binary search over the shared-memory table probes, at every level whose
// step is a multiple of 32 words, ONE bank for the whole warp (measured: 139 wavefronts per warp and element, 9 ms at
// 50 000 haplotypes).  Instead x is hashed by a MONOTONE cell function and only the thresholds sharing x's cell are
// compared: every threshold in a lower cell is <= ... < x's cell, every threshold in a higher cell is > x.
//   key(v)  = the order-preserving integer image of a float (sign bit flipped for positive values, all bits for
//             negative ones; -0 canonicalised to +0 first);
//   x < 1/2:  cell = (key(x) - key(tab[0])) >> shiftA                       in [0, CELLS/2)
//   x >= 1/2: cell = CELLS - 1 - ((key(1 - x) - key(1 - tab[K-1])) >> shiftB)  in [CELLS/2, CELLS)
// Both halves are compositions of monotone maps (fl(1 - x) is monotone in x), so cell() is monotone on all floats.
// The cells are log-spaced towards 0 AND towards 1, where base probabilities and therefore split thresholds crowd;
// equal-width cells (first version, 3.8 ms) and cells log-spaced towards 0 only (second, 3.2 ms: 169 instructions per
// warp and element, long divergent searches inside the few cells of [0.5, 1)) leave most thresholds in a few cells.
// lut[c] = first threshold of cell c | number of thresholds in it << 16, built on the host with the same function.
__host__ __device__ __forceinline__ uint32_t gbt_float_key(float x) {
    x = x + 0.0f;   // -0 -> +0 (exact); NaN never reaches here
#ifdef __CUDA_ARCH__
    const uint32_t b = __float_as_uint(x);
#else
    uint32_t b;
    memcpy(&b, &x, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__host__ __device__ __forceinline__ uint32_t gbt_rank_cell(const GbtRankCells& c, float x) {   // tmin <= x < tmax
    constexpr uint32_t H = GBT_RANK_CELLS / 2;
    if (x < 0.5f) {
        const uint32_t d = gbt_float_key(x) - c.kminA;
        const uint32_t q = d >> c.shiftA;
        return q < H - 1 ? q : H - 1;
    }
#ifdef __CUDA_ARCH__
    const float u = __fsub_rn(1.0f, x);
#else
    const float u = 1.0f - x;
#endif
    const uint32_t d = gbt_float_key(u) - c.kminB;
    const uint32_t q = d >> c.shiftB;
    return GBT_RANK_CELLS - 1 - (q < H - 1 ? q : H - 1);
}

// tab = the sorted thresholds followed by GBT_RANK_PAD NaN sentinels (`NaN <= x` is false): the four entries from the
// cell's first threshold on are compared without a branch -- thresholds of higher cells are > x, so reading past the
// cell's own entries is harmless -- and only an x that passes all four keeps scanning (rare).  x outside
// [tab[0], tab[K-1]) takes its cell from the clamped value and its comparisons from x itself: below the table nothing
// is <= x (rank 0), above it everything is (the scan runs from the last cell's first entry to the sentinel: rank K).
constexpr int GBT_RANK_PAD = 4;
__device__ __forceinline__ uint32_t gbt_rank_lut(const float* __restrict__ tab, const uint32_t* __restrict__ lut,
                                                 const GbtRankCells& c, float x) {
    const float xs = fminf(fmaxf(x, c.tmin), c.tmax);   // NaN -> tmin (its rank is replaced below)
    const uint32_t lo = lut[gbt_rank_cell(c, xs)] & 0xffffu;
    const float* t = tab + lo;
    uint32_t r = lo + (t[0] <= x) + (t[1] <= x) + (t[2] <= x) + (t[3] <= x);
    if (t[3] <= x) {
        while (tab[r] <= x) r++;
    }
    return (x != x) ? 0xFFFFu : r;
}

// One CTA pass = one tile of 32 haplotypes x RANK_EL padded elements.  Reads are coalesced along a haplotype's row
// (warp = haplotype, lane = element, 8 independent loads in flight per lane), ranks are transposed through shared
// memory, and the tile leaves as ONE contiguous 16 KB run of R2 (16 bytes per thread).
constexpr int RANK_EL = 256;
constexpr int RANK_THREADS = 1024;
constexpr int RANK_TS = 34;   // u16 per tile row (32 haplotypes + 2: rows 17 words apart, conflict-free transposition)
template <bool TAB_SMEM>
__global__ void __launch_bounds__(RANK_THREADS, 2)
gbt_rank_tile_kernel(const float* __restrict__ thr, const uint32_t* __restrict__ lut_g, int K, GbtRankCells cells,
                     const float* __restrict__ B, int64_t N, int W, int A, int S, uint16_t* __restrict__ R2) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint16_t* tile = reinterpret_cast<uint16_t*>(smem);
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem + RANK_EL * RANK_TS * 2);
    float* tab_s = reinterpret_cast<float*>(smem + RANK_EL * RANK_TS * 2 + GBT_RANK_CELLS * 4);
    for (int i = threadIdx.x; i < GBT_RANK_CELLS; i += blockDim.x) lut[i] = __ldg(lut_g + i);
    if (TAB_SMEM)   // (the global copy carries the same sentinels)
        for (int i = threadIdx.x; i < K + GBT_RANK_PAD; i += blockDim.x) tab_s[i] = __ldg(thr + i);
    const float* tab = TAB_SMEM ? tab_s : thr;
    const int pad = (S + 1) / 2, Wp = W + S - 1, E = Wp * A;
    const int e_lo = pad * A, e_hi = (pad + W) * A;            // padded elements [e_lo, e_hi) are B's own, in order
    const int runs = (E + RANK_EL - 1) / RANK_EL;
    const int64_t nhb = (N + 31) / 32, items = nhb * runs;
    const int lane = threadIdx.x & 31, h = threadIdx.x >> 5;   // h = haplotype of the block this warp reads
    const int dj = 32 / A, da = 32 - dj * A;                   // (slot, class) advance of 32 elements
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t hb = it / runs;
        const int e0 = (int)(it - hb * runs) * RANK_EL;
        const int64_t n = hb * 32 + h;
        const float* bn = B + n * (int64_t)W * A;
        float x[RANK_EL / 32];
        if (e0 >= e_lo && e0 + RANK_EL <= e_hi) {              // interior tile: one contiguous run of the row
            const float* src = bn + (e0 - e_lo) + lane;
#pragma unroll
            for (int q = 0; q < RANK_EL / 32; q++) x[q] = (n < N) ? __ldg(src + q * 32) : 0.f;
        } else {
            int j = (e0 + lane) / A, a = (e0 + lane) - j * A;
#pragma unroll
            for (int q = 0; q < RANK_EL / 32; q++) {
                const int e = e0 + q * 32 + lane;
                x[q] = 0.f;
                if (n < N && e < E) x[q] = __ldg(bn + spad_to_orig(j, W, pad) * A + a);
                j += dj;
                a += da;
                if (a >= A) { a -= A; j++; }
            }
        }
        __syncthreads();   // the previous tile has left shared memory (and, first time round, lut / tab are in place)
#pragma unroll
        for (int q = 0; q < RANK_EL / 32; q++) {
            // lanes beyond N / elements beyond E rank 0.f like any value: harmless, the first are never read back as
            // data of a real haplotype and the second are not written below
            uint16_t r = (uint16_t)gbt_rank_lut(tab, lut, cells, x[q]);
            if (n >= N) r = 0;
            tile[(q * 32 + lane) * RANK_TS + h] = r;
        }
        __syncthreads();
        {
            // thread t: element t / 4 of the tile, haplotypes 8 * (t % 4) .. + 7 -> 16 bytes
            const int el = threadIdx.x >> 2, part = threadIdx.x & 3;
            if (e0 + el < E) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(tile + el * RANK_TS + part * 8);
                uint4 v;
                v.x = src[0]; v.y = src[1]; v.z = src[2]; v.w = src[3];
                *reinterpret_cast<uint4*>(R2 + ((hb * E + e0 + el) * 32 + part * 8)) = v;
            }
        }
    }
}

// model-create half of the rank pass: the cell table (see above).  Returns 0 / non-zero like the C ABI.
int gbt_rank_lut_build(const float* tab_host, int K, GbtRankCells* cells, uint32_t** lut_dev, float** tab_dev) {
    std::vector<uint32_t> lut(GBT_RANK_CELLS, 0u);
    GbtRankCells c{};
    c.tmin = K ? tab_host[0] : 0.f;
    c.tmax = K ? tab_host[K - 1] : 0.f;
    if (K > 0) {
        constexpr uint32_t H = GBT_RANK_CELLS / 2;
        for (int i = 0; i < K; i++)
            GNX_REQUIRE(tab_host[i] == tab_host[i] && (i == 0 || tab_host[i] >= tab_host[i - 1]), "gnx_gbt_model_create: threshold table is not sorted");
        // lower half: keys of [tab[0], 1/2); upper half: keys of 1 - x for x in [1/2, tab[K-1]], i.e. of [1 - tab[K-1], 1/2]
        c.kminA = gbt_float_key(c.tmin);
        const uint32_t khalf = gbt_float_key(0.5f);
        const uint32_t spanA = khalf > c.kminA ? khalf - c.kminA : 0u;
        while ((spanA >> c.shiftA) >= H) c.shiftA++;
        c.kminB = gbt_float_key(1.0f - c.tmax);
        const uint32_t spanB = khalf > c.kminB ? khalf - c.kminB : 0u;
        while ((spanB >> c.shiftB) >= H) c.shiftB++;
        std::vector<uint32_t> cnt(GBT_RANK_CELLS, 0u);
        uint32_t prev = 0;
        for (int i = 0; i < K; i++) {
            // the last threshold is never looked up (x >= tmax returns K first) but must keep the cells monotone
            const uint32_t cell = gbt_rank_cell(c, tab_host[i]);
            GNX_REQUIRE(cell < (uint32_t)GBT_RANK_CELLS && cell >= prev, "gnx_gbt_model_create: rank cells are not monotone");
            prev = cell;
            cnt[cell]++;
        }
        uint32_t start = 0;
        for (int k = 0; k < GBT_RANK_CELLS; k++) {
            lut[k] = start | (cnt[k] << 16);
            start += cnt[k];
        }
    }
    *cells = c;
    GNX_CUDA(cudaMalloc((void**)lut_dev, GBT_RANK_CELLS * sizeof(uint32_t)));
    GNX_CUDA(cudaMemcpy(*lut_dev, lut.data(), GBT_RANK_CELLS * sizeof(uint32_t), cudaMemcpyHostToDevice));
    std::vector<float> padded(tab_host, tab_host + K);
    padded.resize((size_t)K + GBT_RANK_PAD, nanf(""));
    GNX_CUDA(cudaMalloc((void**)tab_dev, padded.size() * sizeof(float)));
    GNX_CUDA(cudaMemcpy(*tab_dev, padded.data(), padded.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// ---------------------------------------------------------------- K4b: tile walk
// (n & TILE_FMASK) | lane4 as ONE opaque LOP3: left to the compiler it is re-associated into mask + (row base + lane4),
// a second (per-thread) add, instead of staying a per-thread offset next to the uniform row base of the load.
__device__ __forceinline__ uint32_t gnx_feat_off(uint32_t n, uint32_t lane4) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(n), "n"(TILE_FMASK), "r"(lane4));
    return d;
}

// One tree for the two adjacent rows of a warp.  rau / rbu = shared address of the rows' first tile word minus the bias of the
// feature field (warp-uniform), lane4 = lane * 4, rec = shared address of the tree's 128-byte record
// (warp-uniform).  The root's feature of row b is the low half of the pair word row a loads.
__device__ __forceinline__ void gbt_tile_tree2(uint32_t rau, uint32_t rbu, uint32_t lane4, uint32_t t0, uint32_t t1, uint32_t t2,
                                               uint32_t a0, uint32_t rec, float& la, float& lb) {
    const uint32_t x0 = gnx_lds_u32(rau + (a0 + lane4));
    const bool ba0 = x0 > t0, bb0 = (x0 << 17) > t0;
    const uint32_t na1 = ba0 ? t2 : t1, nb1 = bb0 ? t2 : t1;
    uint32_t oa = ba0 ? 32u : 0u, ob = bb0 ? 32u : 0u;
    gnx_add_if_gt(oa, gnx_lds_u32(rau + gnx_feat_off(na1, lane4)), na1, 16u);
    gnx_add_if_gt(ob, gnx_lds_u32(rbu + gnx_feat_off(nb1, lane4)), nb1, 16u);
    const uint32_t na2 = gnx_lds_u32(rec + oa), nb2 = gnx_lds_u32(rec + ob);
    gnx_add_if_gt(oa, gnx_lds_u32(rau + gnx_feat_off(na2, lane4)), na2, 8u);
    gnx_add_if_gt(ob, gnx_lds_u32(rbu + gnx_feat_off(nb2, lane4)), nb2, 8u);
    const uint32_t na3 = gnx_lds_u32(rec + 4u + oa), nb3 = gnx_lds_u32(rec + 4u + ob);
    gnx_add_if_gt(oa, gnx_lds_u32(rau + gnx_feat_off(na3, lane4)), na3, 4u);
    gnx_add_if_gt(ob, gnx_lds_u32(rbu + gnx_feat_off(nb3, lane4)), nb3, 4u);
    la = gnx_lds_f32(rec + 64u + oa);
    lb = gnx_lds_f32(rec + 64u + ob);
}

template <typename TOPT> struct TileTopLoad;
template <> struct TileTopLoad<GbtTileTop> {   // one 16-byte uniform load: nodes 0, 1, 2 and node 0's feature offset
    static __device__ __forceinline__ uint4 get(const GbtTileTop& t, int i) { return t.q[i]; }
};

template <int AT, typename TOPT>
__global__ void __launch_bounds__(RK_THREADS, 1)
gbt_smooth_tile_kernel(const __grid_constant__ TOPT topc, GbtDev m, const unsigned char* __restrict__ forest_img, size_t forest_bytes,
                       uint32_t tile_bytes, const uint16_t* __restrict__ R2, const float* __restrict__ B, int64_t N, int W, int nseg, int Lseg,
                       float* __restrict__ proba, int32_t* __restrict__ label) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int A = AT ? AT : m.A;
    constexpr int AMAX = AT ? AT : GBT_MAX_A;
    // shared memory: [ forest records | rank tile ] (the records first, the tile at least GBT_TILE_FBIAS in: row bases are
    // `tile - GBT_TILE_FBIAS`, which must not fall below the shared window).  A warp always walks rows 2 * warp and 2 * warp + 1 of the tile; when the
    // segment is shorter those rows are clamped to its last row (valid addresses, results not written).
    const uint32_t tile_off = max((uint32_t)forest_bytes, GBT_TILE_FBIAS);
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem + tile_off);
    {
        const uint4* src = reinterpret_cast<const uint4*>(forest_img);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (size_t i = threadIdx.x; i < forest_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    const uint32_t forest_s = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t tile_s = forest_s + tile_off;
    const int pad = (m.S + 1) / 2;
    const int Wp = W + m.S - 1;
    const int Lslots = Lseg + m.S - 1;
    const int rounds = m.T / A;
    // a warp reduction lands in a uniform register: the compiler then knows that the warp index (and the row bases
    // derived from it) are warp-uniform and addresses feature loads as [lane part + uniform row base]
    const int warp = (int)__reduce_max_sync(0xffffffffu, threadIdx.x >> 5), lane = threadIdx.x & 31;
    const uint32_t lane4 = (uint32_t)lane * 4u;
    const int64_t nhb = (N + 31) / 32;
    for (int64_t t = blockIdx.x; t < nhb * nseg; t += gridDim.x) {
        const int64_t hb = t / nseg;
        const int sg = (int)(t - hb * nseg);
        const int64_t n = hb * 32 + lane;
        const int w0 = sg * Lseg;
        __syncthreads();
        // stage: the tile's slots are one contiguous run of R2; a uint4 holds 8 lanes of one element.  Tile word =
        // rank << 17 | rank of the same class one slot on (A elements further; 0 behind the last slot of the chromosome)
        int saw_nan = 0;
        {
            const int slots = min(Lslots, Wp - w0);
            const uint4* src = reinterpret_cast<const uint4*>(R2 + ((hb * Wp + w0) * (int64_t)A) * 32);
            const int nq = slots * A * 4;
            const int nq2 = min(Lslots + 1, Wp - w0) * A * 4;   // the slot behind the tile still belongs to this haplotype block
            for (int q = threadIdx.x; q < nq; q += blockDim.x) {
                const uint4 v = __ldg(src + q);
                const uint4 u = (q + 4 * A < nq2) ? __ldg(src + q + 4 * A) : make_uint4(0u, 0u, 0u, 0u);
                uint4 lo, hi;
                lo.x = (v.x << 17) | (u.x & 0xffffu); lo.y = ((v.x >> 16) << 17) | (u.x >> 16);
                lo.z = (v.y << 17) | (u.y & 0xffffu); lo.w = ((v.y >> 16) << 17) | (u.y >> 16);
                hi.x = (v.z << 17) | (u.z & 0xffffu); hi.y = ((v.z >> 16) << 17) | (u.z >> 16);
                hi.z = (v.w << 17) | (u.w & 0xffffu); hi.w = ((v.w >> 16) << 17) | (u.w >> 16);
                // NaN marker 0xFFFF in either half of any word
                // (also in the slot behind the tile: its ranks are the low halves of the last slot's words)
                saw_nan |= ((v.x & 0xffffu) == 0xffffu) | ((v.x >> 16) == 0xffffu) | ((v.y & 0xffffu) == 0xffffu) | ((v.y >> 16) == 0xffffu) |
                           ((v.z & 0xffffu) == 0xffffu) | ((v.z >> 16) == 0xffffu) | ((v.w & 0xffffu) == 0xffffu) | ((v.w >> 16) == 0xffffu) |
                           ((u.x & 0xffffu) == 0xffffu) | ((u.x >> 16) == 0xffffu) | ((u.y & 0xffffu) == 0xffffu) | ((u.y >> 16) == 0xffffu) |
                           ((u.z & 0xffffu) == 0xffffu) | ((u.z >> 16) == 0xffffu) | ((u.w & 0xffffu) == 0xffffu) | ((u.w >> 16) == 0xffffu);
                uint4* dst = reinterpret_cast<uint4*>(tile + (size_t)q * 8);
                dst[0] = lo;
                dst[1] = hi;
            }
        }
        const int slow = __syncthreads_or(saw_nan);
        // The walk runs in warp-uniform control flow whatever the data (uniform registers for the record and row
        // bases): rows beyond the segment / chromosome are walked and not written; a tile holding NaN is walked too
        // (harmlessly: every address is valid) and then redone by the generic traversal.
        // (through a warp reduction, like `warp`: the row bases below stay in uniform registers and every feature load
        // is [lane part + uniform row base])
        const int last = (int)__reduce_max_sync(0xffffffffu, min(Lseg, W - w0) - 1);
        const int wla = 2 * warp, wlb = 2 * warp + 1;
        {
            const uint32_t rau = tile_s - GBT_TILE_FBIAS + (uint32_t)(min(wla, last) * A * 128);
            const uint32_t rbu = tile_s - GBT_TILE_FBIAS + (uint32_t)(min(wlb, last) * A * 128);
            float pa[AMAX], pb[AMAX];
#pragma unroll
            for (int c = 0; c < AMAX; c++) pa[c] = pb[c] = 0.f;
            int tb = 0;
            uint32_t rec = forest_s;
#pragma unroll 1
            for (int rd = 0; rd < rounds; rd++) {
#pragma unroll
                for (int c = 0; c < AMAX; c++) {
                    if (c < A) {
                        const uint4 tp = TileTopLoad<TOPT>::get(topc, tb + c);
                        float la, lb;
                        gbt_tile_tree2(rau, rbu, lane4, tp.x, tp.y, tp.z, tp.w, rec + c * 128, la, lb);
                        pa[c] = GNX_FADD(pa[c], la);
                        pb[c] = GNX_FADD(pb[c], lb);
                    }
                }
                tb += A;
                rec += 128 * A;
            }
            if (n < N && !slow) {
                if (wla <= last) {
                    const int64_t ra = n * W + w0 + wla;
                    gbt_finish<AT>(m, pa, proba ? proba + ra * A : nullptr, label ? label + ra : nullptr);
                }
                if (wlb <= last) {
                    const int64_t rb = n * W + w0 + wlb;
                    gbt_finish<AT>(m, pb, proba ? proba + rb * A : nullptr, label ? label + rb : nullptr);
                }
            }
        }
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
                    bp[e] = (j < Wp) ? __ldg(B + (nh * W + spad_to_orig(j, W, pad)) * A + a) : 0.f;
                }
                __syncthreads();
                for (int wl = threadIdx.x; wl < Lseg && w0 + wl < W; wl += blockDim.x) {
                    float psum[AMAX];
                    gbt_eval_row<AT>(m, m.nodes, m.leaves, bp + (size_t)wl * A, psum);
                    const int64_t row = nh * W + w0 + wl;
                    gbt_finish<AT>(m, psum, proba ? proba + row * A : nullptr, label ? label + row : nullptr);
                }
            }
        }
    }
}

int gbt_tile_smooth(const gnx_gbt* m, const float* B_dev, int64_t N, int W, float* proba_dev, int32_t* label_dev, cudaStream_t st) {
    const int A = m->d.A, S = m->d.S;
    const size_t smem_max = 227 * 1024;
    const size_t slot_bytes = (size_t)A * 128;
    const size_t tile_off = std::max<size_t>(m->tile_forest_bytes, GBT_TILE_FBIAS);
    if (tile_off + 64 >= smem_max) return -1;
    const size_t room = smem_max - tile_off;
    const int64_t fit = (int64_t)(room / slot_bytes) - (S - 1);
    const int Lmax = (int)std::min<int64_t>(fit, 2 * TILE_WARPS);
    if (Lmax < TILE_WARPS && Lmax < W) return -1;   // a tile must give every warp a window
    const int nseg = (int)ceil_div(W, std::min(Lmax, W));
    const int Lseg = (int)ceil_div(W, nseg);
    const int Wp = W + S - 1;
    const int64_t nhb = ceil_div(N, 32);
    const size_t tile_bytes = (size_t)(Lseg + S - 1) * slot_bytes;
    const size_t smem = tile_bytes + tile_off;
    // the rank scratch comes from the device's stream-ordered pool; keep freed blocks in the pool across calls
    {
        static bool pool_kept[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !pool_kept[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_kept[dev] = true;
        }
    }
    uint16_t* R2 = nullptr;
    GNX_CUDA(cudaMallocAsync((void**)&R2, (size_t)nhb * Wp * A * 32 * sizeof(uint16_t), st));
    if (m->profile) GNX_CUDA(cudaEventRecord(m->ev[0], st));
    {
        const size_t fixed = (size_t)RANK_EL * RANK_TS * 2 + GBT_RANK_CELLS * 4;
        const size_t tab_bytes = (size_t)(m->d.K + GBT_RANK_PAD) * 4;
        const int in_smem = fixed + tab_bytes <= 110 * 1024;   // two CTAs per SM
        const size_t rsm = fixed + (in_smem ? tab_bytes : 0);
        const int64_t items = nhb * ceil_div((int64_t)Wp * A, RANK_EL);
        const int grid = (int)std::min<int64_t>(items, (int64_t)sm_count() * 2);
        if (in_smem) {
            GNX_CUDA(cudaFuncSetAttribute(gbt_rank_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(110 * 1024)));
            gbt_rank_tile_kernel<true><<<grid, RANK_THREADS, rsm, st>>>(m->rank_tab, m->rank_lut, m->d.K, m->rank_cells, B_dev, N, W, A, S, R2);
        } else {
            GNX_CUDA(cudaFuncSetAttribute(gbt_rank_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(110 * 1024)));
            gbt_rank_tile_kernel<false><<<grid, RANK_THREADS, rsm, st>>>(m->rank_tab, m->rank_lut, m->d.K, m->rank_cells, B_dev, N, W, A, S, R2);
        }
        GNX_CUDA(cudaGetLastError());
    }
    if (m->profile) GNX_CUDA(cudaEventRecord(m->ev[1], st));
    const int64_t tiles = nhb * nseg;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count());
#define LAUNCHT(AT, TOPT, TOPV)                                                                                                \
    do {                                                                                                                       \
        GNX_CUDA(cudaFuncSetAttribute(gbt_smooth_tile_kernel<AT, TOPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gbt_smooth_tile_kernel<AT, TOPT><<<grid, RK_THREADS, smem, st>>>(TOPV, m->d, m->tile_forest, m->tile_forest_bytes, (uint32_t)tile_bytes, \
                                                                         R2, B_dev, N, W, nseg, Lseg, proba_dev, label_dev);  \
    } while (0)
#define CALLT(AT) LAUNCHT(AT, GbtTileTop, *m->h_tiletop)
    switch (A) {
        case 2: CALLT(2); break;
        case 3: CALLT(3); break;
        case 4: CALLT(4); break;
        case 5: CALLT(5); break;
        case 6: CALLT(6); break;
        case 7: CALLT(7); break;
        case 8: CALLT(8); break;
        default: CALLT(0); break;
    }
#undef CALLT
#undef LAUNCHT
    GNX_CUDA(cudaGetLastError());
    if (m->profile) GNX_CUDA(cudaEventRecord(m->ev[2], st));
    GNX_CUDA(cudaFreeAsync(R2, st));
    return 0;
}

}  // namespace gnx
