// lr_base_tc.cu -- K1 on the 5th-generation tensor cores (sm_100a only).
//
// One persistent, warp-specialised CTA per SM:
//   warp 0      TMA producer: X chunk [256 haplotypes x 128 SNPs] int8 (32 KB, 128B swizzle)
//               + one weight tile [64 limb columns x 128 SNPs] per window covering the chunk,
//               placed in slot order inside a 32 KB weight stage
//   warp 1      MMA issuer: tcgen05.mma.kind::i8, M=128 (x2 haplotype halves), K=32, N=64 per
//               window -- the windows covering a chunk sit in consecutive TMEM slots and are
//               issued as ONE instruction of N = 64..256 so the X tile is read from shared
//               memory once per step; accumulators int32 in TMEM, 2 halves x 4 slots x 64 columns
//   warp 2      TMEM allocator
//   warps 4-19  epilogue (two groups of 8 warps on alternate windows): tcgen05.ld -> limb recombination (int64) -> float64 sigmoid /
//               normalise -> B[n, w, :] (float32 or float64)
// A CTA walks the SNP axis left to right for its 256 haplotypes; consecutive windows
// overlap by 2*ctx SNPs, so each staged X chunk feeds the <= 4 windows that cover it
// and X is read from HBM exactly once per window block.
//
// Replaces src/Base/base.py:146-180 + sklearn LogisticRegression.predict_proba
// (src/Base/models.py:12-21).
#include <cuda.h>

#include <stdlib.h>

#include <algorithm>

#include "lr_base.cuh"

namespace gnx {

namespace tc {

constexpr int TILE_HAPS = 256;
constexpr int X_STAGE_BYTES = TILE_HAPS * LR_KC;  // 32 KB
constexpr int N_SLOTS = 4;                         // live windows (TMEM accumulator slots)
constexpr int W_TILE_BYTES = LR_TILE_BYTES;        // 8 KB: one window's weights for one chunk
constexpr int W_STAGE_BYTES = N_SLOTS * W_TILE_BYTES;  // 32 KB: the weights of every window covering a chunk, in slot order
constexpr int SX = 4;
constexpr int SW = 3;
constexpr int HALF_COLS = N_SLOTS * LR_NCOLS;      // TMEM columns of one haplotype half: 4 slots x 64
constexpr int TMEM_COLS = 512;
constexpr int N_EPI_GROUPS = 2;  // epilogue groups take alternate windows (the float64 epilogue is latency-bound)
constexpr int N_THREADS = 32 * (4 + 8 * N_EPI_GROUPS);
constexpr int EPI_WARP0 = 4;
constexpr int N_EPI_WARPS = 8;
constexpr int SMEM_BYTES = 1024 + SX * X_STAGE_BYTES + SW * W_STAGE_BYTES + 512;
constexpr uint32_t META_END = 0xffffffffu;  // producer -> MMA issuer: no more chunks

constexpr uint64_t HINT_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t HINT_EVICT_LAST = 0x14F0000000000000ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, both operands K-major
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int32_t* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

struct Sched {
    int n_htiles, n_wblocks, wb;  // haplotype tiles, window blocks, windows per block
    int n_units;
};

template <int APAD, typename OutT>
__global__ void __launch_bounds__(N_THREADS, 1)
lr_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, LrDev m, Sched sc,
             int64_t N, OutT* __restrict__ B) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* smem_x = smem;
    unsigned char* smem_w = smem + SX * X_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SX * X_STAGE_BYTES + SW * W_STAGE_BYTES);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + SX;
    uint64_t* w_full = x_empty + SX;
    uint64_t* w_empty = w_full + SW;
    uint64_t* acc_full = w_empty + SW;
    uint64_t* acc_empty = acc_full + N_SLOTS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + N_SLOTS);
    // per weight stage: first slot | n windows << 2 | first-chunk mask << 5 | last-chunk mask << 9
    volatile uint32_t* meta = tmem_slot + 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SX; i++) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < SW; i++) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < N_SLOTS; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], N_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------- TMA producer
        // Warp-uniform loop, one elected lane issues.  Per 128-SNP chunk: the X tile (256
        // haplotypes x 128 SNPs) and, into one weight stage, the tile of every window that covers
        // the chunk, each at the position of its TMEM slot so that windows in consecutive slots
        // form one contiguous B operand.  The per-chunk schedule is one 16-byte table entry.
        int xs = 0, ws = 0;
        uint32_t xph = 0, wph = 0;
        uint32_t wbase = 0;  // running window counter of this CTA (TMEM slot = counter mod 4)
        for (int u = blockIdx.x; u < sc.n_units; u += gridDim.x) {
            const int wblk = u / sc.n_htiles, ht = u - wblk * sc.n_htiles;
            const int w_lo = wblk * sc.wb, w_hi = min(m.W, w_lo + sc.wb);
            const int kb = __ldg(m.k0 + w_lo), ke = __ldg(m.kend + w_hi - 1);
            const int hap0 = ht * TILE_HAPS;
            uint4 sch = __ldg(m.chunk_sched + kb);
            int cw0 = __ldg(m.chunk_w0 + kb);
            for (int k = kb; k < ke; k++) {
                mbar_wait(&x_empty[xs], xph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&x_full[xs], X_STAGE_BYTES);
                    tma_load_2d(&tmX, &x_full[xs], smem_x + xs * X_STAGE_BYTES, k * LR_KC, hap0, HINT_EVICT_FIRST);
                }
                __syncwarp();
                if (++xs == SX) { xs = 0; xph ^= 1; }
                const uint4 cur = sch;
                const int c0 = cw0;
                if (k + 1 < ke) {  // prefetch the next chunk's entry
                    sch = __ldg(m.chunk_sched + k + 1);
                    cw0 = __ldg(m.chunk_w0 + k + 1);
                }
                const uint32_t ent[4] = {cur.x, cur.y, cur.z, cur.w};
                // windows of this block covering the chunk: indices [ia, iz) of the entry
                const int ia = max(0, w_lo - c0);
                int iz = min(4, w_hi - c0);
#pragma unroll
                for (int i = 3; i >= 0; i--)
                    if (!(ent[i] & 4u) && iz > i) iz = i;
                const uint32_t slot0 = (wbase + (uint32_t)(c0 + ia - w_lo)) & (N_SLOTS - 1);
                uint32_t mt = slot0 | ((uint32_t)(iz - ia) << 2);
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (i >= ia && i < iz) mt |= ((ent[i] & 1u) << (5 + i - ia)) | (((ent[i] >> 1) & 1u) << (9 + i - ia));
                mbar_wait(&w_empty[ws], wph ^ 1);
                if (elect_one()) {
                    meta[ws] = mt;
                    if (m.dbg & 4) {
                        mbar_arrive(&w_full[ws]);
                    } else {
                        mbar_expect_tx(&w_full[ws], (uint32_t)(iz - ia) * W_TILE_BYTES);
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            if (i >= ia && i < iz)
                                tma_load_2d(&tmW, &w_full[ws], smem_w + ws * W_STAGE_BYTES + ((slot0 + i - ia) & (N_SLOTS - 1)) * W_TILE_BYTES, 0,
                                            (int)(ent[i] >> 3) * LR_NCOLS, HINT_EVICT_LAST);
                    }
                }
                __syncwarp();
                if (++ws == SW) { ws = 0; wph ^= 1; }
            }
            wbase += (uint32_t)(w_hi - w_lo);
        }
        mbar_wait(&w_empty[ws], wph ^ 1);
        if (elect_one()) {
            meta[ws] = META_END;
            mbar_arrive(&w_full[ws]);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // All 32 lanes run the loop with warp-uniform control flow and operands (descriptor and
        // address arithmetic stays on the uniform datapath); one elected lane issues the tcgen05
        // instructions.  Per chunk and 32-SNP step the windows in consecutive TMEM slots with the
        // same accumulate flag are ONE instruction of N = 64 x (windows): the X tile is fetched
        // from shared memory once for all of them.
        // idesc: D=S32 (2<<4), A=S8 (1<<7), B=S8 (1<<10), K-major both, M=128 (8<<24); N added per instruction
        constexpr uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
        int xs = 0, ws = 0;
        uint32_t xph = 0, wph = 0, slot_use = 0;  // bit s of slot_use: parity of the uses of TMEM slot s
        const bool no_mma = (m.dbg & 1) != 0;
        const uint32_t smem_x_u32 = smem_u32(smem_x), smem_w_u32 = smem_u32(smem_w);
        for (;;) {
            mbar_wait(&w_full[ws], wph);
            const uint32_t mt = __shfl_sync(0xffffffffu, meta[ws], 0);
            if (mt == META_END) break;
            const uint32_t slot0 = mt & 3u, nwin = (mt >> 2) & 7u, fmask = (mt >> 5) & 15u, lmask = (mt >> 9) & 15u;
            mbar_wait(&x_full[xs], xph);
            for (uint32_t i = 0; i < nwin; i++)
                if ((fmask >> i) & 1u) {
                    const uint32_t slot = (slot0 + i) & (N_SLOTS - 1);
                    mbar_wait(&acc_empty[slot], ((slot_use >> slot) & 1u) ^ 1u);
                    slot_use ^= 1u << slot;
                }
            tc_fence_after();
            const uint64_t xdesc = make_desc(smem_x_u32 + xs * X_STAGE_BYTES);
            const uint64_t wdesc = make_desc(smem_w_u32 + ws * W_STAGE_BYTES);
            if (elect_one()) {
                if (!no_mma) {
#pragma unroll
                    for (int j = 0; j < LR_KC / 32; j++) {
                        // segments of windows [i0, i1): consecutive slots without wrap, same accumulate flag
                        uint32_t i0 = 0;
                        while (i0 < nwin) {
                            const uint32_t s0 = (slot0 + i0) & (N_SLOTS - 1);
                            const uint32_t acc0 = (j == 0 && ((fmask >> i0) & 1u)) ? 0u : 1u;
                            uint32_t i1 = i0 + 1;
                            while (i1 < nwin && s0 + (i1 - i0) < N_SLOTS && ((j == 0 && ((fmask >> i1) & 1u)) ? 0u : 1u) == acc0) i1++;
                            const uint32_t idesc = idesc0 | (((i1 - i0) * LR_NCOLS >> 3) << 17);
                            const uint64_t bdesc = wdesc + (uint64_t)((s0 * W_TILE_BYTES + j * 32) >> 4);
#pragma unroll
                            for (int h = 0; h < 2; h++)
                                mma_i8(tmem_base + h * HALF_COLS + s0 * LR_NCOLS, xdesc + (uint64_t)((h * (128 * LR_KC) + j * 32) >> 4), bdesc,
                                       idesc, acc0);
                            i0 = i1;
                        }
                    }
                }
                tc_commit(&w_empty[ws]);
                tc_commit(&x_empty[xs]);
                for (uint32_t i = 0; i < nwin; i++)
                    if ((lmask >> i) & 1u) tc_commit(&acc_full[(slot0 + i) & (N_SLOTS - 1)]);
            }
            __syncwarp();
            if (++ws == SW) { ws = 0; wph ^= 1; }
            if (++xs == SX) { xs = 0; xph ^= 1; }
        }
    } else if (warp >= EPI_WARP0) {
        // -------------------------------------------------------------- epilogue
        const int e = (warp - EPI_WARP0) & 7, grp = (warp - EPI_WARP0) >> 3;
        const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
        const int half = e >> 2;
        uint32_t wbase = 0;
        for (int u = blockIdx.x; u < sc.n_units; u += gridDim.x) {
            const int wblk = u / sc.n_htiles, ht = u - wblk * sc.n_htiles;
            const int w_lo = wblk * sc.wb, w_hi = min(m.W, w_lo + sc.wb);
            const int64_t n = (int64_t)ht * TILE_HAPS + half * 128 + quad * 32 + lane;
            for (int w = w_lo; w < w_hi; w++) {
                const uint32_t widx = wbase + (uint32_t)(w - w_lo);
                if ((int)(widx % N_EPI_GROUPS) != grp) continue;
                const uint32_t slot = widx & (N_SLOTS - 1);
                mbar_wait(&acc_full[slot], (widx >> 2) & 1);
                tc_fence_after();
                int32_t acc[LR_NCOLS];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + half * HALF_COLS + slot * LR_NCOLS;
                tmem_ld32(taddr, acc);
                tmem_ld32(taddr + 32, acc + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[slot]);
                if (n < N && !(m.dbg & 2)) lr_epilogue_store<APAD, OutT>(acc, m, w, B + (n * m.W + w) * m.A);
            }
            wbase += (uint32_t)(w_hi - w_lo);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<encode_tiled_fn>(p);
    return fn;
}

static int make_map_u8_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride, uint32_t box_inner,
                          uint32_t box_rows) {
    encode_tiled_fn enc = get_encode();
    GNX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t gdim[2] = {inner, rows};
    cuuint64_t gstride[1] = {row_stride};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GNX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) inner=%llu rows=%llu stride=%llu", (int)r, (unsigned long long)inner,
                (unsigned long long)rows, (unsigned long long)row_stride);
    return 0;
}

}  // namespace tc

bool lr_tc_supported(const gnx_lr* m, const int8_t* X, int64_t ldX) {
    int maxlive = 0;
    for (int v : m->h_chunk_wn) maxlive = std::max(maxlive, v);
    return maxlive <= tc::N_SLOTS && (ldX % 16 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
}

// Window-block count that best fills `ctas` persistent CTAs (tail effect vs. the
// 2*ctx SNPs re-read at every block seam).
static tc::Sched choose_sched(const gnx_lr* m, int64_t N, int ctas) {
    tc::Sched best{};
    double best_score = -1.0;
    const int W = m->d.W;
    const int n_ht = (int)ceil_div(N, tc::TILE_HAPS);
    for (int nb = 1; nb <= std::min(W, 256); nb++) {
        const int wb = (int)ceil_div(W, nb);
        const int nbe = (int)ceil_div(W, wb);
        const int64_t units = (int64_t)n_ht * nbe;
        const double eff = (double)units / (double)(ceil_div(units, ctas) * ctas);
        const double seam = (double)(wb * m->d.M) / (double)(wb * m->d.M + 2 * m->d.ctx + LR_KC);
        const double score = eff * seam;
        if (score > best_score + 1e-9) {
            best_score = score;
            best.n_htiles = n_ht;
            best.n_wblocks = nbe;
            best.wb = wb;
            best.n_units = (int)units;
        }
    }
    return best;
}

template <int APAD, typename OutT>
static int launch_t(const CUtensorMap& tmX, const CUtensorMap& tmW, const gnx_lr* m, const tc::Sched& sc, int grid, int64_t N, OutT* B,
                    cudaStream_t st) {
    GNX_CUDA(cudaFuncSetAttribute(tc::lr_tc_kernel<APAD, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    tc::lr_tc_kernel<APAD, OutT><<<grid, tc::N_THREADS, tc::SMEM_BYTES, st>>>(tmX, tmW, m->d, sc, N, B);
    GNX_CUDA(cudaGetLastError());
    return 0;
}

int lr_launch_tc(const gnx_lr* mc, const int8_t* X, int64_t N, int64_t ldX, void* B, bool f64, cudaStream_t st) {
    gnx_lr* m = const_cast<gnx_lr*>(mc);
    if (!lr_tc_supported(m, X, ldX)) {
        // windows narrower than a 128-SNP chunk (more than 4 live accumulators) or a
        // haplotype matrix TMA cannot address: same exact arithmetic on the CUDA cores
        return lr_launch_dp4a_any(m, X, N, ldX, B, f64, st);
    }
    if (!m->tmap_w_ready) {
        if (tc::make_map_u8_2d(reinterpret_cast<CUtensorMap*>(m->tmap_w), m->d.wt, LR_KC, (uint64_t)m->n_tiles * LR_NCOLS, LR_KC, LR_KC, LR_NCOLS))
            return 1;
        m->tmap_w_ready = true;
    }
    alignas(64) CUtensorMap tmX;
    if (tc::make_map_u8_2d(&tmX, X, (uint64_t)m->d.C, (uint64_t)N, (uint64_t)ldX, LR_KC, tc::TILE_HAPS)) return 1;
    const int sms = sm_count();
    GNX_REQUIRE(sms > 0, "no SMs?");
    {
        const char* e = getenv("GNX_LR_DBG");
        m->d.dbg = e ? atoi(e) : 0;
    }
    const tc::Sched sc = choose_sched(m, N, sms);
    const int grid = std::min(sms, sc.n_units);
    const CUtensorMap& tmW = *reinterpret_cast<const CUtensorMap*>(m->tmap_w);
    if (m->d.apad == 8)
        return f64 ? launch_t<8, double>(tmX, tmW, m, sc, grid, N, (double*)B, st) : launch_t<8, float>(tmX, tmW, m, sc, grid, N, (float*)B, st);
    return f64 ? launch_t<16, double>(tmX, tmW, m, sc, grid, N, (double*)B, st) : launch_t<16, float>(tmX, tmW, m, sc, grid, N, (float*)B, st);
}

}  // namespace gnx
