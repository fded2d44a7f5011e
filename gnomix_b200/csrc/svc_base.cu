// svc_base.cu -- K2 + K3: CovRSKBase.predict_proba for all windows.
//
// Replaces, per window, sklearn.svm.SVC(kernel=CovRSK_DP_triangular_numbers, probability=
// True).predict_proba as configured at src/Base/models.py:195-215 and driven by
// Base.predict_proba_vectorized (src/Base/base.py:146-180):
//   K2  the string kernel  src/Base/string_kernel.py:91-111
//         K(x, y) = sum over positions of cov_tri, i.e. (closed form)
//         K(x, y) = sum_{m in Ms} #{length-m substrings on which x == y everywhere}
//       evaluated on 2-bit planes of the int8 haplotypes: z = match word (32 SNPs),
//       R_m = positions ending an all-match stretch of length >= m, K += popc(R_m) for the
//       short lengths (doubling chain 1,2,4,8,16,32) and a run-length table for the long ones
//       (a maximal run of length L adds sum_{m in Ms, m > 32, m <= L} (L - m + 1)).
//   K3  libsvm svm_predict_probability on the precomputed kernel row: pairwise decision
//       values in support-vector order, Platt sigmoid, pairwise coupling
//       (multiclass_probability) -- oracle/gnx_oracle.c orc_svc_proba, same operation order.
//
// Layout: one launch pair per window.  K2: a CTA owns 64 query haplotypes of the window
// (bit planes built once in shared memory with warp ballots straight from the int8 matrix,
// reflect pad applied by index) and streams the window's support vectors through shared
// memory in chunks; lanes = queries, support vector words are broadcasts.  The kernel row
// is written transposed (Kt[sv][hap]) so that K2's stores and K3's loads are coalesced.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "common.cuh"
#include "host_pack.h"

namespace gnx {

constexpr int SVC_MAX_A = 16;
constexpr int SVC_Q = 64;        // query haplotypes per CTA
constexpr int SVC_CHUNK = 48;    // support vectors staged per CTA (3 CTAs per SM fit in shared memory)
constexpr int SVC_THREADS = 256;

struct SvcWin {
    int64_t lo;       // first padded column of the window
    int len, nw, nwp; // features, 32-bit words, padded (odd) word stride
    int nsv;
    const uint32_t* planes;   // [nsv][2][nw]
    const double* coef;       // [A-1][nsv]
    const int32_t* n_support; // [A]
    const double *intercept, *probA, *probB;  // [P]
};

struct SvcDev {
    int A, P, W, fast, small_mask, min_big;
    int64_t C, M, ctx;
    const int32_t* gbig;   // [maxlen+1] run-length table of the lengths > 32
    const uint8_t* ohe;    // [maxlen+2] 1 iff length in Ms (generic path)
};

__device__ __forceinline__ int64_t svc_pad_to_orig(int64_t p, int64_t C, int64_t ctx) {
    if (p < ctx) return ctx - 1 - p;
    if (p >= ctx + C) return 2 * C + ctx - 1 - p;
    return p - ctx;
}

// ------------------------------------------------------------------------------ K2
// grid = (ceil(N / 64), ceil(nsv / 64)).  Kt != NULL: Kt[s * ldK + n]; else Krow[n * nsv + s].
template <int SM_T>
__global__ void __launch_bounds__(SVC_THREADS, 3)
svc_kernel_window(SvcDev m, SvcWin w, const uint32_t* __restrict__ QP, int64_t qp_words, int64_t N, int32_t* __restrict__ Kt,
                  int64_t ldK, int32_t* __restrict__ Krow) {
    extern __shared__ __align__(16) uint32_t sm[];
    uint32_t* qp = sm;                               // [64][2][nwp]
    uint32_t* sp = sm + (size_t)SVC_Q * 2 * w.nwp;   // [SVC_CHUNK][2][nw]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n0 = (int64_t)blockIdx.x * SVC_Q;
    const int nw = w.nw, nwp = w.nwp;

    // ---- query bit planes of this window: an unaligned bit range of the packed padded planes
    {
        const int64_t wi0 = w.lo >> 5;
        const int sh = (int)(w.lo & 31);
        for (int i = threadIdx.x; i < SVC_Q * 2 * nw; i += SVC_THREADS) {
            const int q = i / (2 * nw), rem = i - q * 2 * nw;
            const int pl = rem / nw, j = rem - pl * nw;
            const int64_t n = n0 + q;
            uint32_t v = 0u;
            if (n < N) {
                const uint32_t* src = QP + ((size_t)n * 2 + pl) * qp_words + wi0 + j;
                v = __funnelshift_r(__ldg(src), __ldg(src + 1), sh);
            }
            qp[(q * 2 + pl) * nwp + j] = v;
        }
    }
    const uint32_t tail = (w.len & 31) ? ((1u << (w.len & 31)) - 1u) : 0xffffffffu;
    const uint32_t* q0a = qp + ((lane) * 2) * nwp;
    const uint32_t* q1a = qp + ((lane + 32) * 2) * nwp;

    {   // one support-vector chunk per CTA (grid.y)
        const int s0 = blockIdx.y * SVC_CHUNK;
        const int ns = min(SVC_CHUNK, w.nsv - s0);
        __syncthreads();
        {
            const uint32_t* src = w.planes + (size_t)s0 * 2 * nw;
            for (int i = threadIdx.x; i < ns * 2 * nw; i += SVC_THREADS) sp[i] = __ldg(src + i);
        }
        __syncthreads();
        // warp handles support vectors warp*6 .. warp*6+5 of the chunk, 3 at a time, x 2 queries per lane
        constexpr int SB = 3;
        for (int sb = warp * (SVC_CHUNK / 8); sb < (warp + 1) * (SVC_CHUNK / 8) && sb < ns; sb += SB) {
            int K[2][SB];
            uint32_t zp[2][SB], r2p[2][SB], r4p[2][SB], r8p[2][SB], r16p[2][SB];
            int run[2][SB];
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < SB; b++) {
                    K[a][b] = 0;
                    zp[a][b] = r2p[a][b] = r4p[a][b] = r8p[a][b] = r16p[a][b] = 0u;
                    run[a][b] = 0;
                }
            for (int j = 0; j < nw; j++) {
                uint32_t xq[2][2];
                xq[0][0] = q0a[j]; xq[0][1] = q0a[nwp + j];
                xq[1][0] = q1a[j]; xq[1][1] = q1a[nwp + j];
                const uint32_t msk = (j == nw - 1) ? tail : 0xffffffffu;
#pragma unroll
                for (int b = 0; b < SB; b++) {
                    const int s = min(sb + b, ns - 1);
                    const uint32_t y0 = sp[(s * 2 + 0) * nw + j], y1 = sp[(s * 2 + 1) * nw + j];
#pragma unroll
                    for (int a = 0; a < 2; a++) {
                        const int smask = (SM_T >= 0) ? SM_T : m.small_mask;
                        const uint32_t z = ~((xq[a][0] ^ y0) | (xq[a][1] ^ y1)) & msk;
                        int k = 0;
                        if (smask & 1) k += __popc(z);
                        const uint32_t r2 = z & __funnelshift_l(zp[a][b], z, 1);
                        if (smask & 2) k += __popc(r2);
                        const uint32_t r4 = r2 & __funnelshift_l(r2p[a][b], r2, 2);
                        if (smask & 4) k += __popc(r4);
                        const uint32_t r8 = r4 & __funnelshift_l(r4p[a][b], r4, 4);
                        if (smask & 8) k += __popc(r8);
                        if (smask & 48) {
                            const uint32_t r16 = r8 & __funnelshift_l(r8p[a][b], r8, 8);
                            if (smask & 16) k += __popc(r16);
                            if (smask & 32) k += __popc(r16 & __funnelshift_l(r16p[a][b], r16, 16));
                            r16p[a][b] = r16;
                        }
                        zp[a][b] = z; r2p[a][b] = r2; r4p[a][b] = r4; r8p[a][b] = r8;
                        // maximal runs longer than 32 cross a word boundary: branch-free bookkeeping of the
                        // run that is open at the top of the word; the table lookup only when a long run ends
                        const uint32_t nz = ~z;
                        const int t1 = __clz(__brev(nz));          // trailing ones (32 if the word is all ones)
                        const int L = run[a][b] + t1;
                        const bool open = (nz == 0u);
                        if (!open && L >= m.min_big) k += __ldg(m.gbig + L);
                        run[a][b] = open ? L : __clz(nz);          // leading ones carry into the next word
                        K[a][b] += k;
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < SB; b++)
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    int k = K[a][b];
                    if (run[a][b] >= m.min_big) k += __ldg(m.gbig + run[a][b]);
                    const int64_t n = n0 + lane + 32 * a;
                    const int s = s0 + sb + b;
                    if (n < N && sb + b < ns) {
                        if (Kt) Kt[(int64_t)s * ldK + n] = k;
                        else Krow[n * w.nsv + s] = k;
                    }
                }
        }
    }
}

// ------------------------------------------------------------------------------ K2, production form
// The same sums for the reference's own Ms list -- lengths {1, 4, 8} below 33 (every CovSample(., 0.6, 1.0, 37) list
// starts 1, 4, 8, 39, ...) -- for a GROUP of windows per launch (blockIdx.z), rebuilt around the two pipes the first
// kernel saturates together (ncu, profiles/r2_ncu_svc_summary.txt: 43 instructions and 5 quarter-rate POPC / FLO per
// word pair):
//   * counts: popc(z) + popc(r4) + popc(r8) summed over the words of a window is a population count of 3 * nw words;
//     they go through a carry-save adder tree (Harley-Seal: bit-sliced accumulators ones / twos / fours, 10 CSAs = 20
//     LOP3 per 4 words) and only the two carries that leave a 4-word block are counted with POPC: 0.5 instead of 3
//     quarter-rate instructions per word;
//   * long runs: the open run is kept as its START position, so an all-match word costs nothing but the predicate;
//     the run that closes in a word ends at its first mismatch, popc(z & (~z - 1)) SNPs in, and the next one starts
//     after the last mismatch (FLO); its table g(L) = sum_{m in Ms, 32 < m <= L} (L - m + 1) sits in shared memory;
//   * the word loop is unrolled by 4 with every shared-memory address a base register + immediate; the tail mask and
//     one all-mismatch word behind the window (it closes a run that reaches the window's end) are folded into the
//     operands once per word instead of once per pair.
// Results are the same integers as svc_kernel_window<13> (GNX_SVC_KERNEL=0 selects that one; tests compare both).
__device__ __forceinline__ void svc_csa(uint32_t& acc, uint32_t a, uint32_t b, uint32_t& carry) {
    const uint32_t u = acc ^ a;
    carry = (acc & a) | (u & b);
    acc = u ^ b;
}

struct SvcPair {
    uint32_t zp, r2p, r4p;        // previous word of the doubling chain
    uint32_t ones, twos, fours;   // bit-sliced counters of the carry-save tree
    uint32_t hold_in, hold_t, hold_f;
    int start, k;
};

// one word of one (query, support vector) pair; phase = word index within the 4-word block (compile time)
__device__ __forceinline__ int svc_lds_s32(uint32_t saddr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

template <int PHASE>
__device__ __forceinline__ void svc_word(SvcPair& s, uint32_t nz, int base, int min_big, uint32_t gb) {
    const uint32_t z = ~nz;
    const uint32_t r2 = z & __funnelshift_l(s.zp, z, 1);
    const uint32_t r4 = r2 & __funnelshift_l(s.r2p, r2, 2);
    const uint32_t r8 = r4 & __funnelshift_l(s.r4p, r4, 4);
    s.zp = z; s.r2p = r2; s.r4p = r4;
    // run bookkeeping
    const int t1 = __popc(z & (nz - 1u));                 // matches before the first mismatch (32 when there is none)
    const int L = base + t1 - s.start;
    const bool closed = nz != 0u;
    // branch-free: the run table is 0 below the first long length and at entry 0 (gb = its shared address); a closing
    // run has 0 <= L <= window length
    s.k += svc_lds_s32(gb + 4u * (uint32_t)(closed ? L : 0));
    const int ns = base + 32 - __clz(nz);                 // position behind the last mismatch
    s.start = closed ? ns : s.start;
    // carry-save tree over the inputs z, r4, r8 of four consecutive words
    uint32_t c;
    if (PHASE == 0) {
        svc_csa(s.ones, z, r4, s.hold_t);
        s.hold_in = r8;
    } else if (PHASE == 1) {
        svc_csa(s.ones, s.hold_in, z, c);
        uint32_t c2;
        svc_csa(s.ones, r4, r8, c2);
        uint32_t f1;
        svc_csa(s.twos, s.hold_t, c, f1);
        s.hold_t = c2;
        s.hold_f = f1;
    } else if (PHASE == 2) {
        svc_csa(s.ones, z, r4, c);
        uint32_t f2, e1;
        svc_csa(s.twos, s.hold_t, c, f2);
        svc_csa(s.fours, s.hold_f, f2, e1);
        s.k += __popc(e1) << 3;
        s.hold_in = r8;
    } else {
        svc_csa(s.ones, s.hold_in, z, c);
        uint32_t c2, f3;
        svc_csa(s.ones, r4, r8, c2);
        svc_csa(s.twos, c, c2, f3);
        s.k += __popc(f3) << 2;
    }
}

// four words (one carry-save block) of the SB x 2 pairs a warp holds; MASKED blocks apply the window's tail mask /
// the all-mismatch words behind the window, interior blocks need neither
template <int SB, bool MASKED>
__device__ __forceinline__ void svc_block(SvcPair (&st)[2][SB], const uint32_t* __restrict__ pa0, const uint32_t* __restrict__ pa1,
                                          const uint32_t* __restrict__ pb0, const uint32_t* __restrict__ pb1,
                                          const uint32_t* const (&py)[SB][2], int j0, int nw, uint32_t tail, int min_big,
                                          uint32_t gb) {
#pragma unroll
    for (int ph = 0; ph < 4; ph++) {
        const int j = j0 + ph;
        uint32_t keep = 0xffffffffu;
        if (MASKED) keep = (j < nw - 1) ? 0xffffffffu : ((j == nw - 1) ? tail : 0u);
        uint32_t x0[2], x1[2];
        x0[0] = pa0[ph]; x1[0] = pa1[ph];
        x0[1] = pb0[ph]; x1[1] = pb1[ph];
        if (MASKED) { x1[0] |= ~keep; x1[1] |= ~keep; }
#pragma unroll
        for (int b = 0; b < SB; b++) {
            const uint32_t y0 = py[b][0][ph];
            uint32_t y1 = py[b][1][ph];
            if (MASKED) y1 &= keep;
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const uint32_t nz = (x0[a] ^ y0) | (x1[a] ^ y1);
                if (ph == 0) svc_word<0>(st[a][b], nz, j * 32, min_big, gb);
                else if (ph == 1) svc_word<1>(st[a][b], nz, j * 32, min_big, gb);
                else if (ph == 2) svc_word<2>(st[a][b], nz, j * 32, min_big, gb);
                else svc_word<3>(st[a][b], nz, j * 32, min_big, gb);
            }
        }
    }
}

constexpr int SVC_SB = 2;   // support vectors a warp walks at once (x 2 queries per lane)

// SVC_CCHUNK support vectors per CTA, compiled for SVC_CCTAS CTAs per SM (instantiated as <48, 3>: 80 registers)
template <int SVC_CCHUNK, int SVC_CCTAS>
__global__ void __launch_bounds__(SVC_THREADS, SVC_CCTAS)
svc_kernel_csa(SvcDev m, const SvcWin* __restrict__ wins, int w0, const uint32_t* __restrict__ QP, int64_t qp_words, int64_t N,
               int32_t* __restrict__ Kt, int64_t ldK, int64_t kt_win_stride, int32_t* __restrict__ Krow) {
    extern __shared__ __align__(16) uint32_t sm[];
    const SvcWin w = wins[w0 + blockIdx.z];
    const int s0 = blockIdx.y * SVC_CCHUNK;
    if (s0 >= w.nsv) return;
    const int nw = w.nw;
    const int nblk = (nw + 1 + 3) / 4;   // the window's words plus at least one all-mismatch word behind it
    const int rs = (4 * nblk) | 1;       // row stride in words: whole blocks are addressable, odd for the banks
    uint32_t* qp = sm;                               // [64][2][rs]
    uint32_t* sp = qp + (size_t)SVC_Q * 2 * rs;      // [SVC_CCHUNK][2][rs]
    int32_t* gb = reinterpret_cast<int32_t*>(sp + (size_t)SVC_CCHUNK * 2 * rs);   // [len + 1]
    const uint32_t gb_s = (uint32_t)__cvta_generic_to_shared(gb);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n0 = (int64_t)blockIdx.x * SVC_Q;
    if (Kt) Kt += (int64_t)blockIdx.z * kt_win_stride;
    const int ns = min(SVC_CCHUNK, w.nsv - s0);
    {
        const int64_t wi0 = w.lo >> 5;
        const int sh = (int)(w.lo & 31);
        for (int i = threadIdx.x; i < SVC_Q * 2 * rs; i += SVC_THREADS) {
            const int row = i / rs, j = i - row * rs;
            const int64_t n = n0 + (row >> 1);
            uint32_t v = 0u;
            if (n < N && j < nw) {
                const uint32_t* src = QP + ((size_t)n * 2 + (row & 1)) * qp_words + wi0 + j;
                v = __funnelshift_r(__ldg(src), __ldg(src + 1), sh);
            }
            qp[i] = v;
        }
        const uint32_t* src = w.planes + (size_t)s0 * 2 * nw;
        for (int i = threadIdx.x; i < SVC_CCHUNK * 2 * rs; i += SVC_THREADS) {
            const int row = i / rs, j = i - row * rs;
            sp[i] = (row < ns * 2 && j < nw) ? __ldg(src + (size_t)row * nw + j) : 0u;
        }
        for (int i = threadIdx.x; i <= w.len; i += SVC_THREADS) gb[i] = __ldg(m.gbig + i);
    }
    __syncthreads();
    const uint32_t tail = (w.len & 31) ? ((1u << (w.len & 31)) - 1u) : 0xffffffffu;
    const int min_big = m.min_big;
    const int nfull = (nw - 1) / 4;      // blocks whose four words all lie before the window's last word
    constexpr int SB = SVC_SB;
    for (int sb = warp * (SVC_CCHUNK / 8); sb < (warp + 1) * (SVC_CCHUNK / 8) && sb < ns; sb += SB) {
        SvcPair st[2][SB];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < SB; b++) st[a][b] = SvcPair{0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0, 0};
        const uint32_t* pa0 = qp + (size_t)(lane * 2) * rs;
        const uint32_t* pa1 = pa0 + rs;
        const uint32_t* pb0 = qp + (size_t)((lane + 32) * 2) * rs;
        const uint32_t* pb1 = pb0 + rs;
        const uint32_t* py[SB][2];
#pragma unroll
        for (int b = 0; b < SB; b++) {
            py[b][0] = sp + (size_t)((sb + b) * 2) * rs;   // rows behind the chunk's last vector are zero
            py[b][1] = py[b][0] + rs;
        }
        int blk = 0;
#pragma unroll 1
        for (; blk < nfull; blk++) {
            svc_block<SB, false>(st, pa0, pa1, pb0, pb1, py, blk * 4, nw, tail, min_big, gb_s);
            pa0 += 4; pa1 += 4; pb0 += 4; pb1 += 4;
#pragma unroll
            for (int b = 0; b < SB; b++) { py[b][0] += 4; py[b][1] += 4; }
        }
#pragma unroll 1
        for (; blk < nblk; blk++) {
            svc_block<SB, true>(st, pa0, pa1, pb0, pb1, py, blk * 4, nw, tail, min_big, gb_s);
            pa0 += 4; pa1 += 4; pb0 += 4; pb1 += 4;
#pragma unroll
            for (int b = 0; b < SB; b++) { py[b][0] += 4; py[b][1] += 4; }
        }
#pragma unroll
        for (int b = 0; b < SB; b++)
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const SvcPair& s = st[a][b];
                const int k = s.k + __popc(s.ones) + 2 * __popc(s.twos) + 4 * __popc(s.fours);
                const int64_t n = n0 + lane + 32 * a;
                if (n < N && sb + b < ns) {
                    if (Kt) Kt[(int64_t)(s0 + sb + b) * ldK + n] = k;
                    else Krow[n * w.nsv + s0 + sb + b] = k;
                }
            }
    }
}

// Bit planes of the reflect-padded haplotypes (src/Base/base.py:41-44), once per predict call:
// QP[n][plane][word], bit b of word j = plane bit of padded column 32 j + b; one spare zero word.
__global__ void svc_pack_kernel(SvcDev m, const int8_t* __restrict__ X, int64_t N, int64_t ldX, uint32_t* __restrict__ QP,
                                int64_t qp_words) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t padded = m.C + 2 * m.ctx;
    const int64_t total = N * qp_words;
    for (int64_t t = warp; t < total; t += nwarps) {
        const int64_t n = t / qp_words, j = t - n * qp_words;
        const int64_t p = j * 32 + lane;
        int v = 0;
        if (p < padded) v = X[n * ldX + svc_pad_to_orig(p, m.C, m.ctx)];
        const uint32_t b0 = __ballot_sync(0xffffffffu, v & 1), b1 = __ballot_sync(0xffffffffu, v & 2);
        if (lane == 0) {
            QP[((size_t)n * 2 + 0) * qp_words + j] = b0;
            QP[((size_t)n * 2 + 1) * qp_words + j] = b1;
        }
    }
}

// Generic path (any Ms): the reference's DP verbatim per (query, support vector) pair --
// tri = current match-run length, cov += [tri in Ms], K += cov (string_kernel.py:94-100).
__global__ void svc_kernel_window_generic(SvcDev m, SvcWin w, const int8_t* __restrict__ X, int64_t N, int64_t ldX,
                                          int32_t* __restrict__ Kt, int64_t ldK, int32_t* __restrict__ Krow) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (n >= N) return;
    const uint32_t* y0 = w.planes + (size_t)s * 2 * w.nw;
    const uint32_t* y1 = y0 + w.nw;
    int tri = 0, cov = 0, k = 0;
    for (int p = 0; p < w.len; p++) {
        const int x = X[n * ldX + svc_pad_to_orig(w.lo + p, m.C, m.ctx)] & 3;
        const int y = ((y0[p >> 5] >> (p & 31)) & 1) | (((y1[p >> 5] >> (p & 31)) & 1) << 1);
        if (x == y) {
            tri++;
            cov += m.ohe[tri];
            k += cov;
        } else {
            tri = 0;
            cov = 0;
        }
    }
    if (Kt) Kt[(int64_t)s * ldK + n] = k;
    else Krow[n * w.nsv + s] = k;
}

// ------------------------------------------------------------------------------ K3
__device__ void svc_multiclass_probability(int k, const double* r, double* p) {
    int t, j, iter = 0, max_iter = k > 100 ? k : 100;
    double Q[SVC_MAX_A * SVC_MAX_A], Qp[SVC_MAX_A], pQp, eps = 0.005 / k;
    for (t = 0; t < k; t++) {
        p[t] = 1.0 / k;
        Q[t * k + t] = 0;
        for (j = 0; j < t; j++) {
            Q[t * k + t] += r[j * k + t] * r[j * k + t];
            Q[t * k + j] = Q[j * k + t];
        }
        for (j = t + 1; j < k; j++) {
            Q[t * k + t] += r[j * k + t] * r[j * k + t];
            Q[t * k + j] = -r[j * k + t] * r[t * k + j];
        }
    }
    for (iter = 0; iter < max_iter; iter++) {
        pQp = 0;
        for (t = 0; t < k; t++) {
            Qp[t] = 0;
            for (j = 0; j < k; j++) Qp[t] += Q[t * k + j] * p[j];
            pQp += p[t] * Qp[t];
        }
        double max_error = 0;
        for (t = 0; t < k; t++) {
            double error = fabs(Qp[t] - pQp);
            if (error > max_error) max_error = error;
        }
        if (max_error < eps) break;
        for (t = 0; t < k; t++) {
            double diff = (-Qp[t] + pQp) / Q[t * k + t];
            p[t] += diff;
            pQp = (pQp + diff * (diff * Q[t * k + t] + 2 * Qp[t])) / (1 + diff) / (1 + diff);
            for (j = 0; j < k; j++) {
                Qp[j] = (Qp[j] + diff * Q[t * k + j]) / (1 + diff);
                p[j] /= (1 + diff);
            }
        }
    }
}

// K3a: thread = one (haplotype, class pair): the pairwise decision value in libsvm's order (class i's
// support vectors, then class j's), Platt sigmoid -> r_ij into R[p][n]
__global__ void __launch_bounds__(128)
svc_pair_window(SvcDev m, const SvcWin* __restrict__ wins, int w0, const int32_t* __restrict__ Kt, int64_t ldK, int64_t kt_win_stride,
                int64_t N, double* __restrict__ R) {
    __shared__ int s_start[SVC_MAX_A + 1];
    const SvcWin w = wins[w0 + blockIdx.z];
    Kt += (int64_t)blockIdx.z * kt_win_stride;
    R += (int64_t)blockIdx.z * m.P * ldK;
    if (threadIdx.x == 0) {
        s_start[0] = 0;
        for (int c = 0; c < m.A; c++) s_start[c + 1] = s_start[c] + w.n_support[c];
    }
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (n >= N) return;
    // pair index p -> (i, j), i < j, libsvm order
    int i = 0, rem = p;
    while (rem >= m.A - 1 - i) { rem -= m.A - 1 - i; i++; }
    const int j = i + 1 + rem;
    double sum = 0.0;
    const double* ci = w.coef + (size_t)(j - 1) * w.nsv;
    const double* cj = w.coef + (size_t)i * w.nsv;
    for (int s = s_start[i]; s < s_start[i + 1]; s++) sum += __ldg(ci + s) * (double)Kt[(int64_t)s * ldK + n];
    for (int s = s_start[j]; s < s_start[j + 1]; s++) sum += __ldg(cj + s) * (double)Kt[(int64_t)s * ldK + n];
    const double dec = sum + w.intercept[p];
    const double fApB = dec * w.probA[p] + w.probB[p];
    double v;
    if (fApB >= 0)
        v = gnx_exp(-fApB) / (1.0 + gnx_exp(-fApB));
    else
        v = 1.0 / (1 + gnx_exp(fApB));
    if (v < 1e-7) v = 1e-7;
    if (v > 1 - 1e-7) v = 1 - 1e-7;
    R[(int64_t)p * ldK + n] = v;
}

// K3b: thread = one haplotype: pairwise coupling of the P pair probabilities
__global__ void __launch_bounds__(128)
svc_couple_window(SvcDev m, int w0, const double* __restrict__ R, int64_t ldK, int64_t N, double* __restrict__ B) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int wi = w0 + blockIdx.y;
    R += (int64_t)blockIdx.y * m.P * ldK;
    const int k = m.A;
    double pair[SVC_MAX_A * SVC_MAX_A];
    int p = 0;
    for (int i = 0; i < k; i++)
        for (int j = i + 1; j < k; j++) {
            const double v = R[(int64_t)p * ldK + n];
            pair[i * k + j] = v;
            pair[j * k + i] = 1 - v;
            p++;
        }
    double* out = B + (n * m.W + wi) * k;
    if (k == 2) {
        out[0] = pair[1];
        out[1] = pair[2];
    } else {
        double pr[SVC_MAX_A];
        svc_multiclass_probability(k, pair, pr);
        for (int c = 0; c < k; c++) out[c] = pr[c];
    }
}

}  // namespace gnx

struct gnx_svc {
    gnx::SvcDev d;
    std::vector<gnx::SvcWin> win;
    gnx::SvcWin* d_win;   // the same descriptors on the device (window groups index them by blockIdx.z)
    int use_csa;          // 1: svc_kernel_csa for the reference's Ms list (GNX_SVC_KERNEL=0: first kernel)
    int device;
    int max_nsv;
    void* d_blob;
};

namespace gnx {
void svc_dims(const gnx_svc* m, int64_t* C, int* W, int* A) { *C = m->d.C; *W = m->d.W; *A = m->d.A; }
}  // namespace gnx

using namespace gnx;

extern "C" {

int gnx_svc_model_create(gnx_svc_t** out, int A, int64_t C, int64_t M, int64_t ctx, const int8_t* sv, const int32_t* n_support,
                         const double* dual_coef, const double* intercept, const double* probA, const double* probB,
                         const int32_t* Ms, int n_ms) {
    GNX_REQUIRE(out != nullptr, "gnx_svc_model_create: out is NULL");
    *out = nullptr;
    GNX_REQUIRE(A >= 2 && A <= SVC_MAX_A, "gnx_svc_model_create: A=%d unsupported (2..%d)", A, SVC_MAX_A);
    GNX_REQUIRE(C > 0 && M > 0 && M <= C && ctx >= 0 && ctx <= C, "gnx_svc_model_create: bad geometry");
    GNX_REQUIRE(sv && n_support && dual_coef && intercept && probA && probB && Ms && n_ms > 0, "gnx_svc_model_create: NULL array");
    if (require_blackwell()) return 1;
    const int64_t W = C / M, rem = C - M * W, M_ = M + 2 * ctx;
    const int P = A * (A - 1) / 2;
    const int64_t maxlen = M_ + rem;
    // run-length tables
    std::vector<uint8_t> ohe(maxlen + 2, 0);
    int small_mask = 0, min_big = 0x7fffffff;
    bool fast = true;
    for (int i = 0; i < n_ms; i++) {
        const int mm = Ms[i];
        GNX_REQUIRE(mm >= 1, "gnx_svc_model_create: Ms[%d]=%d", i, mm);
        if (mm <= maxlen) ohe[mm] = 1;
        if (mm <= 32) {
            if ((mm & (mm - 1)) == 0) small_mask |= mm;
            else fast = false;
        } else {
            min_big = std::min(min_big, mm);
        }
    }
    std::vector<int32_t> gbig(maxlen + 1, 0);
    for (int64_t L = 1; L <= maxlen; L++) {
        int64_t g = 0;
        for (int i = 0; i < n_ms; i++)
            if (Ms[i] > 32 && Ms[i] <= L) g += L - Ms[i] + 1;
        gbig[L] = (int32_t)g;
    }
    // per-window packing
    gnx_svc* m = new gnx_svc();
    m->win.resize(W);
    std::vector<uint32_t> planes;
    std::vector<double> coef;
    std::vector<size_t> plane_off(W), coef_off(W);
    int64_t coef_in = 0;
    int max_nsv = 0;
    std::vector<int64_t> sv_offs(W);
    {
        int64_t sv_off = 0;
        size_t pl = 0;
        for (int64_t w = 0; w < W; w++) {
            int nsv = 0;
            for (int c = 0; c < A; c++) {
                if (n_support[w * A + c] < 0) {
                    delete m;
                    set_error("gnx_svc_model_create: negative n_support");
                    return 2;
                }
                nsv += n_support[w * A + c];
            }
            const int len = (int)((w == W - 1) ? (M_ + rem) : M_);
            const int nw = (len + 31) / 32;
            SvcWin& sw = m->win[w];
            sw.lo = (w == W - 1) ? (C + 2 * ctx - (M_ + rem)) : w * M;
            sw.len = len; sw.nw = nw; sw.nwp = nw | 1; sw.nsv = nsv;
            plane_off[w] = pl;
            pl += (size_t)nsv * 2 * nw;
            sv_offs[w] = sv_off;
            sv_off += (int64_t)nsv * len;
            coef_off[w] = coef.size();
            coef.insert(coef.end(), dual_coef + coef_in, dual_coef + coef_in + (int64_t)(A - 1) * nsv);
            coef_in += (int64_t)(A - 1) * nsv;
            max_nsv = std::max(max_nsv, nsv);
        }
        planes.assign(pl, 0u);
    }
    // bit planes of the support vectors, windows in parallel on the host worker pool (1.7e9 values for chr1 with 700
    // vectors per window: 8.7 s as a scalar loop on one thread), each row through the vector pack of host_pack.cpp
    std::atomic<long long> bad_window{-1};
    parallel_for(W, 0, [&](int64_t w) {
        const SvcWin& sw = m->win[w];
        const int len = sw.len, nw = sw.nw, groups = (len + 63) / 64;
        std::vector<uint64_t> tmp((size_t)2 * groups);
        uint32_t* pw = planes.data() + plane_off[w];
        for (int s = 0; s < sw.nsv; s++) {
            if (pack_row_best(sv + sv_offs[w] + (int64_t)s * len, len, tmp.data(), groups)) {
                bad_window.store((long long)w);
                return;
            }
            uint32_t* p0 = pw + ((size_t)s * 2 + 0) * nw;
            uint32_t* p1 = pw + ((size_t)s * 2 + 1) * nw;
            for (int j = 0; j < nw; j++) {
                p0[j] = (uint32_t)(tmp[2 * (j >> 1)] >> (32 * (j & 1)));
                p1[j] = (uint32_t)(tmp[2 * (j >> 1) + 1] >> (32 * (j & 1)));
            }
        }
    });
    if (bad_window.load() >= 0) {
        const long long bw = bad_window.load();
        delete m;
        set_error("gnx_svc_model_create: a support vector value outside 0..3 (window %lld)", bw);
        return 2;
    }
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t o_pl = 0, o_cf = al(planes.size() * 4), o_ns = o_cf + al(coef.size() * 8), o_ic = o_ns + al((size_t)W * A * 4),
                 o_pa = o_ic + al((size_t)W * P * 8), o_pb = o_pa + al((size_t)W * P * 8), o_gb = o_pb + al((size_t)W * P * 8),
                 o_oh = o_gb + al(gbig.size() * 4), total = o_oh + al(ohe.size());
    char* blob = nullptr;
    cudaError_t e = cudaMalloc((void**)&blob, total);
    if (e != cudaSuccess) {
        delete m;
        set_error("gnx_svc_model_create: cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e));
        return 1;
    }
    bool ok = true;
    ok &= cudaMemcpy(blob + o_pl, planes.data(), planes.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_cf, coef.data(), coef.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_ns, n_support, (size_t)W * A * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_ic, intercept, (size_t)W * P * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_pa, probA, (size_t)W * P * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_pb, probB, (size_t)W * P * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_gb, gbig.data(), gbig.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMemcpy(blob + o_oh, ohe.data(), ohe.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        cudaFree(blob);
        delete m;
        set_error("gnx_svc_model_create: H2D copy failed");
        return 1;
    }
    for (int64_t w = 0; w < W; w++) {
        SvcWin& sw = m->win[w];
        sw.planes = reinterpret_cast<const uint32_t*>(blob + o_pl) + plane_off[w];
        sw.coef = reinterpret_cast<const double*>(blob + o_cf) + coef_off[w];
        sw.n_support = reinterpret_cast<const int32_t*>(blob + o_ns) + w * A;
        sw.intercept = reinterpret_cast<const double*>(blob + o_ic) + w * P;
        sw.probA = reinterpret_cast<const double*>(blob + o_pa) + w * P;
        sw.probB = reinterpret_cast<const double*>(blob + o_pb) + w * P;
    }
    m->d_win = nullptr;
    if (cudaMalloc((void**)&m->d_win, sizeof(SvcWin) * (size_t)W) != cudaSuccess ||
        cudaMemcpy(m->d_win, m->win.data(), sizeof(SvcWin) * (size_t)W, cudaMemcpyHostToDevice) != cudaSuccess) {
        if (m->d_win) cudaFree(m->d_win);
        cudaFree(blob);
        delete m;
        set_error("gnx_svc_model_create: window table copy failed");
        return 1;
    }
    {
        const char* e2 = getenv("GNX_SVC_KERNEL");
        m->use_csa = !(e2 && e2[0] == '0');
    }
    cudaGetDevice(&m->device);
    m->d_blob = blob;
    m->max_nsv = max_nsv;
    m->d = SvcDev{A, P, (int)W, fast ? 1 : 0, small_mask, min_big, C, M, ctx, reinterpret_cast<const int32_t*>(blob + o_gb),
                  reinterpret_cast<const uint8_t*>(blob + o_oh)};
    *out = m;
    return 0;
}

void gnx_svc_model_destroy(gnx_svc_t* m) {
    if (!m) return;
    if (m->d_blob) cudaFree(m->d_blob);
    if (m->d_win) cudaFree(m->d_win);
    delete m;
}

static int64_t svc_qp_words(const gnx_svc_t* m) { return (m->d.C + 2 * m->d.ctx + 31) / 32 + 1; }

static int svc_pack(const gnx_svc_t* m, const int8_t* X, int64_t N, int64_t ldX, uint32_t* QP, cudaStream_t st) {
    const int64_t total_warps = N * svc_qp_words(m);
    const int grid = (int)std::min<int64_t>(ceil_div(total_warps, 8), (int64_t)sm_count() * 16);
    svc_pack_kernel<<<grid, 256, 0, st>>>(m->d, X, N, ldX, QP, svc_qp_words(m));
    GNX_CUDA(cudaGetLastError());
    return 0;
}

static bool svc_uses_csa(const gnx_svc_t* m) { return m->d.fast && m->d.small_mask == 13 && m->use_csa; }

// windows [w0, w0 + g) in one launch of the production kernel
static int svc_launch_csa(const gnx_svc_t* m, int w0, int g, const uint32_t* QP, int64_t N, int32_t* Kt, int64_t ldK, int64_t kt_win,
                          int32_t* Krow, cudaStream_t st) {
    int nsv_max = 0, nw_max = 0, len_max = 0;
    for (int w = w0; w < w0 + g; w++) {
        nsv_max = std::max(nsv_max, m->win[w].nsv);
        nw_max = std::max(nw_max, m->win[w].nw);
        len_max = std::max(len_max, m->win[w].len);
    }
    if (nsv_max == 0) return 0;
    const size_t rs_max = (size_t)((4 * ((nw_max + 1 + 3) / 4)) | 1);
    // <48, 3>: 80 registers, 3 CTAs per SM.  <32, 4> (64 registers, 4 CTAs) measured 13 % slower: the kernel is bound by
    // the integer pipe, not by latency, and the smaller chunk stages the 64 queries once per 32 instead of 48 vectors
    constexpr int chunk = 48;
    const size_t smem = ((size_t)SVC_Q * 2 * rs_max + (size_t)chunk * 2 * rs_max + (size_t)len_max + 2) * 4;
    GNX_REQUIRE(smem <= 227 * 1024, "gnx_svc: window of %d SNPs too long for shared memory", len_max);
    dim3 grid((unsigned)ceil_div(N, SVC_Q), (unsigned)ceil_div(nsv_max, chunk), (unsigned)g);
    GNX_CUDA(cudaFuncSetAttribute(svc_kernel_csa<chunk, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    svc_kernel_csa<chunk, 3><<<grid, SVC_THREADS, smem, st>>>(m->d, m->d_win, w0, QP, svc_qp_words(m), N, Kt, ldK, kt_win, Krow);
    GNX_CUDA(cudaGetLastError());
    return 0;
}

static int svc_launch_kernel(const gnx_svc_t* m, int w, const int8_t* X, const uint32_t* QP, int64_t N, int64_t ldX, int32_t* Kt, int64_t ldK,
                             int32_t* Krow, cudaStream_t st) {
    const SvcWin& sw = m->win[w];
    if (sw.nsv == 0) return 0;
    if (svc_uses_csa(m)) return svc_launch_csa(m, w, 1, QP, N, Kt, ldK, 0, Krow, st);
    if (m->d.fast) {
        const size_t smem = ((size_t)SVC_Q * 2 * sw.nwp + (size_t)SVC_CHUNK * 2 * sw.nw) * 4;
        GNX_REQUIRE(smem <= 227 * 1024, "gnx_svc: window of %d SNPs too long for shared memory", sw.len);
        GNX_CUDA(cudaFuncSetAttribute(svc_kernel_window<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)ceil_div(N, SVC_Q), (unsigned)ceil_div(sw.nsv, SVC_CHUNK));
        if (m->d.small_mask == 13) {  // lengths {1, 4, 8} below 33: every CovSample(., 0.6, 1.0, 37) list
            GNX_CUDA(cudaFuncSetAttribute(svc_kernel_window<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            svc_kernel_window<13><<<grid, SVC_THREADS, smem, st>>>(m->d, sw, QP, svc_qp_words(m), N, Kt, ldK, Krow);
        } else {
            svc_kernel_window<-1><<<grid, SVC_THREADS, smem, st>>>(m->d, sw, QP, svc_qp_words(m), N, Kt, ldK, Krow);
        }
    } else {
        dim3 grid((unsigned)ceil_div(N, 128), (unsigned)sw.nsv);
        svc_kernel_window_generic<<<grid, 128, 0, st>>>(m->d, sw, X, N, ldX, Kt, ldK, Krow);
    }
    GNX_CUDA(cudaGetLastError());
    return 0;
}

int gnx_svc_kernel_window(const gnx_svc_t* m, int w, const int8_t* X_dev, int64_t N, int64_t ldX, int32_t* K_dev, void* stream) {
    GNX_REQUIRE(m != nullptr, "gnx_svc_kernel_window: NULL model");
    GNX_REQUIRE(w >= 0 && w < m->d.W, "gnx_svc_kernel_window: window %d outside 0..%d", w, m->d.W - 1);
    GNX_REQUIRE(N >= 0 && ldX >= m->d.C, "gnx_svc_kernel_window: bad shape");
    if (N == 0) return 0;
    GNX_REQUIRE(X_dev && K_dev, "gnx_svc_kernel_window: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* QP = nullptr;
    if (m->d.fast) {
        GNX_CUDA(cudaMallocAsync((void**)&QP, sizeof(uint32_t) * (size_t)N * 2 * svc_qp_words(m), st));
        if (svc_pack(m, X_dev, N, ldX, QP, st)) return 1;
    }
    const int rc = svc_launch_kernel(m, w, X_dev, QP, N, ldX, nullptr, 0, K_dev, st);
    if (QP) cudaFreeAsync(QP, st);
    return rc;
}

int gnx_svc_predict(const gnx_svc_t* m, const int8_t* X_dev, int64_t N, int64_t ldX, double* B_dev, void* stream) {
    GNX_REQUIRE(m != nullptr, "gnx_svc_predict: NULL model");
    GNX_REQUIRE(N >= 0 && ldX >= m->d.C, "gnx_svc_predict: bad shape N=%lld ldX=%lld", (long long)N, (long long)ldX);
    if (N == 0) return 0;
    GNX_REQUIRE(X_dev && B_dev, "gnx_svc_predict: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ldK = (N + 31) & ~int64_t(31);
    const int W = m->d.W;
    const bool csa = svc_uses_csa(m);
    // windows per launch: enough CTAs to make the tail wave irrelevant, bounded by ~1 GB of kernel-value scratch
    const int64_t kt_win = (int64_t)std::max(1, m->max_nsv) * ldK;   // int32 elements per window
    int G = 1;
    if (csa) G = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(16, W), (int64_t(1) << 28) / kt_win));
    const size_t kt_bytes = (sizeof(int32_t) * (size_t)kt_win * G + 255) & ~size_t(255);
    const size_t r_bytes = (sizeof(double) * (size_t)m->d.P * ldK * G + 255) & ~size_t(255);
    const size_t qp_bytes = m->d.fast ? sizeof(uint32_t) * (size_t)N * 2 * svc_qp_words(m) : 0;
    char* scratch = nullptr;
    GNX_CUDA(cudaMallocAsync((void**)&scratch, kt_bytes + r_bytes + qp_bytes + 256, st));
    int32_t* Kt = reinterpret_cast<int32_t*>(scratch);
    double* R = reinterpret_cast<double*>(scratch + kt_bytes);
    uint32_t* QP = m->d.fast ? reinterpret_cast<uint32_t*>(scratch + kt_bytes + r_bytes) : nullptr;
    int rc = 0;
    if (QP) rc = svc_pack(m, X_dev, N, ldX, QP, st);
    for (int w0 = 0; w0 < W && rc == 0; w0 += G) {
        const int g = std::min(G, W - w0);
        if (csa) {
            rc = svc_launch_csa(m, w0, g, QP, N, Kt, ldK, kt_win, nullptr, st);
            if (rc) break;
        } else {
            rc = svc_launch_kernel(m, w0, X_dev, QP, N, ldX, Kt, ldK, nullptr, st);
            if (rc) break;
        }
        dim3 gp((unsigned)ceil_div(N, 128), (unsigned)m->d.P, (unsigned)g);
        svc_pair_window<<<gp, 128, 0, st>>>(m->d, m->d_win, w0, Kt, ldK, kt_win, N, R);
        dim3 gc((unsigned)ceil_div(N, 128), (unsigned)g);
        svc_couple_window<<<gc, 128, 0, st>>>(m->d, w0, R, ldK, N, B_dev);
        if (cudaGetLastError() != cudaSuccess) {
            set_error("gnx_svc_predict: launch failed (window %d)", w0);
            rc = 1;
        }
    }
    cudaFreeAsync(scratch, st);
    return rc;
}

}  // extern "C"
