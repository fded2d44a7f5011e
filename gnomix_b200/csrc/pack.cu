// pack.cu -- device inverse of host_pack.cpp: two 64-bit planes per 64 SNPs -> the int8
// matrix [n, ldX] K1's TMA descriptor addresses.  Pure streaming (reads 1/4 byte, writes
// 1 byte per SNP); one thread restores 16 SNPs and stores them as one 16-byte vector.
#include <algorithm>

#include "common.cuh"
#include "host_pack.h"

namespace gnx {

// nibble n -> four bytes holding bit 0..3 of n (shifted copies 0,7,14,21 never overlap)
__device__ __forceinline__ uint32_t spread4(uint32_t n) { return (n * 0x00204081u) & 0x01010101u; }

__global__ void __launch_bounds__(256)
unpack_kernel(const uint64_t* __restrict__ packed, int64_t n, int64_t pitch_words, int64_t vecs_per_row,
              int8_t* __restrict__ X, int64_t ldX) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // 16-SNP vector within the row
    if (v >= vecs_per_row) return;
    const int64_t g = v >> 2;
    const int sh = (int)(v & 3) * 16;
    for (int64_t r = blockIdx.y; r < n; r += gridDim.y) {
        const uint64_t* row = packed + r * pitch_words;
        const uint32_t p0 = (uint32_t)(__ldg(row + 2 * g) >> sh) & 0xffffu;
        const uint32_t p1 = (uint32_t)(__ldg(row + 2 * g + 1) >> sh) & 0xffffu;
        uint4 o;
        o.x = spread4(p0 & 15u) | (spread4(p1 & 15u) << 1);
        o.y = spread4((p0 >> 4) & 15u) | (spread4((p1 >> 4) & 15u) << 1);
        o.z = spread4((p0 >> 8) & 15u) | (spread4((p1 >> 8) & 15u) << 1);
        o.w = spread4(p0 >> 12) | (spread4(p1 >> 12) << 1);
        *reinterpret_cast<uint4*>(X + r * ldX + v * 16) = o;
    }
}

int unpack_rows(const uint64_t* packed_dev, int64_t n, int64_t pitch_words, int64_t C, int8_t* X_dev, int64_t ldX,
                cudaStream_t st) {
    if (n == 0) return 0;
    // whole 16-byte vectors that fit the row: columns [0, min(ldX, 64 * groups)) rounded down to 16
    const int64_t cols = std::min<int64_t>(ldX, 32 * pitch_words);
    const int64_t vecs = cols / 16;
    GNX_REQUIRE(vecs * 16 >= C, "gnx_unpack: ldX=%lld / pitch too small for C=%lld", (long long)ldX, (long long)C);
    dim3 grid((unsigned)ceil_div(vecs, 256), (unsigned)std::min<int64_t>(n, 32768));
    unpack_kernel<<<grid, 256, 0, st>>>(packed_dev, n, pitch_words, vecs, X_dev, ldX);
    GNX_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gnx

extern "C" int gnx_unpack_dev(const uint64_t* packed_dev, int64_t n, int64_t pitch_words, int64_t C, int8_t* X_dev,
                              int64_t ldX, void* stream) {
    GNX_REQUIRE(n >= 0 && C >= 0, "gnx_unpack_dev: bad shape");
    GNX_REQUIRE(pitch_words >= 2 * gnx::ceil_div(C, 64), "gnx_unpack_dev: pitch_words=%lld too small for C=%lld",
                (long long)pitch_words, (long long)C);
    GNX_REQUIRE(ldX % 16 == 0 && ((uintptr_t)X_dev & 15) == 0, "gnx_unpack_dev: X_dev / ldX must be 16-byte aligned");
    if (n == 0) return 0;
    GNX_REQUIRE(packed_dev && X_dev, "gnx_unpack_dev: NULL buffer");
    if (gnx::require_blackwell()) return 1;
    return gnx::unpack_rows(packed_dev, n, pitch_words, C, X_dev, ldX, (cudaStream_t)stream);
}
