// host_pack.h -- 2-bit-plane packing of the int8 haplotype matrix on the host cores
// (host_pack.cpp) and its device-side inverse (pack.cu); used by gnx_infer_host.
#pragma once

#include <stdint.h>

#include <functional>

namespace gnx {

// Packs rows [0, n) of X (int8 [n, ldX], C valid columns) into out (row pitch
// out_pitch_words 64-bit words, >= 2 * ceil(C / 64)) with `threads` host threads
// (<= 0: all cores this process may run on, or GNX_HOST_THREADS).  Returns 1 if any
// value was outside 0..3 (the packed form is then lossy and must not be used).
// *isa receives 0 scalar / 1 AVX2 / 2 AVX-512BW.
int pack_rows(const int8_t* X, int64_t n, int64_t ldX, int64_t C, uint64_t* out, int64_t out_pitch_words, int threads,
              int* isa);
// One row with the best instruction set of this CPU: x[0..C) -> out[0..2*groups) (groups >= ceil(C/64), the rest
// zero); returns non-zero when a value was outside 0..3.
unsigned pack_row_best(const int8_t* x, int64_t C, uint64_t* out, int64_t groups);
int host_threads_default();
// Runs fn(0) .. fn(items - 1) on the library's persistent host worker pool (`threads` <= 0: default count);
// returns when all are done.  One job at a time.
void parallel_for(int64_t items, int threads, const std::function<void(int64_t)>& fn);

}  // namespace gnx
