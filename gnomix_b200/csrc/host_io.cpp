// host_io.cpp -- output side of run_inference on the host cores: the body of the .fb file
// (reference src/postprocess.py:100-126, write_fb).  One line per window holding the
// probabilities of every haplotype and ancestry as text; the reference formats them with
// numpy's str (shortest round-trip digits; positional for 1e-4 <= |x| < 1e16, scientific
// otherwise), one Python object per number.  Here lines are formatted in parallel with
// std::to_chars (the same shortest-digit strings) and written in order.
#include <stdio.h>
#include <string.h>

#include <charconv>
#include <cmath>
#include <string>
#include <vector>

#include "host_pack.h"

namespace gnx {

void set_error(const char* fmt, ...);

// numpy's str() of a float32 / float64 scalar (numpy/_core/src/multiarray/scalartypes.c.src,
// *type_str_either -> Dragon4 unique mode): returns the number of chars written (<= 32).
template <typename T>
static inline int fmt_np(char* out, T v) {
    if (std::isnan(v)) {
        memcpy(out, "nan", 3);
        return 3;
    }
    if (std::isinf(v)) {
        if (v < 0) { memcpy(out, "-inf", 4); return 4; }
        memcpy(out, "inf", 3);
        return 3;
    }
    const T a = v < 0 ? -v : v;
    // positional range of numpy's str: [1e-4, 1e16) for float64, [1e-4, 1e6) for float32 (numpy 2.x)
    const long double upper = sizeof(T) == 4 ? 1.e6L : 1.e16L;
    if (a == 0 || ((long double)a < upper && (long double)a >= 1.e-4L)) {
        char* e = std::to_chars(out, out + 32, v, std::chars_format::fixed).ptr;
        bool dot = false;
        for (char* p = out; p < e; p++) dot |= (*p == '.');
        if (!dot) { *e++ = '.'; *e++ = '0'; }
        return (int)(e - out);
    }
    return (int)(std::to_chars(out, out + 32, v, std::chars_format::scientific).ptr - out);
}

template <typename T>
static int write_fb_body(FILE* f, const T* proba, int64_t N, int64_t W, int64_t A, const char* const* prefixes, int threads) {
    if (threads <= 0) threads = host_threads_default();
    const int64_t batch = std::min<int64_t>(W, threads);
    size_t max_prefix = 0;
    for (int64_t l = 0; l < W; l++) max_prefix = std::max(max_prefix, strlen(prefixes[l]));
    const size_t cap = max_prefix + (size_t)N * A * 33 + 2;
    std::vector<std::vector<char>> bufs(batch);
    std::vector<size_t> lens(batch);
    for (auto& b : bufs) b.resize(cap);
    for (int64_t l0 = 0; l0 < W; l0 += batch) {
        const int64_t nb = std::min(batch, W - l0);
        parallel_for(nb, threads, [&](int64_t k) {
            const int64_t l = l0 + k;
            char* p = bufs[k].data();
            const size_t pl = strlen(prefixes[l]);
            memcpy(p, prefixes[l], pl);
            p += pl;
            for (int64_t n = 0; n < N; n++) {
                const T* src = proba + (n * W + l) * A;
                for (int64_t a = 0; a < A; a++) {
                    p += fmt_np<T>(p, src[a]);
                    *p++ = '\t';
                }
            }
            if (N * A > 0) p--;  // the last tab becomes the newline
            *p++ = '\n';
            lens[k] = (size_t)(p - bufs[k].data());
        });
        for (int64_t k = 0; k < nb; k++)
            if (fwrite(bufs[k].data(), 1, lens[k], f) != lens[k]) return 1;
    }
    return 0;
}

}  // namespace gnx

// Appends (or writes) the W body lines of a .fb file: line l = prefixes[l] + tab-separated
// str(proba[n, l, a]) for n in 0..N-1, a in 0..A-1 + newline.  proba is row-major [N, W, A],
// float32 or float64 (is_f64).
extern "C" int gnx_write_fb_body(const char* path, int append, const void* proba, int is_f64, int64_t N, int64_t W, int64_t A,
                                 const char* const* prefixes, int threads) {
    if (!path || !prefixes || N < 0 || W < 0 || A < 0 || (!proba && N * W * A > 0)) {
        gnx::set_error("gnx_write_fb_body: bad arguments");
        return 2;
    }
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) {
        gnx::set_error("gnx_write_fb_body: cannot open %s", path);
        return 1;
    }
    static thread_local std::vector<char> iobuf;
    iobuf.resize(size_t(8) << 20);
    setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
    const int rc = is_f64 ? gnx::write_fb_body<double>(f, static_cast<const double*>(proba), N, W, A, prefixes, threads)
                          : gnx::write_fb_body<float>(f, static_cast<const float*>(proba), N, W, A, prefixes, threads);
    const int rc2 = fclose(f);
    if (rc || rc2) {
        gnx::set_error("gnx_write_fb_body: write to %s failed", path);
        return 1;
    }
    return 0;
}

// str() of n float32 / float64 values as numpy prints them, '\n'-separated, into out (cap bytes);
// returns the length written or -1 if out is too small (test hook for the formatter).
extern "C" int64_t gnx_format_floats(const void* v, int is_f64, int64_t n, char* out, int64_t cap) {
    int64_t pos = 0;
    for (int64_t i = 0; i < n; i++) {
        if (pos + 34 > cap) return -1;
        pos += is_f64 ? gnx::fmt_np<double>(out + pos, static_cast<const double*>(v)[i])
                      : gnx::fmt_np<float>(out + pos, static_cast<const float*>(v)[i]);
        out[pos++] = '\n';
    }
    return pos;
}

namespace gnx {

// starts of the '\n'-separated fields of a blob holding `count` fields
static bool split_lines(const char* blob, int64_t len, int64_t count, std::vector<int64_t>& start) {
    start.resize((size_t)count + 1);
    int64_t p = 0;
    for (int64_t i = 0; i < count; i++) {
        start[i] = p;
        const char* nl = p < len ? static_cast<const char*>(memchr(blob + p, '\n', (size_t)(len - p))) : nullptr;
        if (!nl) return false;
        p = (nl - blob) + 1;
    }
    start[count] = p;
    return true;
}

}  // namespace gnx

// Appends (or writes) the records of the phased VCF that npy_to_vcf produces (reference src/utils.py:247-329):
// record j = CHROM POS ID REF ALT QUAL PASS . GT, then for every sample i "hap(2i)|hap(2i+1)" with one character
// '0' + value per haplotype.  The five string columns arrive as blobs of newline-terminated fields (n_rec fields
// each), POS as int64, the haplotypes as int8 [n_hap][ld] (column j = record j).
extern "C" int gnx_write_vcf_body(const char* path, int append, int64_t n_rec, int64_t n_hap, const int8_t* hap, int64_t ld,
                                  const int64_t* pos, const char* chrom, int64_t chrom_len, const char* id, int64_t id_len,
                                  const char* ref, int64_t ref_len, const char* alt, int64_t alt_len, const char* qual,
                                  int64_t qual_len, int threads) {
    if (!path || n_rec < 0 || n_hap < 0 || (n_hap & 1) || ld < n_rec || (n_rec > 0 && (!pos || !chrom || !id || !ref || !alt || !qual)) ||
        (n_rec * n_hap > 0 && !hap)) {
        gnx::set_error("gnx_write_vcf_body: bad arguments");
        return 2;
    }
    std::vector<int64_t> s_chrom, s_id, s_ref, s_alt, s_qual;
    if (!gnx::split_lines(chrom, chrom_len, n_rec, s_chrom) || !gnx::split_lines(id, id_len, n_rec, s_id) ||
        !gnx::split_lines(ref, ref_len, n_rec, s_ref) || !gnx::split_lines(alt, alt_len, n_rec, s_alt) ||
        !gnx::split_lines(qual, qual_len, n_rec, s_qual)) {
        gnx::set_error("gnx_write_vcf_body: a string column has fewer than %lld fields", (long long)n_rec);
        return 2;
    }
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) {
        gnx::set_error("gnx_write_vcf_body: cannot open %s", path);
        return 1;
    }
    if (threads <= 0) threads = gnx::host_threads_default();
    const int64_t RB = 256;                                   // records per task: 256 contiguous bytes of every haplotype row
    const int64_t n_samp = n_hap / 2;
    const int64_t nblk = (n_rec + RB - 1) / RB;
    const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(nblk, threads));
    std::vector<std::vector<char>> bufs((size_t)batch);
    int rc = 0;
    for (int64_t b0 = 0; b0 < nblk && !rc; b0 += batch) {
        const int64_t nb = std::min(batch, nblk - b0);
        gnx::parallel_for(nb, threads, [&](int64_t k) {
            const int64_t r0 = (b0 + k) * RB, rn = std::min(RB, n_rec - r0);
            size_t fixed = 0;
            for (int64_t r = r0; r < r0 + rn; r++)
                fixed += (size_t)((s_chrom[r + 1] - s_chrom[r]) + (s_id[r + 1] - s_id[r]) + (s_ref[r + 1] - s_ref[r]) +
                                  (s_alt[r + 1] - s_alt[r]) + (s_qual[r + 1] - s_qual[r])) + 40;
            std::vector<char>& buf = bufs[(size_t)k];
            buf.resize(fixed + (size_t)rn * (size_t)(4 * n_samp + 1));
            // genotype text of the block, record-major: gen[r][4 * i .. 4 * i + 3] = '\t' a '|' b
            std::vector<char> gen((size_t)rn * (size_t)(4 * n_samp));
            for (int64_t i = 0; i < n_samp; i++) {
                const int8_t* ha = hap + (2 * i) * ld + r0;
                const int8_t* hb = hap + (2 * i + 1) * ld + r0;
                for (int64_t r = 0; r < rn; r++) {
                    char* g = gen.data() + (size_t)r * (size_t)(4 * n_samp) + 4 * i;
                    g[0] = '\t';
                    g[1] = (char)('0' + ha[r]);
                    g[2] = '|';
                    g[3] = (char)('0' + hb[r]);
                }
            }
            char* p = buf.data();
            auto put = [&](const char* blob, const std::vector<int64_t>& st, int64_t r) {
                const size_t l = (size_t)(st[r + 1] - st[r] - 1);
                memcpy(p, blob + st[r], l);
                p += l;
                *p++ = '\t';
            };
            for (int64_t r = r0; r < r0 + rn; r++) {
                put(chrom, s_chrom, r);
                p = std::to_chars(p, p + 24, (long long)pos[r]).ptr;
                *p++ = '\t';
                put(id, s_id, r);
                put(ref, s_ref, r);
                put(alt, s_alt, r);
                put(qual, s_qual, r);
                memcpy(p, "PASS\t.\tGT", 9);
                p += 9;
                memcpy(p, gen.data() + (size_t)(r - r0) * (size_t)(4 * n_samp), (size_t)(4 * n_samp));
                p += 4 * n_samp;
                *p++ = '\n';
            }
            buf.resize((size_t)(p - buf.data()));
        });
        for (int64_t k = 0; k < nb; k++)
            if (fwrite(bufs[(size_t)k].data(), 1, bufs[(size_t)k].size(), f) != bufs[(size_t)k].size()) rc = 1;
    }
    if (fclose(f) || rc) {
        gnx::set_error("gnx_write_vcf_body: write to %s failed", path);
        return 1;
    }
    return 0;
}
