// host_vcf.cpp -- input side of run_inference on the host cores: VCF(.gz) -> genotype calls.
//
// Replaces the scikit-allel call behind read_vcf (reference src/utils.py:55-81; the fields the
// reference reads: calldata/GT, variants/POS|REF|ALT|CHROM|ID|QUAL, samples).  The reference's
// own notebook names this step as the largest part of an inference run (demo.ipynb:236).  The
// file is read whole (plain text directly; BGZF members inflated in parallel; any other gzip file
// through one zlib stream), the line index is built once, and records are parsed in parallel on the library's worker pool straight into the
// int8 [records, samples, 2] genotype block.  Semantics follow gnomix_b200/io.py (which follows
// allel): '##' lines skipped, '#CHROM' names the samples, records with fewer than 10 columns
// skipped, optional CHROM filter, GT = text before the first ':' of a sample column, alleles
// split on '|' or '/', '.' or empty = -1, haploid calls leave the second allele -1, ALT padded
// to three entries, QUAL '.' = NaN.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/gnx.h"
#include "host_pack.h"

namespace gnx {
void set_error(const char* fmt, ...);
}

struct gnx_vcf {
    char* text = nullptr;           // the whole (inflated) file, newline-terminated; malloc'ed (no zero fill)
    size_t text_len = 0, text_cap = 0;
    std::vector<std::string> samples;
    int64_t n_rec = 0, n_smp = 0;
    int threads = 0;
    std::vector<int64_t> ls;        // line starts of candidate records
    std::vector<uint8_t> keep;      // per candidate line: a record we keep
    std::vector<int64_t> blk_count; // kept records before each block of lines
    bool parsed = false;            // pass B done (string spans known)
    // string columns as (offset, length) into text: CHROM, ID, REF, ALT (whole column)
    std::vector<int64_t> off[4];
    std::vector<int32_t> len[4];
    ~gnx_vcf() { free(text); }
};

namespace {

inline const char* find_tab(const char* p, const char* e) {
    const char* t = static_cast<const char*>(memchr(p, '\t', (size_t)(e - p)));
    return t ? t : e;
}

// one allele token [p, q): '.' or empty -> -1, digits -> value (saturated to 127), anything else -> -1
inline int8_t allele(const char* p, const char* q) {
    if (p == q || *p == '.') return -1;
    int v = 0;
    for (; p < q; p++) {
        if (*p < '0' || *p > '9') return -1;
        v = v * 10 + (*p - '0');
        if (v > 127) v = 127;
    }
    return (int8_t)v;
}

// sample column [p, e) -> two alleles
inline void parse_gt(const char* p, const char* e, int8_t* out) {
    const char* c = static_cast<const char*>(memchr(p, ':', (size_t)(e - p)));
    if (c) e = c;
    // fast path: a|b with single characters
    if (e - p == 3 && (p[1] == '|' || p[1] == '/')) {
        out[0] = (p[0] >= '0' && p[0] <= '9') ? (int8_t)(p[0] - '0') : (int8_t)-1;
        out[1] = (p[2] >= '0' && p[2] <= '9') ? (int8_t)(p[2] - '0') : (int8_t)-1;
        return;
    }
    const char* s = p;
    while (s < e && *s != '|' && *s != '/') s++;
    out[0] = allele(p, s);
    if (s >= e) {  // haploid call
        out[1] = -1;
        return;
    }
    const char* s2 = s + 1;
    const char* t = s2;
    while (t < e && *t != '|' && *t != '/') t++;
    out[1] = allele(s2, t);
}

// BGZF (bgzip / tabix files): a series of gzip members of at most 64 KB each, every header carrying an extra
// field "BC" with the member's compressed size, every trailer its inflated size -- so the members can be
// found by hopping over the headers and inflated independently, in parallel.  Returns false if the file is
// not BGZF (the caller falls back to a single gzread stream).
struct BgzfBlock {
    size_t src, csize;   // deflate payload inside the file image
    size_t dst, isize;
};

bool bgzf_index(const unsigned char* f, size_t n, std::vector<BgzfBlock>& blocks, size_t& total) {
    size_t p = 0;
    total = 0;
    blocks.clear();
    while (p < n) {
        if (n - p < 28 || f[p] != 0x1f || f[p + 1] != 0x8b || f[p + 2] != 8 || !(f[p + 3] & 4)) return false;
        const size_t xlen = f[p + 10] | ((size_t)f[p + 11] << 8);
        if (n - p < 12 + xlen + 8) return false;
        size_t bsize = 0, x = p + 12;
        const size_t xe = p + 12 + xlen;
        while (x + 4 <= xe) {
            const size_t slen = f[x + 2] | ((size_t)f[x + 3] << 8);
            if (f[x] == 'B' && f[x + 1] == 'C' && slen == 2 && x + 6 <= xe) bsize = (f[x + 4] | ((size_t)f[x + 5] << 8)) + 1;
            x += 4 + slen;
        }
        if (bsize == 0 || bsize < 12 + xlen + 8 || p + bsize > n) return false;
        const size_t isize = f[p + bsize - 4] | ((size_t)f[p + bsize - 3] << 8) | ((size_t)f[p + bsize - 2] << 16) | ((size_t)f[p + bsize - 1] << 24);
        if (isize > 65536) return false;
        blocks.push_back(BgzfBlock{p + 12 + xlen, bsize - (12 + xlen) - 8, total, isize});
        total += isize;
        p += bsize;
    }
    return !blocks.empty();
}

bool bgzf_inflate(const unsigned char* f, const std::vector<BgzfBlock>& blocks, char* out, int threads) {
    std::atomic<int> bad{0};
    const int64_t per = 64;  // members per task
    const int64_t nb = (int64_t)blocks.size();
    gnx::parallel_for((nb + per - 1) / per, threads, [&](int64_t t) {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) {
            bad.store(1);
            return;
        }
        for (int64_t b = t * per; b < std::min(nb, (t + 1) * per); b++) {
            const BgzfBlock& k = blocks[(size_t)b];
            if (k.isize == 0) continue;
            inflateReset(&zs);
            zs.next_in = const_cast<unsigned char*>(f + k.src);
            zs.avail_in = (uInt)k.csize;
            zs.next_out = reinterpret_cast<unsigned char*>(out + k.dst);
            zs.avail_out = (uInt)k.isize;
            const int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0) bad.store(1);
        }
        inflateEnd(&zs);
    });
    return bad.load() == 0;
}

}  // namespace

extern "C" {

int gnx_vcf_open(gnx_vcf_t** out, const char* path, const char* chm, int threads) {
    if (!out || !path) {
        gnx::set_error("gnx_vcf_open: NULL argument");
        return 2;
    }
    *out = nullptr;
    gnx_vcf* v = new gnx_vcf();
    v->threads = threads;
    size_t n = 0;
    {
        // plain text is read directly; gzip / bgzip (magic 1f 8b) is inflated by zlib
        FILE* fp = fopen(path, "rb");
        if (!fp) {
            delete v;
            gnx::set_error("gnx_vcf_open: cannot open %s", path);
            return 1;
        }
        unsigned char magic[2] = {0, 0};
        const size_t mg = fread(magic, 1, 2, fp);
        fseek(fp, 0, SEEK_END);
        const long fsize = std::max<long>(ftell(fp), 0);
        fseek(fp, 0, SEEK_SET);
        const bool gz = (mg == 2 && magic[0] == 0x1f && magic[1] == 0x8b);
        auto grow = [&](size_t cap) -> bool {
            char* p = static_cast<char*>(realloc(v->text, cap));
            if (!p) return false;
            v->text = p;
            v->text_cap = cap;
            return true;
        };
        bool ok = true;
        if (!gz) {
            ok = grow((size_t)fsize + 2);
            if (ok) n = fread(v->text, 1, (size_t)fsize, fp);
            fclose(fp);
        } else {
            // BGZF members inflate in parallel; any other gzip file goes through one zlib stream
            std::vector<unsigned char> img((size_t)fsize);
            const size_t got_img = fread(img.data(), 1, (size_t)fsize, fp);
            fclose(fp);
            std::vector<BgzfBlock> blocks;
            size_t total = 0;
            bool done = false;
            if (got_img == (size_t)fsize && bgzf_index(img.data(), img.size(), blocks, total)) {
                ok = grow(total + 2);
                if (ok && bgzf_inflate(img.data(), blocks, v->text, threads)) {
                    n = total;
                    done = true;
                }
            }
            if (ok && !done) {
                img.clear();
                img.shrink_to_fit();
                gzFile f = gzopen(path, "rb");
                ok = f != nullptr && grow(std::max<size_t>(size_t(64) << 20, (size_t)fsize * 8));  // genotype text inflates ~6x
                if (f) {
                    gzbuffer(f, 1u << 20);
                    while (ok) {
                        if (v->text_cap - n < (size_t(8) << 20)) ok = grow(v->text_cap * 2);
                        if (!ok) break;
                        const int got = gzread(f, v->text + n, (unsigned)std::min<size_t>(v->text_cap - n - 2, size_t(1) << 30));
                        if (got < 0) ok = false;
                        if (got <= 0) break;
                        n += (size_t)got;
                    }
                    gzclose(f);
                }
            }
        }
        if (!ok) {
            delete v;
            gnx::set_error("gnx_vcf_open: cannot read %s (I/O error or out of memory)", path);
            return 1;
        }
    }
    if (n == 0 || v->text[n - 1] != '\n') v->text[n++] = '\n';
    v->text_len = n;
    const char* base = v->text;
    const char* end = base + n;

    // ---- line index (records only), header
    std::vector<int64_t>& ls = v->ls;
    {
        const char* p = base;
        while (p < end) {
            const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            if (p[0] == '#') {
                if (nl - p > 6 && !memcmp(p, "#CHROM", 6)) {
                    const char* e = nl;
                    if (e > p && e[-1] == '\r') e--;
                    const char* q = p;
                    int col = 0;
                    v->samples.clear();
                    while (q <= e) {
                        const char* t = find_tab(q, e);
                        if (col >= 9) v->samples.emplace_back(q, t);
                        col++;
                        q = t + 1;
                    }
                }
            } else if (nl > p) {
                ls.push_back(p - base);
            }
            p = nl + 1;
        }
    }
    v->n_smp = (int64_t)v->samples.size();
    const int64_t nl_total = (int64_t)ls.size();
    const size_t chm_len = chm ? strlen(chm) : 0;

    // ---- pass A: which lines are records we keep (>= 10 columns, CHROM filter)
    std::vector<uint8_t>& keep = v->keep;
    keep.assign((size_t)nl_total, 0);
    const int64_t blk = 4096;
    const int64_t nblk = (nl_total + blk - 1) / blk;
    std::vector<int64_t>& blk_count = v->blk_count;
    blk_count.assign((size_t)nblk + 1, 0);
    gnx::parallel_for(nblk, threads, [&](int64_t b) {
        int64_t cnt = 0;
        for (int64_t i = b * blk; i < std::min(nl_total, (b + 1) * blk); i++) {
            const char* p = base + ls[i];
            const char* e = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            const char* t = find_tab(p, e);
            if (chm && !((size_t)(t - p) == chm_len && !memcmp(p, chm, chm_len))) continue;
            int tabs = 0;
            const char* q = p;
            while (tabs < 9) {
                q = static_cast<const char*>(memchr(q, '\t', (size_t)(e - q)));
                if (!q) break;
                tabs++;
                q++;
            }
            if (tabs < 9) continue;
            keep[i] = 1;
            cnt++;
        }
        blk_count[b + 1] = cnt;
    });
    for (int64_t b = 0; b < nblk; b++) blk_count[b + 1] += blk_count[b];
    v->n_rec = blk_count[nblk];
    *out = v;
    return 0;
}

void gnx_vcf_close(gnx_vcf_t* v) { delete v; }

int64_t gnx_vcf_num_records(const gnx_vcf_t* v) { return v ? v->n_rec : -1; }
int64_t gnx_vcf_num_samples(const gnx_vcf_t* v) { return v ? v->n_smp : -1; }

/* Pass B: parses the kept records straight into the caller's arrays -- gt [records][samples][2] int8
 * (-1 = missing), pos [records] int32, qual [records] float32; any may be NULL -- and records where the
 * string columns sit. */
int gnx_vcf_copy(gnx_vcf_t* v, int8_t* gt, int32_t* pos, float* qual) {
    if (!v) return 2;
    const int64_t R = v->n_rec, S = v->n_smp;
    const char* base = v->text;
    const char* end = base + v->text_len;
    const int64_t nl_total = (int64_t)v->ls.size();
    const int64_t blk = 4096;
    const int64_t nblk = (nl_total + blk - 1) / blk;
    for (int k = 0; k < 4; k++) {
        v->off[k].resize((size_t)R);
        v->len[k].resize((size_t)R);
    }
    std::atomic<int> bad_pos{0};
    gnx::parallel_for(nblk, v->threads, [&](int64_t b) {
        int64_t r = v->blk_count[b];
        for (int64_t i = b * blk; i < std::min(nl_total, (b + 1) * blk); i++) {
            if (!v->keep[i]) continue;
            const char* p = base + v->ls[i];
            const char* e = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            if (e > p && e[-1] == '\r') e--;
            const char* col[10];
            const char* q = p;
            for (int c = 0; c < 9; c++) {
                col[c] = q;
                q = find_tab(q, e) + 1;
            }
            col[9] = q;
            auto span = [&](int c, int k) {
                v->off[k][r] = col[c] - base;
                v->len[k][r] = (int32_t)((col[c + 1] - 1) - col[c]);
            };
            span(0, 0);  // CHROM
            span(2, 1);  // ID
            span(3, 2);  // REF
            span(4, 3);  // ALT
            if (pos) {
                long long pv = 0;
                const char* s = col[1];
                const char* se = col[2] - 1;
                bool ok = s < se;
                for (; s < se; s++) {
                    if (*s < '0' || *s > '9') { ok = false; break; }
                    pv = pv * 10 + (*s - '0');
                    if (pv > 2147483647LL) { ok = false; break; }
                }
                if (!ok) bad_pos.store(1);
                pos[r] = (int32_t)pv;
            }
            if (qual) {
                const char* s = col[5];
                const char* se = col[6] - 1;
                if (se - s == 0 || (se - s == 1 && *s == '.')) {
                    qual[r] = NAN;
                } else {
                    char tmp[64];
                    const size_t l = std::min<size_t>((size_t)(se - s), sizeof tmp - 1);
                    memcpy(tmp, s, l);
                    tmp[l] = 0;
                    qual[r] = strtof(tmp, nullptr);
                }
            }
            if (gt) {
                int8_t* g = gt + (size_t)r * S * 2;
                const char* s = col[9];
                int64_t k = 0;
                for (; k < S && s <= e; k++) {
                    const char* t = find_tab(s, e);
                    parse_gt(s, t, g + 2 * k);
                    s = t + 1;
                }
                for (; k < S; k++) g[2 * k] = g[2 * k + 1] = -1;   // short record: missing calls
            }
            r++;
        }
    });
    v->parsed = true;
    if (bad_pos.load()) {
        gnx::set_error("gnx_vcf_copy: a POS column is not a non-negative 32-bit integer");
        return 1;
    }
    return 0;
}

/* String columns: field 0 CHROM, 1 ID, 2 REF, 3 ALT (comma-separated as in the file), 4 sample names.
 * Writes every string followed by a newline into buf (if cap suffices); returns the bytes needed. */
int64_t gnx_vcf_strings(gnx_vcf_t* v, int field, char* buf, int64_t cap) {
    if (!v || field < 0 || field > 4) return -1;
    if (field < 4 && !v->parsed && gnx_vcf_copy(v, nullptr, nullptr, nullptr)) return -1;
    int64_t need = 0;
    if (field == 4) {
        for (const auto& s : v->samples) need += (int64_t)s.size() + 1;
        if (buf && cap >= need) {
            int64_t o = 0;
            for (size_t i = 0; i < v->samples.size(); i++) {
                memcpy(buf + o, v->samples[i].data(), v->samples[i].size());
                o += (int64_t)v->samples[i].size();
                buf[o++] = '\n';
            }
        }
        return need;
    }
    for (int64_t r = 0; r < v->n_rec; r++) need += v->len[field][r] + 1;
    if (buf && cap >= need) {
        int64_t o = 0;
        for (int64_t r = 0; r < v->n_rec; r++) {
            memcpy(buf + o, v->text + v->off[field][r], (size_t)v->len[field][r]);
            o += v->len[field][r];
            buf[o++] = '\n';
        }
    }
    return need;
}


/* Genotype block -> aligned haplotype matrix (reference src/utils.py:104-159, vcf_to_npy after the SNP
 * intersection): X[2s + h][fmt_idx[k]] = GT[vcf_idx[k]][s][h] for k < n_idx, flipped 0 <-> 1 where
 * swap[k] (reference alleles disagree), everything that is not 0 / 1 and every column outside fmt_idx
 * = miss_fill.  gt [R][S][2] int8, X [2S][ldX] int8 (C columns written).  Parallel over blocks of 64
 * haplotype rows: each record contributes one 64-byte read per block. */
int gnx_vcf_to_haplotypes(const int8_t* gt, int64_t R, int64_t S, const int64_t* vcf_idx, const int64_t* fmt_idx,
                          const uint8_t* swap, int64_t n_idx, int64_t C, int miss_fill, int8_t* X, int64_t ldX, int threads) {
    if (R < 0 || S < 0 || n_idx < 0 || C < 0 || ldX < C || (n_idx > 0 && (!gt || !vcf_idx || !fmt_idx)) || (2 * S * C > 0 && !X)) {
        gnx::set_error("gnx_vcf_to_haplotypes: bad arguments");
        return 2;
    }
    for (int64_t k = 0; k < n_idx; k++)
        if (vcf_idx[k] < 0 || vcf_idx[k] >= R || fmt_idx[k] < 0 || fmt_idx[k] >= C) {
            gnx::set_error("gnx_vcf_to_haplotypes: index %lld out of range", (long long)k);
            return 2;
        }
    const int64_t H = 2 * S;
    const int64_t nblk = (H + 63) / 64;
    const int8_t mf = (int8_t)miss_fill;
    gnx::parallel_for(nblk, threads, [&](int64_t b) {
        const int64_t h0 = b * 64, hn = std::min<int64_t>(64, H - h0);
        for (int64_t h = 0; h < hn; h++) memset(X + (h0 + h) * ldX, mf, (size_t)C);
        for (int64_t k = 0; k < n_idx; k++) {
            const int8_t* src = gt + vcf_idx[k] * H + h0;
            int8_t* dst = X + h0 * ldX + fmt_idx[k];
            const bool sw = swap && swap[k];
            for (int64_t h = 0; h < hn; h++) {
                const int8_t v = src[h];
                dst[h * ldX] = (v == 0 || v == 1) ? (int8_t)(sw ? 1 - v : v) : mf;
            }
        }
    });
    return 0;
}

/* vcf_to_npy straight into 2-bit planes (include/gnx.h).  Blocks of 64 haplotype rows; with ascending fmt_idx (the
 * usual case: both position lists are sorted) a 64 x 4096 int8 tile that stays in L2 is filled like the int8 matrix
 * would be and packed row by row with the pack kernels of host_pack.cpp; otherwise bits are set one by one. */
int gnx_vcf_to_haplotypes_packed(const int8_t* gt, int64_t R, int64_t S, const int64_t* vcf_idx, const int64_t* fmt_idx,
                                 const uint8_t* swap, int64_t n_idx, int64_t C, int miss_fill, uint64_t* packed, int64_t pitch_words,
                                 int threads) {
    if (R < 0 || S < 0 || n_idx < 0 || C < 0 || pitch_words < 2 * ((C + 63) / 64) || (n_idx > 0 && (!gt || !vcf_idx || !fmt_idx)) ||
        (2 * S * C > 0 && !packed) || miss_fill < 0 || miss_fill > 3) {
        gnx::set_error("gnx_vcf_to_haplotypes_packed: bad arguments");
        return 2;
    }
    bool sorted = true;
    for (int64_t k = 0; k < n_idx; k++) {
        if (vcf_idx[k] < 0 || vcf_idx[k] >= R || fmt_idx[k] < 0 || fmt_idx[k] >= C) {
            gnx::set_error("gnx_vcf_to_haplotypes_packed: index %lld out of range", (long long)k);
            return 2;
        }
        if (k && fmt_idx[k] < fmt_idx[k - 1]) sorted = false;
    }
    const int64_t H = 2 * S;
    const int64_t nblk = (H + 63) / 64;
    const int8_t mf = (int8_t)miss_fill;
    const int64_t groups = pitch_words / 2;
    constexpr int64_t TC = 4096;
    gnx::parallel_for(nblk, threads, [&](int64_t b) {
        const int64_t h0 = b * 64, hn = std::min<int64_t>(64, H - h0);
        if (sorted) {
            std::vector<int8_t> tile((size_t)64 * TC);
            int64_t kk = 0;
            for (int64_t c0 = 0; c0 < C || c0 == 0; c0 += TC) {
                const int64_t cn = std::min(TC, C - c0);
                const bool last = c0 + TC >= C;
                for (int64_t h = 0; h < hn; h++) memset(tile.data() + h * TC, mf, (size_t)std::max<int64_t>(cn, 0));
                for (; kk < n_idx && fmt_idx[kk] < c0 + cn; kk++) {
                    const int8_t* src = gt + vcf_idx[kk] * H + h0;
                    int8_t* dst = tile.data() + (fmt_idx[kk] - c0);
                    const bool sw = swap && swap[kk];
                    for (int64_t h = 0; h < hn; h++) {
                        const int8_t v = src[h];
                        dst[h * TC] = (v == 0 || v == 1) ? (int8_t)(sw ? 1 - v : v) : mf;
                    }
                }
                const int64_t g0 = c0 / 64;
                const int64_t gn = last ? groups - g0 : TC / 64;
                for (int64_t h = 0; h < hn; h++)
                    gnx::pack_row_best(tile.data() + h * TC, std::max<int64_t>(cn, 0), packed + (h0 + h) * pitch_words + 2 * g0, gn);
                if (last) break;
            }
            return;
        }
        const uint64_t f0 = (mf & 1) ? ~0ull : 0ull, f1 = (mf & 2) ? ~0ull : 0ull;
        for (int64_t h = 0; h < hn; h++) {
            uint64_t* row = packed + (h0 + h) * pitch_words;
            for (int64_t g = 0; g < groups; g++) {
                const int64_t left = C - 64 * g;
                const uint64_t m = left >= 64 ? ~0ull : (left > 0 ? ((~0ull) >> (64 - left)) : 0ull);
                row[2 * g] = f0 & m;
                row[2 * g + 1] = f1 & m;
            }
        }
        for (int64_t k = 0; k < n_idx; k++) {
            const int8_t* src = gt + vcf_idx[k] * H + h0;
            const int64_t g = fmt_idx[k] / 64;
            const uint64_t bit = 1ull << (fmt_idx[k] % 64);
            const bool sw = swap && swap[k];
            for (int64_t h = 0; h < hn; h++) {
                const int8_t v = src[h];
                const unsigned o = (v == 0 || v == 1) ? (unsigned)(sw ? 1 - v : v) : (unsigned)mf;
                uint64_t* w = packed + (h0 + h) * pitch_words + 2 * g;
                w[0] = (w[0] & ~bit) | ((o & 1u) ? bit : 0ull);
                w[1] = (w[1] & ~bit) | ((o & 2u) ? bit : 0ull);
            }
        }
    });
    return 0;
}

}  // extern "C"
