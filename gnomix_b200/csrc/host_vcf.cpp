// host_vcf.cpp -- input side of run_inference on the host cores: VCF(.gz) -> genotype calls.
//
// Replaces the scikit-allel call behind read_vcf (reference src/utils.py:55-81; the fields the
// reference reads: calldata/GT, variants/POS|REF|ALT|CHROM|ID|QUAL, samples).  The reference's
// own notebook names this step as the largest part of an inference run (demo.ipynb:236).  The
// file is inflated with zlib (plain, gzip and bgzip all go through gzread), the line index is
// built once, and records are parsed in parallel on the library's worker pool straight into the
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

#include <atomic>
#include <string>
#include <vector>

#include "../../include/gnx.h"
#include "host_pack.h"

namespace gnx {
void set_error(const char* fmt, ...);
}

struct gnx_vcf {
    std::vector<char> text;
    std::vector<std::string> samples;
    int64_t n_rec = 0, n_smp = 0;
    std::vector<int8_t> gt;         // [n_rec][n_smp][2]
    std::vector<int32_t> pos;
    std::vector<float> qual;
    // string columns as (offset, length) into text: CHROM, ID, REF, ALT (whole column)
    std::vector<int64_t> off[4];
    std::vector<int32_t> len[4];
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

}  // namespace

extern "C" {

int gnx_vcf_open(gnx_vcf_t** out, const char* path, const char* chm, int threads) {
    if (!out || !path) {
        gnx::set_error("gnx_vcf_open: NULL argument");
        return 2;
    }
    *out = nullptr;
    gnx_vcf* v = new gnx_vcf();
    std::vector<char>& text = v->text;
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
        const long fsize = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        const bool gz = (mg == 2 && magic[0] == 0x1f && magic[1] == 0x8b);
        if (!gz) {
            text.resize((size_t)std::max<long>(fsize, 0) + 2);
            n = fread(text.data(), 1, (size_t)std::max<long>(fsize, 0), fp);
            fclose(fp);
        } else {
            fclose(fp);
            gzFile f = gzopen(path, "rb");
            if (!f) {
                delete v;
                gnx::set_error("gnx_vcf_open: cannot open %s", path);
                return 1;
            }
            gzbuffer(f, 1u << 20);
            text.resize(std::max<size_t>(size_t(64) << 20, (size_t)fsize * 8));  // VCF genotype text inflates ~6x
            for (;;) {
                if (text.size() - n < (size_t(8) << 20)) text.resize(text.size() * 2);
                const int got = gzread(f, text.data() + n, (unsigned)std::min<size_t>(text.size() - n, size_t(1) << 30));
                if (got < 0) {
                    gzclose(f);
                    delete v;
                    gnx::set_error("gnx_vcf_open: read error in %s", path);
                    return 1;
                }
                if (got == 0) break;
                n += (size_t)got;
            }
            gzclose(f);
        }
    }
    if (n == 0 || text[n - 1] != '\n') text[n++] = '\n';
    text.resize(n);
    const char* base = text.data();
    const char* end = base + n;

    // ---- line index (records only), header
    std::vector<int64_t> ls;  // line starts of candidate records
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
    std::vector<uint8_t> keep((size_t)nl_total, 0);
    const int64_t blk = 4096;
    const int64_t nblk = (nl_total + blk - 1) / blk;
    std::vector<int64_t> blk_count((size_t)nblk + 1, 0);
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
    const int64_t R = blk_count[nblk], S = v->n_smp;
    v->n_rec = R;
    v->gt.assign((size_t)R * S * 2, -1);
    v->pos.resize((size_t)R);
    v->qual.resize((size_t)R);
    for (int k = 0; k < 4; k++) {
        v->off[k].resize((size_t)R);
        v->len[k].resize((size_t)R);
    }
    std::atomic<int> bad_pos{0};

    // ---- pass B: parse
    gnx::parallel_for(nblk, threads, [&](int64_t b) {
        int64_t r = blk_count[b];
        for (int64_t i = b * blk; i < std::min(nl_total, (b + 1) * blk); i++) {
            if (!keep[i]) continue;
            const char* p = base + ls[i];
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
            {
                long long pv = 0;
                const char* s = col[1];
                const char* se = col[2] - 1;
                bool ok = s < se;
                for (; s < se; s++) {
                    if (*s < '0' || *s > '9') { ok = false; break; }
                    pv = pv * 10 + (*s - '0');
                }
                if (!ok || pv > 2147483647LL) bad_pos.store(1);
                v->pos[r] = (int32_t)pv;
            }
            {
                const char* s = col[5];
                const char* se = col[6] - 1;
                if (se - s == 0 || (se - s == 1 && *s == '.')) {
                    v->qual[r] = NAN;
                } else {
                    char tmp[64];
                    const size_t l = std::min<size_t>((size_t)(se - s), sizeof tmp - 1);
                    memcpy(tmp, s, l);
                    tmp[l] = 0;
                    v->qual[r] = strtof(tmp, nullptr);
                }
            }
            int8_t* g = v->gt.data() + (size_t)r * S * 2;
            const char* s = col[9];
            for (int64_t k = 0; k < S && s <= e; k++) {
                const char* t = find_tab(s, e);
                parse_gt(s, t, g + 2 * k);
                s = t + 1;
            }
            r++;
        }
    });
    if (bad_pos.load()) {
        delete v;
        gnx::set_error("gnx_vcf_open: a POS column of %s is not a non-negative 32-bit integer", path);
        return 1;
    }
    *out = v;
    return 0;
}

void gnx_vcf_close(gnx_vcf_t* v) { delete v; }

int64_t gnx_vcf_num_records(const gnx_vcf_t* v) { return v ? v->n_rec : -1; }
int64_t gnx_vcf_num_samples(const gnx_vcf_t* v) { return v ? v->n_smp : -1; }

/* gt [records][samples][2] int8 (-1 = missing), pos [records] int32, qual [records] float32; any may be NULL */
int gnx_vcf_copy(const gnx_vcf_t* v, int8_t* gt, int32_t* pos, float* qual) {
    if (!v) return 2;
    if (gt && !v->gt.empty()) memcpy(gt, v->gt.data(), v->gt.size());
    if (pos && v->n_rec) memcpy(pos, v->pos.data(), sizeof(int32_t) * (size_t)v->n_rec);
    if (qual && v->n_rec) memcpy(qual, v->qual.data(), sizeof(float) * (size_t)v->n_rec);
    return 0;
}

/* String columns: field 0 CHROM, 1 ID, 2 REF, 3 ALT (comma-separated as in the file), 4 sample names.
 * Writes every string followed by a newline into buf (if cap suffices); returns the bytes needed. */
int64_t gnx_vcf_strings(const gnx_vcf_t* v, int field, char* buf, int64_t cap) {
    if (!v || field < 0 || field > 4) return -1;
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
            memcpy(buf + o, v->text.data() + v->off[field][r], (size_t)v->len[field][r]);
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

}  // extern "C"
