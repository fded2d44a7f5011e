"""Input side of gnomix.py's run_inference (reference src/utils.py:55-182): VCF -> the aligned
int8 haplotype block the Base stage consumes, and the genetic map.  scikit-allel is not a
dependency: `read_vcf` returns the same dictionary keys the reference reads from allel
(calldata/GT, variants/POS|REF|ALT|CHROM|ID|QUAL, samples)."""
from __future__ import annotations

import gzip

import numpy as np


def _open(path):
    return gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")


def _parse_gt_block(tails, n_samples):
    """tails: list of bytes, the sample section of each record.  Fast path when every record is
    plain phased/unphased diploid `a|b` with single-digit alleles; otherwise per-field parsing."""
    L = 4 * n_samples - 1
    if all(len(t) == L for t in tails):
        a = np.frombuffer(b"".join(tails), dtype=np.uint8).reshape(len(tails), L)
        sep, tab = a[:, 1::4], a[:, 3::4]
        if np.all((sep == ord("|")) | (sep == ord("/"))) and (tab.size == 0 or np.all(tab == ord("\t"))):
            g = np.stack([a[:, 0::4], a[:, 2::4]], axis=2).astype(np.int16) - ord("0")
            ok = ((g >= 0) & (g <= 9)) | (g == ord(".") - ord("0"))
            if np.all(ok):
                g[g == ord(".") - ord("0")] = -1
                return g.astype(np.int8)
    out = np.full((len(tails), n_samples, 2), -1, dtype=np.int8)
    for r, t in enumerate(tails):
        for s, fld in enumerate(t.split(b"\t")[:n_samples]):
            gt = fld.split(b":", 1)[0].replace(b"/", b"|").split(b"|")
            for h in range(min(2, len(gt))):
                if gt[h] != b"." and gt[h] != b"":
                    out[r, s, h] = int(gt[h])
            if len(gt) == 1:  # haploid call: allel leaves the second allele missing
                out[r, s, 1] = -1
    return out


def _strings(lib, h, field, count):
    import ctypes as C
    need = int(lib.gnx_vcf_strings(h, field, None, 0))
    buf = C.create_string_buffer(max(need, 1))
    lib.gnx_vcf_strings(h, field, buf, need)
    parts = buf.raw[:need].decode().split("\n")[:count]
    out = np.empty(count, dtype=object)
    out[:] = parts
    return out


def read_vcf(vcf_file, chm=None, fields=None, verbose=False):
    """src/utils.py:55-81 through the library's parallel parser (gnx_vcf_open, csrc/host_vcf.cpp):
    the same dictionary read_vcf_py builds, several times faster on cohort-sized files."""
    import ctypes as C
    from . import _lib
    lib = _lib.lib()
    h = C.c_void_p()
    _lib.check(lib.gnx_vcf_open(C.byref(h), str(vcf_file).encode(), None if chm is None else str(chm).encode(), 0), "gnx_vcf_open")
    try:
        R, S = int(lib.gnx_vcf_num_records(h)), int(lib.gnx_vcf_num_samples(h))
        if R == 0:
            if chm is None:
                print("No data found in vcf file {}".format(vcf_file))
                return None
            print('Found no data in vcf file {} in region labeled "{}". Using all data from vcf instead...'.format(vcf_file, chm))
            return read_vcf(vcf_file, None, fields, verbose)
        gt = np.empty((R, S, 2), dtype=np.int8)
        pos = np.empty(R, dtype=np.int32)
        qual = np.empty(R, dtype=np.float32)
        _lib.check(lib.gnx_vcf_copy(h, gt.ctypes.data, pos.ctypes.data, qual.ctypes.data), "gnx_vcf_copy")
        alts = _strings(lib, h, 3, R)
        alt = np.full((R, 3), "", dtype=object)
        alt[:, 0] = alts
        for i in np.flatnonzero(np.char.find(alts.astype(str), ",") >= 0):     # multi-allelic records (rare)
            alt[i, :] = ""
            for j, v in enumerate(alts[i].split(",")[:3]):
                alt[i, j] = v
        data = {
            "samples": _strings(lib, h, 4, S),
            "calldata/GT": gt,
            "variants/CHROM": _strings(lib, h, 0, R),
            "variants/POS": pos,
            "variants/ID": _strings(lib, h, 1, R),
            "variants/REF": _strings(lib, h, 2, R),
            "variants/ALT": alt,
            "variants/QUAL": qual,
        }
    finally:
        lib.gnx_vcf_close(h)
    if verbose:
        print("File read:", R, "SNPs for", S, "individuals")
    return data


def read_vcf_py(vcf_file, chm=None, fields=None, verbose=False):
    """src/utils.py:55-81, pure Python/numpy (the cross-check of the native parser).  `chm` selects
    records whose CHROM equals it (allel's region=); if none match, the whole file is used, with the
    reference's message."""
    chroms, poss, ids, refs, alts, quals, tails = [], [], [], [], [], [], []
    samples = None
    want = None if chm is None else str(chm).encode()
    with _open(vcf_file) as f:
        for line in f:
            if line.startswith(b"##"):
                continue
            if line.startswith(b"#CHROM"):
                samples = line.rstrip(b"\r\n").split(b"\t")[9:]
                continue
            parts = line.rstrip(b"\r\n").split(b"\t", 9)
            if len(parts) < 10:
                continue
            if want is not None and parts[0] != want:
                continue
            chroms.append(parts[0]); poss.append(parts[1]); ids.append(parts[2]); refs.append(parts[3])
            alts.append(parts[4]); quals.append(parts[5])
            tails.append(parts[9] if parts[8] == b"GT" else b"\t".join(fld.split(b":", 1)[0] for fld in parts[9].split(b"\t")))
    if not chroms:
        if chm is None:
            print("No data found in vcf file {}".format(vcf_file))
            return None
        print('Found no data in vcf file {} in region labeled "{}". Using all data from vcf instead...'.format(vcf_file, chm))
        return read_vcf_py(vcf_file, None, fields, verbose)
    n = len(samples)
    gt = np.concatenate([_parse_gt_block(tails[i:i + 20000], n) for i in range(0, len(tails), 20000)], axis=0)
    alt = np.full((len(alts), 3), "", dtype=object)
    for i, a in enumerate(alts):
        for j, v in enumerate(a.decode().split(",")[:3]):
            alt[i, j] = v
    data = {
        "samples": np.array([s.decode() for s in samples], dtype=object),
        "calldata/GT": gt,
        "variants/CHROM": np.array([c.decode() for c in chroms], dtype=object),
        "variants/POS": np.array([int(p) for p in poss], dtype=np.int32),
        "variants/ID": np.array([i.decode() for i in ids], dtype=object),
        "variants/REF": np.array([r.decode() for r in refs], dtype=object),
        "variants/ALT": alt,
        "variants/QUAL": np.array([np.nan if q in (b".", b"") else float(q) for q in quals], dtype=np.float32),
    }
    if verbose:
        print("File read:", gt.shape[0], "SNPs for", n, "individuals")
    return data


def snp_intersection(pos1, pos2, verbose=False):
    """src/utils.py:83-101."""
    assert len(pos2) != 0, "No SNPs of specified chromosome found in query file."
    intersection, idx1, idx2 = np.intersect1d(pos1, pos2, return_indices=True)
    if verbose:
        print("- Number of SNPs from model:", len(pos1))
        print("- Number of SNPs from file:", len(pos2))
        print("- Number of intersecting SNPs:", len(intersection))
        print("- Percentage of model SNPs covered by query file: ", round(len(intersection) / len(pos1), 4) * 100, "%", sep="")
    return idx1, idx2


def vcf_to_npy(vcf_data, snp_pos_fmt=None, snp_ref_fmt=None, miss_fill=2, return_idx=False, verbose=True):
    """src/utils.py:104-159: align to the model's SNP positions (missing positions = miss_fill),
    flip 0/1 where the reference alleles disagree, anything that is not 0/1 -> miss_fill; int8
    [2 * individuals, C] with rows 2i, 2i+1 = individual i.  The transposing gather runs on the
    library's host threads (gnx_vcf_to_haplotypes); vcf_to_npy_py is the numpy statement of it."""
    from . import _lib
    gt = vcf_data["calldata/GT"]
    if not (isinstance(gt, np.ndarray) and gt.dtype == np.int8 and gt.ndim == 3 and gt.shape[2] == 2):
        return vcf_to_npy_py(vcf_data, snp_pos_fmt, snp_ref_fmt, miss_fill, return_idx, verbose)
    gt = np.ascontiguousarray(gt)
    R, S, _ = gt.shape
    if snp_pos_fmt is not None:
        fmt_idx, vcf_idx = snp_intersection(snp_pos_fmt, vcf_data["variants/POS"], verbose=verbose)
        C = len(snp_pos_fmt)
    else:
        fmt_idx = vcf_idx = np.arange(R)
        C = R
    swap = None
    if snp_ref_fmt is not None:
        swap = np.asarray(vcf_data["variants/REF"])[vcf_idx] != np.asarray(snp_ref_fmt)[fmt_idx]
        if swap.any() and verbose:
            print("- Found ", int(swap.sum()), " (", round(np.mean(swap) * 100, 4), "%) different reference variants. Adjusting...", sep="")
        swap = np.ascontiguousarray(swap, dtype=np.uint8)
    vi = np.ascontiguousarray(vcf_idx, dtype=np.int64)
    fi = np.ascontiguousarray(fmt_idx, dtype=np.int64)
    mat = np.empty((2 * S, C), dtype=np.int8)
    _lib.check(_lib.lib().gnx_vcf_to_haplotypes(gt.ctypes.data, R, S, vi.ctypes.data, fi.ctypes.data,
                                                None if swap is None else swap.ctypes.data, len(vi), C, int(miss_fill),
                                                mat.ctypes.data, C, 0), "gnx_vcf_to_haplotypes")
    if return_idx:
        return mat, (vcf_idx if snp_pos_fmt is not None else np.arange(2 * S)), (fmt_idx if snp_pos_fmt is not None else np.arange(2 * S))
    return mat


class PackedHaplotypes:
    """Haplotype matrix [N, C] with values 0..3 held as two bit planes per 64 SNPs (the gnx_pack_rows_host layout:
    row = groups of {u64 bit 0, u64 bit 1}) in page-locked host memory: a quarter of the int8 matrix's bytes, and
    what gnx_infer_host_ex takes as is (x_packed).  `words` is the uint64 view [N, pitch_words]."""

    def __init__(self, N, C, pinned=True):
        import ctypes as Ct
        from . import _lib
        self.N, self.C = int(N), int(C)
        self.pitch_words = 2 * ((self.C + 127) // 128 * 2)      # 128-SNP multiple, as the device rows are pitched
        nbytes = self.N * self.pitch_words * 8
        self._ptr = None
        if pinned and nbytes:
            ptr = Ct.c_void_p()
            _lib.check(_lib.lib().gnx_host_alloc_pinned(Ct.byref(ptr), nbytes), "gnx_host_alloc_pinned")
            self._ptr = ptr
            buf = (Ct.c_uint64 * (nbytes // 8)).from_address(ptr.value)
            self.words = np.frombuffer(buf, dtype=np.uint64).reshape(self.N, self.pitch_words)
        else:
            self.words = np.zeros((self.N, self.pitch_words), dtype=np.uint64)

    @property
    def shape(self):
        return (self.N, self.C)

    def __len__(self):
        return self.N

    def __del__(self):
        try:
            if self._ptr is not None:
                from . import _lib
                self.words = None
                _lib.lib().gnx_host_free_pinned(self._ptr)
                self._ptr = None
        except Exception:
            pass

    @classmethod
    def from_numpy(cls, X, pinned=True):
        """int8 [N, C] with values in 0..3 -> packed (host cores, gnx_pack_rows_host)."""
        import ctypes as Ct
        from . import _lib
        X = np.ascontiguousarray(X, dtype=np.int8)
        out = cls(X.shape[0], X.shape[1], pinned=pinned)
        bad = Ct.c_int(0)
        rc = _lib.lib().gnx_pack_rows_host(X.ctypes.data, X.shape[0], X.strides[0] if X.shape[0] else X.shape[1], X.shape[1],
                                           out.words.ctypes.data, out.pitch_words, 0, Ct.byref(bad))
        if rc != 0 or bad.value:
            raise ValueError("haplotype values outside 0..3 cannot be packed to 2 bits")
        return out

    def to_numpy(self):
        """the int8 matrix back (numpy; for checks)."""
        w = np.ascontiguousarray(self.words).reshape(self.N, -1, 2)
        bits = np.unpackbits(w.view(np.uint8).reshape(self.N, -1, 2, 8), axis=-1, bitorder="little").reshape(self.N, -1, 2, 64)
        X = (bits[:, :, 0, :] | (bits[:, :, 1, :] << 1)).reshape(self.N, -1).astype(np.int8)
        return np.ascontiguousarray(X[:, :self.C])


def vcf_to_packed(vcf_data, snp_pos_fmt=None, snp_ref_fmt=None, miss_fill=2, return_idx=False, verbose=True, pinned=True):
    """vcf_to_npy (src/utils.py:104-159) with the result written straight into 2-bit planes in pinned memory
    (gnx_vcf_to_haplotypes_packed): the matrix Gnomix.predict_host / the driver ship to the GPU without packing
    or materialising int8.  `PackedHaplotypes.to_numpy()` equals vcf_to_npy's matrix."""
    from . import _lib
    gt = vcf_data["calldata/GT"]
    assert isinstance(gt, np.ndarray) and gt.ndim == 3 and gt.shape[2] == 2, "calldata/GT must be [records, samples, 2]"
    gt = np.ascontiguousarray(gt, dtype=np.int8)
    R, S, _ = gt.shape
    if snp_pos_fmt is not None:
        fmt_idx, vcf_idx = snp_intersection(snp_pos_fmt, vcf_data["variants/POS"], verbose=verbose)
        C = len(snp_pos_fmt)
    else:
        fmt_idx = vcf_idx = np.arange(R)
        C = R
    swap = None
    if snp_ref_fmt is not None:
        swap = np.asarray(vcf_data["variants/REF"])[vcf_idx] != np.asarray(snp_ref_fmt)[fmt_idx]
        if swap.any() and verbose:
            print("- Found ", int(swap.sum()), " (", round(np.mean(swap) * 100, 4), "%) different reference variants. Adjusting...", sep="")
        swap = np.ascontiguousarray(swap, dtype=np.uint8)
    vi = np.ascontiguousarray(vcf_idx, dtype=np.int64)
    fi = np.ascontiguousarray(fmt_idx, dtype=np.int64)
    out = PackedHaplotypes(2 * S, C, pinned=pinned)
    _lib.check(_lib.lib().gnx_vcf_to_haplotypes_packed(gt.ctypes.data, R, S, vi.ctypes.data, fi.ctypes.data,
                                                       None if swap is None else swap.ctypes.data, len(vi), C, int(miss_fill),
                                                       out.words.ctypes.data, out.pitch_words, 0), "gnx_vcf_to_haplotypes_packed")
    if return_idx:
        return out, (vcf_idx if snp_pos_fmt is not None else np.arange(2 * S)), (fmt_idx if snp_pos_fmt is not None else np.arange(2 * S))
    return out


def vcf_to_npy_py(vcf_data, snp_pos_fmt=None, snp_ref_fmt=None, miss_fill=2, return_idx=False, verbose=True):
    """src/utils.py:104-159 in numpy, statement by statement."""
    data = vcf_data["calldata/GT"]
    chm_len, n_ind, _ = data.shape
    data = data.reshape(chm_len, n_ind * 2).T
    mat = data
    vcf_idx, fmt_idx = np.arange(n_ind * 2), np.arange(n_ind * 2)
    if snp_pos_fmt is not None:
        fmt_idx, vcf_idx = snp_intersection(snp_pos_fmt, vcf_data["variants/POS"], verbose=verbose)
        mat = np.full((n_ind * 2, len(snp_pos_fmt)), miss_fill, dtype=np.int8)
        mat[:, fmt_idx] = data[:, vcf_idx]
    else:
        mat = np.array(mat, dtype=np.int8)
    if snp_ref_fmt is not None:
        swap = np.asarray(vcf_data["variants/REF"])[vcf_idx] != np.asarray(snp_ref_fmt)[fmt_idx]
        if swap.any() and verbose:
            print("- Found ", int(swap.sum()), " (", round(np.mean(swap) * 100, 4), "%) different reference variants. Adjusting...", sep="")
        fmt_swap_idx = np.array(fmt_idx)[swap]
        mat[:, fmt_swap_idx] = (mat[:, fmt_swap_idx] - 1) * (-1)
    mat[np.logical_and(mat != 0, mat != 1)] = miss_fill
    mat = mat.astype(np.int8)
    if return_idx:
        return mat, vcf_idx, fmt_idx
    return mat


def read_genetic_map(genetic_map_path, chm=None, header=None):
    """src/utils.py:161-182."""
    import pandas as pd
    df = pd.read_csv(genetic_map_path, delimiter="\t", header=header, comment="#", dtype=str)
    df.columns = ["chm", "pos", "pos_cm"]
    try:
        df = df.astype({"chm": str, "pos": int, "pos_cm": float})
    except ValueError:
        if header is None:
            print("WARNING: Something wrong with genetic map format. Trying with header...")
            return read_genetic_map(genetic_map_path, chm=chm, header=0)
        raise Exception("Genetic map format not understood.")
    if chm is not None:
        chm = str(chm)
        df = df[df.chm == "chr" + chm] if len(df[df.chm == chm]) == 0 else df[df.chm == chm]
    return df
