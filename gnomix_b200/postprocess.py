"""Output side of gnomix.py's run_inference (reference src/postprocess.py:25-126): the window
table and the .msp / .fb writers, byte-identical to the reference's files (tests/golden/meta.npz
holds files written by the reference's own functions) but without its per-SNP Python loop and
its [W, N] string-matrix detour."""
from __future__ import annotations

import numpy as np


def get_meta_data(chm, model_pos, query_pos, n_wind, wind_size, gen_map_df):
    """src/postprocess.py:25-67.  Returns the same pandas DataFrame (all-string cells, as the
    reference's mixed np.array produces): chm, spos, epos, sgpos, egpos, n snps."""
    import pandas as pd
    model_pos = np.asarray(model_pos)
    query_pos = np.asarray(query_pos)
    model_chm_len = len(model_pos)
    chm_array = [chm] * n_wind
    starts = np.arange(0, model_chm_len, wind_size)
    spos_idx = starts[:-1]
    epos_idx = np.concatenate([starts[1:-1], np.array([model_chm_len])]) - 1
    spos = model_pos[spos_idx]
    epos = model_pos[epos_idx]
    # linear interpolation of the genetic map, clamped to its end points (interp1d with fill_value)
    gpos, gcm = np.asarray(gen_map_df.pos, dtype=np.float64), np.asarray(gen_map_df.pos_cm, dtype=np.float64)
    sgpos = np.round(_interp_like_scipy(gpos, gcm, spos), 5)
    egpos = np.round(_interp_like_scipy(gpos, gcm, epos), 5)
    # query SNPs per window: the reference walks query_pos once, counting positions <= epos[w]
    cum = np.searchsorted(query_pos, epos[:n_wind - 1], side="right") if _is_sorted(query_pos) else _walk_counts(query_pos, epos, n_wind)
    n_snps = np.zeros_like(epos)
    n_snps[:n_wind - 1] = np.diff(np.concatenate([[0], cum]))
    n_snps[n_wind - 1] = len(query_pos) - (cum[-1] if n_wind > 1 else 0)
    meta_data = np.array([chm_array, spos, epos, sgpos, egpos, n_snps]).T
    df = pd.DataFrame(meta_data)
    df.columns = ["chm", "spos", "epos", "sgpos", "egpos", "n snps"]
    return df


def _is_sorted(a):
    return len(a) < 2 or bool(np.all(a[1:] >= a[:-1]))


def _walk_counts(query_pos, epos, n_wind):
    q, out = 0, []
    for w in range(n_wind - 1):
        while q < len(query_pos) and query_pos[q] <= epos[w]:
            q += 1
        out.append(q)
    return np.array(out, dtype=np.int64)


def _interp_like_scipy(x, y, xq):
    """scipy.interpolate.interp1d(x, y, fill_value=(y[0], y[-1]), bounds_error=False) on sorted x:
    slope form y_lo + slope * (xq - x_lo), which is what interp1d's linear kernel evaluates."""
    xq = np.asarray(xq, dtype=np.float64)
    idx = np.clip(np.searchsorted(x, xq), 1, len(x) - 1)
    lo, hi = idx - 1, idx
    slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
    out = slope * (xq - x[lo]) + y[lo]
    out = np.where(xq < x[0], y[0], out)
    out = np.where(xq > x[-1], y[-1], out)
    return out


def write_msp(msp_prefix, meta_data, pred_labels, populations, query_samples):
    """src/postprocess.py:84-98."""
    pred_labels = np.asarray(pred_labels)
    meta = np.asarray(meta_data).astype(str)
    W = meta.shape[0]
    with open(msp_prefix + ".msp", "wb") as f:
        f.write(("#Subpopulation order/codes: " + "\t".join([str(pop) + "=" + str(i) for i, pop in enumerate(populations)]) + "\n").encode())
        f.write(("#" + "\t".join(meta_data.columns) + "\t").encode())
        f.write(("\t".join([str(s) for s in np.concatenate([[s + ".0", s + ".1"] for s in query_samples])]) + "\n").encode())
        lab = pred_labels.T                                             # [W, N]
        if lab.size and lab.min() >= 0 and lab.max() <= 9:
            # single-digit labels: build each line's label section as bytes directly
            N = lab.shape[1]
            row = np.empty((W, 2 * N), dtype=np.uint8)
            row[:, 0::2] = ord("\t")
            row[:, 1::2] = lab.astype(np.uint8) + ord("0")
            for l in range(W):
                f.write("\t".join(meta[l]).encode())
                f.write(row[l].tobytes())
                f.write(b"\n")
        else:
            ls = lab.astype(str)
            for l in range(W):
                f.write(("\t".join(meta[l]) + "\t" + "\t".join(ls[l]) + "\n").encode())


def write_fb(fb_prefix, meta_data, proba, ancestry, query_samples):
    """src/postprocess.py:100-126 (same header, same shortest-repr number formatting).  The body --
    one line per window with N*A probabilities as text -- is formatted by the library's host
    threads (gnx_write_fb_body, csrc/host_io.cpp) instead of one Python object per number."""
    import ctypes as C
    from . import _lib
    proba = np.asarray(proba)
    if proba.dtype not in (np.float32, np.float64):
        proba = proba.astype(np.float64)
    proba = np.ascontiguousarray(proba)
    n_rows = meta_data.shape[0]
    N, W, A = proba.shape
    assert W == n_rows, "proba has %d windows, the window table %d" % (W, n_rows)
    pp = np.round(np.mean(np.array(meta_data[["spos", "epos"]], dtype=int), axis=1)).astype(int)
    gp = np.mean(np.array(meta_data[["sgpos", "egpos"]], dtype=float), axis=1).astype(float)
    chm = np.asarray(meta_data["chm"]).astype(str)
    header = ["chromosome", "physical position", "genetic_position", "genetic_marker_index"]
    header += [":::".join([q, h, a]) for q in query_samples for h in ["hap1", "hap2"] for a in ancestry]
    path = fb_prefix + ".fb"
    with open(path, "w") as f:
        f.write("#reference_panel_population:\t")
        f.write("\t".join(ancestry) + "\n")
        f.write("\t".join(header) + "\n")
    prefixes = [("\t".join([chm[l], str(pp[l]), repr(float(gp[l])), "."]) + "\t").encode() for l in range(n_rows)]
    arr = (C.c_char_p * n_rows)(*prefixes)
    _lib.check(_lib.lib().gnx_write_fb_body(path.encode(), 1, proba.ctypes.data, int(proba.dtype == np.float64), N, W, A,
                                            C.cast(arr, C.c_void_p), 0), "gnx_write_fb_body")


def msp_to_lai(msp_file, positions, lai_file=None):
    """src/postprocess.py:128-160 (BETA in the reference): per-SNP ancestry by repeating every window's row
    `n snps` times; same DataFrame and, with lai_file, the same file."""
    import pandas as pd
    positions = np.asarray(positions)
    msp_df = pd.read_csv(msp_file, sep="\t", comment="#", header=None)
    data_window = np.array(msp_df.iloc[:, 6:])
    n_reps = msp_df.iloc[:, 5].to_numpy()
    assert np.sum(n_reps) == len(positions)
    data_snp = np.repeat(data_window, n_reps, axis=0)
    pos_lower_bound = int(msp_df.iloc[0, 1])
    pos_upper_bound = int(msp_df.iloc[-1, 2])
    extrapolating_lo = np.sum(positions < pos_lower_bound)
    extrapolating_hi = np.sum(positions > pos_upper_bound)
    if extrapolating_lo > 0:
        print("WARNING: Extrapolating ancestry inference for {} SNPs (lower bound is position {})".format(extrapolating_lo, pos_lower_bound))
    if extrapolating_hi > 0:
        print("WARNING: Extrapolating ancestry inference for {} SNPs (upper bound is position {})".format(extrapolating_hi, pos_upper_bound))
    with open(msp_file) as f:
        first_line = f.readline()
        second_line = f.readline()
    samples = second_line[:-1].split("\t")[6:]
    df = pd.DataFrame(data_snp, columns=samples, index=positions)
    if lai_file is not None:
        with open(lai_file, "w") as f:
            f.write(first_line)
        df.to_csv(lai_file, sep="\t", mode="a", index_label="position")
    return df


def get_bed_data(msp_df, sample, pop_order=None):
    """src/postprocess.py:162-193: one row per maximal run of equal ancestry (change points found at once
    instead of a Python loop over windows)."""
    anc = np.asarray(msp_df[sample])
    start = np.concatenate([[0], np.flatnonzero(anc[1:] != anc[:-1]) + 1])
    stop = np.concatenate([start[1:] - 1, [len(anc) - 1]])
    label = (lambda v: v) if pop_order is None else (lambda v: pop_order[v])
    return {
        "chm": np.asarray(msp_df["#chm"])[start].astype(int),
        "spos": np.asarray(msp_df["spos"])[start].astype(int),
        "epos": np.asarray(msp_df["epos"])[stop].astype(int),
        "ancestry": [label(v) for v in anc[start]],
        "sgpos": list(np.asarray(msp_df["sgpos"])[start]),
        "egpos": list(np.asarray(msp_df["egpos"])[stop]),
    }


def msp_to_bed(msp_file, root, pop_order=None):
    """src/postprocess.py:195-210: one .bed per haplotype column.  Like the reference, the header line is
    split without stripping its newline, so the last haplotype's file name ends in "\\n.bed"."""
    import os
    import pandas as pd
    with open(msp_file) as f:
        _ = f.readline()
        second_line = f.readline()
    header = second_line.split("\t")
    msp_df = pd.read_csv(msp_file, sep="\t", comment="#", names=header)
    for sample in header[6:]:
        sample_file_name = os.path.join(root, sample.replace(".", "_") + ".bed")
        pd.DataFrame(get_bed_data(msp_df, sample, pop_order=pop_order)).to_csv(sample_file_name, sep="\t", index=False)
